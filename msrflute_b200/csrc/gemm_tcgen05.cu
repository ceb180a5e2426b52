// bf16 GEMM on the 5th-generation tensor cores: C[g] = A[g] (M x K) * B[g]^T (N x K)  (+bias, +ReLU)
//
// Both operands are K-major ("TN"), the layout every Linear / conv-as-GEMM / LSTM projection in this framework
// produces.  Two kernels: a PERSISTENT one (default for N > 64: one CTA per SM walks tiles, accumulator double-buffered
// in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1, BLOCK_N 128 or 256) and the original
// one-tile-per-CTA kernel below (N <= 64, and the reference the persistent kernel is tested against).
// Structure of the simple kernel (one 128 x BLOCK_N output tile per CTA, 192 threads):
//
//   warp 0      TMA producer  : cp.async.bulk.tensor.3d (global -> 128B-swizzled smem), 4-stage mbarrier ring
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BLOCK_N x 16),
//                               fp32 accumulator in TMEM; tcgen05.commit releases smem stages / signals the epilogue
//   warps 2..5  epilogue      : tcgen05.ld 32x32b (TMEM -> registers), + bias, ReLU, convert, 16-byte global stores
//
// M/N/K tails are handled by TMA out-of-bounds zero fill and store predication.  Batched via grid.z (3-D tensor maps).
// No CUTLASS: descriptors and PTX are spelled out below (layouts follow the PTX ISA "tcgen05 matrix descriptors").
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace flute {
namespace gemm {
using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kThreads = 192;

// Instruction descriptor, kind::f16: [4,6) D format (1 = f32) | [7,10) A format (1 = bf16) | [10,13) B format (1 = bf16) |
//   bit 15 / 16: A / B major (0 = K-major) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(umma_n >> 3) << 17) |
         (static_cast<uint32_t>(umma_m >> 4) << 24);
}

constexpr int kMnBlockBytes = BLOCK_K * 128;      // one MN-major block: 64 reduction rows x 64 bf16
// MN-major SWIZZLE_128B operand: 64 contiguous M/N elements per 128-byte row, 8 reduction rows per 1024-byte atom
// (stride byte offset), 64-wide M/N blocks kMnBlockBytes apart (leading byte offset)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(kMnBlockBytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct EpilogueParams {
  void* C;              // [G, M, N] bf16 or fp32, row-major
  const float* bias;    // [N] or nullptr
  int M, N, K;
  long long c_batch_stride;
  int relu;
  int out_fp32;
  // operand majorness (persistent kernel): 0 = K-major tile (rows = M/N index, 64 reduction elements per 128-byte row);
  // 1 = MN-major: the tensor is stored [reduction, M or N] (e.g. dY / X / W as they sit in memory for wgrad / dgrad) and
  // a tile is BLOCK/64 blocks of [64 reduction rows x 64 M/N elements] — no transposed copy is ever made
  int a_mn, b_mn;
};

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const EpilogueParams ep) {
  using L = SmemLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y, g = blockIdx.z;
  const int num_k_blocks = (ep.K + BLOCK_K - 1) / BLOCK_K;
  constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;      // power of two >= 32

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(empty_bar + s, phase ^ 1);
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        mbar_expect_tx(full_bar + s, L::kStageBytes);
        tma_load_3d(sa, &map_a, full_bar + s, kb * BLOCK_K, m_blk * BLOCK_M, g);
        tma_load_3d(sb, &map_b, full_bar + s, kb * BLOCK_K, n_blk * BLOCK_N, g);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single elected lane)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(full_bar + s, phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
        const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(b_addr);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                   (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar + s);                     // smem stage reusable once these MMAs retire
      }
      umma_commit(tmem_full_bar);                       // accumulator complete
    }
  } else {
    // ===================================================== epilogue: TMEM -> registers -> global
    const int q = warp & 3;                             // TMEM lane quadrant this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m_blk * BLOCK_M + q * 32 + lane;
    const long long c_off = static_cast<long long>(g) * ep.c_batch_stride + static_cast<long long>(row) * ep.N;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), v);
      const int col0 = n_blk * BLOCK_N + c0;
      if (row < ep.M && col0 < ep.N) {
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          if (ep.bias != nullptr && col0 + j < ep.N) x += ep.bias[col0 + j];
          if (ep.relu) x = fmaxf(x, 0.f);
          f[j] = x;
        }
        const bool full = (col0 + 32 <= ep.N);
        if (ep.out_fp32) {
          float* dst = reinterpret_cast<float*>(ep.C) + c_off + col0;
          if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32 && col0 + j < ep.N; ++j) dst[j] = f[j];
          }
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.C) + c_off + col0;
          if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < ep.N; ++j) dst[j] = __float2bfloat16(f[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------ persistent kernel
// One CTA per SM walks tiles (m fastest, so CTAs running side by side share the B tile in L2).  The accumulator is
// double-buffered in TMEM (2 x BLOCK_N columns): while the four epilogue warps drain tile i (tcgen05.ld -> bias/ReLU ->
// global), the MMA warp already accumulates tile i+1, and the TMA warp runs up to kStagesP k-blocks ahead across tile
// boundaries.  BLOCK_N = 256 halves the shared-memory operand traffic per MMA (a 128x128x16 UMMA reads 8 KB per 64
// tensor clocks = the whole smem bandwidth; 128x256x16 reads 12 KB per 128).
template <int BLOCK_N>
struct SmemLayoutP {
  static constexpr int kStagesP = BLOCK_N == 256 ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = kStagesP * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                               const EpilogueParams ep, const int m_tiles, const int n_tiles, const int num_tiles) {
  using L = SmemLayoutP<BLOCK_N>;
  constexpr int S = L::kStagesP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * L::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;          // [2] accumulator stage ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;          // [2] accumulator stage drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k_blocks = (ep.K + BLOCK_K - 1) / BLOCK_K;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;                       // 256 or 512: power of two

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, 4);                                  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_mn = m_tiles * n_tiles;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int kit = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = tile / tiles_mn, r = tile - g * tiles_mn, n_blk = r / m_tiles, m_blk = r - n_blk * m_tiles;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++kit) {
          const int s = kit % S;
          mbar_wait(empty_bar + s, ((kit / S) & 1) ^ 1);
          uint8_t* sa = smem + s * L::kStageBytes;
          mbar_expect_tx(full_bar + s, L::kStageBytes);
          if (ep.a_mn) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_3d(sa + j * kMnBlockBytes, &map_a, full_bar + s, m_blk * BLOCK_M + 64 * j, kb * BLOCK_K, g);
          } else {
            tma_load_3d(sa, &map_a, full_bar + s, kb * BLOCK_K, m_blk * BLOCK_M, g);
          }
          if (ep.b_mn) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_3d(sa + L::kABytes + j * kMnBlockBytes, &map_b, full_bar + s, n_blk * BLOCK_N + 64 * j, kb * BLOCK_K, g);
          } else {
            tma_load_3d(sa + L::kABytes, &map_b, full_bar + s, kb * BLOCK_K, n_blk * BLOCK_N, g);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single elected lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N) | (ep.a_mn ? (1u << 15) : 0u) | (ep.b_mn ? (1u << 16) : 0u);
      const uint64_t a_step = ep.a_mn ? 128u : 2u, b_step = ep.b_mn ? 128u : 2u;     // descriptor units (16 B) per UMMA_K
      int kit = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(acc_empty + acc, ((t >> 1) & 1) ^ 1);             // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb, ++kit) {
          const int s = kit % S;
          mbar_wait(full_bar + s, (kit / S) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint64_t adesc = ep.a_mn ? make_smem_desc_mn(a_addr) : make_smem_desc(a_addr);
          const uint64_t bdesc = ep.b_mn ? make_smem_desc_mn(a_addr + L::kABytes) : make_smem_desc(a_addr + L::kABytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar + s);
        }
        umma_commit(acc_full + acc);
      }
    }
  } else {
    // ===================================================== epilogue warps (2..5): TMEM -> registers -> global
    const int q = warp & 3;
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int g = tile / tiles_mn, r = tile - g * tiles_mn, n_blk = r / m_tiles, m_blk = r - n_blk * m_tiles;
      const int acc = t & 1;
      mbar_wait(acc_full + acc, (t >> 1) & 1);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const long long c_off = static_cast<long long>(g) * ep.c_batch_stride + static_cast<long long>(row) * ep.N;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N + c0), v);
        const int col0 = n_blk * BLOCK_N + c0;
        if (row < ep.M && col0 < ep.N) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]);
            if (ep.bias != nullptr && col0 + j < ep.N) x += ep.bias[col0 + j];
            if (ep.relu) x = fmaxf(x, 0.f);
            f[j] = x;
          }
          const bool full = (col0 + 32 <= ep.N);
          if (ep.out_fp32) {
            float* dst = reinterpret_cast<float*>(ep.C) + c_off + col0;
            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32 && col0 + j < ep.N; ++j) dst[j] = f[j];
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.C) + c_off + col0;
            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(dst + j) = pk;
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < ep.N; ++j) dst[j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();                                            // our tcgen05.ld's are complete (wait::ld inside)
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult qres;
    void* p = nullptr;
    FLUTE_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    TORCH_CHECK(qres == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 3-D map over a [G, R, K] bf16 row-major tensor; box = [1, box_rows, BLOCK_K], 128-byte swizzle, zero OOB fill.
static CUtensorMap make_map(const void* ptr, int64_t G, int64_t R, int64_t K, int64_t batch_stride_elems, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(R), static_cast<cuuint64_t>(G)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 2, static_cast<cuuint64_t>(batch_stride_elems) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", static_cast<int>(r));
  return m;
}

// 3-D map over a [G, R (reduction), C (M or N)] bf16 row-major tensor consumed MN-major; box = [1, 64 reduction rows, 64]
static CUtensorMap make_map_mn(const void* ptr, int64_t G, int64_t R, int64_t C, int64_t batch_stride_elems) {
  CUtensorMap m;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(R), static_cast<cuuint64_t>(G)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(batch_stride_elems) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(BLOCK_K), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (MN-major) failed with code ", static_cast<int>(r));
  return m;
}

template <int BLOCK_N>
static void launch(const CUtensorMap& ma, const CUtensorMap& mb, const EpilogueParams& ep, int G, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  dim3 grid((ep.M + BLOCK_M - 1) / BLOCK_M, (ep.N + BLOCK_N - 1) / BLOCK_N, G);
  gemm_bf16_tn_kernel<BLOCK_N><<<grid, kThreads, L::kTotal, stream>>>(ma, mb, ep);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}


template <int BLOCK_N>
static void launch_persistent(const CUtensorMap& ma, const CUtensorMap& mb, const EpilogueParams& ep, int G, cudaStream_t stream) {
  using L = SmemLayoutP<BLOCK_N>;
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    int dev = 0;
    FLUTE_CUDA_CHECK(cudaGetDevice(&dev));
    FLUTE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int m_tiles = (ep.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (ep.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles * G;
  const int grid = std::min(num_tiles, sms);
  gemm_bf16_tn_persistent_kernel<BLOCK_N><<<grid, kThreads, L::kTotal, stream>>>(ma, mb, ep, m_tiles, n_tiles, num_tiles);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// 0 = auto (persistent, BLOCK_N by shape), 1 = one-tile-per-CTA kernel, 2 = persistent BLOCK_N 128, 3 = persistent BLOCK_N 256
static int g_gemm_impl = 0;

}  // namespace gemm

// C[M, N] = A · Bᵀ with either operand given in its OTHER storage order (2-D, persistent kernel, N > 64):
//   a_mn: a is stored [K, M] (reduction-major rows, e.g. dY for wgrad);   b_mn: b is stored [K, N] (e.g. W for dgrad).
// The MN-major UMMA descriptors read those tiles in place, so linear-layer backward passes need no transposed copies.
torch::Tensor gemm_bf16_mn(torch::Tensor a, torch::Tensor b, bool a_mn, bool b_mn, bool out_fp32) {
  using namespace gemm;
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.scalar_type() == torch::kBFloat16 && b.scalar_type() == torch::kBFloat16 &&
              a.dim() == 2 && b.dim() == 2 && a.is_contiguous() && b.is_contiguous(), "gemm_bf16_mn: 2-D contiguous bf16");
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0);
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == K, "reduction length mismatch");
  TORCH_CHECK(N > 64 && a.size(1) % 8 == 0 && b.size(1) % 8 == 0, "gemm_bf16_mn: N > 64 and 16-byte row pitches");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(a.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(b.data_ptr()) % 16 == 0);
  const c10::cuda::CUDAGuard guard(a.device());
  FLUTE_CUDA_CHECK(cudaSetDevice(a.device().index()));       // driver entry points need a current context on this thread
  auto stream = at::cuda::getCurrentCUDAStream();
  torch::Tensor c = torch::empty({M, N}, a.options().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16));
  if (M == 0) return c;
  const int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const bool can256 = N >= 256 && (m_tiles * ((N + 255) / 256) >= 148 || K >= 2048);
  const int block_n = can256 ? 256 : 128;
  CUtensorMap ma = a_mn ? make_map_mn(a.data_ptr(), 1, K, M, K * M) : make_map(a.data_ptr(), 1, M, K, M * K, BLOCK_M);
  CUtensorMap mb = b_mn ? make_map_mn(b.data_ptr(), 1, K, N, K * N) : make_map(b.data_ptr(), 1, N, K, N * K, block_n);
  EpilogueParams ep{};
  ep.C = c.data_ptr();
  ep.bias = nullptr;
  ep.M = static_cast<int>(M); ep.N = static_cast<int>(N); ep.K = static_cast<int>(K);
  ep.c_batch_stride = M * N;
  ep.relu = 0;
  ep.out_fp32 = out_fp32 ? 1 : 0;
  ep.a_mn = a_mn ? 1 : 0; ep.b_mn = b_mn ? 1 : 0;
  if (block_n == 256) launch_persistent<256>(ma, mb, ep, 1, stream); else launch_persistent<128>(ma, mb, ep, 1, stream);
  return c;
}

// a: [M, K] or [G, M, K] bf16 ; b: [N, K] or [G, N, K] bf16 (a 2-D b is shared by all batches) -> C [.., M, N]
torch::Tensor gemm_bf16_tn(torch::Tensor a, torch::Tensor b, c10::optional<torch::Tensor> bias, bool relu, bool out_fp32) {
  using namespace gemm;
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.scalar_type() == torch::kBFloat16 && b.scalar_type() == torch::kBFloat16,
              "gemm_bf16_tn: bf16 CUDA tensors expected");
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous(), "gemm_bf16_tn: contiguous (K-major) operands expected");
  const bool batched = a.dim() == 3;
  TORCH_CHECK((a.dim() == 2 || a.dim() == 3) && (b.dim() == 2 || b.dim() == a.dim()), "bad ranks");
  const int64_t G = batched ? a.size(0) : 1;
  const int64_t M = a.size(-2), K = a.size(-1), N = b.size(-2);
  TORCH_CHECK(b.size(-1) == K, "K mismatch");
  TORCH_CHECK(K % 8 == 0, "K must be a multiple of 8 (16-byte TMA row pitch)");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(a.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(b.data_ptr()) % 16 == 0);
  const c10::cuda::CUDAGuard guard(a.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto opts = a.options().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16);
  torch::Tensor c = batched ? torch::empty({G, M, N}, opts) : torch::empty({M, N}, opts);
  if (M == 0 || N == 0) return c;
  const int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  // persistent kernel: BLOCK_N 256 when that still gives every SM a tile, else 128; the simple kernel keeps N <= 64
  int block_n = N <= 64 ? 64 : 128;
  bool persistent = g_gemm_impl != 1 && N > 64;
  if (persistent) {
    // 256-wide tiles halve the operand traffic per MMA; take them when every SM still gets a tile, or when K is long
    // enough that mainloop efficiency beats tile-count balance (measured: 4096x768x3072 572 vs 510 TFLOP/s)
    const bool can256 = N >= 256 && (m_tiles * ((N + 255) / 256) * G >= 148 || K >= 2048);
    block_n = g_gemm_impl == 3 ? 256 : g_gemm_impl == 2 ? 128 : (can256 ? 256 : 128);
  }
  const int64_t b_bstride = (b.dim() == 3) ? N * K : 0;
  CUtensorMap ma = make_map(a.data_ptr(), G, M, K, M * K, BLOCK_M);
  CUtensorMap mb = make_map(b.data_ptr(), b.dim() == 3 ? G : 1, N, K, b_bstride == 0 ? N * K : b_bstride, block_n);
  EpilogueParams ep{};
  ep.C = c.data_ptr();
  ep.bias = nullptr;
  torch::Tensor bias_f;
  if (bias.has_value()) {
    bias_f = bias->to(torch::kFloat32).contiguous();
    TORCH_CHECK(bias_f.numel() == N, "bias size mismatch");
    ep.bias = bias_f.data_ptr<float>();
  }
  ep.M = static_cast<int>(M); ep.N = static_cast<int>(N); ep.K = static_cast<int>(K);
  ep.c_batch_stride = M * N;
  ep.relu = relu ? 1 : 0;
  ep.out_fp32 = out_fp32 ? 1 : 0;
  int batches = static_cast<int>(G);
  if (b.dim() == 2 && G > 1) {
    // a 2-D (shared) B: run the batches as one tall GEMM (A is [G*M, K] contiguous) — identical result, one launch
    ma = make_map(a.data_ptr(), 1, G * M, K, G * M * K, BLOCK_M);
    ep.M = static_cast<int>(G * M);
    ep.c_batch_stride = 0;
    batches = 1;
  }
  if (persistent) {
    if (block_n == 256) launch_persistent<256>(ma, mb, ep, batches, stream); else launch_persistent<128>(ma, mb, ep, batches, stream);
  } else if (block_n == 64) {
    launch<64>(ma, mb, ep, batches, stream);
  } else {
    launch<128>(ma, mb, ep, batches, stream);
  }
  return c;
}

void gemm_set_impl(int64_t impl) {
  TORCH_CHECK(impl >= 0 && impl <= 3, "impl: 0 auto, 1 simple, 2 persistent/128, 3 persistent/256");
  gemm::g_gemm_impl = static_cast<int>(impl);
}

}  // namespace flute
