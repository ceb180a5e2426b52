// Multi-head self-attention, forward + backward, on the 5th-generation tensor cores.
//
//   O = dropout(softmax(Q K^T * scale + key_bias)) V          Q, K, V: [B, H, S, 64] bf16 (any strides, d contiguous)
//
// Replaces torch SDPA (library flash kernels) in the BERT encoder (/root/reference/experiments/mlm_bert/model.py:119-125
// instantiates HF BertSelfAttention) for head_dim 64 and S <= 512 — the whole federated MLM task.
//
// Forward, one CTA per (128-query block, head, batch), 6 warps:
//   warp 0   TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64, 128-byte swizzle) through a 2-stage ring;
//            rows past the sequence end are zero-filled by the tensor map (no host padding)
//   warp 1   one thread issues tcgen05.mma kind::f16: S_j = Q K_j^T (128 x 128 x 64, fp32 in TMEM), and once the softmax
//            warps have published P_j: O_j = P_j V_j (128 x 64 x 128; V is consumed as an MN-MAJOR B operand straight
//            from its row-major tile — no transpose anywhere)
//   warps 2-5  thread == query row: tcgen05.ld of its S row, online softmax (running max / sum in registers), Philox
//            dropout, bf16 P row written into the 128B-swizzled K-major layout the next MMA reads, O accumulated in
//            registers across key blocks (rescaled when the max moves), final normalisation, bf16 store, log-sum-exp
// Backward, one CTA per (128-key block, head, batch): K, V resident in shared memory; per query block
//   S = Q K^T and dP = dO V^T into TMEM -> thread-per-row: P = exp(S - lse), dS = P o (dP - rowsum(dO o O)) with the
//   SAME dropout mask recomputed from the Philox counter -> P^T, dS, dS^T written as K-major A operands ->
//   dV += P^T dO, dK += dS^T Q (accumulated in TMEM over the query blocks), dQ = dS K (fp32 atomics when S > 128).
// TMEM budget: forward 128 (S) + 64 (O_j); backward 128 + 128 + 3 x 64 = 448 of 512 columns.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace flute {
namespace attn {
using namespace tc;

constexpr int kThreads = 192;
constexpr int BQ = 128, BK = 128, HD = 64;
constexpr int TILE_BYTES = 128 * 128;          // 128 rows x 64 bf16
constexpr float kLog2e = 1.4426950408889634f;

struct AttnP {
  CUtensorMap maps[4];           // fwd: Q, K, V;  bwd: Q, K, V, dO      (4-D: d, s, h, b) — kernel-parameter space, so a
                                 // captured CUDA graph carries them (no host->device copy to replay)
  __nv_bfloat16* out;            // fwd: O [B, S, H, 64]
  float* lse;                    // [B, H, S] log-sum-exp of the scaled, biased scores (natural log)
  const float* kbias;            // [B, S] additive key bias (nullable)
  // backward
  const __nv_bfloat16* o_in;     // O  [B, S, H, 64]
  const __nv_bfloat16* do_in;    // dO [B, S, H, 64] (contiguous copy)
  float* dq;                     // [B, S, H, 64] fp32 (zero-initialised when S > 128)
  __nv_bfloat16* dk;             // [B, S, H, 64]
  __nv_bfloat16* dv;
  int B, H, S;
  float scale;                   // 1 / sqrt(d)
  float p_drop;                  // dropout probability (0: none)
  const long long* seed_ptr;     // device counter value used by this call (graph replays read the live value)
};

// ---- dropout: Philox4x32-10 keyed by the seed, counter = (row index, group of 8 keys) -> 8 16-bit uniforms -------------
__device__ __forceinline__ uint4 philox(unsigned long long seed, unsigned long long ctr_lo, unsigned ctr_hi) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr_lo), c1 = static_cast<uint32_t>(ctr_lo >> 32), c2 = ctr_hi, c3 = 0x5eed5eedu;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask bits of keys [k8*8, k8*8 + 8) of global row `row` (= ((b * H + h) * S + q)): bit i set = keep key k8*8 + i
__device__ __forceinline__ uint32_t keep_bits8(unsigned long long seed, unsigned long long row, int k8, uint32_t thr16) {
  const uint4 r = philox(seed, row, static_cast<unsigned>(k8));
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thr16) ? 1u : 0u;   m |= ((r.x >> 16) >= thr16) ? 2u : 0u;
  m |= ((r.y & 0xFFFFu) >= thr16) ? 4u : 0u;   m |= ((r.y >> 16) >= thr16) ? 8u : 0u;
  m |= ((r.z & 0xFFFFu) >= thr16) ? 16u : 0u;  m |= ((r.z >> 16) >= thr16) ? 32u : 0u;
  m |= ((r.w & 0xFFFFu) >= thr16) ? 64u : 0u;  m |= ((r.w >> 16) >= thr16) ? 128u : 0u;
  return m;
}
__host__ __device__ inline uint32_t drop_threshold(float p) {
  const float t = p * 65536.f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : static_cast<uint32_t>(t + 0.5f));
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major SWIZZLE_128B descriptor (rows of 128 B, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return make_smem_desc(addr); }
// MN-major SWIZZLE_128B descriptor of a [K rows][64 bf16] tile: 64 contiguous MN elements per 128-byte row, 8 K-rows per
// 1024-byte atom (stride byte offset), next 64-wide MN block `lbo` bytes away (unused when N == 64)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor: D fp32, A/B bf16, optional MN-major B
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }     // the 4 compute warps
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// store 8 consecutive bf16 (one 16-byte chunk) of row `row`, columns [col8*8, col8*8+8) of a 128-column K-major operand
// made of two [128 rows x 128 B] swizzled slabs
__device__ __forceinline__ void store_row_chunk(uint8_t* base, int row, int col8, uint4 v) {
  const int slab = col8 >> 3, chunk = col8 & 7;
  *reinterpret_cast<uint4*>(base + slab * TILE_BYTES + sw128_offset(row, chunk)) = v;
}
// store ONE bf16 at (row, col) of the same two-slab layout (transposed writes)
__device__ __forceinline__ void store_elem(uint8_t* base, int row, int col, __nv_bfloat16 v) {
  const int slab = col >> 6, c = col & 63;
  *reinterpret_cast<__nv_bfloat16*>(base + slab * TILE_BYTES + sw128_offset(row, c >> 3) + (c & 7) * 2) = v;
}

// ================================================================================================================ forward
constexpr int FWD_SMEM = TILE_BYTES /*Q*/ + 2 * 2 * TILE_BYTES /*K,V x 2 stages*/ + 2 * TILE_BYTES /*P*/ + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1) attn_fwd_kernel(const __grid_constant__ AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TILE_BYTES;                 // stage s: K at sKV + s*2*TILE, V right behind it
  uint8_t* sP = sKV + 4 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;             // 1
  uint64_t* kv_full = bars + 1;        // 2
  uint64_t* kv_empty = bars + 3;       // 2
  uint64_t* s_full = bars + 5;         // 1
  uint64_t* p_ready = bars + 6;        // 1 (128 arrivals)
  uint64_t* o_full = bars + 7;         // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int nkv = (S + BK - 1) / BK;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane < 3) prefetch_tmap(p.maps + lane);
  if (warp == 1) tmem_alloc(tmem_ptr, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(sQ, p.maps + 0, q_full, 0, qb * BQ, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty + s, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + s, 2 * TILE_BYTES);
        tma_load_4d(sKV + s * 2 * TILE_BYTES, p.maps + 1, kv_full + s, 0, j * BK, h, b);
        tma_load_4d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, p.maps + 2, kv_full + s, 0, j * BK, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = idesc_bf16(128, 128, false);
      constexpr uint32_t idO = idesc_bf16(128, 64, true);
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_full + s, (j >> 1) & 1);
        tc_fence_after();
        const uint64_t dq = desc_kmajor(smem_u32(sQ));
        const uint64_t dk = desc_kmajor(smem_u32(sKV + s * 2 * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, dq + 2 * k, dk + 2 * k, idS, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_ready, j & 1);                 // P_j is in shared memory (and S_j has been consumed)
        tc_fence_after();
        const uint32_t pa = smem_u32(sP);
        const uint64_t dv = desc_mnmajor(smem_u32(sKV + s * 2 * TILE_BYTES + TILE_BYTES), TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dp = desc_kmajor(pa + (k >> 2) * TILE_BYTES) + 2 * (k & 3);
          umma_f16(tO, dp, dv + static_cast<uint64_t>(128 * k), idO, k > 0 ? 1u : 0u);
        }
        umma_commit(kv_empty + s);
        umma_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ softmax warps
    const int qd = warp & 3;
    const int r = qd * 32 + lane;                 // row inside the query block
    const int q = qb * BQ + r;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float sc = p.scale * kLog2e;            // scores in log2 units
    const unsigned long long grow = (static_cast<unsigned long long>(b) * p.H + h) * S + q;
    const uint32_t thr = drop_threshold(p.p_drop);
    const unsigned long long seed = thr > 0 ? static_cast<unsigned long long>(*p.seed_ptr) : 0ull;
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    const float* kb = p.kbias != nullptr ? p.kbias + static_cast<long long>(b) * S : nullptr;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum of the scaled, biased scores
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_off + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kk = j * BK + c * 32 + i;
          float x = __uint_as_float(v[i]) * sc;
          if (kb != nullptr && kk < S) x += kb[kk] * kLog2e;
          x = kk < S ? x : -INFINITY;
          m_blk = fmaxf(m_blk, x);
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;          // fully masked row so far
      const float alpha = exp2f(m_run - m_use);                       // m_run = -inf -> 0
      if (j > 0) {
        // fold in the previous block's P V (computed relative to the previous maximum), then rescale
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tO + lane_off + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(v[i])) * alpha;
        }
      }
      l_run *= alpha;
      m_run = m_new;
      // pass 2: probabilities (unnormalised), dropout, bf16 P row -> shared memory
      float l_blk = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_off + c * 32, v);
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kk = j * BK + c * 32 + i;
          float x = __uint_as_float(v[i]) * sc;
          if (kb != nullptr && kk < S) x += kb[kk] * kLog2e;
          const float e = kk < S ? exp2f(x - m_use) : 0.f;
          l_blk += e;
          pr[i] = e;
        }
        if (thr > 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t keep = keep_bits8(seed, grow, (j * BK + c * 32) / 8 + g, thr);
#pragma unroll
            for (int i = 0; i < 8; ++i) pr[g * 8 + i] = ((keep >> i) & 1u) ? pr[g * 8 + i] * inv_keep : 0.f;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16(pr[g * 8 + 0], pr[g * 8 + 1]); w.y = pack_bf16(pr[g * 8 + 2], pr[g * 8 + 3]);
          w.z = pack_bf16(pr[g * 8 + 4], pr[g * 8 + 5]); w.w = pack_bf16(pr[g * 8 + 6], pr[g * 8 + 7]);
          store_row_chunk(sP, r, c * 4 + g, w);
        }
      }
      l_run += l_blk;
      tc_fence_before();
      fence_proxy_async();                         // generic-proxy writes of P -> visible to the tensor core
      mbar_arrive(p_ready);
    }
    // last block's P V
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_off + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(v[i])) * inv_l;
    }
    if (q < S) {
      __nv_bfloat16* dst = p.out + ((static_cast<long long>(b) * S + q) * p.H + h) * HD;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        w.x = pack_bf16(o_acc[g * 8 + 0], o_acc[g * 8 + 1]); w.y = pack_bf16(o_acc[g * 8 + 2], o_acc[g * 8 + 3]);
        w.z = pack_bf16(o_acc[g * 8 + 4], o_acc[g * 8 + 5]); w.w = pack_bf16(o_acc[g * 8 + 6], o_acc[g * 8 + 7]);
        reinterpret_cast<uint4*>(dst)[g] = w;
      }
      // natural-log log-sum-exp of the scaled scores (what the backward pass subtracts)
      p.lse[(static_cast<long long>(b) * p.H + h) * S + q] =
          l_run > 0.f ? (m_run + log2f(l_run)) * (1.f / kLog2e) : -INFINITY;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

// =============================================================================================================== backward
constexpr int BWD_SMEM = 4 * TILE_BYTES /*K V Q dO*/ + 3 * 2 * TILE_BYTES /*P^T dS dS^T*/ + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1) attn_bwd_kernel(const __grid_constant__ AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ = sV + TILE_BYTES;
  uint8_t* sdO = sQ + TILE_BYTES;
  uint8_t* sPT = sdO + TILE_BYTES;               // P^T  [key][query]  (A of dV)
  uint8_t* sdS = sPT + 2 * TILE_BYTES;           // dS   [query][key]  (A of dQ)
  uint8_t* sdST = sdS + 2 * TILE_BYTES;          // dS^T [key][query]  (A of dK)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdST + 2 * TILE_BYTES);
  uint64_t* kv_full = bars;            // 1
  uint64_t* qdo_full = bars + 1;       // 1
  uint64_t* qdo_empty = bars + 2;      // 1
  uint64_t* sdp_full = bars + 3;       // 1
  uint64_t* ds_ready = bars + 4;       // 128 arrivals
  uint64_t* dq_full = bars + 5;        // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int nq = (S + BQ - 1) / BQ;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1); mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1); mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 128); mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane < 4) prefetch_tmap(p.maps + lane);
  if (warp == 1) tmem_alloc(tmem_ptr, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320, tdQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_4d(sK, p.maps + 1, kv_full, 0, kb_i * BK, h, b);
      tma_load_4d(sV, p.maps + 2, kv_full, 0, kb_i * BK, h, b);
      for (int i = 0; i < nq; ++i) {
        mbar_wait(qdo_empty, (i & 1) ^ 1);
        mbar_expect_tx(qdo_full, 2 * TILE_BYTES);
        tma_load_4d(sQ, p.maps + 0, qdo_full, 0, i * BQ, h, b);
        tma_load_4d(sdO, p.maps + 3, qdo_full, 0, i * BQ, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = idesc_bf16(128, 128, false);
      constexpr uint32_t idG = idesc_bf16(128, 64, true);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nq; ++i) {
        mbar_wait(qdo_full, i & 1);
        tc_fence_after();
        const uint64_t dQd = desc_kmajor(smem_u32(sQ)), dKd = desc_kmajor(smem_u32(sK));
        const uint64_t dOd = desc_kmajor(smem_u32(sdO)), dVd = desc_kmajor(smem_u32(sV));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, dQd + 2 * k, dKd + 2 * k, idS, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdP, dOd + 2 * k, dVd + 2 * k, idS, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
        mbar_wait(ds_ready, i & 1);
        tc_fence_after();
        const uint64_t bdO = desc_mnmajor(smem_u32(sdO), TILE_BYTES);     // [query][d]: N = d, K = query
        const uint64_t bQ = desc_mnmajor(smem_u32(sQ), TILE_BYTES);
        const uint64_t bK = desc_mnmajor(smem_u32(sK), TILE_BYTES);       // [key][d]:   N = d, K = key
        const uint32_t aPT = smem_u32(sPT), adS = smem_u32(sdS), adST = smem_u32(sdST);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tdV, desc_kmajor(aPT + (k >> 2) * TILE_BYTES) + 2 * (k & 3), bdO + static_cast<uint64_t>(128 * k), idG,
                   (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tdK, desc_kmajor(adST + (k >> 2) * TILE_BYTES) + 2 * (k & 3), bQ + static_cast<uint64_t>(128 * k), idG,
                   (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tdQ, desc_kmajor(adS + (k >> 2) * TILE_BYTES) + 2 * (k & 3), bK + static_cast<uint64_t>(128 * k), idG,
                   k > 0 ? 1u : 0u);
        umma_commit(qdo_empty);
        umma_commit(dq_full);
      }
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float sc = p.scale * kLog2e;
    const uint32_t thr = drop_threshold(p.p_drop);
    const unsigned long long seed = thr > 0 ? static_cast<unsigned long long>(*p.seed_ptr) : 0ull;
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    const float* kb = p.kbias != nullptr ? p.kbias + static_cast<long long>(b) * S : nullptr;
    const bool atomic_dq = gridDim.x > 1;

    for (int i = 0; i < nq; ++i) {
      const int q = i * BQ + r;
      const bool qv = q < S;
      // D = rowsum(dO o O), log-sum-exp (log2 units) of this thread's query row
      float Dsum = 0.f, lse2 = 0.f;
      if (qv) {
        const long long ro = ((static_cast<long long>(b) * S + q) * p.H + h) * HD;
        const uint4* po = reinterpret_cast<const uint4*>(p.o_in + ro);
        const uint4* pd = reinterpret_cast<const uint4*>(p.do_in + ro);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 a = po[g], d = pd[g];
          const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = __bfloat1622float2(a2[e]), fd = __bfloat1622float2(d2[e]);
            Dsum += fa.x * fd.x + fa.y * fd.y;
          }
        }
        lse2 = p.lse[(static_cast<long long>(b) * p.H + h) * S + q] * kLog2e;
      }
      const unsigned long long grow = (static_cast<unsigned long long>(b) * p.H + h) * S + q;
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32(tS + lane_off + c * 32, vs);
        tmem_ld_32x32(tdP + lane_off + c * 32, vd);
        float pd_[32], ds_[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int kk = kb_i * BK + c * 32 + j;
          float x = __uint_as_float(vs[j]) * sc;
          if (kb != nullptr && kk < S) x += kb[kk] * kLog2e;
          const bool ok = qv && kk < S && lse2 != -INFINITY;
          const float pr = ok ? exp2f(x - lse2) : 0.f;
          pd_[j] = pr;
          ds_[j] = __uint_as_float(vd[j]);
        }
        if (thr > 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t keep = keep_bits8(seed, grow, (kb_i * BK + c * 32) / 8 + g, thr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float mk = ((keep >> j) & 1u) ? inv_keep : 0.f;
              const float pr = pd_[g * 8 + j];
              ds_[g * 8 + j] = pr * (ds_[g * 8 + j] * mk - Dsum) * p.scale;
              pd_[g * 8 + j] = pr * mk;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) ds_[j] = pd_[j] * (ds_[j] - Dsum) * p.scale;
        }
        // dS row chunk (K-major [query][key]) + the transposed copies [key][query]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16(ds_[g * 8 + 0], ds_[g * 8 + 1]); w.y = pack_bf16(ds_[g * 8 + 2], ds_[g * 8 + 3]);
          w.z = pack_bf16(ds_[g * 8 + 4], ds_[g * 8 + 5]); w.w = pack_bf16(ds_[g * 8 + 6], ds_[g * 8 + 7]);
          store_row_chunk(sdS, r, c * 4 + g, w);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          store_elem(sPT, c * 32 + j, r, __float2bfloat16_rn(pd_[j]));
          store_elem(sdST, c * 32 + j, r, __float2bfloat16_rn(ds_[j]));
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(ds_ready);
      // dQ of this query block
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tdQ + lane_off + c * 32, v);
        if (qv) {
          float* dst = p.dq + ((static_cast<long long>(b) * S + q) * p.H + h) * HD + c * 32;
          if (atomic_dq) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
      tc_fence_before();
    }
    // dK, dV of this key block (complete once the last dq_full fired: same commit group)
    const int kk = kb_i * BK + r;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t src = t == 0 ? tdK : tdV;
      __nv_bfloat16* base = t == 0 ? p.dk : p.dv;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(src + lane_off + c * 32, v);
        if (kk < S) {
          uint4* dst = reinterpret_cast<uint4*>(base + ((static_cast<long long>(b) * S + kk) * p.H + h) * HD + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            w.y = pack_bf16(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            w.z = pack_bf16(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            w.w = pack_bf16(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            dst[g] = w;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// keep-mask [B, H, S, S] (uint8) of the kernels' dropout stream: the test oracle applies exactly this mask
__global__ void attn_mask_kernel(unsigned char* __restrict__ out, int B, int H, int S, const long long* seed_ptr, float p) {
  const unsigned long long seed = static_cast<unsigned long long>(*seed_ptr);
  const long long row = blockIdx.x;                     // (b * H + h) * S + q
  const uint32_t thr = drop_threshold(p);
  for (int k8 = threadIdx.x; k8 < (S + 7) / 8; k8 += blockDim.x) {
    const uint32_t keep = thr > 0 ? keep_bits8(seed, static_cast<unsigned long long>(row), k8, thr) : 0xFFu;
    for (int i = 0; i < 8 && k8 * 8 + i < S; ++i) out[row * S + k8 * 8 + i] = (keep >> i) & 1u;
  }
}

// ================================================================================================================== host
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult qres;
    void* ptr = nullptr;
    FLUTE_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    TORCH_CHECK(qres == cudaDriverEntryPointSuccess && ptr != nullptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// t: [B, H, S, 64] view (d contiguous) -> 4-D map (d, s, h, b), box 64 x 128 x 1 x 1, 128-byte swizzle, zero OOB fill
// cuTensorMapEncodeTiled is a DRIVER entry point: on a thread that has not touched the runtime yet (autograd worker
// threads call the backward) no context is current and the encode call faults — bind the primary context first.
static void bind_context(const torch::Tensor& t) { FLUTE_CUDA_CHECK(cudaSetDevice(t.device().index())); }

static CUtensorMap make_map(const torch::Tensor& t) {
  TORCH_CHECK(t.dim() == 4 && t.size(3) == HD && t.stride(3) == 1 && t.scalar_type() == torch::kBFloat16,
              "attention operands: [B, H, S, 64] bf16 with contiguous head dimension");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0 && (t.stride(0) * 2) % 16 == 0 && (t.stride(1) * 2) % 16 == 0 &&
              (t.stride(2) * 2) % 16 == 0, "attention operands: 16-byte aligned strides");
  CUtensorMap m;
  cuuint64_t gd[4] = {static_cast<cuuint64_t>(HD), static_cast<cuuint64_t>(t.size(2)), static_cast<cuuint64_t>(t.size(1)),
                      static_cast<cuuint64_t>(t.size(0))};
  cuuint64_t gs[3] = {static_cast<cuuint64_t>(t.stride(2) * 2), static_cast<cuuint64_t>(t.stride(1) * 2),
                      static_cast<cuuint64_t>(t.stride(0) * 2)};
  cuuint32_t bx[4] = {HD, 128, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.data_ptr(), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", static_cast<int>(r));
  return m;
}

}  // namespace attn

bool attention_supported(int64_t S, int64_t D) { return D == attn::HD && S >= 1 && S <= 512; }

// returns {O [B, S, H, 64] bf16, lse [B, H, S] fp32}
std::vector<torch::Tensor> attention_fwd(torch::Tensor q, torch::Tensor k, torch::Tensor v, c10::optional<torch::Tensor> key_bias,
                                         double scale, double p_drop, torch::Tensor seed) {
  using namespace attn;
  TORCH_CHECK(q.is_cuda() && q.sizes() == k.sizes() && q.sizes() == v.sizes());
  const int B = static_cast<int>(q.size(0)), H = static_cast<int>(q.size(1)), S = static_cast<int>(q.size(2));
  TORCH_CHECK(attention_supported(S, q.size(3)), "attention_fwd: head_dim 64, S <= 512");
  const c10::cuda::CUDAGuard guard(q.device());
  bind_context(q);
  auto out = torch::empty({B, S, H, HD}, q.options());
  auto lse = torch::empty({B, H, S}, q.options().dtype(torch::kFloat32));
  AttnP p{};
  p.maps[0] = make_map(q); p.maps[1] = make_map(k); p.maps[2] = make_map(v); p.maps[3] = p.maps[0];
  p.out = reinterpret_cast<__nv_bfloat16*>(out.data_ptr());
  p.lse = lse.data_ptr<float>();
  if (key_bias.has_value()) {
    TORCH_CHECK(key_bias->is_cuda() && key_bias->scalar_type() == torch::kFloat32 && key_bias->is_contiguous() &&
                key_bias->numel() == static_cast<int64_t>(B) * S, "key_bias: [B, S] fp32");
    p.kbias = key_bias->data_ptr<float>();
  }
  p.B = B; p.H = H; p.S = S;
  p.scale = static_cast<float>(scale); p.p_drop = static_cast<float>(p_drop);
  TORCH_CHECK(seed.is_cuda() && seed.scalar_type() == torch::kInt64 && seed.numel() == 1, "seed: 1-element int64 CUDA tensor");
  p.seed_ptr = reinterpret_cast<const long long*>(seed.data_ptr<int64_t>());
  static bool attr = false;
  if (!attr) {
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  attn_fwd_kernel<<<dim3((S + BQ - 1) / BQ, H, B), kThreads, FWD_SMEM, at::cuda::getCurrentCUDAStream()>>>(p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {out, lse};
}

// returns {dQ [B, S, H, 64] fp32, dK, dV [B, S, H, 64] bf16}
std::vector<torch::Tensor> attention_bwd(torch::Tensor q, torch::Tensor k, torch::Tensor v, torch::Tensor o, torch::Tensor lse,
                                         torch::Tensor d_o, c10::optional<torch::Tensor> key_bias, double scale, double p_drop,
                                         torch::Tensor seed) {
  using namespace attn;
  const int B = static_cast<int>(q.size(0)), H = static_cast<int>(q.size(1)), S = static_cast<int>(q.size(2));
  TORCH_CHECK(attention_supported(S, q.size(3)));
  TORCH_CHECK(o.is_contiguous() && o.sizes() == torch::IntArrayRef({B, S, H, HD}) && d_o.is_contiguous() &&
              d_o.sizes() == o.sizes() && d_o.scalar_type() == torch::kBFloat16 && lse.is_contiguous(),
              "attention_bwd: O / dO [B, S, H, 64] contiguous bf16");
  const c10::cuda::CUDAGuard guard(q.device());
  bind_context(q);
  auto do_view = d_o.permute({0, 2, 1, 3});                     // [B, H, S, 64] view of the [B, S, H, 64] buffer
  const int nkb = (S + BK - 1) / BK;
  auto dq = nkb > 1 ? torch::zeros({B, S, H, HD}, q.options().dtype(torch::kFloat32))
                    : torch::empty({B, S, H, HD}, q.options().dtype(torch::kFloat32));
  auto dk = torch::empty({B, S, H, HD}, q.options());
  auto dv = torch::empty({B, S, H, HD}, q.options());
  AttnP p{};
  p.maps[0] = make_map(q); p.maps[1] = make_map(k); p.maps[2] = make_map(v); p.maps[3] = make_map(do_view);
  p.lse = lse.data_ptr<float>();
  p.o_in = reinterpret_cast<const __nv_bfloat16*>(o.data_ptr());
  p.do_in = reinterpret_cast<const __nv_bfloat16*>(d_o.data_ptr());
  p.dq = dq.data_ptr<float>();
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk.data_ptr());
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv.data_ptr());
  if (key_bias.has_value()) p.kbias = key_bias->data_ptr<float>();
  p.B = B; p.H = H; p.S = S;
  p.scale = static_cast<float>(scale); p.p_drop = static_cast<float>(p_drop);
  TORCH_CHECK(seed.is_cuda() && seed.scalar_type() == torch::kInt64 && seed.numel() == 1, "seed: 1-element int64 CUDA tensor");
  p.seed_ptr = reinterpret_cast<const long long*>(seed.data_ptr<int64_t>());
  attn_bwd_kernel<<<dim3(nkb, H, B), kThreads, BWD_SMEM, at::cuda::getCurrentCUDAStream()>>>(p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {dq, dk, dv};
}

torch::Tensor attention_dropout_mask(int64_t B, int64_t H, int64_t S, double p_drop, torch::Tensor seed) {
  const c10::cuda::CUDAGuard guard(seed.device());
  auto out = torch::empty({B, H, S, S}, seed.options().dtype(torch::kUInt8));
  attn::attn_mask_kernel<<<static_cast<unsigned>(B * H * S), 64, 0, at::cuda::getCurrentCUDAStream()>>>(
      out.data_ptr<unsigned char>(), static_cast<int>(B), static_cast<int>(H), static_cast<int>(S),
      reinterpret_cast<const long long*>(seed.data_ptr<int64_t>()), static_cast<float>(p_drop));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

}  // namespace flute
