// SlotNet: the flagship client step (VERDICT r1 "next round" item 1) as a short static program of
// TMA-fed tcgen05 kernels over NHWC activations.
//
// Every convolution / linear layer of all S simulated clients ("slots") is ONE launch of `sn_gemm_kernel`:
//
//   warp 0   : TMA producer — one thread issues cp.async.bulk.tensor (5-D tiled maps over [S,B,H,W,C] activations:
//              a filter tap is ONE box whose start coordinate is shifted by (kh - pad, kw - pad); out-of-bounds
//              elements are zero-filled by the TMA unit = padding.  Stride-2 layers read one of four "parity" maps
//              (base pointer offset by (py, px), strides doubled), so no traversal strides are needed.)  Weights live
//              in the per-slot parameter arena as [Cout, live taps, Cin] and are fetched with a 4-D map.
//   warp 1   : allocates TMEM and issues tcgen05.mma.kind::tf32 (UMMA 128 x TN x 8, fp32 operands straight from the
//              arenas — no conversion pass) into a TMEM accumulator; tcgen05.commit releases stages.
//   warps 2-5: epilogue — tcgen05.ld, then one of
//                E_STORE : (+bias) (+skip gradient) -> NHWC store
//                E_GNFWD : GroupNorm(2 ch / group, per-group affine) statistics INSIDE the tile (a tile holds whole
//                          images), normalise + affine + residual + ReLU; writes z (pre-norm), y and (mean, rstd)
//                E_GNBWD : (dgrad) + skip gradient, ReLU mask, GroupNorm backward (two in-tile reductions), writes dz
//                          for the previous conv, the masked gradient (next skip) and atomically dgamma / dbeta
//                E_WGRAD : weight-gradient tile straight into the gradient arena (store, or fp32 atomics for split-K)
//
// Three GEMM forms, all on 128-byte-swizzled shared-memory tiles written by TMA:
//   fprop  Y[pix, co]        = sum_{tap,ci} X[pix + tap, ci]   * W[co, tap, ci]     A K-major,  B K-major
//   dgrad  dX[pix, ci]       = sum_{tap,co} dY[pix - tap, co]  * W[co, tap, ci]     A K-major,  B MN-major
//   wgrad  dW[co, (tap,ci)]  = sum_{pix}    X[pix + tap, ci]   * dY[pix, co]        A MN-major, B MN-major
// (MN-major operands use the transpose bits of the instruction descriptor; the canonical MN-major SWIZZLE_128B atom —
// 8 k-rows x 128 B of 32 consecutive M/N elements — is exactly what a TMA box of 32 channels x k pixels writes.)
//
// The few non-GEMM layers are small NHWC kernels in this file: stem im2col, stem GroupNorm+ReLU+max-pool (fwd / bwd,
// one CTA per image), stand-alone GroupNorm backward (down-sample branches), softmax-CE with bias gradient.
// Replaces cuDNN at /root/reference/experiments/cv_resnet_fedcifar100/model.py:28,119,158 and
// group_normalization.py:59-84.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <algorithm>
#include <queue>
#include <vector>
#include "common.cuh"
#include "tcgen05.cuh"

namespace py = pybind11;

namespace flute {
namespace sn {
using namespace tc;

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };
enum { E_STORE = 0, E_GNFWD = 1, E_GNBWD = 2, E_WGRAD = 3 };

constexpr int kThreads = 192;
constexpr int TM = 128;              // tile rows (UMMA M)
constexpr int kMaxStages = 12;       // pipeline depth is a launch parameter (bytes in flight vs. TMA latency)
constexpr int A_BYTES = TM * 128;    // 128 rows x 32 fp32
constexpr int BLK_BYTES = 32 * 128;  // one MN-major block: 32 k-rows x 32 fp32
constexpr int MAX_TAPS = 12;
constexpr int SCRATCH_BYTES = 2 * 4 * 64 * 4;   // epilogue cross-warp reductions, double buffered

struct GemmP {
  int mode, epi;
  int S, B;
  // ---- row space (fprop: output pixels, dgrad: input pixels [of one parity class]) -------------------------------
  int H, W;                  // spatial size of the row space
  int bw, bh, bb;            // TMA box in pixels: rows = bw * bh * bb (bw == W)
  int tiles_y;               // H / bh (only > 1 when bb == 1)
  int row_tiles;             // row tiles per slot and per class
  int ncls;                  // dgrad of a stride-2 conv: 4 parity classes, else 1
  int cls_py[4], cls_px[4], cls_tap0[4], cls_nt[4];
  int outH, outW;            // full spatial size of the output tensor (== H, W unless ncls == 4)
  int ntaps;                 // taps (ncls == 1)
  int Cred;                  // channels reduced per tap (fprop: Cin, dgrad: Cout)
  int TN, N;                 // tile width / total columns (fprop: Cout, dgrad: Cin, wgrad: Cout)
  signed char tap_dx[MAX_TAPS], tap_dy[MAX_TAPS], tap_map[MAX_TAPS], tap_w[MAX_TAPS];
  // ---- wgrad -----------------------------------------------------------------------------------------------------
  int Cin, Kw;               // Kw = ntaps * Cin = length of one filter row in the arena
  int kchunks, ksplit;       // 32-pixel K chunks per slot; split-K factor
  int kbw, kbh, kbb, kH;     // K-chunk pixel box (product 32) and dy height
  // ---- tensor maps: [0..3] A (activation, by parity), [4] B ------------------------------------------------------
  const CUtensorMap* maps;
  // ---- epilogue --------------------------------------------------------------------------------------------------
  float* out;                // E_STORE: output; E_GNFWD: y; E_GNBWD: dz; E_WGRAD: gradient arena + tensor offset
  float* out2;               // E_GNFWD: z (pre-norm);  E_GNBWD: masked gradient "tm" (may be null)
  float* stats;              // [S*B, C/2, 2] mean, rstd (E_GNFWD writes, E_GNBWD reads)
  const float* res;          // E_GNFWD: residual (nullable); E_STORE / E_GNBWD: skip gradient (nullable)
  const float* yprev;        // E_GNBWD: post-activation output of the layer being differentiated (ReLU mask)
  const float* zprev;        // E_GNBWD: its pre-norm input
  const float* Warena;       // per-slot parameters (gamma / beta / bias)
  float* Garena;             // per-slot gradients
  long long arena_stride;    // floats between slots (P)
  long long gamma_off, beta_off, bias_off;   // offsets inside a slot row (bias_off < 0: none)
  int relu;
  float eps;
  int skip_cls;              // E_STORE with ncls == 4: class that receives `res` (compact half-resolution), -1 = none
  // shared-memory descriptor parameters per operand: leading / stride byte offset, descriptor step per UMMA_K (16-byte
  // units), layout type (2 = SWIZZLE_128B for K-major tiles, 1 = SWIZZLE_128B_BASE32B: the only layout tcgen05 accepts
  // for MN-major 32-bit operands — 4 k-rows x 128 B atoms, 32-byte chunks XORed with the row index; TMA writes it with
  // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
  int a_lbo, a_sbo, a_kstep, a_layout;
  int b_lbo, b_sbo, b_kstep, b_layout;
  int stages;                // TMA -> MMA ring depth (2..kMaxStages)
  int nacc;                  // independent TMEM accumulators (1, 2 or 4): UMMA k-step j of a stage accumulates into
                             // accumulator j % nacc, the epilogue adds them — shortens the dependent-MMA chain
  int dbg;                   // measurement hooks: bit 0 = skip the MMAs, bit 1 = skip the TMA loads
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
         "r"(c4)
      : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_128B, explicit leading / stride byte offsets
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 32 accumulator columns of this thread's row, summed over the `nacc` independent accumulators (TN columns apart)
__device__ __forceinline__ void tmem_ld_acc(uint32_t taddr, int TN, int nacc, uint32_t (&v)[32]) {
  tmem_ld_32x32(taddr, v);
  for (int a = 1; a < nacc; ++a) {
    uint32_t w[32];
    tmem_ld_32x32(taddr + static_cast<uint32_t>(a * TN), w);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
  }
}

// all-reduce of two 16-value arrays over the NL lanes of an image segment (fully unrolled: these epilogues run with
// one warp per scheduler, so every avoided instruction is latency off the critical path of the launch)
template <int NL>
__device__ __forceinline__ void seg_allreduce2_t(float (&a)[16], float (&b)[16]) {
#pragma unroll
  for (int o = 1; o < NL; o <<= 1) {
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      a[g] += __shfl_xor_sync(0xffffffffu, a[g], o);
      b[g] += __shfl_xor_sync(0xffffffffu, b[g], o);
    }
  }
}
__device__ __forceinline__ void seg_allreduce2(float (&a)[16], float (&b)[16], int nl) {
  switch (nl) {
    case 32: seg_allreduce2_t<32>(a, b); break;
    case 16: seg_allreduce2_t<16>(a, b); break;
    case 8: seg_allreduce2_t<8>(a, b); break;
    case 4: seg_allreduce2_t<4>(a, b); break;
    case 2: seg_allreduce2_t<2>(a, b); break;
    default: break;
  }
}

// Per-CTA pipeline state shared by the per-launch kernel and the persistent step kernel: the TMA -> MMA ring keeps
// running across tiles (stage / parity derive from a running iteration count), the accumulator barrier flips once per tile.
struct TileCtx {
  uint8_t* smem;             // ring: kStages x stage_bytes
  float* scratch;
  uint64_t* full_bar; uint64_t* empty_bar;
  uint64_t* acc_bar;         // [nbuf] accumulator ready (MMA -> epilogue)
  uint64_t* acc_free;        // [nbuf] accumulator drained (epilogue -> MMA), nbuf == 2 only
  int nbuf, acc_stride;      // TMEM accumulators per CTA and their distance in columns
  uint32_t tmem_base;
  int kStages, stage_bytes;
  uint32_t it_base;          // ring iterations consumed by earlier tiles of this CTA
  uint32_t acc_phase;        // tiles with work done by this CTA so far (parity of acc_bar)
};

// One 128 x TN output tile (bx = row tile / class / k-split, by = column tile, slot).  Returns the number of ring
// iterations it consumed (identical for every thread of the CTA).
template <int EPI>
__device__ __forceinline__ int gemm_tile(const GemmP& p, const int bx, const int by, const int slot, const TileCtx& c) {
  uint8_t* const smem = c.smem;
  const int stage_bytes = c.stage_bytes;
  const int kStages = c.kStages;
  float* const scratch = c.scratch;
  uint64_t* const full_bar = c.full_bar;
  uint64_t* const empty_bar = c.empty_bar;
  const int abuf = static_cast<int>(c.acc_phase % static_cast<uint32_t>(c.nbuf));
  const uint32_t aphase = (c.acc_phase / static_cast<uint32_t>(c.nbuf)) & 1u;
  uint64_t* const acc_bar = c.acc_bar + abuf;
  const uint32_t tmem_base = c.tmem_base + static_cast<uint32_t>(abuf * c.acc_stride);
  const uint32_t it_base = c.it_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = by * p.TN;

  // ---- tile decode -------------------------------------------------------------------------------------------------
  int cls = 0, tile = bx, tap0 = 0, nt = p.ntaps;
  int rb = 0, ks = 0;                       // wgrad: row block, k split
  if (p.mode == WGRAD) {
    rb = tile / p.ksplit;
    ks = tile - rb * p.ksplit;
  } else if (p.ncls == 4) {
    cls = tile / p.row_tiles;
    tile -= cls * p.row_tiles;
    tap0 = p.cls_tap0[cls];
    nt = p.cls_nt[cls];
  }
  int b0 = 0, y0 = 0;
  if (p.mode != WGRAD) {
    if (p.tiles_y > 1) { b0 = tile / p.tiles_y; y0 = (tile - b0 * p.tiles_y) * p.bh; }
    else b0 = tile * p.bb;
  }
  const int rows = p.bw * p.bh * p.bb;
  int total_its;
  int pc_begin = 0, pc_end = 0;
  if (p.mode == WGRAD) {
    const int per = (p.kchunks + p.ksplit - 1) / p.ksplit;
    pc_begin = ks * per;
    pc_end = min(p.kchunks, pc_begin + per);
    total_its = max(0, pc_end - pc_begin);
  } else {
    total_its = nt * ((p.Cred + 31) / 32);
  }

  if (warp == 0) {
    // =============================================================================================== TMA producer
    if (lane == 0 && total_its > 0) {
      const CUtensorMap* mapB = p.maps + 4;
      int it = 0;
      if (p.mode == WGRAD) {
        const int per_img = (p.kbb > 1) ? 1 : (p.kH / p.kbh);      // y-chunks per image when a chunk is part of an image
        int nA = 0;
        for (int i = 0; i < 4; ++i) nA += (rb * TM + 32 * i < p.Kw) ? 1 : 0;
        const uint32_t tx = static_cast<uint32_t>((nA + p.TN / 32) * BLK_BYTES);
        for (int pc = pc_begin; pc < pc_end; ++pc, ++it) {
          const uint32_t gi = it_base + static_cast<uint32_t>(it);
          const int s = static_cast<int>(gi % kStages);
          mbar_wait(empty_bar + s, ((gi / kStages) & 1) ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + A_BYTES;
          int kb0, ky0;
          if (p.kbb > 1) { kb0 = pc * p.kbb; ky0 = 0; }
          else { kb0 = pc / per_img; ky0 = (pc - kb0 * per_img) * p.kbh; }
          if (p.dbg & 2) { mbar_expect_tx(full_bar + s, 0); continue; }
          mbar_expect_tx(full_bar + s, tx);
          for (int i = 0; i < 4; ++i) {
            const int r = rb * TM + 32 * i;
            if (r < p.Kw) {
              const int t = r / p.Cin, c0 = r - t * p.Cin;
              tma_load_5d(sa + i * BLK_BYTES, p.maps + p.tap_map[t], full_bar + s, c0, p.tap_dx[t], ky0 + p.tap_dy[t], kb0, slot);
            }
          }
          for (int j = 0; j < p.TN / 32; ++j)
            tma_load_5d(sb + j * BLK_BYTES, mapB, full_bar + s, n0 + 32 * j, 0, ky0, kb0, slot);
        }
      } else {
        const uint32_t tx = static_cast<uint32_t>(rows * 128 + p.TN * 128);
        const int Cred = p.Cred, mode = p.mode, nb = p.TN / 32, dbg = p.dbg;
        const CUtensorMap* maps0 = p.maps;
        for (int ti = 0; ti < nt; ++ti) {
          const int t = tap0 + ti;
          const CUtensorMap* mapA = maps0 + p.tap_map[t];
          const int cx = p.tap_dx[t], cy = y0 + p.tap_dy[t], wt = p.tap_w[t];
          for (int c0 = 0; c0 < Cred; c0 += 32, ++it) {
            const uint32_t gi = it_base + static_cast<uint32_t>(it);
            const int s = static_cast<int>(gi % kStages);
            mbar_wait(empty_bar + s, ((gi / kStages) & 1) ^ 1);
            uint8_t* sa = smem + s * stage_bytes;
            uint8_t* sb = sa + A_BYTES;
            if (dbg & 2) { mbar_expect_tx(full_bar + s, 0); continue; }
            mbar_expect_tx(full_bar + s, tx);
            tma_load_5d(sa, mapA, full_bar + s, c0, cx, cy, b0, slot);
            if (mode == FPROP) {
              tma_load_4d(sb, mapB, full_bar + s, c0, wt, n0, slot);                  // [TN co rows][32 ci]
            } else {
              for (int j = 0; j < nb; ++j)                                            // MN-major: [32 co rows][32 ci]
                tma_load_4d(sb + j * BLK_BYTES, mapB, full_bar + s, n0 + 32 * j, wt, c0, slot);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================================================== MMA issuer
    if (lane == 0 && total_its > 0) {
      const bool a_mn = p.mode == WGRAD, b_mn = p.mode != FPROP;
      const uint32_t idesc = make_idesc_fmt(TM, p.TN, 2u) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
      const uint64_t a_step = static_cast<uint64_t>(p.a_kstep), b_step = static_cast<uint64_t>(p.b_kstep);
      const uint32_t a_lbo = p.a_lbo, a_sbo = p.a_sbo, a_layout = p.a_layout;
      const uint32_t b_lbo = p.b_lbo, b_sbo = p.b_sbo, b_layout = p.b_layout;
      const int nacc = p.nacc, accTN = p.TN, dbg = p.dbg;
      if (c.nbuf > 1) {                                   // the epilogue of two tiles ago has drained this accumulator
        mbar_wait(c.acc_free + abuf, aphase ^ 1u);
        tc_fence_after();
      }
      for (int it = 0; it < total_its; ++it) {
        const uint32_t gi = it_base + static_cast<uint32_t>(it);
        const int s = static_cast<int>(gi % kStages);
        mbar_wait(full_bar + s, (gi / kStages) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint64_t adesc = make_desc(a_addr, a_lbo, a_sbo, a_layout);
        const uint64_t bdesc = make_desc(a_addr + A_BYTES, b_lbo, b_sbo, b_layout);
        if (!(dbg & 1)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int acc = k & (nacc - 1);
            umma_tf32(tmem_base + static_cast<uint32_t>(acc * accTN), adesc + a_step * k, bdesc + b_step * k, idesc,
                      (it > 0 || k >= nacc) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar + s);
      }
      umma_commit(acc_bar);
    }
  } else if (total_its > 0) {
    // =============================================================================================== epilogue
    mbar_wait(acc_bar, aphase);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;                     // tile row of this thread
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int TNc = p.TN, Ntot = p.N;

    if (EPI == E_WGRAD) {
      const int row = rb * TM + r;                   // (tap, ci) flat index
      const int Kw = p.Kw;
      const bool atomic = p.ksplit > 1;
      float* dst0 = p.out + static_cast<long long>(slot) * p.arena_stride + row + static_cast<long long>(n0) * Kw;
#pragma unroll 1
      for (int c0 = 0; c0 < TNc; c0 += 32) {
        uint32_t v[32];
        tmem_ld_acc(taddr + c0, TNc, p.nacc, v);
        const int ncol = min(32, Ntot - (n0 + c0));            // valid columns of this chunk
        if (row < Kw && ncol > 0) {
          float* d = dst0 + static_cast<long long>(c0) * Kw;
          if (ncol == 32) {
            if (atomic) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { atomicAdd(d, __uint_as_float(v[j])); d += Kw; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { *d = __uint_as_float(v[j]); d += Kw; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < ncol) { if (atomic) atomicAdd(d, __uint_as_float(v[j])); else *d = __uint_as_float(v[j]); }
              d += Kw;
            }
          }
        }
      }
    } else {
      // ---- row -> pixel ----------------------------------------------------------------------------------------
      const int hw_box = p.bh * p.bw;
      const int bi = r / hw_box, rem = r - bi * hw_box, yy = rem / p.bw, xx = rem - yy * p.bw;
      const int b = b0 + bi;
      const bool valid = r < rows && b < p.B;
      int oy = y0 + yy, ox = xx;
      if (p.ncls == 4) { oy = 2 * oy + p.cls_py[cls]; ox = 2 * ox + p.cls_px[cls]; }
      const long long img = static_cast<long long>(slot) * p.B + b;
      const long long pix = (img * p.outH + oy) * p.outW + ox;
      const long long row_off = pix * Ntot;
      const float* Ws = p.Warena + static_cast<long long>(slot) * p.arena_stride;

      if (EPI == E_STORE) {
        const bool add = p.res != nullptr && (p.ncls == 1 || p.skip_cls == cls);
        // compact skip (ncls == 4): the half-resolution tensor [S,B,H,W,N] indexed by this class's own grid
        const long long skip_off = p.ncls == 4 ? ((img * p.H + (y0 + yy)) * p.W + xx) * Ntot : row_off;
        const bool has_bias = p.bias_off >= 0;
#pragma unroll 1
        for (int c0 = 0; c0 < TNc; c0 += 32) {
          uint32_t v[32];
          tmem_ld_acc(taddr + c0, TNc, p.nacc, v);
          const int col = n0 + c0;
          if (valid && col < Ntot) {
            if (col + 32 <= Ntot) {
              float* o_ptr = p.out + row_off + col;
              const float* b_ptr = Ws + (has_bias ? p.bias_off + col : 0);
              const float* s_ptr = add ? p.res + skip_off + col : nullptr;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                       __uint_as_float(v[j + 3]));
                if (has_bias) {
                  const float4 bv = *reinterpret_cast<const float4*>(b_ptr + j);
                  o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                }
                if (add) {
                  const float4 sv = *reinterpret_cast<const float4*>(s_ptr + j);
                  o.x += sv.x; o.y += sv.y; o.z += sv.z; o.w += sv.w;
                }
                *reinterpret_cast<float4*>(o_ptr + j) = o;
              }
            } else {
              for (int j = 0; j < 32 && col + j < Ntot; ++j) {
                float o = __uint_as_float(v[j]);
                if (has_bias) o += Ws[p.bias_off + col + j];
                if (add) o += p.res[skip_off + col + j];
                p.out[row_off + col + j] = o;
              }
            }
          }
        }
      } else {
        // ---- GroupNorm epilogues: a tile holds whole images (bh == H), 2 channels per group ------------------------
        const int HW = p.H * p.W;
        const int nl = HW < 32 ? HW : 32;            // lanes of one image inside a warp
        const int wpi = HW > 32 ? HW / 32 : 1;       // warps per image
        const int q0 = (q / wpi) * wpi;              // first warp (quadrant) of this thread's image
        const float inv_n = 1.f / static_cast<float>(2 * HW);
        const int G2 = Ntot / 2;                     // groups
        float* Gs = p.Garena + static_cast<long long>(slot) * p.arena_stride;
        const float* gam = Ws + p.gamma_off;
        const float* bet = Ws + p.beta_off;
        const bool has_res = p.res != nullptr, relu = p.relu != 0;
        int chunk = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < TNc; c0 += 32, ++chunk) {
          uint32_t vr[32];
          tmem_ld_acc(taddr + c0, TNc, p.nacc, vr);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? __uint_as_float(vr[j]) : 0.f;
          const int col = n0 + c0;                   // N is a multiple of 32 for every GroupNorm layer
          const int g0 = col >> 1;
          float* red = scratch + (chunk & 1) * (4 * 64);
          float ga[16];
#pragma unroll
          for (int g = 0; g < 16; g += 4) {
            const float4 t = *reinterpret_cast<const float4*>(gam + g0 + g);
            ga[g] = t.x; ga[g + 1] = t.y; ga[g + 2] = t.z; ga[g + 3] = t.w;
          }
          if (EPI == E_GNFWD) {
            float s1[16], s2[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              s1[g] = v[2 * g] + v[2 * g + 1];
              s2[g] = v[2 * g] * v[2 * g] + v[2 * g + 1] * v[2 * g + 1];
            }
            seg_allreduce2(s1, s2, nl);
            if (wpi > 1) {
              if (lane == 0) {
#pragma unroll
                for (int g = 0; g < 16; ++g) { red[q * 64 + g] = s1[g]; red[q * 64 + 16 + g] = s2[g]; }
              }
              epi_bar();
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                float a = 0.f, c = 0.f;
                for (int w = 0; w < wpi; ++w) { a += red[(q0 + w) * 64 + g]; c += red[(q0 + w) * 64 + 16 + g]; }
                s1[g] = a; s2[g] = c;
              }
            }
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              const float mean = s1[g] * inv_n;
              const float var = fmaxf(s2[g] * inv_n - mean * mean, 0.f);
              s1[g] = mean; s2[g] = rsqrtf(var + p.eps);
            }
            if (valid) {
              if ((r % HW) == 0) {                   // one thread per image writes the statistics
                float4* sp = reinterpret_cast<float4*>(p.stats + (img * G2 + g0) * 2);
#pragma unroll
                for (int g = 0; g < 16; g += 2) sp[g >> 1] = make_float4(s1[g], s2[g], s1[g + 1], s2[g + 1]);
              }
              float be[16];
#pragma unroll
              for (int g = 0; g < 16; g += 4) {
                const float4 t = *reinterpret_cast<const float4*>(bet + g0 + g);
                be[g] = t.x; be[g + 1] = t.y; be[g + 2] = t.z; be[g + 3] = t.w;
              }
              float* z_ptr = p.out2 + row_off + col;
              float* y_ptr = p.out + row_off + col;
              const float* r_ptr = has_res ? p.res + row_off + col : nullptr;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                *reinterpret_cast<float4*>(z_ptr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int g = (j + e) >> 1;
                  o[e] = (v[j + e] - s1[g]) * (s2[g] * ga[g]) + be[g];
                }
                if (has_res) {
                  const float4 rv = *reinterpret_cast<const float4*>(r_ptr + j);
                  o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
                }
                if (relu) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                *reinterpret_cast<float4*>(y_ptr + j) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          } else {
            // ---- E_GNBWD: v = d(loss)/d(post-activation output of the previous layer), main branch -------------------
            float xh[32];
            if (valid) {
              const float* s_ptr = has_res ? p.res + row_off + col : nullptr;
              const float* y_ptr = p.yprev + row_off + col;
              const float* z_ptr = p.zprev + row_off + col;
              float* t_ptr = p.out2 != nullptr ? p.out2 + row_off + col : nullptr;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (has_res) {
                  const float4 sv = *reinterpret_cast<const float4*>(s_ptr + j);
                  v[j] += sv.x; v[j + 1] += sv.y; v[j + 2] += sv.z; v[j + 3] += sv.w;
                }
                if (relu) {
                  const float4 yv = *reinterpret_cast<const float4*>(y_ptr + j);
                  v[j] = yv.x > 0.f ? v[j] : 0.f; v[j + 1] = yv.y > 0.f ? v[j + 1] : 0.f;
                  v[j + 2] = yv.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = yv.w > 0.f ? v[j + 3] : 0.f;
                }
                if (t_ptr != nullptr)
                  *reinterpret_cast<float4*>(t_ptr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                const float4 zv = *reinterpret_cast<const float4*>(z_ptr + j);
                xh[j] = zv.x; xh[j + 1] = zv.y; xh[j + 2] = zv.z; xh[j + 3] = zv.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) xh[j] = 0.f;
            }
            float rs[16];
            {
              const float4* sp = reinterpret_cast<const float4*>(p.stats + (img * G2 + g0) * 2);
#pragma unroll
              for (int g = 0; g < 16; g += 2) {
                float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) st = sp[g >> 1];
                rs[g] = st.y; rs[g + 1] = st.w;
                xh[2 * g] = (xh[2 * g] - st.x) * st.y;         xh[2 * g + 1] = (xh[2 * g + 1] - st.x) * st.y;
                xh[2 * g + 2] = (xh[2 * g + 2] - st.z) * st.w; xh[2 * g + 3] = (xh[2 * g + 3] - st.z) * st.w;
              }
            }
            // per-group sums of d and d * xhat: butterfly in INCREASING distance — after log2(nl) steps every lane
            // holds its image's sums (→ dz), after all 5 steps the warp totals (→ dgamma / dbeta, one atomic per warp)
            float sa[16], sbv[16], ia[16], ib[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              sa[g] = v[2 * g] + v[2 * g + 1];                                   // sum d
              sbv[g] = v[2 * g] * xh[2 * g] + v[2 * g + 1] * xh[2 * g + 1];     // sum d * xhat
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              if (o == nl) {
#pragma unroll
                for (int g = 0; g < 16; ++g) { ia[g] = sa[g]; ib[g] = sbv[g]; }
              }
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                sa[g] += __shfl_xor_sync(0xffffffffu, sa[g], o);
                sbv[g] += __shfl_xor_sync(0xffffffffu, sbv[g], o);
              }
            }
            if (nl == 32) {
#pragma unroll
              for (int g = 0; g < 16; ++g) { ia[g] = sa[g]; ib[g] = sbv[g]; }
            }
            if (lane == 0) {
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                atomicAdd(Gs + p.gamma_off + g0 + g, sbv[g]);
                atomicAdd(Gs + p.beta_off + g0 + g, sa[g]);
              }
            }
            if (wpi > 1) {
              if (lane == 0) {
#pragma unroll
                for (int g = 0; g < 16; ++g) { red[q * 64 + g] = ia[g]; red[q * 64 + 16 + g] = ib[g]; }
              }
              epi_bar();
#pragma unroll
              for (int g = 0; g < 16; ++g) {
                float a = 0.f, c = 0.f;
                for (int w = 0; w < wpi; ++w) { a += red[(q0 + w) * 64 + g]; c += red[(q0 + w) * 64 + 16 + g]; }
                ia[g] = a; ib[g] = c;
              }
            }
            if (valid) {
              float* o_ptr = p.out + row_off + col;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int g = (j + e) >> 1;
                  o[e] = rs[g] * ga[g] * (v[j + e] - (ia[g] + xh[j + e] * ib[g]) * inv_n);
                }
                *reinterpret_cast<float4*>(o_ptr + j) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    if (c.nbuf > 1) {
      __syncwarp();
      if (lane == 0) mbar_arrive(c.acc_free + abuf);
    }
  }
  return total_its;
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 2) sn_gemm_kernel(const __grid_constant__ GemmP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  TileCtx c;
  c.smem = smem;
  c.stage_bytes = A_BYTES + p.TN * 128;
  c.kStages = p.stages;
  c.scratch = reinterpret_cast<float*>(smem + c.kStages * c.stage_bytes);
  c.full_bar = reinterpret_cast<uint64_t*>(smem + c.kStages * c.stage_bytes + SCRATCH_BYTES);
  c.empty_bar = c.full_bar + kMaxStages;
  c.acc_bar = c.empty_bar + kMaxStages;
  c.acc_free = c.acc_bar;
  c.nbuf = 1;
  c.acc_stride = 0;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(c.acc_bar + 1);
  c.it_base = 0;
  c.acc_phase = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.kStages; ++s) {
      mbar_init(c.full_bar + s, 1);
      mbar_init(c.empty_bar + s, 1);
    }
    mbar_init(c.acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane < 5) prefetch_tmap(p.maps + lane);
  if (warp == 1) tmem_alloc(tmem_ptr, static_cast<uint32_t>(p.TN * p.nacc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) ran while the
  // previous kernel of the chain was still executing.  Wait for it to complete (its writes are visible afterwards),
  // THEN let the next kernel start its own prologue — releasing only after the wait means at most two kernels of the
  // chain are ever co-resident, and every CTA of this grid is already resident when the successor's CTAs arrive.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  gemm_tile<EPI>(p, blockIdx.x, blockIdx.y, blockIdx.z, c);
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(c.tmem_base, static_cast<uint32_t>(p.TN * p.nacc));
  }
}

// ====================================================================================================== small kernels
// Stem im2col: x [N, 3, 32, 32] (arbitrary strides) -> A [N, 16, 16, 160], column k = (kh * 7 + kw) * 3 + c, k >= 147 zero.
__global__ void sn_im2col_stem_kernel(const float* __restrict__ x, long long sN, long long sC, long long sH, long long sW,
                                      float* __restrict__ out, int N) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;     // one float4 of a row
  const long long total = static_cast<long long>(N) * 256 * 40;
  if (idx >= total) return;
  const int q4 = static_cast<int>(idx % 40);
  const long long pixi = idx / 40;
  const int ox = static_cast<int>(pixi & 15), oy = static_cast<int>((pixi >> 4) & 15);
  const long long n = pixi >> 8;
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = q4 * 4 + e;
    float val = 0.f;
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3, kh = tap / 7, kw = tap - kh * 7;
      const int iy = 2 * oy - 3 + kh, ix = 2 * ox - 3 + kw;
      if (iy >= 0 && iy < 32 && ix >= 0 && ix < 32) val = __ldg(x + n * sN + c * sC + iy * sH + ix * sW);
    }
    o[e] = val;
  }
  *reinterpret_cast<float4*>(out + idx * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

constexpr int kStemFwdSmem = 256 * 32 * 4;
constexpr int kStemBwdSmem = 256 * 32 * 4 + 64 * 32 * 4 + 64 * 32;
struct StemP {
  const float* z;        // [N, 16, 16, 64] stem conv output
  float* stats;          // [N, 32, 2]
  float* pooled;         // [N, 8, 8, 64]
  unsigned char* arg;    // [N, 8, 8, 64] index 0..8 of the arg-max inside the 3x3 window
  const float* dpool;    // backward: [N, 8, 8, 64]
  float* dz;             // backward: [N, 16, 16, 64]
  const float* Warena; float* Garena; long long arena_stride, gamma_off, beta_off;
  int B; float eps;
};

// GroupNorm(32 groups of 2) + ReLU + max-pool 3x3 / 2 / pad 1: one CTA per (image, half of the channels) — groups are
// pairs of adjacent channels, so the two halves are independent.  256 threads = 32 channels x 8 pixel lanes; the CTA's
// [256 px][32 ch] slice of z is staged in shared memory.
__global__ void __launch_bounds__(256) sn_stem_fwd_kernel(const StemP p) {
  extern __shared__ float sm[];                 // [256 px][32 ch]
  __shared__ float s_mean[16], s_rstd[16];
  const int n = blockIdx.x, half = blockIdx.y, tid = threadIdx.x, slot = n / p.B;
  const int c = tid & 31, pl = tid >> 5, cg = half * 32 + c;          // channel in the CTA / pixel lane / global channel
  const float* zsrc = p.z + static_cast<long long>(n) * 256 * 64 + half * 32;
  for (int i = tid; i < 256 * 8; i += 256) {                            // 8 float4 per pixel row of 32 channels
    const int px = i >> 3, q4 = i & 7;
    reinterpret_cast<float4*>(sm)[i] = *reinterpret_cast<const float4*>(zsrc + px * 64 + q4 * 4);
  }
  __syncthreads();
  {
    // 8 warps x 2 groups each: lanes walk the 256 pixels
    const int warp = tid >> 5, lane = tid & 31;
    for (int gi = 0; gi < 2; ++gi) {
      const int g = warp * 2 + gi;
      float s1 = 0.f, s2 = 0.f;
      for (int px = lane; px < 256; px += 32) {
        const float2 v = *reinterpret_cast<const float2*>(sm + px * 32 + 2 * g);
        s1 += v.x + v.y; s2 += v.x * v.x + v.y * v.y;
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) {
        const float mean = s1 * (1.f / 512.f), var = fmaxf(s2 * (1.f / 512.f) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        s_mean[g] = mean; s_rstd[g] = rstd;
        *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(n) * 32 + half * 16 + g) * 2) = make_float2(mean, rstd);
      }
    }
  }
  __syncthreads();
  const float* Ws = p.Warena + static_cast<long long>(slot) * p.arena_stride;
  const int g = c >> 1, gg = cg >> 1;
  const float ga = Ws[p.gamma_off + gg], be = Ws[p.beta_off + gg], mean = s_mean[g], rstd = s_rstd[g];
  const float sc = rstd * ga, sh = be - mean * rstd * ga;
  for (int op = pl; op < 64; op += 8) {
    const int oy = op >> 3, ox = op & 7;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = 2 * oy - 1 + k / 3, ix = 2 * ox - 1 + k % 3;
      if (iy >= 0 && iy < 16 && ix >= 0 && ix < 16) {
        const float y = fmaxf(sm[(iy * 16 + ix) * 32 + c] * sc + sh, 0.f);
        if (y > best) { best = y; bi = k; }
      }
    }
    const long long o = (static_cast<long long>(n) * 64 + op) * 64 + cg;
    p.pooled[o] = best;
    p.arg[o] = static_cast<unsigned char>(bi);
  }
}

// backward of the same: d(pooled) -> dz of the stem conv, dgamma / dbeta.  Same decomposition; d(pooled) and the
// arg-max bytes of the CTA's channels are staged in shared memory, dy (after the ReLU mask) lives in shared memory.
__global__ void __launch_bounds__(256) sn_stem_bwd_kernel(const StemP p) {
  extern __shared__ float sm[];                 // [256 px][32 ch] dy | [64 px][32 ch] dpool | [64][32] arg bytes
  float* s_dp = sm + 256 * 32;
  unsigned char* s_arg = reinterpret_cast<unsigned char*>(s_dp + 64 * 32);
  __shared__ float s_a[16], s_b[16], r1[256], r2[256];
  const int n = blockIdx.x, half = blockIdx.y, tid = threadIdx.x, slot = n / p.B;
  const int c = tid & 31, pl = tid >> 5, cg = half * 32 + c, g = c >> 1, gg = cg >> 1;
  const float* Ws = p.Warena + static_cast<long long>(slot) * p.arena_stride;
  float* Gs = p.Garena + static_cast<long long>(slot) * p.arena_stride;
  const float* z = p.z + static_cast<long long>(n) * 256 * 64 + half * 32;
  for (int i = tid; i < 64 * 32; i += 256) {
    const int op = i >> 5, cc = i & 31;
    const long long o = (static_cast<long long>(n) * 64 + op) * 64 + half * 32 + cc;
    s_dp[i] = p.dpool[o];
    s_arg[i] = p.arg[o];
  }
  __syncthreads();
  const float2 st = *reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(n) * 32 + gg) * 2);
  const float ga = Ws[p.gamma_off + gg], be = Ws[p.beta_off + gg];
  // gather formulation of max-pool backward: input pixel (iy, ix) receives from the <= 4 windows that contain it
  float dsum_g = 0.f, dsum_b = 0.f;
  for (int ip = pl; ip < 256; ip += 8) {
    const int iy = ip >> 4, ix = ip & 15;
    float d = 0.f;
    const int oy_lo = iy >> 1, oy_hi = min(7, (iy + 1) >> 1);
    const int ox_lo = ix >> 1, ox_hi = min(7, (ix + 1) >> 1);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const int k = (iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1));
        const int o = (oy * 8 + ox) * 32 + c;
        if (s_arg[o] == k) d += s_dp[o];
      }
    const float xh = (z[ip * 64 + c] - st.x) * st.y;
    d = (xh * ga + be) > 0.f ? d : 0.f;           // ReLU mask
    sm[ip * 32 + c] = d;
    dsum_g += d * xh;
    dsum_b += d;
  }
  r1[tid] = dsum_g; r2[tid] = dsum_b;
  __syncthreads();
  if (tid < 16) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; ++k)
      for (int e = 0; e < 2; ++e) { a += r1[k * 32 + 2 * tid + e]; b += r2[k * 32 + 2 * tid + e]; }
    s_a[tid] = a; s_b[tid] = b;                 // a = sum d*xhat, b = sum d
    atomicAdd(Gs + p.gamma_off + half * 16 + tid, a);
    atomicAdd(Gs + p.beta_off + half * 16 + tid, b);
  }
  __syncthreads();
  const float A = s_a[g] * ga * (1.f / 512.f), Bm = s_b[g] * ga * (1.f / 512.f);
  float* dz = p.dz + static_cast<long long>(n) * 256 * 64 + half * 32;
  for (int ip = pl; ip < 256; ip += 8) {
    const float xh = (z[ip * 64 + c] - st.x) * st.y;
    dz[ip * 64 + c] = st.y * (sm[ip * 32 + c] * ga - Bm - xh * A);
  }
}

// Stand-alone GroupNorm (2 ch / group) backward on NHWC, thread == (image, group), loops over the image's pixels.
//   d = din (+ add1) (+ add2 scattered from the half-resolution grid at even pixels);  d *= (y > 0) when y given;
//   tm <- d (optional);  dz <- GroupNorm backward of d through (z, stats, gamma);  dgamma / dbeta atomics.
struct GnBwdP {
  const float* din; const float* add1; const float* add2; const float* y; const float* z; const float* stats;
  float* tm; float* dz;
  const float* Warena; float* Garena; long long arena_stride, gamma_off, beta_off;
  int N, B, H, W, C;
};
// t = global thread index (whole warps share a block): used by the stand-alone kernel and by the persistent step kernel
__device__ __forceinline__ void gn_bwd_block(const GnBwdP& p, const int PS, const long long t) {
  // lane = pixel lane * (32 / PS) + group: PS (power of two <= 8, <= H*W) lanes share one (image, group) and split its
  // pixels; for a fixed pixel lane the warp's 32 / PS groups are adjacent channels (one full 32-byte sector at PS = 8)
  const int G2 = p.C / 2;
  const int lane = threadIdx.x & 31, gs = 32 / PS;
  const int pl = lane / gs;
  const long long idx = (t >> 5) * gs + (lane - pl * gs);
  const bool live = idx < static_cast<long long>(p.N) * G2;
  const long long idc = live ? idx : 0;
  const int g = static_cast<int>(idc % G2);
  const long long n = idc / G2;
  const int slot = static_cast<int>(n / p.B), HW = p.H * p.W;
  const float ga = p.Warena[static_cast<long long>(slot) * p.arena_stride + p.gamma_off + g];
  const float2 st = *reinterpret_cast<const float2*>(p.stats + (n * G2 + g) * 2);
  const long long base = n * HW * p.C + 2 * g;
  const int H2 = p.H / 2, W2 = p.W / 2;
  float sg = 0.f, sb = 0.f;
  if (live) {
    for (int px = pl; px < HW; px += PS) {
      const long long o = base + static_cast<long long>(px) * p.C;
      float2 d = *reinterpret_cast<const float2*>(p.din + o);
      if (p.add1 != nullptr) { const float2 a = *reinterpret_cast<const float2*>(p.add1 + o); d.x += a.x; d.y += a.y; }
      if (p.add2 != nullptr) {
        const int yy = px / p.W, xx = px - yy * p.W;
        if (((yy | xx) & 1) == 0) {
          const float2 a = *reinterpret_cast<const float2*>(p.add2 + (n * H2 * W2 + (yy >> 1) * W2 + (xx >> 1)) * p.C + 2 * g);
          d.x += a.x; d.y += a.y;
        }
      }
      if (p.y != nullptr) {
        const float2 yv = *reinterpret_cast<const float2*>(p.y + o);
        d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f;
      }
      if (p.tm != nullptr) *reinterpret_cast<float2*>(p.tm + o) = d;
      else *reinterpret_cast<float2*>(p.dz + o) = d;                   // stash (dz is rewritten below)
      const float2 zv = *reinterpret_cast<const float2*>(p.z + o);
      sg += d.x * (zv.x - st.x) * st.y + d.y * (zv.y - st.x) * st.y;
      sb += d.x + d.y;
    }
  }
  for (int o = gs; o < 32; o <<= 1) {                // the PS lanes of one (image, group) are gs lanes apart
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (!live) return;
  if (pl == 0) {
    float* Gs = p.Garena + static_cast<long long>(slot) * p.arena_stride;
    atomicAdd(Gs + p.gamma_off + g, sg);
    atomicAdd(Gs + p.beta_off + g, sb);
  }
  const float inv_n = 1.f / static_cast<float>(2 * HW);
  const float A = sg * ga * inv_n, Bm = sb * ga * inv_n;
  const float* dsrc = p.tm != nullptr ? p.tm : p.dz;
  for (int px = pl; px < HW; px += PS) {
    const long long o = base + static_cast<long long>(px) * p.C;
    const float2 d = *reinterpret_cast<const float2*>(dsrc + o);
    const float2 zv = *reinterpret_cast<const float2*>(p.z + o);
    const float x0 = (zv.x - st.x) * st.y, x1 = (zv.y - st.x) * st.y;
    *reinterpret_cast<float2*>(p.dz + o) = make_float2(st.y * (d.x * ga - Bm - x0 * A), st.y * (d.y * ga - Bm - x1 * A));
  }
}
__global__ void __launch_bounds__(128) sn_gn_bwd_kernel(const GnBwdP p, const int PS) {
  gn_bwd_block(p, PS, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
}

// softmax cross-entropy (mean over the slot's batch) forward + backward, bias gradient, per-slot loss accumulation
struct CeP {
  const float* logits; const long long* labels; float* dlogits; float* loss; float* Garena;
  long long arena_stride, bias_off; int B, C;
};
__global__ void __launch_bounds__(256) sn_ce_kernel(const CeP p) {
  const int row = blockIdx.x, slot = row / p.B, tid = threadIdx.x;
  const float* x = p.logits + static_cast<long long>(row) * p.C;
  float m = -INFINITY;
  for (int j = tid; j < p.C; j += 256) m = fmaxf(m, x[j]);
  __shared__ float sred[8];
  m = warp_max(m);
  if ((tid & 31) == 0) sred[tid >> 5] = m;
  __syncthreads();
  m = sred[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, sred[i]);
  __syncthreads();
  float s = 0.f;
  for (int j = tid; j < p.C; j += 256) s += __expf(x[j] - m);
  s = warp_sum(s);
  if ((tid & 31) == 0) sred[tid >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < 8; ++i) s += sred[i];
  const int label = static_cast<int>(p.labels[row]);
  const float inv_b = 1.f / static_cast<float>(p.B), inv_s = 1.f / s;
  float* Gs = p.Garena + static_cast<long long>(slot) * p.arena_stride + p.bias_off;
  for (int j = tid; j < p.C; j += 256) {
    const float d = (__expf(x[j] - m) * inv_s - (j == label ? 1.f : 0.f)) * inv_b;
    p.dlogits[static_cast<long long>(row) * p.C + j] = d;
    atomicAdd(Gs + j, d);
  }
  if (tid == 0) atomicAdd(p.loss + slot, (logf(s) + m - x[label]) * inv_b);
}

// ============================================================================================ persistent step kernel
// One cooperative launch runs a whole dependency-ordered range of the program (the 20 forward convolutions, or the
// whole backward pass from the classifier to layer1): the host groups ops into PHASES (an op's phase is one more than
// the latest phase that produced one of its inputs), cuts every op into its 128 x TN tiles, balances the tiles of a
// phase over the resident CTAs (longest-processing-time first) and uploads the per-CTA work lists.  Inside the kernel a
// CTA keeps its TMEM allocation, mbarrier ring and tensor-map cache across tiles and meets the other CTAs at a grid
// barrier between phases — ~1.5 us instead of a kernel boundary, and independent ops (down-sample branch next to conv1,
// every weight gradient next to the data-gradient chain) share a phase.  Replaces ~70 dependent launches per local step.
constexpr int kMegaStages = 3;
constexpr int kMegaStageBytes = A_BYTES + 64 * 128;
constexpr int kMaxPhaseOps = 6;
struct MegaOp {
  int kind;                  // 0 gemm, 4 GroupNorm backward
  int ps;                    // gn: pixel lanes
  int pad[2];
  GemmP g;
  GnBwdP n;
};
struct MegaP {
  const MegaOp* ops;
  const int* phase_ops;      // [nphases][kMaxPhaseOps] op indices of the phase (-1: unused)
  const int4* items;         // {op slot inside the phase, bx, by, slot}
  const int* cta_start;      // [nphases][ncta + 1] offsets into items
  int nphases, ncta;
  unsigned* bar;             // [0] phase barrier, [1] exit counter (the last CTA out resets both)
};
constexpr int kMegaOpBytes = ((static_cast<int>(sizeof(MegaOp)) + 15) / 16) * 16;
constexpr int kMegaSmem = kMegaStages * kMegaStageBytes + SCRATCH_BYTES + (2 * kMaxStages + 4) * 8 + 16 +
                          kMaxPhaseOps * kMegaOpBytes + 1024;
constexpr int kEpiThreads = 128;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __maxnreg__(160) sn_mega_kernel(const __grid_constant__ MegaP mp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  TileCtx c;
  c.smem = smem;
  c.stage_bytes = kMegaStageBytes;
  c.kStages = kMegaStages;
  c.scratch = reinterpret_cast<float*>(smem + kMegaStages * kMegaStageBytes);
  c.full_bar = reinterpret_cast<uint64_t*>(smem + kMegaStages * kMegaStageBytes + SCRATCH_BYTES);
  c.empty_bar = c.full_bar + kMaxStages;
  c.acc_bar = c.empty_bar + kMaxStages;          // [2]
  c.acc_free = c.acc_bar + 2;                    // [2]
  c.nbuf = 2;
  c.acc_stride = 64;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(c.acc_free + 2);
  uint8_t* optab = reinterpret_cast<uint8_t*>(tmem_ptr) + 16;     // the phase's op records
  c.it_base = 0;
  c.acc_phase = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMegaStages; ++s) {
      mbar_init(c.full_bar + s, 1);
      mbar_init(c.empty_bar + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(c.acc_bar + b, 1);
      mbar_init(c.acc_free + b, kEpiThreads / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 128u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_ptr;

  const unsigned G = gridDim.x;
  for (int ph = 0; ph < mp.nphases; ++ph) {
    // ---- this phase's op records -> shared memory (every role reads its parameters from there) -----------------------
    const int* pops = mp.phase_ops + ph * kMaxPhaseOps;
    for (int k = 0; k < kMaxPhaseOps; ++k) {
      const int oi = __ldg(pops + k);
      if (oi < 0) break;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(mp.ops + oi);
      uint32_t* dst = reinterpret_cast<uint32_t*>(optab + k * kMegaOpBytes);
      for (int w = threadIdx.x; w < static_cast<int>(sizeof(MegaOp) / 4); w += kThreads) dst[w] = __ldg(src + w);
    }
    __syncthreads();
    if (warp == 0) {
      const int k = lane / 5, m = lane - k * 5;
      if (k < kMaxPhaseOps && __ldg(pops + k) >= 0) {
        const MegaOp* o = reinterpret_cast<const MegaOp*>(optab + k * kMegaOpBytes);
        if (o->kind == 0) prefetch_tmap(o->g.maps + m);
      }
    }
    const int* cs = mp.cta_start + static_cast<long long>(ph) * (mp.ncta + 1) + blockIdx.x;
    const int i0 = __ldg(cs), i1 = __ldg(cs + 1);
    // Roles walk the CTA's tile list independently: the producer runs ahead as far as the ring allows, the MMA thread
    // as far as the two accumulators allow, the epilogue warps drain behind them.
    for (int i = i0; i < i1; ++i) {
      const int4 item = __ldg(mp.items + i);
      const MegaOp* o = reinterpret_cast<const MegaOp*>(optab + item.x * kMegaOpBytes);
      if (o->kind == 0) {
        const GemmP& p = o->g;
        int its;
        switch (p.epi) {
          case E_STORE: its = gemm_tile<E_STORE>(p, item.y, item.z, item.w, c); break;
          case E_GNFWD: its = gemm_tile<E_GNFWD>(p, item.y, item.z, item.w, c); break;
          case E_GNBWD: its = gemm_tile<E_GNBWD>(p, item.y, item.z, item.w, c); break;
          default:      its = gemm_tile<E_WGRAD>(p, item.y, item.z, item.w, c); break;
        }
        c.it_base += static_cast<uint32_t>(its);
        c.acc_phase += its > 0 ? 1u : 0u;
      } else if (warp >= 2) {
        gn_bwd_block(o->n, o->ps, static_cast<long long>(item.y) * kEpiThreads + (threadIdx.x - 64));
      }
    }
    // global results of this phase become visible to the async proxy (TMA reads of the next phase) of every CTA
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (ph + 1 < mp.nphases) {
      // ---- grid barrier --------------------------------------------------------------------------------------------
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(mp.bar, 1u);
        const unsigned target = G * static_cast<unsigned>(ph + 1);
        // every CTA of the grid is resident (2 per SM by construction), so this wait is short; the watchdog turns a
        // scheduling surprise into a launch failure instead of a hung device
        const long long t0 = clock64();
        while (ld_acquire_u32(mp.bar) < target) {
          if (clock64() - t0 > 4000000000LL) __trap();
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(mp.bar + 1, 1u);
    if (prev == G - 1) {               // every CTA has left its last barrier: safe to re-arm for the next launch
      mp.bar[0] = 0u;
      mp.bar[1] = 0u;
      __threadfence();
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(c.tmem_base, 128u);
  }
}

// ====================================================================================================== host side
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

struct Op {
  int kind;                 // 0 gemm, 1 im2col, 2 stem fwd, 3 stem bwd, 4 gn bwd, 5 ce, 6 zero
  GemmP g; dim3 grid; int smem;
  // im2col
  const float* x; long long sN, sC, sH, sW; float* out; int N;
  StemP stem; GnBwdP gn; CeP ce;
  float* zero_ptr; long long zero_bytes;
  int side;                 // 1: may run on the side stream (weight gradients)
};

class Program {
 public:
  Program() {}
  // tensor maps ---------------------------------------------------------------------------------------------------
  // returns the index of the encoded map; dims / strides (bytes, rank-1 entries) / box innermost first
  // swizzle: 0 = SWIZZLE_128B (K-major operand tiles), 1 = SWIZZLE_128B_ATOM_32B (MN-major 32-bit operand tiles)
  int64_t add_map(int64_t ptr, std::vector<int64_t> dims, std::vector<int64_t> strides, std::vector<int64_t> box,
                  int64_t swizzle) {
    const int rank = static_cast<int>(dims.size());
    TORCH_CHECK(rank >= 3 && rank <= 5 && static_cast<int>(strides.size()) == rank - 1 && static_cast<int>(box.size()) == rank);
    CUtensorMap m;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = static_cast<cuuint64_t>(dims[i]); bx[i] = static_cast<cuuint32_t>(box[i]); es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = static_cast<cuuint64_t>(strides[i]);
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, reinterpret_cast<void*>(ptr), gd, gs, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", static_cast<int>(r), " (rank ", rank, ")");
    host_maps_.push_back(m);
    return static_cast<int64_t>(host_maps_.size()) - 1;
  }

  void add_gemm(py::dict d) {
    Op op{};
    op.kind = 0;
    GemmP& g = op.g;
    auto I = [&](const char* k, int def = 0) { return d.contains(k) ? d[k].cast<int>() : def; };
    auto L = [&](const char* k, long long def = 0) { return d.contains(k) ? d[k].cast<long long>() : def; };
    auto P = [&](const char* k) { return d.contains(k) ? reinterpret_cast<float*>(d[k].cast<long long>()) : nullptr; };
    g.mode = I("mode"); g.epi = I("epi"); g.S = I("S"); g.B = I("B");
    g.H = I("H", 1); g.W = I("W", 1); g.bw = I("bw", 1); g.bh = I("bh", 1); g.bb = I("bb", 1);
    g.tiles_y = I("tiles_y", 1); g.row_tiles = I("row_tiles", 1); g.ncls = I("ncls", 1);
    g.outH = I("outH", g.H); g.outW = I("outW", g.W);
    g.ntaps = I("ntaps", 1); g.Cred = I("Cred", 32); g.TN = I("TN", 64); g.N = I("N", 64);
    g.Cin = I("Cin", 32); g.Kw = I("Kw", 32); g.kchunks = I("kchunks", 1); g.ksplit = I("ksplit", 1);
    g.kbw = I("kbw", 1); g.kbh = I("kbh", 1); g.kbb = I("kbb", 1); g.kH = I("kH", 1);
    g.relu = I("relu", 0); g.eps = d.contains("eps") ? d["eps"].cast<float>() : 1e-5f; g.skip_cls = I("skip_cls", -1);
    auto fill = [&](const char* k, signed char* dst, int n) {
      for (int i = 0; i < n; ++i) dst[i] = 0;
      if (d.contains(k)) {
        auto v = d[k].cast<std::vector<int>>();
        TORCH_CHECK(static_cast<int>(v.size()) <= n, k, " too long");
        for (size_t i = 0; i < v.size(); ++i) dst[i] = static_cast<signed char>(v[i]);
      }
    };
    fill("tap_dx", g.tap_dx, MAX_TAPS); fill("tap_dy", g.tap_dy, MAX_TAPS);
    fill("tap_map", g.tap_map, MAX_TAPS); fill("tap_w", g.tap_w, MAX_TAPS);
    auto fill4 = [&](const char* k, int* dst) {
      for (int i = 0; i < 4; ++i) dst[i] = 0;
      if (d.contains(k)) { auto v = d[k].cast<std::vector<int>>(); for (size_t i = 0; i < v.size() && i < 4; ++i) dst[i] = v[i]; }
    };
    fill4("cls_py", g.cls_py); fill4("cls_px", g.cls_px); fill4("cls_tap0", g.cls_tap0); fill4("cls_nt", g.cls_nt);
    g.out = P("out"); g.out2 = P("out2"); g.stats = P("stats"); g.res = P("res"); g.yprev = P("yprev"); g.zprev = P("zprev");
    g.Warena = P("Warena"); g.Garena = P("Garena");
    g.arena_stride = L("arena_stride"); g.gamma_off = L("gamma_off"); g.beta_off = L("beta_off"); g.bias_off = L("bias_off", -1);
    {
      // operand descriptors: {lbo, sbo, kstep, layout}; defaults by mode
      const bool a_mn = g.mode == WGRAD, b_mn = g.mode != FPROP;
      std::vector<int> da = d.contains("a_desc") ? d["a_desc"].cast<std::vector<int>>()
                                                 : (a_mn ? std::vector<int>{BLK_BYTES, 512, 64, 1} : std::vector<int>{16, 1024, 2, 2});
      std::vector<int> db = d.contains("b_desc") ? d["b_desc"].cast<std::vector<int>>()
                                                 : (b_mn ? std::vector<int>{BLK_BYTES, 512, 64, 1} : std::vector<int>{16, 1024, 2, 2});
      TORCH_CHECK(da.size() == 4 && db.size() == 4, "a_desc / b_desc: {lbo, sbo, kstep, layout}");
      g.a_lbo = da[0]; g.a_sbo = da[1]; g.a_kstep = da[2]; g.a_layout = da[3];
      g.b_lbo = db[0]; g.b_sbo = db[1]; g.b_kstep = db[2]; g.b_layout = db[3];
    }
    auto maps = d["maps"].cast<std::vector<int64_t>>();          // 5 map indices (A0..A3, B); -1 = unused
    TORCH_CHECK(maps.size() == 5, "maps: 5 entries expected");
    op_maps_.push_back(maps);
    TORCH_CHECK(g.TN == 32 || g.TN == 64 || g.TN == 128 || g.TN == 256, "TN must be 32, 64, 128 or 256");
    const int row_tiles_total = g.mode == WGRAD ? I("row_blocks", 1) * g.ksplit : g.row_tiles * g.ncls;
    op.grid = dim3(row_tiles_total, (g.N + g.TN - 1) / g.TN, g.S);
    op.side = I("side", 0);
    {
      // Pipeline depth = shared-memory budget / stage size.  Measured: the K loop is bound by the TMA unit's row rate
      // (~4 cycles per 128-byte row), not by latency, so 3-4 stages are enough; the budget is kept small instead so that
      // an SM can hold a CTA of kernel N, the prologue of kernel N+1 (programmatic dependent launch) and a
      // weight-gradient CTA from the side stream at the same time (84 + 84 + 56 KB).
      const int stage_bytes = A_BYTES + g.TN * 128;
      const long long ctas = static_cast<long long>(op.grid.x) * op.grid.y * op.grid.z;
      (void)ctas;
      const int budget = I("smem_budget", op.side ? 56 * 1024 : 84 * 1024);
      g.stages = std::max(2, std::min(kMaxStages, budget / stage_bytes));
      g.nacc = I("nacc", 1);     // measured: extra accumulators buy nothing (the K loop is not MMA-latency bound)
      TORCH_CHECK((g.nacc == 1 || g.nacc == 2 || g.nacc == 4) && g.TN * g.nacc <= 512, "bad nacc");
      g.dbg = I("dbg", 0);
      op.smem = g.stages * stage_bytes + SCRATCH_BYTES + (2 * kMaxStages + 2) * 8 + 1024;
    }
    ops_.push_back(op);
  }

  void add_im2col(int64_t x, std::vector<int64_t> strides, int64_t out, int64_t N) {
    Op op{};
    op.kind = 1;
    op.x = reinterpret_cast<const float*>(x);
    op.sN = strides[0]; op.sC = strides[1]; op.sH = strides[2]; op.sW = strides[3];
    op.out = reinterpret_cast<float*>(out); op.N = static_cast<int>(N);
    op_maps_.push_back({});
    ops_.push_back(op);
  }

  void add_stem(bool backward, py::dict d) {
    Op op{};
    op.kind = backward ? 3 : 2;
    auto Pf = [&](const char* k) { return d.contains(k) ? reinterpret_cast<float*>(d[k].cast<long long>()) : nullptr; };
    StemP& s = op.stem;
    s.z = Pf("z"); s.stats = Pf("stats"); s.pooled = Pf("pooled");
    s.arg = reinterpret_cast<unsigned char*>(d["arg"].cast<long long>());
    s.dpool = Pf("dpool"); s.dz = Pf("dz"); s.Warena = Pf("Warena"); s.Garena = Pf("Garena");
    s.arena_stride = d["arena_stride"].cast<long long>(); s.gamma_off = d["gamma_off"].cast<long long>();
    s.beta_off = d["beta_off"].cast<long long>(); s.B = d["B"].cast<int>(); s.eps = d["eps"].cast<float>();
    op.N = d["N"].cast<int>();
    op_maps_.push_back({});
    ops_.push_back(op);
  }

  void add_gn_bwd(py::dict d) {
    Op op{};
    op.kind = 4;
    auto Pf = [&](const char* k) { return d.contains(k) ? reinterpret_cast<float*>(d[k].cast<long long>()) : nullptr; };
    GnBwdP& g = op.gn;
    g.din = Pf("din"); g.add1 = Pf("add1"); g.add2 = Pf("add2"); g.y = Pf("y"); g.z = Pf("z"); g.stats = Pf("stats");
    g.tm = Pf("tm"); g.dz = Pf("dz"); g.Warena = Pf("Warena"); g.Garena = Pf("Garena");
    g.arena_stride = d["arena_stride"].cast<long long>(); g.gamma_off = d["gamma_off"].cast<long long>();
    g.beta_off = d["beta_off"].cast<long long>();
    g.N = d["N"].cast<int>(); g.B = d["B"].cast<int>(); g.H = d["H"].cast<int>(); g.W = d["W"].cast<int>(); g.C = d["C"].cast<int>();
    op_maps_.push_back({});
    ops_.push_back(op);
  }

  void add_ce(py::dict d) {
    Op op{};
    op.kind = 5;
    CeP& c = op.ce;
    c.logits = reinterpret_cast<const float*>(d["logits"].cast<long long>());
    c.labels = reinterpret_cast<const long long*>(d["labels"].cast<long long>());
    c.dlogits = reinterpret_cast<float*>(d["dlogits"].cast<long long>());
    c.loss = reinterpret_cast<float*>(d["loss"].cast<long long>());
    c.Garena = reinterpret_cast<float*>(d["Garena"].cast<long long>());
    c.arena_stride = d["arena_stride"].cast<long long>(); c.bias_off = d["bias_off"].cast<long long>();
    c.B = d["B"].cast<int>(); c.C = d["C"].cast<int>();
    op.N = d["rows"].cast<int>();
    op_maps_.push_back({});
    ops_.push_back(op);
  }

  void add_zero(int64_t ptr, int64_t bytes) {
    Op op{};
    op.kind = 6;
    op.zero_ptr = reinterpret_cast<float*>(ptr); op.zero_bytes = bytes;
    op_maps_.push_back({});
    ops_.push_back(op);
  }

  // upload the tensor maps and resolve per-op map pointers
  void finalize() {
    TORCH_CHECK(!finalized_, "finalize() called twice");
    auto opts = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA);
    // per GEMM op: 5 consecutive maps (unused entries replicate map 0 of that op so prefetches stay valid)
    std::vector<CUtensorMap> flat;
    for (size_t i = 0; i < ops_.size(); ++i) {
      if (ops_[i].kind != 0) continue;
      const auto& idx = op_maps_[i];
      int64_t first = -1;
      for (auto v : idx) if (v >= 0 && first < 0) first = v;
      TORCH_CHECK(first >= 0, "gemm op without tensor maps");
      for (int k = 0; k < 5; ++k) flat.push_back(host_maps_[idx[k] >= 0 ? idx[k] : first]);
    }
    maps_dev_ = torch::empty({static_cast<int64_t>(flat.size() * sizeof(CUtensorMap)) + 128}, opts);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(maps_dev_.data_ptr()) + 127) & ~static_cast<uintptr_t>(127));
    FLUTE_CUDA_CHECK(cudaMemcpy(base, flat.data(), flat.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    size_t k = 0;
    int max_smem = 0;
    for (auto& op : ops_) {
      if (op.kind != 0) continue;
      op.g.maps = reinterpret_cast<const CUtensorMap*>(base) + 5 * k;
      ++k;
      max_smem = std::max(max_smem, op.smem);
    }
    if (max_smem > 0) {
      FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_gemm_kernel<E_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_gemm_kernel<E_GNFWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_gemm_kernel<E_GNBWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_gemm_kernel<E_WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemFwdSmem));
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_stem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemBwdSmem));
    finalized_ = true;
  }

  // Weight-gradient launches (side = 1) only feed the optimizer step, never the next backward kernel: fork them onto a
  // second stream (event wait behind the kernel that produced their dY) and join at the end of run().  Inside a CUDA
  // graph capture this becomes a parallel branch, so the critical path of a local step is fwd + dgrad chain only.
  void set_side_stream(bool on) { use_side_ = on; }

  // launch ops [begin, end) on the current stream; returns the number of kernels launched
  int64_t run(int64_t begin, int64_t end) {
    TORCH_CHECK(finalized_, "finalize() first");
    cudaStream_t main_stream = at::cuda::getCurrentCUDAStream();
    if (end < 0 || end > static_cast<int64_t>(ops_.size())) end = static_cast<int64_t>(ops_.size());
    if (use_side_ && side_stream_ == nullptr) {
      FLUTE_CUDA_CHECK(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
      fork_events_.resize(ops_.size(), nullptr);
      FLUTE_CUDA_CHECK(cudaEventCreateWithFlags(&join_event_, cudaEventDisableTiming));
    }
    bool side_used = false;
    int64_t n = 0;
    for (int64_t i = begin; i < end; ++i) {
      if (use_mega_) {
        const Mega* hit = nullptr;
        for (const auto& m : megas_) if (m.begin == i && m.end <= end) hit = &m;
        if (hit != nullptr) {
          launch_mega(*hit, main_stream);
          i = hit->end - 1;
          ++n;
          continue;
        }
      }
      const Op& op = ops_[i];
      cudaStream_t stream = main_stream;
      if (use_side_ && op.side) {
        if (fork_events_[i] == nullptr) FLUTE_CUDA_CHECK(cudaEventCreateWithFlags(&fork_events_[i], cudaEventDisableTiming));
        FLUTE_CUDA_CHECK(cudaEventRecord(fork_events_[i], main_stream));
        FLUTE_CUDA_CHECK(cudaStreamWaitEvent(side_stream_, fork_events_[i], 0));
        stream = side_stream_;
        side_used = true;
      }
      switch (op.kind) {
        case 0:
          launch_gemm(op, stream);
          break;
        case 1: {
          const long long total = static_cast<long long>(op.N) * 256 * 40;
          sn_im2col_stem_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(op.x, op.sN, op.sC, op.sH, op.sW,
                                                                                              op.out, op.N);
          break;
        }
        case 2:
          sn_stem_fwd_kernel<<<dim3(op.N, 2), 256, kStemFwdSmem, stream>>>(op.stem);
          break;
        case 3:
          sn_stem_bwd_kernel<<<dim3(op.N, 2), 256, kStemBwdSmem, stream>>>(op.stem);
          break;
        case 4: {
          const int hw = op.gn.H * op.gn.W;
          const int PS = hw >= 8 ? 8 : hw;             // H, W are powers of two
          const long long groups = static_cast<long long>(op.gn.N) * (op.gn.C / 2);
          const long long warps = (groups + (32 / PS) - 1) / (32 / PS);
          sn_gn_bwd_kernel<<<static_cast<unsigned>((warps + 3) / 4), 128, 0, stream>>>(op.gn, PS);
          break;
        }
        case 5:
          sn_ce_kernel<<<op.N, 256, 0, stream>>>(op.ce);
          break;
        case 6:
          FLUTE_CUDA_CHECK(cudaMemsetAsync(op.zero_ptr, 0, static_cast<size_t>(op.zero_bytes), stream));
          break;
      }
      ++n;
    }
    if (side_used) {
      FLUTE_CUDA_CHECK(cudaEventRecord(join_event_, side_stream_));
      FLUTE_CUDA_CHECK(cudaStreamWaitEvent(main_stream, join_event_, 0));
    }
    FLUTE_CUDA_CHECK(cudaGetLastError());
    return n;
  }

  int64_t num_ops() const { return static_cast<int64_t>(ops_.size()); }
  void set_pdl(bool on) { use_pdl_ = on; }
  void set_mega(bool on) { use_mega_ = on; }

  // Fuse ops [begin, end) (GEMMs and stand-alone GroupNorm backward only) into ONE persistent cooperative launch.
  // phases[i] = dependency level of op begin + i (ops of one phase are mutually independent).  Returns the number of
  // resident CTAs the work was balanced over.
  int64_t add_mega(int64_t begin, int64_t end, std::vector<int> phases, int64_t ctas_per_sm) {
    TORCH_CHECK(finalized_, "finalize() first");
    TORCH_CHECK(begin >= 0 && end <= static_cast<int64_t>(ops_.size()) && begin < end &&
                static_cast<int64_t>(phases.size()) == end - begin, "add_mega: bad range");
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMegaSmem));
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn_mega_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
    int dev = 0, sms = 0, occ = 0;
    FLUTE_CUDA_CHECK(cudaGetDevice(&dev));
    FLUTE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FLUTE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sn_mega_kernel, kThreads, kMegaSmem));
    TORCH_CHECK(occ >= 1, "persistent step kernel does not fit on an SM");
    max_occ_ = occ;
    occ = ctas_per_sm > 0 ? static_cast<int>(ctas_per_sm) : 2;   // see launch_mega() for why the API's answer (1) is not used
    TORCH_CHECK(occ >= 1 && occ <= 2, "persistent step kernel: 1 or 2 CTAs per SM");
    const int ncta = occ * sms;
    int nph = 0;
    for (int v : phases) { TORCH_CHECK(v >= 0, "negative phase"); nph = std::max(nph, v + 1); }

    struct Item { int op, bx, by, slot; float cost; };
    std::vector<int> phase_ops(static_cast<size_t>(nph) * kMaxPhaseOps, -1);
    std::vector<int> phase_nops(nph, 0);
    std::vector<MegaOp> mops;
    std::vector<std::vector<Item>> by_phase(nph);
    for (int64_t i = begin; i < end; ++i) {
      const Op& op = ops_[i];
      MegaOp mo{};
      const int ph_i = phases[i - begin];
      TORCH_CHECK(phase_nops[ph_i] < kMaxPhaseOps, "add_mega: more than ", kMaxPhaseOps, " ops in phase ", ph_i);
      const int oi = phase_nops[ph_i]++;                      // slot inside the phase's op table
      phase_ops[static_cast<size_t>(ph_i) * kMaxPhaseOps + oi] = static_cast<int>(mops.size());
      auto& dst = by_phase[ph_i];
      if (op.kind == 0) {
        mo.kind = 0;
        mo.g = op.g;
        mo.g.nacc = 1;
        TORCH_CHECK(mo.g.TN <= 64, "persistent step kernel: TN <= 64");
        const GemmP& g = mo.g;
        for (unsigned z = 0; z < op.grid.z; ++z)
          for (unsigned y = 0; y < op.grid.y; ++y)
            for (unsigned x = 0; x < op.grid.x; ++x) {
              int its;
              if (g.mode == WGRAD) {
                const int per = (g.kchunks + g.ksplit - 1) / g.ksplit;
                const int ks = static_cast<int>(x) % g.ksplit;
                its = std::max(0, std::min(g.kchunks, ks * per + per) - ks * per);
              } else {
                const int nt = g.ncls == 4 ? g.cls_nt[x / g.row_tiles] : g.ntaps;
                its = nt * ((g.Cred + 31) / 32);
              }
              const float fixed = g.epi == E_WGRAD ? 16.f : (g.epi == E_STORE ? 10.f : 13.f);
              dst.push_back({oi, static_cast<int>(x), static_cast<int>(y), static_cast<int>(z), fixed + static_cast<float>(its)});
            }
      } else if (op.kind == 4) {
        mo.kind = 4;
        mo.n = op.gn;
        const int hw = op.gn.H * op.gn.W;
        mo.ps = hw >= 8 ? 8 : hw;
        const long long groups = static_cast<long long>(op.gn.N) * (op.gn.C / 2);
        const long long warps = (groups + (32 / mo.ps) - 1) / (32 / mo.ps);
        const int blocks = static_cast<int>((warps + (kEpiThreads / 32) - 1) / (kEpiThreads / 32));
        for (int b = 0; b < blocks; ++b) dst.push_back({oi, b, 0, 0, 8.f + 0.5f * static_cast<float>(hw / mo.ps)});
      } else {
        TORCH_CHECK(false, "add_mega: op ", i, " of kind ", op.kind, " cannot run inside the persistent kernel");
      }
      mops.push_back(mo);
    }
    // longest-processing-time-first balancing of every phase over the resident CTAs
    std::vector<int4> items;
    std::vector<int> cta_start(static_cast<size_t>(nph) * (ncta + 1), 0);
    float crit = 0.f;
    for (int ph = 0; ph < nph; ++ph) {
      auto& v = by_phase[ph];
      std::stable_sort(v.begin(), v.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
      std::vector<std::vector<int>> mine(ncta);
      std::priority_queue<std::pair<float, int>, std::vector<std::pair<float, int>>, std::greater<std::pair<float, int>>> pq;
      for (int cta = 0; cta < ncta; ++cta) pq.push({0.f, cta});
      float worst = 0.f;
      for (size_t k = 0; k < v.size(); ++k) {
        auto top = pq.top(); pq.pop();
        mine[top.second].push_back(static_cast<int>(k));
        top.first += v[k].cost;
        worst = std::max(worst, top.first);
        pq.push(top);
      }
      crit += worst;
      for (int cta = 0; cta < ncta; ++cta) {
        // keep tiles of the same op adjacent so the shared op copy is reloaded as rarely as possible
        std::stable_sort(mine[cta].begin(), mine[cta].end(), [&](int a, int b) { return v[a].op < v[b].op; });
        cta_start[static_cast<size_t>(ph) * (ncta + 1) + cta] = static_cast<int>(items.size());
        for (int k : mine[cta]) items.push_back(make_int4(v[k].op, v[k].bx, v[k].by, v[k].slot));
      }
      cta_start[static_cast<size_t>(ph) * (ncta + 1) + ncta] = static_cast<int>(items.size());
    }
    Mega m;
    m.begin = begin; m.end = end; m.grid = ncta; m.nphases = nph; m.est_cost = crit;
    auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA);
    m.ops_dev = torch::empty({static_cast<int64_t>(mops.size() * sizeof(MegaOp))}, u8);
    m.items_dev = torch::empty({static_cast<int64_t>(std::max<size_t>(1, items.size()) * sizeof(int4))}, u8);
    m.cta_dev = torch::empty({static_cast<int64_t>(cta_start.size() * sizeof(int))}, u8);
    m.pops_dev = torch::empty({static_cast<int64_t>(phase_ops.size() * sizeof(int))}, u8);
    FLUTE_CUDA_CHECK(cudaMemcpy(m.pops_dev.data_ptr(), phase_ops.data(), phase_ops.size() * sizeof(int), cudaMemcpyHostToDevice));
    m.p.phase_ops = reinterpret_cast<const int*>(m.pops_dev.data_ptr());
    m.bar_dev = torch::zeros({64}, u8);
    FLUTE_CUDA_CHECK(cudaMemcpy(m.ops_dev.data_ptr(), mops.data(), mops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice));
    FLUTE_CUDA_CHECK(cudaMemcpy(m.items_dev.data_ptr(), items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
    FLUTE_CUDA_CHECK(cudaMemcpy(m.cta_dev.data_ptr(), cta_start.data(), cta_start.size() * sizeof(int), cudaMemcpyHostToDevice));
    m.p.ops = reinterpret_cast<const MegaOp*>(m.ops_dev.data_ptr());
    m.p.items = reinterpret_cast<const int4*>(m.items_dev.data_ptr());
    m.p.cta_start = reinterpret_cast<const int*>(m.cta_dev.data_ptr());
    m.p.nphases = nph; m.p.ncta = ncta;
    m.p.bar = reinterpret_cast<unsigned*>(m.bar_dev.data_ptr());
    megas_.push_back(m);
    use_mega_ = true;
    return ncta;
  }
  // {begin, end, phases, ctas, estimated critical path in ring iterations} of every fused range
  std::vector<std::vector<double>> mega_info() const {
    std::vector<std::vector<double>> r;
    for (const auto& m : megas_)
      r.push_back({static_cast<double>(m.begin), static_cast<double>(m.end), static_cast<double>(m.nphases),
                   static_cast<double>(m.grid), static_cast<double>(m.est_cost), static_cast<double>(max_occ_),
                   static_cast<double>(kMegaSmem)});
    return r;
  }

 private:
  struct Mega {
    int64_t begin, end; int grid, nphases; float est_cost;
    torch::Tensor ops_dev, items_dev, cta_dev, bar_dev, pops_dev;
    MegaP p;
  };
  void launch_mega(const Mega& m, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(m.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(kMegaSmem);
    cfg.stream = stream;
    // Not a cooperative launch: the occupancy API reports 1 CTA/SM for every kernel that allocates tensor memory
    // (measured: even 64-thread / 40 KB configurations of this kernel and the per-launch GEMM kernels, which ncu shows
    // running 2 per SM), so cudaLaunchCooperativeKernel would cap the grid at 148.  Two CTAs of 80 KB / 160 registers x
    // 192 threads / 128 TMEM columns fit on an SM and nothing else runs on this stream's SMs while the step kernel does
    // (ordinary stream order, no programmatic overlap), so all 296 CTAs become resident; the in-kernel watchdog traps
    // instead of hanging if that assumption is ever violated.
    const cudaError_t e = cudaLaunchKernelEx(&cfg, sn_mega_kernel, m.p);
    TORCH_CHECK(e == cudaSuccess, "persistent step kernel launch failed (", m.grid, " CTAs, ", kMegaSmem, " B smem): ",
                cudaGetErrorString(e));
  }
  void launch_gemm(const Op& op, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = op.grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(op.smem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl_ ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    switch (op.g.epi) {
      case E_STORE: e = cudaLaunchKernelEx(&cfg, sn_gemm_kernel<E_STORE>, op.g); break;
      case E_GNFWD: e = cudaLaunchKernelEx(&cfg, sn_gemm_kernel<E_GNFWD>, op.g); break;
      case E_GNBWD: e = cudaLaunchKernelEx(&cfg, sn_gemm_kernel<E_GNBWD>, op.g); break;
      default: e = cudaLaunchKernelEx(&cfg, sn_gemm_kernel<E_WGRAD>, op.g); break;
    }
    FLUTE_CUDA_CHECK(e);
  }

 public:

 private:
  std::vector<CUtensorMap> host_maps_;
  std::vector<std::vector<int64_t>> op_maps_;
  std::vector<Op> ops_;
  torch::Tensor maps_dev_;
  bool finalized_ = false;
  bool use_side_ = false;
  bool use_pdl_ = true;
  bool use_mega_ = false;
  int max_occ_ = 0;
  std::vector<Mega> megas_;
  cudaStream_t side_stream_ = nullptr;
  cudaEvent_t join_event_ = nullptr;
  std::vector<cudaEvent_t> fork_events_;
};

}  // namespace sn

// resource usage / occupancy of the persistent kernel (diagnostics)
static std::vector<int64_t> slotnet_mega_attrs(int64_t smem) {
  cudaFuncAttributes a{};
  FLUTE_CUDA_CHECK(cudaFuncGetAttributes(&a, sn::sn_mega_kernel));
  FLUTE_CUDA_CHECK(cudaFuncSetAttribute(sn::sn_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sn::kMegaSmem));
  int occ = 0;
  FLUTE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sn::sn_mega_kernel, sn::kThreads,
                                                                static_cast<size_t>(smem > 0 ? smem : sn::kMegaSmem)));
  std::vector<int64_t> r = {a.numRegs, static_cast<int64_t>(a.sharedSizeBytes), static_cast<int64_t>(a.localSizeBytes),
                            a.maxThreadsPerBlock, occ, sn::kMegaSmem};
  for (int threads : {64, 128, 192, 256, 384}) {
    int o = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sn::sn_mega_kernel, threads, 40000);
    r.push_back(o);
  }
  size_t avail = 0;
  cudaOccupancyAvailableDynamicSMemPerBlock(&avail, sn::sn_mega_kernel, 2, sn::kThreads);
  r.push_back(static_cast<int64_t>(avail));
  {
    cudaFuncAttributes g{};
    cudaFuncGetAttributes(&g, sn::sn_gemm_kernel<sn::E_GNBWD>);
    cudaFuncSetAttribute(sn::sn_gemm_kernel<sn::E_GNBWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 90000);
    int o = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sn::sn_gemm_kernel<sn::E_GNBWD>, sn::kThreads, 85000);
    r.push_back(g.numRegs);
    r.push_back(o);
  }
  cudaGetLastError();
  return r;
}

void bind_slotnet(py::module_& m) {
  m.def("slotnet_mega_attrs", &slotnet_mega_attrs, py::arg("smem") = 0);
  py::class_<sn::Program>(m, "SlotProgram")
      .def(py::init<>())
      .def("add_map", &sn::Program::add_map, py::arg("ptr"), py::arg("dims"), py::arg("strides"), py::arg("box"),
           py::arg("swizzle") = 0)
      .def("add_gemm", &sn::Program::add_gemm)
      .def("add_im2col", &sn::Program::add_im2col)
      .def("add_stem", &sn::Program::add_stem)
      .def("add_gn_bwd", &sn::Program::add_gn_bwd)
      .def("add_ce", &sn::Program::add_ce)
      .def("add_zero", &sn::Program::add_zero)
      .def("finalize", &sn::Program::finalize)
      .def("run", &sn::Program::run, py::arg("begin") = 0, py::arg("end") = -1)
      .def("set_side_stream", &sn::Program::set_side_stream)
      .def("set_pdl", &sn::Program::set_pdl)
      .def("set_mega", &sn::Program::set_mega)
      .def("add_mega", &sn::Program::add_mega, py::arg("begin"), py::arg("end"), py::arg("phases"), py::arg("ctas_per_sm") = 0)
      .def("mega_info", &sn::Program::mega_info)
      .def("num_ops", &sn::Program::num_ops);
}

}  // namespace flute
