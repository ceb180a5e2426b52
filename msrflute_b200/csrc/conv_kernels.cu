// Slot-batched direct convolution (fprop / dgrad / wgrad) for the multi-client training engine.
//
// In a federated round every simulated client ("slot") has its OWN copy of the weights, so the S forward/backward
// passes of a wave cannot share a weight tile — but they share everything else.  These kernels run one layer for
// all S slots in ONE launch: blockIdx.z walks (slot, split) and the per-slot weight / weight-gradient tensors are
// addressed straight inside the [S, P] parameter arenas (base pointer + slot * arena_stride + tensor_offset), so the
// weight gradient lands in the gradient arena where the fused clip/SGD kernel expects it — no per-client launches,
// no layout conversions, no grad copies.
//
// Shapes in the FL benchmarks are tiny (ResNet-18 on 32x32 inputs, batch 20: M = B*Ho*Wo in {5120, 1280, 320, 80, 20}
// output pixels per slot) and the deep layers are weight-bandwidth / latency bound (a 512x512x3x3 filter is 9.4 MB per
// slot and is used for 20 output pixels).  Formulation: implicit GEMM with on-the-fly im2col gathering,
//     fprop : Y[M, Cout]      = A(x)[M, K]        * W[Cout, K]^T          K = Cin*taps
//     dgrad : dX[Mi, Cin]     = A'(dY)[Mi, K']    * W'[Cin, K']^T         K' = Cout*taps
//     wgrad : dW[Cout, K]    += dY^T[Cout, M]     * A(x)[M, K]
// * 64x64 output tiles, 16-deep chunks staged in shared memory, 256 threads x (4x4) register tiles of fp32 FMAs;
// * software pipelined: the global loads of chunk i+1 are issued into registers before the FMAs of chunk i;
// * TAP PRUNING: filter taps that only ever see zero padding are dropped from the reduction.  On a 1x1 feature map
//   (ResNet-18 layer4 with CIFAR-size inputs: 60 % of all parameters) a 3x3/pad-1 convolution is exactly its centre
//   tap — 9x less work and weight traffic, and the gradient of the other 8 taps is identically zero;
// * when the tile grid would not fill the 148 SMs the reduction is split across blockIdx.z and partial tiles are
//   combined with fp32 atomics (outputs pre-zeroed; the gradient arena is zero between steps).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/types.h>
#include <cstdlib>
#include "common.cuh"
#include "tcgen05.cuh"

namespace flute {
namespace conv {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256, MAX_TAPS = 49;

struct ConvP {
  int S, B, Cin, Hi, Wi, Cout, Ho, Wo, KH, KW, stride, pad;
  int splits;                 // reduction splits (blockIdx.z = slot * splits + split)
  int ntaps;                  // taps that touch real data for at least one output position
  int compact;                // 1: the weight tensor stores ONLY the live taps, [Cout, Cin, ntaps] (compact slot arenas)
  // tcgen05 dgrad of a stride-2 convolution is split into its 4 parity classes (ih+pad, iw+pad mod 2): an input pixel of
  // class (ph, pw) only receives taps with kh = ph, kw = pw (mod 2), so each class is a dense stride-1-like problem with
  // ~1/4 of the taps instead of a 4x larger masked one.  cls = ph * 2 + pw.
  int ncls;                       // 0 = no decomposition, 4 = by parity
  int cls_tile0[5];               // first row-tile (blockIdx.x) of each class, [4] = total
  int cls_h0[2], cls_hc[2];       // first ih of parity ph and how many rows have it
  int cls_w0[2], cls_wc[2];
  unsigned char cls_ntaps[4];
  unsigned char cls_tap_idx[4][16];   // indices into taps[]
  long long w_slot_stride;    // floats between two slots' copies of this weight tensor (= arena row length P)
  unsigned short taps[MAX_TAPS];   // kh << 8 | kw
};

// 4x4 register tile FMA over one staged chunk.  As/Bs are [BK][64+4] (k-major) so a warp reads consecutive columns.
__device__ __forceinline__ void tile_fma(const float (&As)[BK][BM + 4], const float (&Bs)[BK][BN + 4], float (&acc)[4][4],
                                         int tm, int tn) {
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------------------------ fprop
// x [S,B,Cin,Hi,Wi]  w: slot s at w + s*w_slot_stride, [Cout, Cin*KH*KW]  y [S,B,Cout,Ho,Wo]
__global__ void __launch_bounds__(THREADS) conv_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            float* __restrict__ y, const ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Ho * p.Wo, N = p.Cout, KHW = p.KH * p.KW, K = p.Cin * p.ntaps;
  const int Kfull = p.compact ? K : p.Cin * KHW;       // row pitch of the stored filter
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kchunks = (K + BK - 1) / BK;
  const int per = (kchunks + p.splits - 1) / p.splits;
  const int kc_begin = split * per, kc_end = min(kchunks, kc_begin + per);
  const float* xs = x + static_cast<long long>(slot) * p.B * p.Cin * p.Hi * p.Wi;
  const float* ws = w + static_cast<long long>(slot) * p.w_slot_stride;
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  float acc[4][4] = {};
  const int a_m = tid & 63, a_k = tid >> 6;          // A: 64 m (fastest, coalesced x reads) x 4 k per pass, 4 passes
  const int b_k = tid & 15, b_n = tid >> 4;          // B: 16 k (fastest, W rows are K-contiguous) x 16 n per pass
  long long xbase = -1;
  int ih0 = 0, iw0 = 0;
  {
    const int m = m0 + a_m;
    if (m < M) {
      const int b = m / (p.Ho * p.Wo), r = m - b * p.Ho * p.Wo, oh = r / p.Wo, ow = r - oh * p.Wo;
      xbase = static_cast<long long>(b) * p.Cin * p.Hi * p.Wi;
      ih0 = oh * p.stride - p.pad;
      iw0 = ow * p.stride - p.pad;
    }
  }
  float ra[4], rb[4];
  auto load_chunk = [&](int kc) {
    const int k0 = kc * BK;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int k = k0 + a_k + pass * 4;
      float v = 0.f;
      if (xbase >= 0 && k < K) {
        const int ci = k / p.ntaps, t = p.taps[k - ci * p.ntaps];
        const int ih = ih0 + (t >> 8), iw = iw0 + (t & 255);
        if (ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi) v = __ldg(xs + xbase + (static_cast<long long>(ci) * p.Hi + ih) * p.Wi + iw);
      }
      ra[pass] = v;
    }
    const int kb = k0 + b_k;
    int woff = -1;
    if (kb < K) {
      const int ci = kb / p.ntaps, t = p.taps[kb - ci * p.ntaps];
      woff = p.compact ? kb : ci * KHW + (t >> 8) * p.KW + (t & 255);
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int n = n0 + b_n + pass * 16;
      rb[pass] = (woff >= 0 && n < N) ? __ldg(ws + static_cast<long long>(n) * Kfull + woff) : 0.f;
    }
  };
  if (kc_begin < kc_end) load_chunk(kc_begin);
  for (int kc = kc_begin; kc < kc_end; ++kc) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      As[a_k + pass * 4][a_m] = ra[pass];
      Bs[b_k][b_n + pass * 16] = rb[pass];
    }
    __syncthreads();
    if (kc + 1 < kc_end) load_chunk(kc + 1);          // next chunk's global loads fly during the FMAs
    tile_fma(As, Bs, acc, tm, tn);
    __syncthreads();
  }
  float* ys = y + static_cast<long long>(slot) * p.B * p.Cout * p.Ho * p.Wo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm * 4 + i;
    if (m >= M) continue;
    const int b = m / (p.Ho * p.Wo), r = m - b * p.Ho * p.Wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n >= N) continue;
      float* dst = ys + (static_cast<long long>(b) * p.Cout + n) * p.Ho * p.Wo + r;
      if (p.splits > 1) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------------ dgrad
// dx[S,B,Cin,Hi,Wi] = sum_{co,kh,kw} dy[b,co,oh,ow] * w[co,ci,kh,kw]  with  ih = oh*stride - pad + kh
__global__ void __launch_bounds__(THREADS) conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                            float* __restrict__ dx, const ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Hi * p.Wi, N = p.Cin, KHW = p.KH * p.KW, K = p.Cout * p.ntaps;     // reduction over (co, tap)
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kchunks = (K + BK - 1) / BK;
  const int per = (kchunks + p.splits - 1) / p.splits;
  const int kc_begin = split * per, kc_end = min(kchunks, kc_begin + per);
  const float* dys = dy + static_cast<long long>(slot) * p.B * p.Cout * p.Ho * p.Wo;
  const float* ws = w + static_cast<long long>(slot) * p.w_slot_stride;
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  float acc[4][4] = {};
  const int a_m = tid & 63, a_k = tid >> 6;
  const int b_k = tid & 15, b_n = tid >> 4;
  long long dybase = -1;
  int th0 = 0, tw0 = 0;
  {
    const int m = m0 + a_m;
    if (m < M) {
      const int b = m / (p.Hi * p.Wi), r = m - b * p.Hi * p.Wi, ih = r / p.Wi, iw = r - ih * p.Wi;
      dybase = static_cast<long long>(b) * p.Cout * p.Ho * p.Wo;
      th0 = ih + p.pad;
      tw0 = iw + p.pad;
    }
  }
  const int wtaps = p.compact ? p.ntaps : KHW;         // taps stored per (co, ci)
  const int CinKHW = p.Cin * wtaps;
  float ra[4], rb[4];
  auto load_chunk = [&](int kc) {
    const int k0 = kc * BK;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int k = k0 + a_k + pass * 4;
      float v = 0.f;
      if (dybase >= 0 && k < K) {
        const int co = k / p.ntaps, t = p.taps[k - co * p.ntaps];
        const int th = th0 - (t >> 8), tw = tw0 - (t & 255);               // = oh*stride, ow*stride
        if (th >= 0 && tw >= 0) {
          int oh = th, ow = tw;
          bool ok = true;
          if (p.stride != 1) { oh = th / p.stride; ow = tw / p.stride; ok = (oh * p.stride == th) && (ow * p.stride == tw); }
          if (ok && oh < p.Ho && ow < p.Wo) v = __ldg(dys + dybase + (static_cast<long long>(co) * p.Ho + oh) * p.Wo + ow);
        }
      }
      ra[pass] = v;
    }
    const int kb = k0 + b_k;
    long long woff = -1;
    if (kb < K) {
      const int co = kb / p.ntaps, ti = kb - co * p.ntaps, t = p.taps[ti];
      woff = static_cast<long long>(co) * CinKHW + (p.compact ? ti : (t >> 8) * p.KW + (t & 255));
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int ci = n0 + b_n + pass * 16;
      rb[pass] = (woff >= 0 && ci < N) ? __ldg(ws + woff + static_cast<long long>(ci) * wtaps) : 0.f;
    }
  };
  if (kc_begin < kc_end) load_chunk(kc_begin);
  for (int kc = kc_begin; kc < kc_end; ++kc) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      As[a_k + pass * 4][a_m] = ra[pass];
      Bs[b_k][b_n + pass * 16] = rb[pass];
    }
    __syncthreads();
    if (kc + 1 < kc_end) load_chunk(kc + 1);
    tile_fma(As, Bs, acc, tm, tn);
    __syncthreads();
  }
  float* dxs = dx + static_cast<long long>(slot) * p.B * p.Cin * p.Hi * p.Wi;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm * 4 + i;
    if (m >= M) continue;
    const int b = m / (p.Hi * p.Wi), r = m - b * p.Hi * p.Wi;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n >= N) continue;
      float* dst = dxs + (static_cast<long long>(b) * p.Cin + n) * p.Hi * p.Wi + r;
      if (p.splits > 1) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------------ wgrad
// dw[slot][co, k] += sum_m dy[m, co] * A(x)[m, k]     (tile rows = co, tile cols = pruned k, reduction over m)
__global__ void __launch_bounds__(THREADS) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw, const ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];     // dy^T chunk: [m-chunk][co]
  __shared__ __align__(16) float Bs[BK][BN + 4];     // im2col chunk: [m-chunk][k]
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Ho * p.Wo, KHW = p.KH * p.KW, K = p.Cin * p.ntaps, HoWo = p.Ho * p.Wo;
  const int Kfull = p.compact ? K : p.Cin * KHW;
  const int co0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
  const int mchunks = (M + BK - 1) / BK;
  const int per = (mchunks + p.splits - 1) / p.splits;
  const int mc_begin = split * per, mc_end = min(mchunks, mc_begin + per);
  const float* xs = x + static_cast<long long>(slot) * p.B * p.Cin * p.Hi * p.Wi;
  const float* dys = dy + static_cast<long long>(slot) * p.B * p.Cout * HoWo;
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  float acc[4][4] = {};
  const int l_m = tid & 15, l_c = tid >> 4;          // 16 m (fastest: contiguous in dy, nearly so in x) x 16 cols per pass
  int kci[4], kkh[4], kkw[4];                        // this thread's 4 reduction-free k columns, decoded once
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int k = k0 + l_c + pass * 16;
    if (k < K) { kci[pass] = k / p.ntaps; const int t = p.taps[k - kci[pass] * p.ntaps]; kkh[pass] = t >> 8; kkw[pass] = t & 255; }
    else kci[pass] = -1;
  }
  float ra[4], rb[4];
  auto load_chunk = [&](int mc) {
    const int m = mc * BK + l_m;
    const bool mv = m < M;
    int b = 0, oh = 0, ow = 0;
    if (mv) { b = m / HoWo; const int r = m - b * HoWo; oh = r / p.Wo; ow = r - oh * p.Wo; }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int co = co0 + l_c + pass * 16;
      ra[pass] = (mv && co < p.Cout) ? __ldg(dys + (static_cast<long long>(b) * p.Cout + co) * HoWo + oh * p.Wo + ow) : 0.f;
      float v = 0.f;
      if (mv && kci[pass] >= 0) {
        const int ih = oh * p.stride - p.pad + kkh[pass], iw = ow * p.stride - p.pad + kkw[pass];
        if (ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi)
          v = __ldg(xs + ((static_cast<long long>(b) * p.Cin + kci[pass]) * p.Hi + ih) * p.Wi + iw);
      }
      rb[pass] = v;
    }
  };
  if (mc_begin < mc_end) load_chunk(mc_begin);
  for (int mc = mc_begin; mc < mc_end; ++mc) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      As[l_m][l_c + pass * 16] = ra[pass];
      Bs[l_m][l_c + pass * 16] = rb[pass];
    }
    __syncthreads();
    if (mc + 1 < mc_end) load_chunk(mc + 1);
    tile_fma(As, Bs, acc, tm, tn);
    __syncthreads();
  }
  float* dws = dw + static_cast<long long>(slot) * p.w_slot_stride;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + tn * 4 + j;
    if (k >= K) continue;
    const int ci = k / p.ntaps, t = p.taps[k - ci * p.ntaps];
    const int woff = p.compact ? k : ci * KHW + (t >> 8) * p.KW + (t & 255);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + tm * 4 + i;
      if (co < p.Cout) atomicAdd(dws + static_cast<long long>(co) * Kfull + woff, acc[i][j]);   // arena is zero between steps
    }
  }
}

static int pick_splits(long long tiles, int red_chunks) {
  // aim for >= 4 CTAs per SM; keep at least 2 reduction chunks per split
  const long long target = 148LL * 4;
  if (tiles >= target) return 1;
  long long s = (target + tiles - 1) / tiles;
  s = std::min<long long>(s, std::max(1, red_chunks / 2));
  return static_cast<int>(std::max<long long>(1, std::min<long long>(s, 32)));
}

static void set_taps(ConvP& p) {
  TORCH_CHECK(p.KH * p.KW <= MAX_TAPS, "kernel larger than 7x7 is not supported");
  p.ntaps = 0;
  for (int kh = 0; kh < p.KH; ++kh) {
    bool hv = false;
    for (int oh = 0; oh < p.Ho && !hv; ++oh) { const int ih = oh * p.stride - p.pad + kh; hv = ih >= 0 && ih < p.Hi; }
    if (!hv) continue;
    for (int kw = 0; kw < p.KW; ++kw) {
      bool wv = false;
      for (int ow = 0; ow < p.Wo && !wv; ++ow) { const int iw = ow * p.stride - p.pad + kw; wv = iw >= 0 && iw < p.Wi; }
      if (wv) p.taps[p.ntaps++] = static_cast<unsigned short>((kh << 8) | kw);
    }
  }
  TORCH_CHECK(p.ntaps > 0, "convolution never touches its input");
}

static ConvP make_params(int64_t S, int64_t B, int64_t Cin, int64_t Hi, int64_t Wi, int64_t Cout, int64_t KH, int64_t KW,
                         int64_t stride, int64_t pad, int64_t w_slot_stride) {
  ConvP p;
  p.S = static_cast<int>(S); p.B = static_cast<int>(B); p.Cin = static_cast<int>(Cin);
  p.Hi = static_cast<int>(Hi); p.Wi = static_cast<int>(Wi);
  p.Cout = static_cast<int>(Cout); p.KH = static_cast<int>(KH); p.KW = static_cast<int>(KW);
  p.stride = static_cast<int>(stride); p.pad = static_cast<int>(pad);
  p.Ho = (p.Hi + 2 * p.pad - p.KH) / p.stride + 1;
  p.Wo = (p.Wi + 2 * p.pad - p.KW) / p.stride + 1;
  p.w_slot_stride = w_slot_stride;
  p.splits = 1;
  p.compact = 0;
  p.ncls = 0;
  set_taps(p);
  return p;
}

static void check5(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.dim() == 5 && t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), name,
              " must be contiguous fp32 CUDA [S, B, C, H, W]");
}
static void check_arena(const at::Tensor& a, int64_t S) {
  TORCH_CHECK(a.dim() == 2 && a.is_cuda() && a.scalar_type() == at::kFloat && a.is_contiguous() && a.size(0) >= S,
              "arena must be a contiguous fp32 CUDA [S, P] tensor");
}


// ====================================================================================================================
// tcgen05 path: the same three implicit GEMMs on the 5th-generation tensor cores (kind::tf32, fp32 accumulate in TMEM).
//
// The operand tiles of a convolution are gathers (im2col with padding / stride / pruned taps, and weights that live
// inside the parameter arena in [Cout, Cin, KH, KW] order), so TMA cannot fetch them from NCHW tensors; instead eight
// PRODUCER warps gather 128 x 32 (A) and 64 x 32 (B) fp32 tiles straight into the K-major SWIZZLE_128B shared-memory
// layout that tcgen05.mma consumes, fence them to the async proxy and arrive on the stage's mbarrier.  One elected
// thread of warp 8 issues four UMMA 128x64x8 per stage and releases the stage with tcgen05.commit; after the last stage
// the producer warps become the epilogue: tcgen05.ld the fp32 accumulator (thread == tile row) and scatter it to NCHW /
// the gradient arena (plain stores, or fp32 atomics when the reduction is split over blockIdx.z).
//
// Versus the FMA kernels above this removes the 256 FMAs + 32 LDS.128 per thread per 16-deep chunk (the tensor core does
// a 128x64x32 stage asynchronously) and decodes each reduction index once per row instead of once per element:
//   * "row" gathers   (fprop/dgrad A): thread == tile row (output pixel), walks 16 reduction indices (ci, tap) with an
//     incremental decode; consecutive threads read consecutive pixels (coalesced) and write one 16-byte chunk each;
//   * "lane" gathers  (weights, and both wgrad operands): lane == reduction index (contiguous in memory), loop over
//     tile rows; a warp writes one full 128-byte swizzled row per store (bank-conflict free).
namespace tcv {
using namespace tc;

constexpr int TM = 128, TN = 64, TK = 32, T_UMMA_K = 8;
// N-tile width is a template parameter: 64 (2 CTAs/SM), 128 (3 stages, 2 CTAs/SM) or 256 (1 CTA/SM).  A wide tile gathers
// the expensive A operand (im2col rows) ONCE for all output channels instead of once per 64 of them.
constexpr int t_stages(int tn) { return tn == 128 ? 3 : 4; }
constexpr int t_stage_bytes(int tn) { return TM * TK * 4 + tn * TK * 4; }
constexpr int T_PRODUCERS = 256, T_THREADS = 288;
constexpr int T_A_BYTES = TM * TK * 4;
constexpr int T_TABLE_BYTES = TM * 8 + 64 * 8 + 64 * 4;    // wgrad row table + per-tap (offset, kh, kw) table + tap ids
constexpr int t_smem(int tn) { return t_stages(tn) * t_stage_bytes(tn) + T_TABLE_BYTES + 256 + 1024; }
enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

__device__ __forceinline__ void st_elem(uint8_t* tile, int row, int e, float v) {     // element e (0..31) of a tile row
  *reinterpret_cast<float*>(tile + sw128_offset(row, e >> 2) + (e & 3) * 4) = v;
}

// in0/in1/out per mode:  FPROP: x, w, y   DGRAD: dy, w, dx   WGRAD: x, dy, dw(grad arena, accumulated)
template <int MODE, int TNW>
__global__ void __launch_bounds__(T_THREADS, TNW == 256 ? 1 : 2)
conv_tc_kernel(const float* __restrict__ in0, const float* __restrict__ in1, float* __restrict__ out, const ConvP p,
               const int vec_b) {
  constexpr int T_STAGES = t_stages(TNW), T_STAGE_BYTES = t_stage_bytes(TNW), TN = TNW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  int2* row_table = reinterpret_cast<int2*>(smem + T_STAGES * T_STAGE_BYTES);
  int2* tap_table = row_table + TM;                  // [ntaps] {kh * W + kw, kh << 16 | kw}: no LDC / branches in the gathers
  int* tap_gidx = reinterpret_cast<int*>(tap_table + 64);   // index of the tap in p.taps[] (weight addressing)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T_STAGES * T_STAGE_BYTES + T_TABLE_BYTES);
  uint64_t* empty_bar = full_bar + T_STAGES;
  uint64_t* acc_bar = empty_bar + T_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int KHW = p.KH * p.KW, HiWi = p.Hi * p.Wi, HoWo = p.Ho * p.Wo;
  // GEMM view of this mode: R rows x C cols, reduction length RED
  // parity class of this CTA (stride-2 dgrad only): its own row space, tap list and reduction length
  int cls = -1, NT = p.ntaps, cls_h0 = 0, cls_w0 = 0, cls_hc = p.Hi, cls_wc = p.Wi, tile_x = blockIdx.x;
  if (MODE == DGRAD && p.ncls == 4) {
    cls = 0;
    while (cls < 3 && tile_x >= p.cls_tile0[cls + 1]) ++cls;
    tile_x -= p.cls_tile0[cls];
    NT = p.cls_ntaps[cls];
    cls_h0 = p.cls_h0[cls >> 1]; cls_hc = p.cls_hc[cls >> 1];
    cls_w0 = p.cls_w0[cls & 1]; cls_wc = p.cls_wc[cls & 1];
  }
  const int pix_step = cls >= 0 ? 2 : 1;                                 // spacing of this CTA's pixels in the image
  const int R = MODE == FPROP ? p.B * HoWo : MODE == DGRAD ? p.B * cls_hc * cls_wc : p.Cin * p.ntaps;
  const int C = MODE == FPROP ? p.Cout : MODE == DGRAD ? p.Cin : p.Cout;
  const int RED = MODE == FPROP ? p.Cin * p.ntaps : MODE == DGRAD ? p.Cout * NT : p.B * HoWo;
  const int r_tile0 = tile_x * TM, c_tile0 = blockIdx.y * TN;
  const int stages_total = (RED + TK - 1) / TK;
  const int per = (stages_total + p.splits - 1) / p.splits;
  const int st_begin = split * per, st_end = min(stages_total, st_begin + per);
  const int nst = max(0, st_end - st_begin);

  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) {
      mbar_init(full_bar + s, T_PRODUCERS);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr, TN);
  if (tid >= TM && tid < TM + NT) {
    // .x = the tap's separable address offset: fprop  x[.., oh*s-pad+kh, ow*s-pad+kw]      -> + kh*Wi + kw
    //                                          dgrad dy[.., (ih+pad-kh)/s, (iw+pad-kw)/s]   -> - (kh/s)*Wo - kw/s
    //      (when s divides ih+pad-kh the quotient is floor((ih+pad)/s) - floor(kh/s));  .y = kh << 16 | kw
    const int gi = cls >= 0 ? p.cls_tap_idx[cls][tid - TM] : tid - TM;
    const int t = p.taps[gi], kh = t >> 8, kw = t & 255;
    const int off = MODE == DGRAD ? -((kh / p.stride) * p.Wo + kw / p.stride) : kh * p.Wi + kw;
    tap_table[tid - TM] = make_int2(off, (kh << 16) | kw);
    tap_gidx[tid - TM] = gi;
  }
  if (MODE == WGRAD && tid < TM) {                   // per-row (ci, tap) decode of this tile, shared by all stages
    const int k = r_tile0 + tid;
    int2 e = make_int2(0, 0x7F7F);                   // kh = kw = 127: always out of bounds -> zero row
    if (k < R) {
      const int ci = k / p.ntaps, t = p.taps[k - ci * p.ntaps];
      e = make_int2(ci * HiWi + (t >> 8) * p.Wi + (t & 255), t);
    }
    row_table[tid] = e;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long in0_slot = static_cast<long long>(slot) * p.B * (MODE == DGRAD ? p.Cout * HoWo : p.Cin * HiWi);
  const float* a_src = in0 + in0_slot;                                           // x (fprop/wgrad) or dy (dgrad)
  const float* b_src = MODE == WGRAD ? in1 + static_cast<long long>(slot) * p.B * p.Cout * HoWo      // dy
                                     : in1 + static_cast<long long>(slot) * p.w_slot_stride;         // weights

  if (tid < T_PRODUCERS) {
    // ------------------------------------------------------------------------------------------ producers
    const int row = tid & (TM - 1), khalf = tid >> 7;
    // row-gather state (fprop / dgrad): this thread's pixel.  Which taps fall inside the image is a property of the
    // pixel alone, so it is decided once (one bit per tap) and the per-element work in the stage loop is: test a bit,
    // add the tap's table offset, load, select.
    unsigned long long tap_ok = 0ull;
    int base = 0;                                      // offset of (b, channel 0, pixel-dependent part)
    if (MODE == FPROP || MODE == DGRAD) {
      const int m = r_tile0 + row;
      if (m < R) {
        const int HW = MODE == FPROP ? HoWo : cls_hc * cls_wc, Wd = MODE == FPROP ? p.Wo : cls_wc;
        const int b = m / HW, r = m - b * HW;
        int ph = r / Wd, pw = r - ph * Wd;
        if (MODE == DGRAD) { ph = cls_h0 + ph * pix_step; pw = cls_w0 + pw * pix_step; }   // pixel of this class
        if (MODE == FPROP) {
          const int h0 = ph * p.stride - p.pad, w0 = pw * p.stride - p.pad;
          base = b * p.Cin * HiWi + h0 * p.Wi + w0;
          for (int t = 0; t < p.ntaps; ++t) {
            const int2 tp = tap_table[t];
            const int ih = h0 + (tp.y >> 16), iw = w0 + (tp.y & 0xFFFF);
            if (static_cast<unsigned>(ih) < static_cast<unsigned>(p.Hi) && static_cast<unsigned>(iw) < static_cast<unsigned>(p.Wi))
              tap_ok |= 1ull << t;
          }
        } else {
          const int h0 = ph + p.pad, w0 = pw + p.pad;
          base = b * p.Cout * HoWo + (h0 / p.stride) * p.Wo + w0 / p.stride;
          for (int t = 0; t < NT; ++t) {
            const int2 tp = tap_table[t];
            const int th = h0 - (tp.y >> 16), tw = w0 - (tp.y & 0xFFFF);          // = oh * stride, ow * stride
            if (th >= 0 && tw >= 0 && th % p.stride == 0 && tw % p.stride == 0 && th / p.stride < p.Ho && tw / p.stride < p.Wo)
              tap_ok |= 1ull << t;
          }
        }
      }
    }
    for (int it = 0; it < nst; ++it) {
      const int s = it % T_STAGES;
      mbar_wait(empty_bar + s, ((it / T_STAGES) & 1) ^ 1);
      uint8_t* sa = smem + s * T_STAGE_BYTES;
      uint8_t* sb = sa + T_A_BYTES;
      const int r0 = (st_begin + it) * TK;
      if (MODE == FPROP || MODE == DGRAD) {
        // ---- A: row gather, 16 reduction indices (channel, tap) of this thread's pixel.  Branch-free: masked
        // elements load element 0 and are zeroed by a select, so all 16 loads are in flight together.
        const int k = r0 + khalf * 16;
        const int c0 = k / NT;                              // ci (fprop) / co (dgrad) of the first index
        const int CS = MODE == FPROP ? HiWi : HoWo;         // channel stride of the gathered tensor
        int ti = k - c0 * NT;
        int cbase = base + c0 * CS;
        const int nvalid = RED - k;                         // reduction indices left (only the last stage is partial)
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool ok = ((tap_ok >> ti) & 1ull) != 0ull && j < nvalid;
          const float xv = __ldg(a_src + (ok ? cbase + tap_table[ti].x : 0));
          v[j] = ok ? xv : 0.f;
          const bool wrap = ++ti == NT;
          ti = wrap ? 0 : ti;
          cbase += wrap ? CS : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(sa + sw128_offset(row, khalf * 4 + q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        // ---- B: weights
        if (MODE == FPROP && vec_b) {
          // unpruned filter with a 16-byte aligned row pitch: the reduction index IS the memory index
          const int chunk = tid & 7;
#pragma unroll
          for (int pass = 0; pass < TN / 32; ++pass) {
            const int brow = (tid >> 3) + pass * 32, n = c_tile0 + brow, kk = r0 + chunk * 4;
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < C && kk < RED) wv = __ldg(reinterpret_cast<const float4*>(b_src + static_cast<long long>(n) * RED + kk));
            *reinterpret_cast<float4*>(sb + sw128_offset(brow, chunk)) = wv;
          }
        } else {
          const int kb = r0 + lane;
          long long woff = -1;
          if (kb < RED) {
            const int cc = kb / NT, tib = kb - cc * NT;
            const int2 tp = tap_table[tib];
            const int wtaps = p.compact ? p.ntaps : KHW;                    // taps stored per (co, ci)
            const int tap_off = p.compact ? tap_gidx[tib] : (tp.y >> 16) * p.KW + (tp.y & 0xFFFF);
            woff = MODE == FPROP ? static_cast<long long>(cc) * wtaps + tap_off
                                 : static_cast<long long>(cc) * p.Cin * wtaps + tap_off;
          }
          const long long row_pitch = MODE == FPROP ? static_cast<long long>(p.Cin) * (p.compact ? p.ntaps : KHW)
                                                    : (p.compact ? p.ntaps : KHW);
#pragma unroll
          for (int i = 0; i < TN / 8; ++i) {
            const int brow = warp * (TN / 8) + i, n = c_tile0 + brow;
            const bool okb = woff >= 0 && n < C;
            const float wv = __ldg(b_src + (okb ? woff + n * row_pitch : 0));
            st_elem(sb, brow, lane, okb ? wv : 0.f);
          }
        }
      } else {
        // ---- wgrad: lane == output pixel m (contiguous in both x and dy), loop over tile rows
        const int m = r0 + lane;
        const bool mv = m < RED;
        int xpart = 0, dypart = 0, ihb = 0, iwb = 0;
        if (mv) {
          const int b = m / HoWo, r = m - b * HoWo, oh = r / p.Wo, ow = r - oh * p.Wo;
          ihb = oh * p.stride - p.pad; iwb = ow * p.stride - p.pad;
          xpart = b * p.Cin * HiWi + ihb * p.Wi + iwb;
          dypart = b * p.Cout * HoWo + r;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int arow = warp * 16 + i;
          const int2 e = row_table[arow];
          const int ih = ihb + (e.y >> 8), iw = iwb + (e.y & 255);
          const bool ok = mv && static_cast<unsigned>(ih) < static_cast<unsigned>(p.Hi) && static_cast<unsigned>(iw) < static_cast<unsigned>(p.Wi);
          const float xv = __ldg(a_src + (ok ? xpart + e.x : 0));
          st_elem(sa, arow, lane, ok ? xv : 0.f);
        }
#pragma unroll
        for (int i = 0; i < TN / 8; ++i) {
          const int brow = warp * (TN / 8) + i, co = c_tile0 + brow;
          const bool okd = mv && co < C;
          const float dv = __ldg(b_src + (okd ? dypart + co * HoWo : 0));
          st_elem(sb, brow, lane, okd ? dv : 0.f);
        }
      }
      fence_proxy_async();                           // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(full_bar + s);
    }
    // ------------------------------------------------------------------------------------------ epilogue
    if (nst > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      const int q = warp & 3, chalf = warp >> 2;     // TMEM lane quadrant of this warp, which half of the columns
      const int rr = r_tile0 + q * 32 + lane;
#pragma unroll 1
      for (int cc0 = 0; cc0 < TN / 2; cc0 += 32) {
      const int col0 = chalf * (TN / 2) + cc0;       // first accumulator column of this chunk
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(col0), v);
      if (rr < R) {
        if (MODE == WGRAD) {
          const int ci = rr / p.ntaps, t = p.taps[rr - ci * p.ntaps];
          float* dst = out + static_cast<long long>(slot) * p.w_slot_stride +
                       (p.compact ? rr : ci * KHW + (t >> 8) * p.KW + (t & 255));
          const long long pitch = p.compact ? static_cast<long long>(R) : static_cast<long long>(p.Cin) * KHW;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = c_tile0 + col0 + j;
            if (co < C) atomicAdd(dst + co * pitch, __uint_as_float(v[j]));
          }
        } else {
          const int HWr = MODE == FPROP ? HoWo : cls_hc * cls_wc;     // rows per image in this CTA's row space
          const int HW = MODE == FPROP ? HoWo : HiWi;                 // pixels per channel plane of the output
          const int b = rr / HWr;
          int r = rr - b * HWr;
          if (MODE == DGRAD && cls >= 0) {
            const int jh = r / cls_wc, jw = r - jh * cls_wc;
            r = (cls_h0 + 2 * jh) * p.Wi + cls_w0 + 2 * jw;
          }
          float* dst = out + (static_cast<long long>(slot) * p.B + b) * C * HW + r;
          const int n_base = c_tile0 + col0;
          dst += static_cast<long long>(n_base) * HW;
          if (p.splits > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_base + j < C) atomicAdd(dst + j * HW, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_base + j < C) dst[j * HW] = __uint_as_float(v[j]);
          }
        }
      }
      }  // column chunks
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_fmt(TM, TN, 2u);
    for (int it = 0; it < nst; ++it) {
      const int s = it % T_STAGES;
      mbar_wait(full_bar + s, (it / T_STAGES) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * T_STAGE_BYTES);
      const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(a_addr + T_A_BYTES);
#pragma unroll
      for (int k = 0; k < TK / T_UMMA_K; ++k)         // 8 tf32 = 32 bytes along K: +2 in the (addr >> 4) field
        umma_tf32(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                  (it | k) != 0 ? 1u : 0u);
      umma_commit(empty_bar + s);
    }
    if (nst > 0) umma_commit(acc_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TN);
  }
}

static int pick_splits_tc(long long tiles, int stages, int ctas_per_sm = 2) {
  // These launches are latency-bound: a CTA costs a fixed prologue/epilogue (barriers, TMEM allocation, tap masks,
  // accumulator read-out) plus ~1.2 us per 32-deep stage, and 2 CTAs are resident per SM (2 x 98 KB smem).  Splitting the
  // reduction shortens the per-CTA chain but a grid of 300 CTAs on 296 slots runs as TWO waves; pick the split count that
  // minimises  waves x (fixed + stages_per_cta x stage)  (+ the zero-fill / atomic combine when splitting at all).
  const double fixed_us = 5.0, stage_us = 1.2, combine_us = 1.5;
  const long long slots = 148LL * ctas_per_sm;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= std::min(32, std::max(1, stages)); ++s) {
    const long long waves = (tiles * s + slots - 1) / slots;
    const int per = (stages + s - 1) / s;
    const double cost = static_cast<double>(waves) * (fixed_us + per * stage_us) + (s > 1 ? combine_us : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

template <int MODE, int TNW>
static void launch_tc_w(const float* in0, const float* in1, float* out, const ConvP& p, int rows, int cols, int vec_b,
                        int row_tiles) {
  static bool configured = false;
  if (!configured) {
    FLUTE_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<MODE, TNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem(TNW)));
    configured = true;
  }
  dim3 grid(row_tiles >= 0 ? row_tiles : (rows + TM - 1) / TM, (cols + TNW - 1) / TNW, p.S * p.splits);
  conv_tc_kernel<MODE, TNW><<<grid, T_THREADS, t_smem(TNW), at::cuda::getCurrentCUDAStream()>>>(in0, in1, out, p, vec_b);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// N-tile width for `cols` output columns: the widest tile that the columns fill (see t_stages)
static int pick_tn(int cols) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = std::getenv("FLUTE_CONV_TN");
    forced = e != nullptr ? std::atoi(e) : 0;
  }
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  return cols >= 256 ? 256 : cols >= 128 ? 128 : 64;
}
static int ctas_per_sm(int tn) { return tn == 256 ? 1 : 2; }

template <int MODE>
static void launch_tc(const float* in0, const float* in1, float* out, const ConvP& p, int rows, int cols, int vec_b,
                      int tn, int row_tiles = -1) {
  if (tn == 256) launch_tc_w<MODE, 256>(in0, in1, out, p, rows, cols, vec_b, row_tiles);
  else if (tn == 128) launch_tc_w<MODE, 128>(in0, in1, out, p, rows, cols, vec_b, row_tiles);
  else launch_tc_w<MODE, 64>(in0, in1, out, p, rows, cols, vec_b, row_tiles);
}

}  // namespace tcv

// 0 = auto (tcgen05 when the GEMM has >= 48 rows / reduction items per slot), 1 = FMA kernels only, 2 = tcgen05 always
static int g_conv_impl = 0;
static int tc_min_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("FLUTE_CONV_TC_MIN_ROWS");
    v = e != nullptr ? std::max(1, std::atoi(e)) : 48;
  }
  return v;
}
static bool use_tc(int work_rows) { return g_conv_impl == 2 || (g_conv_impl == 0 && work_rows >= tc_min_rows()); }

}  // namespace conv

// w_arena: the [S, P] parameter arena; the layer's weight of slot s lives at w_arena + s*P + w_offset
at::Tensor slot_conv_fprop(at::Tensor x, at::Tensor w_arena, int64_t w_offset, int64_t Cout, int64_t KH, int64_t KW,
                           int64_t stride, int64_t pad, bool compact) {
  using namespace conv;
  check5(x, "x");
  check_arena(w_arena, x.size(0));
  ConvP p = make_params(x.size(0), x.size(1), x.size(2), x.size(3), x.size(4), Cout, KH, KW, stride, pad, w_arena.size(1));
  p.compact = compact ? 1 : 0;
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.ntaps;
  if (use_tc(M)) {
    const int tn = tcv::pick_tn(p.Cout);
    const long long tiles = static_cast<long long>((M + tcv::TM - 1) / tcv::TM) * ((p.Cout + tn - 1) / tn) * p.S;
    p.splits = tcv::pick_splits_tc(tiles, (K + tcv::TK - 1) / tcv::TK, tcv::ctas_per_sm(tn));
    auto y = p.splits > 1 ? at::zeros({p.S, p.B, p.Cout, p.Ho, p.Wo}, x.options())
                          : at::empty({p.S, p.B, p.Cout, p.Ho, p.Wo}, x.options());
    const int vec_b = (p.ntaps == p.KH * p.KW || p.compact) && (K % 4 == 0) && ((w_offset % 4) == 0) && (w_arena.size(1) % 4 == 0);
    tcv::launch_tc<tcv::FPROP>(x.data_ptr<float>(), w_arena.data_ptr<float>() + w_offset, y.data_ptr<float>(), p, M, p.Cout, vec_b, tn);
    return y;
  }
  dim3 grid((M + BM - 1) / BM, (p.Cout + BN - 1) / BN, 1);
  p.splits = pick_splits(static_cast<long long>(grid.x) * grid.y * p.S, (K + BK - 1) / BK);
  grid.z = p.S * p.splits;
  auto y = p.splits > 1 ? at::zeros({p.S, p.B, p.Cout, p.Ho, p.Wo}, x.options())
                        : at::empty({p.S, p.B, p.Cout, p.Ho, p.Wo}, x.options());
  conv_fprop_kernel<<<grid, THREADS, 0, at::cuda::getCurrentCUDAStream()>>>(
      x.data_ptr<float>(), w_arena.data_ptr<float>() + w_offset, y.data_ptr<float>(), p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return y;
}

at::Tensor slot_conv_dgrad(at::Tensor dy, at::Tensor w_arena, int64_t w_offset, int64_t Cin, int64_t Hi, int64_t Wi,
                           int64_t KH, int64_t KW, int64_t stride, int64_t pad, bool compact) {
  using namespace conv;
  check5(dy, "dy");
  check_arena(w_arena, dy.size(0));
  ConvP p = make_params(dy.size(0), dy.size(1), Cin, Hi, Wi, dy.size(2), KH, KW, stride, pad, w_arena.size(1));
  p.compact = compact ? 1 : 0;
  TORCH_CHECK(p.Ho == dy.size(3) && p.Wo == dy.size(4), "dy spatial size mismatch");
  const c10::cuda::CUDAGuard guard(dy.device());
  const int M = p.B * p.Hi * p.Wi, K = p.Cout * p.ntaps;
  if (use_tc(M)) {
    const int tn = tcv::pick_tn(p.Cin);
    const int n_tiles = (p.Cin + tn - 1) / tn;
    if (p.stride == 2) {
      // parity decomposition (see ConvP): 4 dense sub-problems instead of one 4x larger masked one
      p.ncls = 4;
      for (int par = 0; par < 2; ++par) {
        const int h0 = ((par - p.pad) % 2 + 2) % 2, w0 = h0;
        p.cls_h0[par] = h0; p.cls_hc[par] = h0 < p.Hi ? (p.Hi - h0 + 1) / 2 : 0;
        p.cls_w0[par] = w0; p.cls_wc[par] = w0 < p.Wi ? (p.Wi - w0 + 1) / 2 : 0;
      }
      int tile0 = 0, max_taps = 0;
      for (int c = 0; c < 4; ++c) {
        const int ph = c >> 1, pw = c & 1;
        int nt = 0;
        for (int t = 0; t < p.ntaps; ++t)
          if (((p.taps[t] >> 8) & 1) == ph && ((p.taps[t] & 255) & 1) == pw && nt < 16) p.cls_tap_idx[c][nt++] = static_cast<unsigned char>(t);
        p.cls_ntaps[c] = static_cast<unsigned char>(nt);
        max_taps = std::max(max_taps, nt);
        p.cls_tile0[c] = tile0;
        const int rows_c = p.B * p.cls_hc[ph] * p.cls_wc[pw];
        if (nt > 0) tile0 += (rows_c + tcv::TM - 1) / tcv::TM;
      }
      p.cls_tile0[4] = tile0;
      auto dx = at::zeros({p.S, p.B, p.Cin, p.Hi, p.Wi}, dy.options());     // classes without taps stay zero
      if (tile0 > 0) {
        p.splits = tcv::pick_splits_tc(static_cast<long long>(tile0) * n_tiles * p.S, (p.Cout * max_taps + tcv::TK - 1) / tcv::TK,
                                       tcv::ctas_per_sm(tn));
        tcv::launch_tc<tcv::DGRAD>(dy.data_ptr<float>(), w_arena.data_ptr<float>() + w_offset, dx.data_ptr<float>(), p, M, p.Cin,
                                   0, tn, tile0);
      }
      return dx;
    }
    const long long tiles = static_cast<long long>((M + tcv::TM - 1) / tcv::TM) * n_tiles * p.S;
    p.splits = tcv::pick_splits_tc(tiles, (K + tcv::TK - 1) / tcv::TK, tcv::ctas_per_sm(tn));
    auto dx = p.splits > 1 ? at::zeros({p.S, p.B, p.Cin, p.Hi, p.Wi}, dy.options())
                           : at::empty({p.S, p.B, p.Cin, p.Hi, p.Wi}, dy.options());
    tcv::launch_tc<tcv::DGRAD>(dy.data_ptr<float>(), w_arena.data_ptr<float>() + w_offset, dx.data_ptr<float>(), p, M, p.Cin, 0, tn);
    return dx;
  }
  dim3 grid((M + BM - 1) / BM, (p.Cin + BN - 1) / BN, 1);
  p.splits = pick_splits(static_cast<long long>(grid.x) * grid.y * p.S, (K + BK - 1) / BK);
  grid.z = p.S * p.splits;
  auto dx = p.splits > 1 ? at::zeros({p.S, p.B, p.Cin, p.Hi, p.Wi}, dy.options())
                         : at::empty({p.S, p.B, p.Cin, p.Hi, p.Wi}, dy.options());
  conv_dgrad_kernel<<<grid, THREADS, 0, at::cuda::getCurrentCUDAStream()>>>(
      dy.data_ptr<float>(), w_arena.data_ptr<float>() + w_offset, dx.data_ptr<float>(), p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return dx;
}

// accumulates into g_arena[s, g_offset : g_offset + Cout*Cin*KH*KW]  (the arena must be zero or hold a partial sum)
void slot_conv_wgrad(at::Tensor x, at::Tensor dy, at::Tensor g_arena, int64_t g_offset, int64_t KH, int64_t KW,
                     int64_t stride, int64_t pad, bool compact) {
  using namespace conv;
  check5(x, "x");
  check5(dy, "dy");
  check_arena(g_arena, x.size(0));
  ConvP p = make_params(x.size(0), x.size(1), x.size(2), x.size(3), x.size(4), dy.size(2), KH, KW, stride, pad,
                        g_arena.size(1));
  p.compact = compact ? 1 : 0;
  TORCH_CHECK(p.Ho == dy.size(3) && p.Wo == dy.size(4), "dy spatial size mismatch");
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.ntaps;
  if (use_tc(M)) {
    const int tn = tcv::pick_tn(p.Cout);
    const long long tiles = static_cast<long long>((K + tcv::TM - 1) / tcv::TM) * ((p.Cout + tn - 1) / tn) * p.S;
    p.splits = tcv::pick_splits_tc(tiles, (M + tcv::TK - 1) / tcv::TK, tcv::ctas_per_sm(tn));
    tcv::launch_tc<tcv::WGRAD>(x.data_ptr<float>(), dy.data_ptr<float>(), g_arena.data_ptr<float>() + g_offset, p, K, p.Cout, 0, tn);
    return;
  }
  dim3 grid((p.Cout + BM - 1) / BM, (K + BN - 1) / BN, 1);
  p.splits = pick_splits(static_cast<long long>(grid.x) * grid.y * p.S, (M + BK - 1) / BK);
  grid.z = p.S * p.splits;
  conv_wgrad_kernel<<<grid, THREADS, 0, at::cuda::getCurrentCUDAStream()>>>(
      x.data_ptr<float>(), dy.data_ptr<float>(), g_arena.data_ptr<float>() + g_offset, p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void slot_conv_set_impl(int64_t impl) {
  TORCH_CHECK(impl >= 0 && impl <= 2, "impl: 0 = auto, 1 = FMA, 2 = tcgen05");
  conv::g_conv_impl = static_cast<int>(impl);
}

}  // namespace flute
