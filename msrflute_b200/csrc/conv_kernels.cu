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
#include "common.cuh"

namespace flute {
namespace conv {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256, MAX_TAPS = 49;

struct ConvP {
  int S, B, Cin, Hi, Wi, Cout, Ho, Wo, KH, KW, stride, pad;
  int splits;                 // reduction splits (blockIdx.z = slot * splits + split)
  int ntaps;                  // taps that touch real data for at least one output position
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
  const int M = p.B * p.Ho * p.Wo, N = p.Cout, KHW = p.KH * p.KW, K = p.Cin * p.ntaps, Kfull = p.Cin * KHW;
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
      woff = ci * KHW + (t >> 8) * p.KW + (t & 255);
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
  const int CinKHW = p.Cin * KHW;
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
      const int co = kb / p.ntaps, t = p.taps[kb - co * p.ntaps];
      woff = static_cast<long long>(co) * CinKHW + (t >> 8) * p.KW + (t & 255);
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int ci = n0 + b_n + pass * 16;
      rb[pass] = (woff >= 0 && ci < N) ? __ldg(ws + woff + static_cast<long long>(ci) * KHW) : 0.f;
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
  const int M = p.B * p.Ho * p.Wo, KHW = p.KH * p.KW, K = p.Cin * p.ntaps, Kfull = p.Cin * KHW, HoWo = p.Ho * p.Wo;
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
    const int woff = ci * KHW + (t >> 8) * p.KW + (t & 255);
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

}  // namespace conv

// w_arena: the [S, P] parameter arena; the layer's weight of slot s lives at w_arena + s*P + w_offset
at::Tensor slot_conv_fprop(at::Tensor x, at::Tensor w_arena, int64_t w_offset, int64_t Cout, int64_t KH, int64_t KW,
                           int64_t stride, int64_t pad) {
  using namespace conv;
  check5(x, "x");
  check_arena(w_arena, x.size(0));
  ConvP p = make_params(x.size(0), x.size(1), x.size(2), x.size(3), x.size(4), Cout, KH, KW, stride, pad, w_arena.size(1));
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.ntaps;
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
                           int64_t KH, int64_t KW, int64_t stride, int64_t pad) {
  using namespace conv;
  check5(dy, "dy");
  check_arena(w_arena, dy.size(0));
  ConvP p = make_params(dy.size(0), dy.size(1), Cin, Hi, Wi, dy.size(2), KH, KW, stride, pad, w_arena.size(1));
  TORCH_CHECK(p.Ho == dy.size(3) && p.Wo == dy.size(4), "dy spatial size mismatch");
  const c10::cuda::CUDAGuard guard(dy.device());
  const int M = p.B * p.Hi * p.Wi, K = p.Cout * p.ntaps;
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
                     int64_t stride, int64_t pad) {
  using namespace conv;
  check5(x, "x");
  check5(dy, "dy");
  check_arena(g_arena, x.size(0));
  ConvP p = make_params(x.size(0), x.size(1), x.size(2), x.size(3), x.size(4), dy.size(2), KH, KW, stride, pad,
                        g_arena.size(1));
  TORCH_CHECK(p.Ho == dy.size(3) && p.Wo == dy.size(4), "dy spatial size mismatch");
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.ntaps;
  dim3 grid((p.Cout + BM - 1) / BM, (K + BN - 1) / BN, 1);
  p.splits = pick_splits(static_cast<long long>(grid.x) * grid.y * p.S, (M + BK - 1) / BK);
  grid.z = p.S * p.splits;
  conv_wgrad_kernel<<<grid, THREADS, 0, at::cuda::getCurrentCUDAStream()>>>(
      x.data_ptr<float>(), dy.data_ptr<float>(), g_arena.data_ptr<float>() + g_offset, p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flute
