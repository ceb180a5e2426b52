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
// output pixels per slot) and the deep layers are weight-bandwidth bound (a 512x512x3x3 filter is 9.4 MB per slot and
// is used for 20 output pixels).  Formulation: implicit GEMM with on-the-fly im2col gathering,
//     fprop : Y[M, Cout]      = A(x)[M, K]        * W[Cout, K]^T          K = Cin*KH*KW
//     dgrad : dX[Mi, Cin]     = A'(dY)[Mi, K']    * W'[Cin, K']^T         K' = Cout*KH*KW
//     wgrad : dW[Cout, K]    += dY^T[Cout, M]     * A(x)[M, K]
// 64x64 output tiles, 16-deep K chunks staged in shared memory, 256 threads x (4x4) register tiles of fp32 FMAs.
// When the tile grid would not fill the 148 SMs the reduction dimension is split across blockIdx.z and partial tiles
// are combined with fp32 atomics (outputs are pre-zeroed by the caller / the gradient arena is zero between steps).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/types.h>
#include "common.cuh"

namespace flute {
namespace conv {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct ConvP {
  int S, B, Cin, Hi, Wi, Cout, Ho, Wo, KH, KW, stride, pad;
  int splits;                 // reduction splits (blockIdx.z = slot * splits + split)
  long long w_slot_stride;    // floats between two slots' copies of this weight tensor (= arena row length P)
};

// 4x4 register tile FMA over one staged K chunk.  As/Bs are [BK][64] (k-major) so a warp reads consecutive columns.
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
                                                            float* __restrict__ y, ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Ho * p.Wo, N = p.Cout, KHW = p.KH * p.KW, K = p.Cin * KHW;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kchunks = (K + BK - 1) / BK;
  const int per = (kchunks + p.splits - 1) / p.splits;
  const int kc_begin = split * per, kc_end = min(kchunks, kc_begin + per);
  const float* xs = x + static_cast<long long>(slot) * p.B * p.Cin * p.Hi * p.Wi;
  const float* ws = w + static_cast<long long>(slot) * p.w_slot_stride;
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  float acc[4][4] = {};
  // each thread stages 4 A elements (same k, 4 consecutive m... we use m fastest for coalesced x reads) and 4 B elements
  const int a_m = tid & 63, a_k = tid >> 6;          // A: 64 m x 4 k per pass, 4 passes
  const int b_k = tid & 15, b_n = tid >> 4;          // B: 16 k x 16 n per pass, 4 passes (k fastest: W rows are K-contiguous)
  int mb = -1, moh = 0, mow = 0;
  {
    const int m = m0 + a_m;
    if (m < M) { mb = m / (p.Ho * p.Wo); const int r = m - mb * p.Ho * p.Wo; moh = r / p.Wo; mow = r - moh * p.Wo; }
  }
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const int k0 = kc * BK;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int kk = a_k + pass * 4, k = k0 + kk;
      float v = 0.f;
      if (mb >= 0 && k < K) {
        const int ci = k / KHW, t = k - ci * KHW, kh = t / p.KW, kw = t - kh * p.KW;
        const int ih = moh * p.stride - p.pad + kh, iw = mow * p.stride - p.pad + kw;
        if (ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi)
          v = __ldg(xs + ((static_cast<long long>(mb) * p.Cin + ci) * p.Hi + ih) * p.Wi + iw);
      }
      As[kk][a_m] = v;
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int nn = b_n + pass * 16, n = n0 + nn, k = k0 + b_k;
      Bs[b_k][nn] = (n < N && k < K) ? __ldg(ws + static_cast<long long>(n) * K + k) : 0.f;
    }
    __syncthreads();
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
                                                            float* __restrict__ dx, ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Hi * p.Wi, N = p.Cin, KHW = p.KH * p.KW, K = p.Cout * KHW;     // reduction over (co, tap)
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
  int mb = -1, mih = 0, miw = 0;
  {
    const int m = m0 + a_m;
    if (m < M) { mb = m / (p.Hi * p.Wi); const int r = m - mb * p.Hi * p.Wi; mih = r / p.Wi; miw = r - mih * p.Wi; }
  }
  const int CinKHW = p.Cin * KHW;
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const int k0 = kc * BK;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int kk = a_k + pass * 4, k = k0 + kk;
      float v = 0.f;
      if (mb >= 0 && k < K) {
        const int co = k / KHW, t = k - co * KHW, kh = t / p.KW, kw = t - kh * p.KW;
        const int th = mih + p.pad - kh, tw = miw + p.pad - kw;          // = oh*stride, ow*stride
        if (th >= 0 && tw >= 0 && th % p.stride == 0 && tw % p.stride == 0) {
          const int oh = th / p.stride, ow = tw / p.stride;
          if (oh < p.Ho && ow < p.Wo)
            v = __ldg(dys + ((static_cast<long long>(mb) * p.Cout + co) * p.Ho + oh) * p.Wo + ow);
        }
      }
      As[kk][a_m] = v;
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int nn = b_n + pass * 16, ci = n0 + nn, k = k0 + b_k;
      float v = 0.f;
      if (ci < N && k < K) {
        const int co = k / KHW, t = k - co * KHW;
        v = __ldg(ws + static_cast<long long>(co) * CinKHW + ci * KHW + t);
      }
      Bs[b_k][nn] = v;
    }
    __syncthreads();
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
// dw[slot][co, k] += sum_m dy[m, co] * A(x)[m, k]     (tile rows = co, tile cols = k, reduction over m)
__global__ void __launch_bounds__(THREADS) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw, ConvP p) {
  __shared__ __align__(16) float As[BK][BM + 4];     // dy^T chunk: [m-chunk][co]
  __shared__ __align__(16) float Bs[BK][BN + 4];     // im2col chunk: [m-chunk][k]
  const int slot = blockIdx.z / p.splits, split = blockIdx.z - slot * p.splits;
  const int M = p.B * p.Ho * p.Wo, KHW = p.KH * p.KW, K = p.Cin * KHW, HoWo = p.Ho * p.Wo;
  const int co0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
  const int mchunks = (M + BK - 1) / BK;
  const int per = (mchunks + p.splits - 1) / p.splits;
  const int mc_begin = split * per, mc_end = min(mchunks, mc_begin + per);
  const float* xs = x + static_cast<long long>(slot) * p.B * p.Cin * p.Hi * p.Wi;
  const float* dys = dy + static_cast<long long>(slot) * p.B * p.Cout * HoWo;
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  float acc[4][4] = {};
  const int l_m = tid & 15, l_c = tid >> 4;          // 16 m (fastest: contiguous in dy and roughly in x) x 16 cols per pass
  // decode this thread's 4 k columns once
  int kci[4], kkh[4], kkw[4];
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int k = k0 + l_c + pass * 16;
    if (k < K) { kci[pass] = k / KHW; const int t = k - kci[pass] * KHW; kkh[pass] = t / p.KW; kkw[pass] = t - kkh[pass] * p.KW; }
    else kci[pass] = -1;
  }
  for (int mc = mc_begin; mc < mc_end; ++mc) {
    const int m = mc * BK + l_m;
    int b = 0, oh = 0, ow = 0;
    const bool mv = m < M;
    if (mv) { b = m / HoWo; const int r = m - b * HoWo; oh = r / p.Wo; ow = r - oh * p.Wo; }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int cc = l_c + pass * 16, co = co0 + cc;
      As[l_m][cc] = (mv && co < p.Cout) ? __ldg(dys + (static_cast<long long>(b) * p.Cout + co) * HoWo + oh * p.Wo + ow) : 0.f;
      float v = 0.f;
      if (mv && kci[pass] >= 0) {
        const int ih = oh * p.stride - p.pad + kkh[pass], iw = ow * p.stride - p.pad + kkw[pass];
        if (ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi)
          v = __ldg(xs + ((static_cast<long long>(b) * p.Cin + kci[pass]) * p.Hi + ih) * p.Wi + iw);
      }
      Bs[l_m][cc] = v;
    }
    __syncthreads();
    tile_fma(As, Bs, acc, tm, tn);
    __syncthreads();
  }
  float* dws = dw + static_cast<long long>(slot) * p.w_slot_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tm * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tn * 4 + j;
      if (k >= K) continue;
      atomicAdd(dws + static_cast<long long>(co) * K + k, acc[i][j]);      // gradient arena is zero between steps
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

static ConvP make_params(const at::Tensor& x, int64_t Cout, int64_t KH, int64_t KW, int64_t stride, int64_t pad,
                         int64_t w_slot_stride) {
  TORCH_CHECK(x.dim() == 5 && x.is_cuda() && x.scalar_type() == at::kFloat && x.is_contiguous(),
              "slot conv: x must be contiguous fp32 CUDA [S, B, C, H, W]");
  ConvP p;
  p.S = static_cast<int>(x.size(0)); p.B = static_cast<int>(x.size(1)); p.Cin = static_cast<int>(x.size(2));
  p.Hi = static_cast<int>(x.size(3)); p.Wi = static_cast<int>(x.size(4));
  p.Cout = static_cast<int>(Cout); p.KH = static_cast<int>(KH); p.KW = static_cast<int>(KW);
  p.stride = static_cast<int>(stride); p.pad = static_cast<int>(pad);
  p.Ho = (p.Hi + 2 * p.pad - p.KH) / p.stride + 1;
  p.Wo = (p.Wi + 2 * p.pad - p.KW) / p.stride + 1;
  p.w_slot_stride = w_slot_stride;
  p.splits = 1;
  return p;
}

}  // namespace conv

// w_base: the [S, P] parameter arena; the layer's weight of slot s lives at w_base + s*P + w_offset
at::Tensor slot_conv_fprop(at::Tensor x, at::Tensor w_arena, int64_t w_offset, int64_t Cout, int64_t KH, int64_t KW,
                           int64_t stride, int64_t pad) {
  using namespace conv;
  TORCH_CHECK(w_arena.dim() == 2 && w_arena.is_cuda() && w_arena.scalar_type() == at::kFloat && w_arena.is_contiguous());
  ConvP p = make_params(x, Cout, KH, KW, stride, pad, w_arena.size(1));
  TORCH_CHECK(w_arena.size(0) >= p.S, "arena has fewer rows than slots");
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.KH * p.KW;
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
  TORCH_CHECK(dy.dim() == 5 && dy.is_cuda() && dy.scalar_type() == at::kFloat && dy.is_contiguous());
  ConvP p;
  p.S = static_cast<int>(dy.size(0)); p.B = static_cast<int>(dy.size(1)); p.Cout = static_cast<int>(dy.size(2));
  p.Ho = static_cast<int>(dy.size(3)); p.Wo = static_cast<int>(dy.size(4));
  p.Cin = static_cast<int>(Cin); p.Hi = static_cast<int>(Hi); p.Wi = static_cast<int>(Wi);
  p.KH = static_cast<int>(KH); p.KW = static_cast<int>(KW); p.stride = static_cast<int>(stride); p.pad = static_cast<int>(pad);
  p.w_slot_stride = w_arena.size(1);
  const c10::cuda::CUDAGuard guard(dy.device());
  const int M = p.B * p.Hi * p.Wi, K = p.Cout * p.KH * p.KW;
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
  TORCH_CHECK(dy.dim() == 5 && dy.is_cuda() && dy.scalar_type() == at::kFloat && dy.is_contiguous());
  TORCH_CHECK(g_arena.dim() == 2 && g_arena.is_cuda() && g_arena.scalar_type() == at::kFloat && g_arena.is_contiguous());
  ConvP p = make_params(x, dy.size(2), KH, KW, stride, pad, g_arena.size(1));
  TORCH_CHECK(p.Ho == dy.size(3) && p.Wo == dy.size(4), "dy spatial size mismatch");
  const c10::cuda::CUDAGuard guard(x.device());
  const int M = p.B * p.Ho * p.Wo, K = p.Cin * p.KH * p.KW;
  dim3 grid((p.Cout + BM - 1) / BM, (K + BN - 1) / BN, 1);
  p.splits = pick_splits(static_cast<long long>(grid.x) * grid.y * p.S, (M + BK - 1) / BK);
  grid.z = p.S * p.splits;
  conv_wgrad_kernel<<<grid, THREADS, 0, at::cuda::getCurrentCUDAStream()>>>(
      x.data_ptr<float>(), dy.data_ptr<float>(), g_arena.data_ptr<float>() + g_offset, p);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flute
