// Small bandwidth-bound kernels of the client / server paths that are not part of the arena step (SURVEY K7, K14, K15,
// K19).  Each replaces a chain of ATen launches with one pass over the data.
//
//   seg_minmax          per-tensor (segment) min / max of a flat gradient arena            (K14, statistics half)
//   quantize_segments   snap to 2^bits uniform levels in [lo, hi] + |g| <= thresh -> 0      (K14, encode half)
//   local_dp            g <- g * min(1, C/||g||) (or C/||g||) + sigma * N(0,1), Philox       (K15)
//   softmax_ce          per-row loss and d(logits) of softmax cross-entropy in one kernel    (K7)
//   cosine_stats        <a,b>, ||a||^2, ||b||^2 in one pass                                  (K19)
//   max_pool2d          NCHW max-pool fwd (+1-byte argmax) / bwd                             (K8)
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/types.h>
#include <cfloat>
#include "common.cuh"

namespace flute {
namespace misc {

constexpr int kThreads = 256;

// order-preserving float <-> uint mapping so atomicMin/atomicMax on unsigned work for floats
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// first segment whose end is > i  (segments are sorted, non-overlapping; gaps = alignment padding)
__device__ __forceinline__ int find_segment(const long long* __restrict__ seg, int nseg, long long i) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (seg[2 * mid] + seg[2 * mid + 1] > i) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------ K14 statistics
// ord[2*s] = min, ord[2*s+1] = max in the ordered-uint domain; one block works on a contiguous chunk, reduces per
// segment in registers while the chunk stays inside one segment and flushes with two atomics when it changes.
__global__ void __launch_bounds__(kThreads) seg_minmax_kernel(const float* __restrict__ g, long long n,
                                                              const long long* __restrict__ seg, int nseg,
                                                              unsigned* __restrict__ ord, long long chunk) {
  const long long begin = blockIdx.x * chunk, end = min(n, begin + chunk);
  __shared__ float smin[kThreads / 32], smax[kThreads / 32];
  long long pos = begin;
  while (pos < end) {
    const int s = find_segment(seg, nseg, pos);
    const long long so = seg[2 * s], se = so + seg[2 * s + 1];
    if (pos >= se) break;                                    // past the last tensor (tail padding)
    if (pos < so) { pos = so; continue; }                    // padding between tensors
    if (so >= end) break;
    const long long stop = min(end, se);
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (long long i = pos + threadIdx.x; i < stop; i += kThreads) {
      const float v = g[i];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { smin[w] = mn; smax[w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < kThreads / 32; ++k) { mn = fminf(mn, smin[k]); mx = fmaxf(mx, smax[k]); }
      atomicMin(ord + 2 * s, f2ord(mn));
      atomicMax(ord + 2 * s + 1, f2ord(mx));
    }
    pos = stop;
  }
}

__global__ void ord_to_float_kernel(const unsigned* __restrict__ ord, float* __restrict__ out, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    out[3 * s] = ord2f(ord[2 * s]);
    out[3 * s + 1] = ord2f(ord[2 * s + 1]);
  }
}

// ------------------------------------------------------------------------------------------------ K14 encode
// stats [nseg, 3] = (lo, hi, thresh).  level = clamp(ceil((g - lo)/w - 1/2), 0, L-1); value = lo + level*w; then
// |g| <= thresh -> 0.  Optionally also emits the level codes (uint8 when bits <= 8, else int32) and keep bits.
template <typename CodeT>
__global__ void __launch_bounds__(kThreads) quantize_kernel(float* __restrict__ g, long long n,
                                                            const long long* __restrict__ seg, int nseg,
                                                            const float* __restrict__ stats, int levels,
                                                            CodeT* __restrict__ codes, unsigned char* __restrict__ keep) {
  const long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n) return;
  const int s = find_segment(seg, nseg, i);
  const long long so = seg[2 * s];
  if (i < so || i >= so + seg[2 * s + 1]) return;          // alignment padding: untouched
  const float lo = stats[3 * s], hi = stats[3 * s + 1], th = stats[3 * s + 2];
  const float v = g[i];
  const float w = (hi - lo) / static_cast<float>(levels - 1);
  float q = lo;
  float idx = 0.f;
  if (w > 0.f) {
    idx = fminf(fmaxf(ceilf((v - lo) / w - 0.5f), 0.f), static_cast<float>(levels - 1));
    q = lo + idx * w;
  }
  const bool kept = fabsf(v) > th;
  g[i] = kept ? q : 0.f;
  if (codes != nullptr) codes[i] = static_cast<CodeT>(idx);
  if (keep != nullptr) keep[i] = kept ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ K15
// stats2 = (sum, sumsq) of g from a previous reduction.  mode 0: clip (scale = min(1, C/norm)), mode 1: normalise
// (scale = C/norm, the reference's Gaussian mechanism).  Noise keyed by element index: layout / GPU-count invariant.
__global__ void __launch_bounds__(kThreads) local_dp_kernel(float* __restrict__ g, long long n,
                                                            const float* __restrict__ sumsq, float max_grad, float sigma,
                                                            int mode, unsigned long long seed) {
  const long long q = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;   // one float4 per thread
  const long long i = q * 4;
  if (i >= n) return;
  const float norm = sqrtf(fmaxf(*sumsq, 0.f));
  float scale = norm > 0.f ? max_grad / norm : 1.f;
  if (mode == 0) scale = fminf(scale, 1.f);
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sigma != 0.f) z = philox_normal4(seed, static_cast<uint64_t>(q));
  const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i + k < n) g[i + k] = g[i + k] * scale + sigma * zz[k];
}

__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float v = g[i];
    acc += v * v;
  }
  const float2 r = block_sum2(acc, 0.f);
  if (threadIdx.x == 0) atomicAdd(out, r.x);
}

// ------------------------------------------------------------------------------------------------ K7
// One warp per row.  loss[r] = logsumexp(x) - x[target];  dx = (softmax(x) - onehot) * grad_scale  (dx may alias x).
__global__ void __launch_bounds__(kThreads) softmax_ce_kernel(const float* __restrict__ x, const long long* __restrict__ tgt,
                                                              float* __restrict__ loss, float* __restrict__ dx,
                                                              int rows, int C, float grad_scale, long long ignore_index) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * C;
  float m = -FLT_MAX;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(xr[c] - m);
  s = warp_sum(s);
  const long long t = tgt[row];
  const bool ignored = t == ignore_index || t < 0 || t >= C;
  if (lane == 0) loss[row] = ignored ? 0.f : (m + __logf(s)) - xr[t];
  if (dx != nullptr) {
    float* dr = dx + static_cast<long long>(row) * C;
    const float inv = ignored ? 0.f : grad_scale / s;
    for (int c = lane; c < C; c += 32) {
      float p = __expf(xr[c] - m) * inv;
      if (!ignored && c == t) p -= grad_scale;
      dr[c] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K19
__global__ void __launch_bounds__(kThreads) cosine_stats_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                long long n, float* __restrict__ out) {
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float x = a[i], y = b[i];
    dot += x * y;
    na += x * x;
    nb += y * y;
  }
  const float2 r1 = block_sum2(dot, na);
  const float2 r2 = block_sum2(nb, 0.f);
  if (threadIdx.x == 0) {
    atomicAdd(out, r1.x);
    atomicAdd(out + 1, r1.y);
    atomicAdd(out + 2, r2.x);
  }
}


// ------------------------------------------------------------------------------------------------ K8 max-pool
// NCHW fp32 max-pool with the window argmax kept as one byte per output; the backward routes dy to that input element
// (windows overlap when stride < kernel, hence atomics on a zeroed dx).  First maximum wins ties, like ATen.
__global__ void __launch_bounds__(kThreads) max_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                unsigned char* __restrict__ arg, long long total, int H, int W,
                                                                int Ho, int Wo, int k, int stride, int pad) {
  const long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= total) return;
  const int ow = static_cast<int>(i % Wo), oh = static_cast<int>((i / Wo) % Ho);
  const long long plane = i / (static_cast<long long>(Wo) * Ho);
  const float* xp = x + plane * H * W;
  float best = -FLT_MAX;
  int bi = 0;
  for (int kh = 0; kh < k; ++kh) {
    const int ih = oh * stride - pad + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int iw = ow * stride - pad + kw;
      if (iw < 0 || iw >= W) continue;
      const float v = xp[ih * W + iw];
      if (v > best) { best = v; bi = kh * k + kw; }
    }
  }
  y[i] = best;
  arg[i] = static_cast<unsigned char>(bi);
}

__global__ void __launch_bounds__(kThreads) max_pool_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ arg,
                                                                float* __restrict__ dx, long long total, int H, int W, int Ho,
                                                                int Wo, int k, int stride, int pad) {
  const long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= total) return;
  const int ow = static_cast<int>(i % Wo), oh = static_cast<int>((i / Wo) % Ho);
  const long long plane = i / (static_cast<long long>(Wo) * Ho);
  const int a = arg[i], ih = oh * stride - pad + a / k, iw = ow * stride - pad + a % k;
  atomicAdd(dx + plane * H * W + ih * W + iw, dy[i]);
}

static void check_flat(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), name, " must be a contiguous fp32 CUDA tensor");
}

}  // namespace misc

// flat [n] fp32; seg [nseg, 2] int64 (offset, size) sorted  ->  stats [nseg, 3] with columns 0/1 = min/max (2 left 0)
at::Tensor seg_minmax(at::Tensor flat, at::Tensor seg) {
  using namespace misc;
  check_flat(flat, "flat");
  TORCH_CHECK(seg.is_cuda() && seg.scalar_type() == at::kLong && seg.dim() == 2 && seg.size(1) == 2 && seg.is_contiguous());
  const c10::cuda::CUDAGuard guard(flat.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int nseg = static_cast<int>(seg.size(0));
  auto ord = at::empty({nseg, 2}, flat.options().dtype(at::kInt));
  ord.select(1, 0).fill_(-1);               // 0xFFFFFFFF = +inf in the ordered domain
  ord.select(1, 1).zero_();                 // 0 = -inf
  auto stats = at::zeros({nseg, 3}, flat.options());
  const long long n = flat.numel();
  const long long chunk = 1 << 16;
  const int blocks = static_cast<int>((n + chunk - 1) / chunk);
  if (blocks > 0)
    seg_minmax_kernel<<<blocks, kThreads, 0, stream>>>(flat.data_ptr<float>(), n,
                                                      reinterpret_cast<const long long*>(seg.data_ptr<int64_t>()), nseg,
                                                      reinterpret_cast<unsigned*>(ord.data_ptr<int>()), chunk);
  ord_to_float_kernel<<<(nseg + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const unsigned*>(ord.data_ptr<int>()),
                                                               stats.data_ptr<float>(), nseg);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return stats;
}

// in-place simulated quantization of `flat`; optionally returns (codes, keep) for the packed wire format
std::vector<at::Tensor> quantize_segments(at::Tensor flat, at::Tensor seg, at::Tensor stats, int64_t bits, bool emit_codes) {
  using namespace misc;
  check_flat(flat, "flat");
  check_flat(stats, "stats");
  TORCH_CHECK(seg.is_cuda() && seg.scalar_type() == at::kLong && seg.is_contiguous() && stats.size(0) == seg.size(0));
  TORCH_CHECK(bits >= 1 && bits <= 16, "quant_bits must be in [1, 16]");
  const c10::cuda::CUDAGuard guard(flat.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const long long n = flat.numel();
  const int nseg = static_cast<int>(seg.size(0)), levels = 1 << bits;
  const int blocks = static_cast<int>((n + kThreads - 1) / kThreads);
  std::vector<at::Tensor> out;
  const long long* segp = reinterpret_cast<const long long*>(seg.data_ptr<int64_t>());
  if (emit_codes) {
    auto keep = at::zeros({n}, flat.options().dtype(at::kByte));
    if (bits <= 8) {
      auto codes = at::zeros({n}, flat.options().dtype(at::kByte));
      if (blocks > 0)
        quantize_kernel<unsigned char><<<blocks, kThreads, 0, stream>>>(flat.data_ptr<float>(), n, segp, nseg, stats.data_ptr<float>(),
                                                                       levels, codes.data_ptr<uint8_t>(), keep.data_ptr<uint8_t>());
      out = {codes, keep};
    } else {
      auto codes = at::zeros({n}, flat.options().dtype(at::kInt));
      if (blocks > 0)
        quantize_kernel<int><<<blocks, kThreads, 0, stream>>>(flat.data_ptr<float>(), n, segp, nseg, stats.data_ptr<float>(), levels,
                                                             codes.data_ptr<int>(), keep.data_ptr<uint8_t>());
      out = {codes, keep};
    }
  } else if (blocks > 0) {
    quantize_kernel<unsigned char><<<blocks, kThreads, 0, stream>>>(flat.data_ptr<float>(), n, segp, nseg, stats.data_ptr<float>(),
                                                                   levels, nullptr, nullptr);
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

// g <- g * scale(||g||, max_grad) + sigma * N(0,1);  returns ||g|| before scaling (device scalar)
at::Tensor local_dp(at::Tensor flat, double max_grad, double sigma, bool clip_only, int64_t seed) {
  using namespace misc;
  check_flat(flat, "flat");
  const c10::cuda::CUDAGuard guard(flat.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const long long n = flat.numel();
  auto sumsq = at::zeros({1}, flat.options());
  if (n == 0) return sumsq;
  const int rblocks = static_cast<int>(std::min<long long>((n + kThreads - 1) / kThreads, 148 * 8));
  sumsq_kernel<<<rblocks, kThreads, 0, stream>>>(flat.data_ptr<float>(), n, sumsq.data_ptr<float>());
  const long long quads = (n + 3) / 4;
  local_dp_kernel<<<static_cast<int>((quads + kThreads - 1) / kThreads), kThreads, 0, stream>>>(
      flat.data_ptr<float>(), n, sumsq.data_ptr<float>(), static_cast<float>(max_grad), static_cast<float>(sigma),
      clip_only ? 0 : 1, static_cast<unsigned long long>(seed));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return sumsq.sqrt();
}

// logits [rows, C] fp32, target [rows] int64 -> (loss [rows], dlogits [rows, C] * grad_scale)
std::vector<at::Tensor> softmax_ce(at::Tensor logits, at::Tensor target, double grad_scale, int64_t ignore_index, bool want_grad) {
  using namespace misc;
  check_flat(logits, "logits");
  TORCH_CHECK(logits.dim() == 2 && target.is_cuda() && target.scalar_type() == at::kLong && target.is_contiguous() &&
              target.numel() == logits.size(0), "softmax_ce: logits [rows, C] fp32, target [rows] int64");
  const c10::cuda::CUDAGuard guard(logits.device());
  const int rows = static_cast<int>(logits.size(0)), C = static_cast<int>(logits.size(1));
  auto loss = at::empty({rows}, logits.options());
  at::Tensor dx = want_grad ? at::empty_like(logits) : at::Tensor();
  if (rows > 0)
    softmax_ce_kernel<<<(rows + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
        logits.data_ptr<float>(), reinterpret_cast<const long long*>(target.data_ptr<int64_t>()), loss.data_ptr<float>(),
        want_grad ? dx.data_ptr<float>() : nullptr, rows, C, static_cast<float>(grad_scale), ignore_index);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  if (want_grad) return {loss, dx};
  return {loss};
}

// -> [3] = (<a,b>, ||a||^2, ||b||^2)
at::Tensor cosine_stats(at::Tensor a, at::Tensor b) {
  using namespace misc;
  check_flat(a, "a");
  check_flat(b, "b");
  TORCH_CHECK(a.numel() == b.numel(), "cosine_stats: size mismatch");
  const c10::cuda::CUDAGuard guard(a.device());
  auto out = at::zeros({3}, a.options());
  const long long n = a.numel();
  if (n > 0) {
    const int blocks = static_cast<int>(std::min<long long>((n + kThreads - 1) / kThreads, 148 * 8));
    cosine_stats_kernel<<<blocks, kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(a.data_ptr<float>(), b.data_ptr<float>(), n,
                                                                                    out.data_ptr<float>());
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

// x [..., H, W] fp32 contiguous -> {y [..., Ho, Wo], argmax (uint8, window index)}
std::vector<at::Tensor> max_pool2d_fwd(at::Tensor x, int64_t k, int64_t stride, int64_t pad) {
  using namespace misc;
  check_flat(x, "x");
  TORCH_CHECK(x.dim() >= 3 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad <= k / 2, "max_pool2d_fwd: bad arguments");
  const int H = static_cast<int>(x.size(-2)), W = static_cast<int>(x.size(-1));
  const int Ho = (H + 2 * static_cast<int>(pad) - static_cast<int>(k)) / static_cast<int>(stride) + 1;
  const int Wo = (W + 2 * static_cast<int>(pad) - static_cast<int>(k)) / static_cast<int>(stride) + 1;
  auto sizes = x.sizes().vec();
  sizes[sizes.size() - 2] = Ho;
  sizes[sizes.size() - 1] = Wo;
  const c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty(sizes, x.options());
  auto arg = at::empty(sizes, x.options().dtype(at::kByte));
  const long long total = y.numel();
  if (total > 0)
    max_pool_fwd_kernel<<<static_cast<int>((total + kThreads - 1) / kThreads), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
        x.data_ptr<float>(), y.data_ptr<float>(), arg.data_ptr<uint8_t>(), total, H, W, Ho, Wo, static_cast<int>(k),
        static_cast<int>(stride), static_cast<int>(pad));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {y, arg};
}

at::Tensor max_pool2d_bwd(at::Tensor dy, at::Tensor arg, int64_t H, int64_t W, int64_t k, int64_t stride, int64_t pad) {
  using namespace misc;
  check_flat(dy, "dy");
  TORCH_CHECK(arg.is_cuda() && arg.scalar_type() == at::kByte && arg.is_contiguous() && arg.numel() == dy.numel());
  auto sizes = dy.sizes().vec();
  const int Ho = static_cast<int>(dy.size(-2)), Wo = static_cast<int>(dy.size(-1));
  sizes[sizes.size() - 2] = H;
  sizes[sizes.size() - 1] = W;
  const c10::cuda::CUDAGuard guard(dy.device());
  auto dx = at::zeros(sizes, dy.options());
  const long long total = dy.numel();
  if (total > 0)
    max_pool_bwd_kernel<<<static_cast<int>((total + kThreads - 1) / kThreads), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
        dy.data_ptr<float>(), arg.data_ptr<uint8_t>(), dx.data_ptr<float>(), total, static_cast<int>(H), static_cast<int>(W), Ho, Wo,
        static_cast<int>(k), static_cast<int>(stride), static_cast<int>(pad));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return dx;
}

}  // namespace flute
