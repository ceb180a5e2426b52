// LayerNorm forward / backward for the transformer tasks (BERT MLM, NRMS): one warp per row of H features, two-pass
// variance in registers/shuffles, affine fused; the backward recomputes x-hat from the saved (mean, rstd), writes dx
// and accumulates d-gamma / d-beta per block in shared memory before ONE atomicAdd per feature per block
// (instead of ATen's separate reduction kernels).  fp32 or bf16 activations, fp32 parameters and statistics.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {
namespace ln {

constexpr int kWarps = 8;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) layer_norm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ b, T* __restrict__ y,
                                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                                     int rows, int H, float eps) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + static_cast<long long>(row) * H;
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s += to_f(xr[i]);
  const float m = warp_sum(s) / H;
  float v = 0.f;
  for (int i = lane; i < H; i += 32) { const float d = to_f(xr[i]) - m; v += d * d; }
  const float r = rsqrtf(warp_sum(v) / H + eps);
  if (lane == 0) { mean[row] = m; rstd[row] = r; }
  T* yr = y + static_cast<long long>(row) * H;
  for (int i = lane; i < H; i += 32) {
    float o = (to_f(xr[i]) - m) * r;
    if (w != nullptr) o = o * w[i] + b[i];
    yr[i] = from_f<T>(o);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma += dy * xhat,  dbeta += dy
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) layer_norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                     const float* __restrict__ w, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, T* __restrict__ dx,
                                                                     float* __restrict__ dw, float* __restrict__ db,
                                                                     int rows, int H, int rows_per_block) {
  extern __shared__ float sm[];                 // [2][H] block-level d-gamma / d-beta
  float* s_dw = sm;
  float* s_db = sm + H;
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * rows_per_block, row1 = min(rows, row0 + rows_per_block);
  for (int row = row0 + warp; row < row1; row += kWarps) {
    const T* xr = x + static_cast<long long>(row) * H;
    const T* gr = dy + static_cast<long long>(row) * H;
    const float m = mean[row], r = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < H; i += 32) {
      const float xh = (to_f(xr[i]) - m) * r, g = to_f(gr[i]) * (w != nullptr ? w[i] : 1.f);
      s1 += g;
      s2 += g * xh;
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
    T* dr = dx + static_cast<long long>(row) * H;
    for (int i = lane; i < H; i += 32) {
      const float xh = (to_f(xr[i]) - m) * r, d = to_f(gr[i]);
      dr[i] = from_f<T>(r * (d * (w != nullptr ? w[i] : 1.f) - s1 - xh * s2));
      if (dw != nullptr) {
        atomicAdd(s_dw + i, d * xh);             // shared-memory atomics: 8 warps of this block share the feature
        atomicAdd(s_db + i, d);
      }
    }
  }
  if (dw != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
      atomicAdd(dw + i, s_dw[i]);
      atomicAdd(db + i, s_db[i]);
    }
  }
}

}  // namespace ln

// x [rows, H] (fp32 or bf16, contiguous) ; weight/bias [H] fp32 or undefined -> {y, mean, rstd}
std::vector<torch::Tensor> layer_norm_fwd(torch::Tensor x, c10::optional<torch::Tensor> weight, c10::optional<torch::Tensor> bias,
                                          double eps) {
  using namespace ln;
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.dim() == 2, "layer_norm_fwd: contiguous [rows, H] CUDA tensor expected");
  const int rows = static_cast<int>(x.size(0)), H = static_cast<int>(x.size(1));
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto y = torch::empty_like(x);
  auto mean = torch::empty({rows}, x.options().dtype(torch::kFloat32));
  auto rstd = torch::empty({rows}, x.options().dtype(torch::kFloat32));
  const float* wp = weight.has_value() ? weight->data_ptr<float>() : nullptr;
  const float* bp = bias.has_value() ? bias->data_ptr<float>() : nullptr;
  TORCH_CHECK((wp == nullptr) == (bp == nullptr), "weight and bias must be given together");
  const int blocks = (rows + kWarps - 1) / kWarps;
  if (rows > 0) {
    if (x.scalar_type() == torch::kFloat32)
      layer_norm_fwd_kernel<float><<<blocks, kWarps * 32, 0, stream>>>(x.data_ptr<float>(), wp, bp, y.data_ptr<float>(),
                                                                       mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, H,
                                                                       static_cast<float>(eps));
    else if (x.scalar_type() == torch::kBFloat16)
      layer_norm_fwd_kernel<__nv_bfloat16><<<blocks, kWarps * 32, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), wp, bp, reinterpret_cast<__nv_bfloat16*>(y.data_ptr()),
          mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, H, static_cast<float>(eps));
    else
      TORCH_CHECK(false, "layer_norm_fwd: fp32 or bf16 only");
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {y, mean, rstd};
}

// -> {dx, dweight, dbias}  (dweight/dbias fp32, zeros when no affine)
std::vector<torch::Tensor> layer_norm_bwd(torch::Tensor dy, torch::Tensor x, c10::optional<torch::Tensor> weight,
                                          torch::Tensor mean, torch::Tensor rstd) {
  using namespace ln;
  TORCH_CHECK(dy.is_cuda() && dy.is_contiguous() && x.is_contiguous() && dy.sizes() == x.sizes() && dy.scalar_type() == x.scalar_type());
  const int rows = static_cast<int>(x.size(0)), H = static_cast<int>(x.size(1));
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto dx = torch::empty_like(x);
  const bool affine = weight.has_value();
  auto dw = torch::zeros({affine ? H : 0}, x.options().dtype(torch::kFloat32));
  auto db = torch::zeros({affine ? H : 0}, x.options().dtype(torch::kFloat32));
  const float* wp = affine ? weight->data_ptr<float>() : nullptr;
  // enough blocks to fill the GPU, enough rows per block to amortise the [2][H] shared reduction
  const int rows_per_block = std::max(kWarps, (rows + 148 * 4 - 1) / (148 * 4));
  const int blocks = (rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = 2 * static_cast<size_t>(H) * sizeof(float);
  TORCH_CHECK(smem <= 48 * 1024, "layer_norm_bwd: hidden size too large for the shared reduction");
  if (rows > 0) {
    if (x.scalar_type() == torch::kFloat32)
      layer_norm_bwd_kernel<float><<<blocks, kWarps * 32, smem, stream>>>(
          dy.data_ptr<float>(), x.data_ptr<float>(), wp, mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr<float>(),
          affine ? dw.data_ptr<float>() : nullptr, affine ? db.data_ptr<float>() : nullptr, rows, H, rows_per_block);
    else if (x.scalar_type() == torch::kBFloat16)
      layer_norm_bwd_kernel<__nv_bfloat16><<<blocks, kWarps * 32, smem, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(dy.data_ptr()), reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), wp,
          mean.data_ptr<float>(), rstd.data_ptr<float>(), reinterpret_cast<__nv_bfloat16*>(dx.data_ptr()),
          affine ? dw.data_ptr<float>() : nullptr, affine ? db.data_ptr<float>() : nullptr, rows, H, rows_per_block);
    else
      TORCH_CHECK(false, "layer_norm_bwd: fp32 or bf16 only");
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {dx, dw, db};
}

}  // namespace flute
