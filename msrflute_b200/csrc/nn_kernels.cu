// Small layer kernels that used to be ATen / cuDNN calls (K3 BatchNorm, K6 embedding, K8 dropout).
//
//   embedding_fwd / embedding_bwd : row gather (16-byte vectors) / scatter-add of gradient rows (fp32 RED), padding row
//                                   skipped — nlp_rnn_fedshakespeare (90 x 8), nlg_gru (10000 x 160), BERT (30522 x 768)
//   dropout_fwd / dropout_bwd     : Philox4x32-10 mask recomputed in the backward from (seed, element index): no mask
//                                   tensor; the seed is read from DEVICE memory so a captured CUDA graph draws a fresh
//                                   mask on every replay (the caller bumps the counter inside the graph)
//   batch_norm_fwd / batch_norm_bwd: training-mode BatchNorm2d on NCHW fp32, one CTA per channel, fused
//                                   affine (+ residual) (+ ReLU) epilogue and running-statistics update — what the
//                                   reference's RESNET actually instantiates (experiments/cv_resnet_fedcifar100/model.py:116)
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <vector>
#include "common.cuh"

namespace flute {
namespace nnk {

constexpr int kT = 256;

// ------------------------------------------------------------------------------------------------ embedding
__global__ void __launch_bounds__(kT) embedding_fwd_kernel(const long long* __restrict__ idx, const float* __restrict__ w,
                                                           float* __restrict__ out, int64_t n_tok, int D, int64_t V) {
  const int64_t total = n_tok * D;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * kT) {
    const int64_t t = i / D;
    const int d = static_cast<int>(i - t * D);
    const long long row = idx[t];
    out[i] = (row >= 0 && row < V) ? __ldg(w + row * D + d) : 0.f;
  }
}
__global__ void __launch_bounds__(kT) embedding_bwd_kernel(const long long* __restrict__ idx, const float* __restrict__ dy,
                                                           float* __restrict__ dw, int64_t n_tok, int D, int64_t V,
                                                           long long padding_idx) {
  const int64_t total = n_tok * D;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * kT) {
    const int64_t t = i / D;
    const int d = static_cast<int>(i - t * D);
    const long long row = idx[t];
    if (row >= 0 && row < V && row != padding_idx) atomicAdd(dw + row * D + d, dy[i]);
  }
}

// ------------------------------------------------------------------------------------------------ dropout
// keep = u > p ; y = keep ? x / (1 - p) : 0        u = uniform(0, 1] of Philox(seed, i / 4)[i % 4]
template <bool kBwd>
__global__ void __launch_bounds__(kT) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p,
                                                     const long long* __restrict__ seed_ptr) {
  const uint64_t seed = static_cast<uint64_t>(*seed_ptr);
  const float scale = 1.f / (1.f - p);
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const int64_t nq = (n + 3) >> 2;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; q < nq; q += static_cast<int64_t>(gridDim.x) * kT) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), 0x44524F50u, 0u), key);
    const float u[4] = {u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w)};
    const int64_t i0 = 4 * q;
    if (i0 + 3 < n && (reinterpret_cast<uintptr_t>(x + i0) & 15) == 0 && (reinterpret_cast<uintptr_t>(y + i0) & 15) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(x + i0);
      *reinterpret_cast<float4*>(y + i0) = make_float4(u[0] > p ? v.x * scale : 0.f, u[1] > p ? v.y * scale : 0.f,
                                                       u[2] > p ? v.z * scale : 0.f, u[3] > p ? v.w * scale : 0.f);
    } else {
      for (int e = 0; e < 4 && i0 + e < n; ++e) y[i0 + e] = u[e] > p ? x[i0 + e] * scale : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm2d (training)
// x [N, C, H, W]: block c reduces its channel over N*H*W, then normalises.  stats[c] = (mean, rstd) saved for backward.
__global__ void __launch_bounds__(kT) batch_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ res,
                                                            float* __restrict__ y, float* __restrict__ stats,
                                                            float* __restrict__ run_mean, float* __restrict__ run_var, int N,
                                                            int C, int HW, float eps, float momentum, int relu) {
  const int c = blockIdx.x;
  const int64_t cnt = static_cast<int64_t>(N) * HW;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = threadIdx.x; i < cnt; i += kT) {
    const int64_t n = i / HW, r = i - n * HW;
    const float v = x[(n * C + c) * HW + r];
    s1 += v; s2 += v * v;
  }
  const float2 tot = block_sum2(s1, s2);
  const float mean = tot.x / static_cast<float>(cnt);
  const float var = fmaxf(tot.y / static_cast<float>(cnt) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    stats[2 * c] = mean; stats[2 * c + 1] = rstd;
    if (run_mean != nullptr) {
      const float unbiased = cnt > 1 ? var * static_cast<float>(cnt) / static_cast<float>(cnt - 1) : var;
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * unbiased;
    }
  }
  const float g = gamma != nullptr ? gamma[c] : 1.f, b = beta != nullptr ? beta[c] : 0.f;
  const float sc = rstd * g, sh = b - mean * rstd * g;
  for (int64_t i = threadIdx.x; i < cnt; i += kT) {
    const int64_t n = i / HW, r = i - n * HW, o = (n * C + c) * HW + r;
    float v = fmaf(x[o], sc, sh);
    if (res != nullptr) v += res[o];
    y[o] = relu ? fmaxf(v, 0.f) : v;
  }
}

// dx, dgamma, dbeta (+ dres = masked dy).  y (post-activation output) is needed only for the ReLU mask.
__global__ void __launch_bounds__(kT) batch_norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ y, const float* __restrict__ gamma,
                                                            const float* __restrict__ stats, float* __restrict__ dx,
                                                            float* __restrict__ dres, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int N, int C, int HW, int relu) {
  const int c = blockIdx.x;
  const int64_t cnt = static_cast<int64_t>(N) * HW;
  const float mean = stats[2 * c], rstd = stats[2 * c + 1];
  float sg = 0.f, sb = 0.f;
  for (int64_t i = threadIdx.x; i < cnt; i += kT) {
    const int64_t n = i / HW, r = i - n * HW, o = (n * C + c) * HW + r;
    float d = dy[o];
    if (relu && !(y[o] > 0.f)) d = 0.f;
    sg += d * (x[o] - mean) * rstd;
    sb += d;
  }
  const float2 tot = block_sum2(sg, sb);
  if (threadIdx.x == 0) { dgamma[c] = tot.x; dbeta[c] = tot.y; }
  const float g = gamma != nullptr ? gamma[c] : 1.f;
  const float inv = 1.f / static_cast<float>(cnt);
  for (int64_t i = threadIdx.x; i < cnt; i += kT) {
    const int64_t n = i / HW, r = i - n * HW, o = (n * C + c) * HW + r;
    float d = dy[o];
    if (relu && !(y[o] > 0.f)) d = 0.f;
    if (dres != nullptr) dres[o] = d;
    const float xh = (x[o] - mean) * rstd;
    dx[o] = g * rstd * (d - tot.y * inv - xh * tot.x * inv);
  }
}

static int grid_for(int64_t n) { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + kT - 1) / kT, 148 * 8))); }

}  // namespace nnk

torch::Tensor embedding_fwd(torch::Tensor idx, torch::Tensor weight) {
  using namespace nnk;
  TORCH_CHECK(idx.is_cuda() && idx.scalar_type() == torch::kInt64 && weight.is_cuda() && weight.scalar_type() == torch::kFloat32 &&
              weight.dim() == 2 && weight.is_contiguous());
  auto ic = idx.contiguous();
  const int64_t n = ic.numel(), D = weight.size(1);
  auto sizes = ic.sizes().vec();
  sizes.push_back(D);
  auto out = torch::empty(sizes, weight.options());
  if (n == 0) return out;
  const c10::cuda::CUDAGuard guard(weight.device());
  embedding_fwd_kernel<<<grid_for(n * D), kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<const long long*>(ic.data_ptr<int64_t>()), weight.data_ptr<float>(), out.data_ptr<float>(), n,
      static_cast<int>(D), weight.size(0));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

torch::Tensor embedding_bwd(torch::Tensor idx, torch::Tensor dy, int64_t V, int64_t padding_idx) {
  using namespace nnk;
  auto ic = idx.contiguous();
  auto dyc = dy.contiguous();
  const int64_t n = ic.numel(), D = dyc.size(-1);
  auto dw = torch::zeros({V, D}, dyc.options());
  if (n == 0) return dw;
  const c10::cuda::CUDAGuard guard(dy.device());
  embedding_bwd_kernel<<<grid_for(n * D), kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<const long long*>(ic.data_ptr<int64_t>()), dyc.data_ptr<float>(), dw.data_ptr<float>(), n,
      static_cast<int>(D), V, static_cast<long long>(padding_idx));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return dw;
}

// seed: 1-element int64 CUDA tensor (read on the device)
torch::Tensor dropout_apply(torch::Tensor x, double p, torch::Tensor seed, bool backward) {
  using namespace nnk;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kFloat32 && seed.is_cuda() && seed.scalar_type() == torch::kInt64 &&
              seed.numel() == 1 && p >= 0.0 && p < 1.0);
  auto xc = x.contiguous();
  auto y = torch::empty_like(xc);
  const int64_t n = xc.numel();
  if (n == 0) return y;
  const c10::cuda::CUDAGuard guard(x.device());
  const int grid = grid_for((n + 3) / 4);
  const auto* sp = reinterpret_cast<const long long*>(seed.data_ptr<int64_t>());
  if (backward)
    dropout_kernel<true><<<grid, kT, 0, at::cuda::getCurrentCUDAStream()>>>(xc.data_ptr<float>(), y.data_ptr<float>(), n,
                                                                           static_cast<float>(p), sp);
  else
    dropout_kernel<false><<<grid, kT, 0, at::cuda::getCurrentCUDAStream()>>>(xc.data_ptr<float>(), y.data_ptr<float>(), n,
                                                                            static_cast<float>(p), sp);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return y;
}

std::vector<torch::Tensor> batch_norm_fwd(torch::Tensor x, c10::optional<torch::Tensor> gamma, c10::optional<torch::Tensor> beta,
                                          c10::optional<torch::Tensor> residual, c10::optional<torch::Tensor> run_mean,
                                          c10::optional<torch::Tensor> run_var, double momentum, double eps, bool relu) {
  using namespace nnk;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kFloat32 && x.dim() == 4, "batch_norm_fwd: fp32 CUDA [N, C, H, W]");
  auto xc = x.contiguous();
  const int N = static_cast<int>(xc.size(0)), C = static_cast<int>(xc.size(1)), HW = static_cast<int>(xc.size(2) * xc.size(3));
  auto y = torch::empty_like(xc);
  auto stats = torch::empty({C, 2}, xc.options());
  torch::Tensor rc;
  if (residual.has_value()) rc = residual->contiguous();
  const c10::cuda::CUDAGuard guard(x.device());
  batch_norm_fwd_kernel<<<C, kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      xc.data_ptr<float>(), gamma.has_value() ? gamma->data_ptr<float>() : nullptr,
      beta.has_value() ? beta->data_ptr<float>() : nullptr, residual.has_value() ? rc.data_ptr<float>() : nullptr,
      y.data_ptr<float>(), stats.data_ptr<float>(), run_mean.has_value() ? run_mean->data_ptr<float>() : nullptr,
      run_var.has_value() ? run_var->data_ptr<float>() : nullptr, N, C, HW, static_cast<float>(eps),
      static_cast<float>(momentum), relu ? 1 : 0);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {y, stats};
}

std::vector<torch::Tensor> batch_norm_bwd(torch::Tensor dy, torch::Tensor x, torch::Tensor y, c10::optional<torch::Tensor> gamma,
                                          torch::Tensor stats, bool relu, bool want_dres) {
  using namespace nnk;
  auto dyc = dy.contiguous();
  auto xc = x.contiguous();
  const int N = static_cast<int>(xc.size(0)), C = static_cast<int>(xc.size(1)), HW = static_cast<int>(xc.size(2) * xc.size(3));
  auto dx = torch::empty_like(xc);
  auto dres = want_dres ? torch::empty_like(xc) : torch::Tensor();
  auto dgamma = torch::empty({C}, xc.options());
  auto dbeta = torch::empty({C}, xc.options());
  const c10::cuda::CUDAGuard guard(x.device());
  batch_norm_bwd_kernel<<<C, kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      dyc.data_ptr<float>(), xc.data_ptr<float>(), y.data_ptr<float>(), gamma.has_value() ? gamma->data_ptr<float>() : nullptr,
      stats.data_ptr<float>(), dx.data_ptr<float>(), want_dres ? dres.data_ptr<float>() : nullptr, dgamma.data_ptr<float>(),
      dbeta.data_ptr<float>(), N, C, HW, relu ? 1 : 0);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  if (want_dres) return {dx, dgamma, dbeta, dres};
  return {dx, dgamma, dbeta};
}

}  // namespace flute
