// Flat-arena kernels for the client side of a federated round.
//
//   fused_client_step   : global-norm clip + gradient sufficient statistics + SGD(momentum, wd) + zero-grad
//                         over [S rows x P] in two launches (reduce, then update) — replaces
//                         clip_grad_norm_ + per-parameter grad.clone().cpu().numpy() + optimizer.step()
//                         of the reference (core/trainer.py:383-391), for S simulated clients at once.
//   clip_and_stats      : the first half only (non-SGD client optimizers).
//   pseudo_grad         : out = (w_global - w_local) * weight  (+ sum / sum^2 of the raw pseudo-gradient).
//   accumulate_pseudo_grad : acc += sum_s weight[s] * (w_global - w_local[s])  — the worker-side half of the
//                         gather: weighting (fedavg.py:80) and aggregation (strategies/utils.py:21-33) in one pass.
//
// All buffers are fp32, rows are 128-byte aligned (parallel/arena.py) so every access is a 16-byte vector.
// These kernels are HBM-bandwidth bound by construction (1 FMA per 4-16 bytes); the only design goals are
// 16-byte coalesced streaming accesses, enough bytes in flight, and one pass over the data.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;  // float4 per thread per iteration -> 64 B in flight per thread

static inline int blocks_for(int64_t n_vec4) {
  // persistent-ish sizing: at most 2 CTAs per SM (148 SMs), at least 1
  int64_t want = (n_vec4 + kThreads * kUnroll - 1) / (kThreads * kUnroll);
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 2)));
}

// ---------------------------------------------------------------------------------------------------- reduce
// partial[s][b] = (sum g, sum g^2) over block b's share of row s
// FedProx: the proximal term (mu/2)||w - w_global||^2 is part of the client loss in the reference
// (core/trainer.py:463-467), i.e. its gradient pmult[j] * (w - w_ref) joins the back-propagated gradient BEFORE clipping
// and the gradient statistics.  Closed form, folded into both passes (pmult carries mu times the reference's per-tensor
// multiplicity, see Trainer.fedprox_multiplicity); no autograd graph of P norms.
__device__ __forceinline__ float4 prox_grad(float4 g, const float4 w, const float4 r, const float4 m) {
  g.x = fmaf(m.x, w.x - r.x, g.x); g.y = fmaf(m.y, w.y - r.y, g.y);
  g.z = fmaf(m.z, w.z - r.z, g.z); g.w = fmaf(m.w, w.w - r.w, g.w);
  return g;
}

__global__ void __launch_bounds__(kThreads) row_reduce_kernel(const float* __restrict__ g, int64_t P,
                                                              float2* __restrict__ partial,
                                                              const float* __restrict__ w = nullptr,
                                                              const float* __restrict__ wref = nullptr,
                                                              const float* __restrict__ pmult = nullptr,
                                                              float* __restrict__ prox_loss = nullptr) {
  const int s = blockIdx.y;
  if (pmult != nullptr) {
    // FedProx variant (one extra read of w and of the two shared rows)
    const float4* gv = reinterpret_cast<const float4*>(g + static_cast<int64_t>(s) * P);
    const float4* wv = reinterpret_cast<const float4*>(w + static_cast<int64_t>(s) * P);
    const float4* rv = reinterpret_cast<const float4*>(wref);
    const float4* mv = reinterpret_cast<const float4*>(pmult);
    const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
      const float4 ww = ld_na(wv + i), rr = __ldg(rv + i), mm = __ldg(mv + i);
      const float4 v = prox_grad(ld_na(gv + i), ww, rr, mm);
      a += (v.x + v.y) + (v.z + v.w);
      b += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      const float dx = ww.x - rr.x, dy = ww.y - rr.y, dz = ww.z - rr.z, dw = ww.w - rr.w;
      c += 0.5f * ((mm.x * dx * dx + mm.y * dy * dy) + (mm.z * dz * dz + mm.w * dw * dw));   // the term's loss value
    }
    const float2 r = block_sum2(a, b);
    if (threadIdx.x == 0) partial[static_cast<int64_t>(s) * gridDim.x + blockIdx.x] = r;
    if (prox_loss != nullptr) {
      const float2 r2 = block_sum2(c, 0.f);
      if (threadIdx.x == 0) atomicAdd(prox_loss + s, r2.x);
    }
    return;
  }
  const float4* gv = reinterpret_cast<const float4*>(g + static_cast<int64_t>(s) * P);
  const int64_t n4 = P >> 2;
  float a = 0.f, b = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = ld_na(gv + i + u * stride);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      a += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      b += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
    }
  }
  for (; i < n4; i += stride) {
    const float4 v = ld_na(gv + i);
    a += (v.x + v.y) + (v.z + v.w);
    b += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  const float2 r = block_sum2(a, b);
  if (threadIdx.x == 0) partial[static_cast<int64_t>(s) * gridDim.x + blockIdx.x] = r;
}

__device__ __forceinline__ float2 fold_partials(const float2* __restrict__ partial, int nb) {
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const float2 p = partial[i];
    a += p.x;
    b += p.y;
  }
  return block_sum2(a, b);
}

// ---------------------------------------------------------------------------------------------------- update
// hyper row: lr, max_norm, wd, momentum.  stats row: sum, sumsq, count, last_norm.
template <bool kMomentum, bool kStepOnly>
__global__ void __launch_bounds__(kThreads)
row_update_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ mom, int64_t P,
                  const float2* __restrict__ partial, int nb_reduce, const float* __restrict__ hyper,
                  float* __restrict__ stats, const int* __restrict__ first_step, float n_logical, bool nesterov,
                  float dampening, bool zero_grad, const float* __restrict__ wref = nullptr,
                  const float* __restrict__ pmult = nullptr) {
  const int s = blockIdx.y;
  const float2 tot = fold_partials(partial + static_cast<int64_t>(s) * nb_reduce, nb_reduce);
  const float lr = hyper[s * 4 + 0], max_norm = hyper[s * 4 + 1], wd = hyper[s * 4 + 2], mu = hyper[s * 4 + 3];
  const float norm = sqrtf(tot.y);
  const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[s * 4 + 0] += coef * tot.x;
    stats[s * 4 + 1] += coef * coef * tot.y;
    stats[s * 4 + 2] += n_logical;
    stats[s * 4 + 3] = norm;
  }
  const bool first = kMomentum && first_step != nullptr && first_step[s] != 0;
  float4* wv = reinterpret_cast<float4*>(w + static_cast<int64_t>(s) * P);
  float4* gv = reinterpret_cast<float4*>(g + static_cast<int64_t>(s) * P);
  float4* mv = kMomentum ? reinterpret_cast<float4*>(mom + static_cast<int64_t>(s) * P) : nullptr;
  const int64_t n4 = P >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    float4 gg = ld_na(gv + i);
    if (kStepOnly) {  // clip_and_stats: scale the gradient in place, nothing else
      gg.x *= coef; gg.y *= coef; gg.z *= coef; gg.w *= coef;
      st_stream(gv + i, gg);
      continue;
    }
    float4 ww = ld_na(wv + i);
    if (pmult != nullptr)
      gg = prox_grad(gg, ww, __ldg(reinterpret_cast<const float4*>(wref) + i), __ldg(reinterpret_cast<const float4*>(pmult) + i));
    float4 d = make_float4(fmaf(coef, gg.x, wd * ww.x), fmaf(coef, gg.y, wd * ww.y), fmaf(coef, gg.z, wd * ww.z),
                           fmaf(coef, gg.w, wd * ww.w));
    if (kMomentum) {
      float4 m = ld_na(mv + i);
      if (first) {
        m = d;
      } else {
        const float k = 1.f - dampening;
        m = make_float4(fmaf(mu, m.x, k * d.x), fmaf(mu, m.y, k * d.y), fmaf(mu, m.z, k * d.z), fmaf(mu, m.w, k * d.w));
      }
      st_stream(mv + i, m);
      d = nesterov ? make_float4(fmaf(mu, m.x, d.x), fmaf(mu, m.y, d.y), fmaf(mu, m.z, d.z), fmaf(mu, m.w, d.w)) : m;
    }
    ww.x = fmaf(-lr, d.x, ww.x); ww.y = fmaf(-lr, d.y, ww.y); ww.z = fmaf(-lr, d.z, ww.z); ww.w = fmaf(-lr, d.w, ww.w);
    st_stream(wv + i, ww);
    if (zero_grad) st_stream(gv + i, zero);
  }
}

// AdamW (HF / reference semantics, msrflute_b200/utils/optimizers AdamW == /root/reference/utils/optimizers/adamW.py):
//   g <- clip(g);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;  w <- w - lr * bias_corr * m / (sqrt(v) + eps);
//   w <- w * (1 - lr * wd)   (decoupled, after the update).   `step` is the 1-based step count of the row (device int32).
__global__ void __launch_bounds__(kThreads)
row_adamw_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t P,
                 const float2* __restrict__ partial, int nb_reduce, const float* __restrict__ hyper,
                 float* __restrict__ stats, const int* __restrict__ step, float n_logical, float b1, float b2, float eps,
                 bool correct_bias, bool zero_grad) {
  const int s = blockIdx.y;
  const float2 tot = fold_partials(partial + static_cast<int64_t>(s) * nb_reduce, nb_reduce);
  const float lr = hyper[s * 4 + 0], max_norm = hyper[s * 4 + 1], wd = hyper[s * 4 + 2];
  const float norm = sqrtf(tot.y);
  const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[s * 4 + 0] += coef * tot.x;
    stats[s * 4 + 1] += coef * coef * tot.y;
    stats[s * 4 + 2] += n_logical;
    stats[s * 4 + 3] = norm;
  }
  const float t = static_cast<float>(step[s]);
  const float step_size = correct_bias ? lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t)) : lr;
  const float decay = wd > 0.f ? 1.f - lr * wd : 1.f;
  float4* wv = reinterpret_cast<float4*>(w + static_cast<int64_t>(s) * P);
  float4* gv = reinterpret_cast<float4*>(g + static_cast<int64_t>(s) * P);
  float4* mv = reinterpret_cast<float4*>(m + static_cast<int64_t>(s) * P);
  float4* vv = reinterpret_cast<float4*>(v + static_cast<int64_t>(s) * P);
  const int64_t n4 = P >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    float4 gg = ld_na(gv + i), ww = ld_na(wv + i), mm = ld_na(mv + i), v2 = ld_na(vv + i);
    gg.x *= coef; gg.y *= coef; gg.z *= coef; gg.w *= coef;
    mm.x = fmaf(b1, mm.x, c1 * gg.x); mm.y = fmaf(b1, mm.y, c1 * gg.y); mm.z = fmaf(b1, mm.z, c1 * gg.z); mm.w = fmaf(b1, mm.w, c1 * gg.w);
    v2.x = fmaf(b2, v2.x, c2 * gg.x * gg.x); v2.y = fmaf(b2, v2.y, c2 * gg.y * gg.y);
    v2.z = fmaf(b2, v2.z, c2 * gg.z * gg.z); v2.w = fmaf(b2, v2.w, c2 * gg.w * gg.w);
    ww.x = (ww.x - step_size * mm.x / (sqrtf(v2.x) + eps)) * decay;
    ww.y = (ww.y - step_size * mm.y / (sqrtf(v2.y) + eps)) * decay;
    ww.z = (ww.z - step_size * mm.z / (sqrtf(v2.z) + eps)) * decay;
    ww.w = (ww.w - step_size * mm.w / (sqrtf(v2.w) + eps)) * decay;
    st_stream(mv + i, mm);
    st_stream(vv + i, v2);
    st_stream(wv + i, ww);
    if (zero_grad) st_stream(gv + i, zero);
  }
}

static void check_rows(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.dim() == 2 && t.is_contiguous(), name,
              " must be a contiguous fp32 CUDA [S, P] tensor");
  TORCH_CHECK(t.size(1) % 4 == 0, name, ": row length must be a multiple of 4 floats (arena padding)");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, name, " must be 16-byte aligned");
}

void fused_client_step(torch::Tensor w, torch::Tensor g, torch::Tensor hyper, torch::Tensor stats,
                       c10::optional<torch::Tensor> mom, c10::optional<torch::Tensor> first_step, int64_t n_logical,
                       bool nesterov, double dampening, bool zero_grad, c10::optional<torch::Tensor> prox_ref,
                       c10::optional<torch::Tensor> prox_mult, c10::optional<torch::Tensor> prox_loss) {
  check_rows(w, "w");
  check_rows(g, "g");
  TORCH_CHECK(w.sizes() == g.sizes(), "w/g shape mismatch");
  const int S = static_cast<int>(w.size(0));
  const int64_t P = w.size(1);
  TORCH_CHECK(hyper.is_cuda() && hyper.numel() == S * 4 && stats.is_cuda() && stats.numel() == S * 4);
  const c10::cuda::CUDAGuard guard(w.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int nb = blocks_for(P >> 2);
  auto partial = torch::empty({S, nb, 2}, w.options());
  const float* pref = nullptr;
  const float* pmul = nullptr;
  if (prox_ref.has_value() && prox_mult.has_value()) {
    TORCH_CHECK(prox_ref->numel() == P && prox_mult->numel() == P && prox_ref->scalar_type() == torch::kFloat32 &&
                prox_mult->scalar_type() == torch::kFloat32 && prox_ref->is_cuda() && prox_mult->is_cuda(),
                "FedProx: reference weights and per-element multipliers must be fp32 CUDA rows of length P");
    pref = prox_ref->data_ptr<float>();
    pmul = prox_mult->data_ptr<float>();
  }
  row_reduce_kernel<<<dim3(nb, S), kThreads, 0, stream>>>(g.data_ptr<float>(), P,
                                                         reinterpret_cast<float2*>(partial.data_ptr<float>()),
                                                         w.data_ptr<float>(), pref, pmul,
                                                         prox_loss.has_value() ? prox_loss->data_ptr<float>() : nullptr);
  const int* fs = first_step.has_value() ? first_step->data_ptr<int>() : nullptr;
  if (mom.has_value()) {
    check_rows(*mom, "mom");
    row_update_kernel<true, false><<<dim3(nb, S), kThreads, 0, stream>>>(
        w.data_ptr<float>(), g.data_ptr<float>(), mom->data_ptr<float>(), P,
        reinterpret_cast<const float2*>(partial.data_ptr<float>()), nb, hyper.data_ptr<float>(), stats.data_ptr<float>(),
        fs, static_cast<float>(n_logical), nesterov, static_cast<float>(dampening), zero_grad, pref, pmul);
  } else {
    row_update_kernel<false, false><<<dim3(nb, S), kThreads, 0, stream>>>(
        w.data_ptr<float>(), g.data_ptr<float>(), nullptr, P, reinterpret_cast<const float2*>(partial.data_ptr<float>()),
        nb, hyper.data_ptr<float>(), stats.data_ptr<float>(), nullptr, static_cast<float>(n_logical), nesterov,
        static_cast<float>(dampening), zero_grad, pref, pmul);
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void fused_client_adamw(torch::Tensor w, torch::Tensor g, torch::Tensor m, torch::Tensor v, torch::Tensor step,
                        torch::Tensor hyper, torch::Tensor stats, int64_t n_logical, double beta1, double beta2, double eps,
                        bool correct_bias, bool zero_grad) {
  check_rows(w, "w"); check_rows(g, "g"); check_rows(m, "m"); check_rows(v, "v");
  TORCH_CHECK(w.sizes() == g.sizes() && w.sizes() == m.sizes() && w.sizes() == v.sizes(), "w/g/m/v shape mismatch");
  const int S = static_cast<int>(w.size(0));
  const int64_t P = w.size(1);
  TORCH_CHECK(hyper.is_cuda() && hyper.numel() == S * 4 && stats.is_cuda() && stats.numel() == S * 4 && step.is_cuda() &&
              step.numel() == S && step.scalar_type() == torch::kInt32);
  const c10::cuda::CUDAGuard guard(w.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int nb = blocks_for(P >> 2);
  auto partial = torch::empty({S, nb, 2}, w.options());
  row_reduce_kernel<<<dim3(nb, S), kThreads, 0, stream>>>(g.data_ptr<float>(), P,
                                                         reinterpret_cast<float2*>(partial.data_ptr<float>()));
  row_adamw_kernel<<<dim3(nb, S), kThreads, 0, stream>>>(
      w.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), P,
      reinterpret_cast<const float2*>(partial.data_ptr<float>()), nb, hyper.data_ptr<float>(), stats.data_ptr<float>(),
      step.data_ptr<int>(), static_cast<float>(n_logical), static_cast<float>(beta1), static_cast<float>(beta2),
      static_cast<float>(eps), correct_bias, zero_grad);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void clip_and_stats(torch::Tensor g, torch::Tensor hyper, torch::Tensor stats, int64_t n_logical) {
  check_rows(g, "g");
  const int S = static_cast<int>(g.size(0));
  const int64_t P = g.size(1);
  const c10::cuda::CUDAGuard guard(g.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int nb = blocks_for(P >> 2);
  auto partial = torch::empty({S, nb, 2}, g.options());
  row_reduce_kernel<<<dim3(nb, S), kThreads, 0, stream>>>(g.data_ptr<float>(), P,
                                                         reinterpret_cast<float2*>(partial.data_ptr<float>()));
  row_update_kernel<false, true><<<dim3(nb, S), kThreads, 0, stream>>>(
      nullptr, g.data_ptr<float>(), nullptr, P, reinterpret_cast<const float2*>(partial.data_ptr<float>()), nb,
      hyper.data_ptr<float>(), stats.data_ptr<float>(), nullptr, static_cast<float>(n_logical), false, 0.f, false);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ pseudo-grad
__global__ void __launch_bounds__(kThreads)
pseudo_grad_kernel(const float* __restrict__ wg, const float* __restrict__ wl, float* __restrict__ out, int64_t P,
                   const float* __restrict__ weight, float* __restrict__ stats) {
  const float wt = weight != nullptr ? *weight : 1.f;
  const float4* a = reinterpret_cast<const float4*>(wg);
  const float4* b = reinterpret_cast<const float4*>(wl);
  float4* o = reinterpret_cast<float4*>(out);
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 x = ld_stream(a + i), y = ld_na(b + i);
    const float4 d = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    s1 += (d.x + d.y) + (d.z + d.w);
    s2 += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
    st_stream(o + i, make_float4(d.x * wt, d.y * wt, d.z * wt, d.w * wt));
  }
  if (stats != nullptr) {
    const float2 r = block_sum2(s1, s2);
    if (threadIdx.x == 0) {
      atomicAdd(stats + 0, r.x);
      atomicAdd(stats + 1, r.y);
    }
  }
}

void pseudo_grad(torch::Tensor wg, torch::Tensor wl, torch::Tensor out, c10::optional<torch::Tensor> weight,
                 c10::optional<torch::Tensor> stats) {
  TORCH_CHECK(wg.is_cuda() && wg.scalar_type() == torch::kFloat32 && wg.numel() == wl.numel() && wg.numel() == out.numel());
  TORCH_CHECK(wg.numel() % 4 == 0, "arena length must be a multiple of 4");
  const c10::cuda::CUDAGuard guard(wg.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  if (stats.has_value()) stats->narrow(0, 0, 2).zero_();
  const int64_t P = wg.numel();
  pseudo_grad_kernel<<<blocks_for(P >> 2), kThreads, 0, stream>>>(
      wg.data_ptr<float>(), wl.data_ptr<float>(), out.data_ptr<float>(), P,
      weight.has_value() ? weight->data_ptr<float>() : nullptr, stats.has_value() ? stats->data_ptr<float>() : nullptr);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// acc[i] += sum_s wts[s] * active[s] * (wg[i] - wl[s][i])
__global__ void __launch_bounds__(kThreads)
accumulate_pg_kernel(float* __restrict__ acc, const float* __restrict__ wg, const float* __restrict__ wl, int64_t P,
                     int S, const float* __restrict__ wts, const int* __restrict__ active) {
  extern __shared__ float s_w[];
  for (int s = threadIdx.x; s < S; s += blockDim.x) s_w[s] = (active == nullptr || active[s] != 0) ? wts[s] : 0.f;
  __syncthreads();
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const float4* g4 = reinterpret_cast<const float4*>(wg);
  float4* a4 = reinterpret_cast<float4*>(acc);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 x = ld_stream(g4 + i);
    float4 r = ld_na(a4 + i);
    for (int s = 0; s < S; ++s) {
      const float wt = s_w[s];
      if (wt == 0.f) continue;
      const float4 y = ld_stream(reinterpret_cast<const float4*>(wl + static_cast<int64_t>(s) * P) + i);
      r.x = fmaf(wt, x.x - y.x, r.x); r.y = fmaf(wt, x.y - y.y, r.y);
      r.z = fmaf(wt, x.z - y.z, r.z); r.w = fmaf(wt, x.w - y.w, r.w);
    }
    st_stream(a4 + i, r);
  }
}

// ---- compact slot arenas -------------------------------------------------------------------------------------------
// A slot arena may hold only the LIVE elements of the model (filter taps that can only ever see zero padding have
// identically zero gradients, so with weight_decay == 0 they never change and need not exist per client).
// map[j] = index of slot element j in the global arena, -1 for alignment padding.
__global__ void __launch_bounds__(kThreads)
slot_scatter_in_kernel(float* __restrict__ W, int64_t Pc, int S, const float* __restrict__ wg, const int* __restrict__ map) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; j < Pc; j += stride) {
    const int m = map[j];
    const float v = m >= 0 ? __ldg(wg + m) : 0.f;
    for (int s = 0; s < S; ++s) W[static_cast<int64_t>(s) * Pc + j] = v;
  }
}

// acc[map[j]] += sum_s wts[s] * active[s] * (wg[map[j]] - wl[s][j])      (map is injective: no atomics)
__global__ void __launch_bounds__(kThreads)
accumulate_pg_mapped_kernel(float* __restrict__ acc, const float* __restrict__ wg, const float* __restrict__ wl, int64_t Pc,
                            int S, const float* __restrict__ wts, const int* __restrict__ active, const int* __restrict__ map) {
  extern __shared__ float s_w[];
  for (int s = threadIdx.x; s < S; s += blockDim.x) s_w[s] = (active == nullptr || active[s] != 0) ? wts[s] : 0.f;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; j < Pc; j += stride) {
    const int m = map[j];
    if (m < 0) continue;
    const float x = __ldg(wg + m);
    float r = acc[m];
    for (int s = 0; s < S; ++s) {
      const float wt = s_w[s];
      if (wt != 0.f) r = fmaf(wt, x - wl[static_cast<int64_t>(s) * Pc + j], r);
    }
    acc[m] = r;
  }
}

void slot_scatter_in(torch::Tensor W, torch::Tensor wg, torch::Tensor map) {
  check_rows(W, "W");
  TORCH_CHECK(map.is_cuda() && map.scalar_type() == torch::kInt32 && map.numel() == W.size(1) && wg.is_cuda() &&
              wg.scalar_type() == torch::kFloat32, "slot_scatter_in: int32 map of the slot row length expected");
  const c10::cuda::CUDAGuard guard(W.device());
  const int64_t Pc = W.size(1);
  slot_scatter_in_kernel<<<blocks_for(Pc), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
      W.data_ptr<float>(), Pc, static_cast<int>(W.size(0)), wg.data_ptr<float>(), map.data_ptr<int>());
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void accumulate_pseudo_grad_mapped(torch::Tensor acc, torch::Tensor wg, torch::Tensor wl, torch::Tensor weights,
                                   c10::optional<torch::Tensor> active, torch::Tensor map) {
  check_rows(wl, "w_local");
  const int S = static_cast<int>(wl.size(0));
  const int64_t Pc = wl.size(1);
  TORCH_CHECK(map.is_cuda() && map.scalar_type() == torch::kInt32 && map.numel() == Pc && acc.numel() == wg.numel() &&
              weights.numel() == S && weights.scalar_type() == torch::kFloat32);
  const c10::cuda::CUDAGuard guard(acc.device());
  accumulate_pg_mapped_kernel<<<blocks_for(Pc), kThreads, S * sizeof(float), at::cuda::getCurrentCUDAStream()>>>(
      acc.data_ptr<float>(), wg.data_ptr<float>(), wl.data_ptr<float>(), Pc, S, weights.data_ptr<float>(),
      active.has_value() ? active->data_ptr<int>() : nullptr, map.data_ptr<int>());
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void accumulate_pseudo_grad(torch::Tensor acc, torch::Tensor wg, torch::Tensor wl, torch::Tensor weights,
                            c10::optional<torch::Tensor> active) {
  check_rows(wl, "w_local");
  const int S = static_cast<int>(wl.size(0));
  const int64_t P = wl.size(1);
  TORCH_CHECK(acc.numel() == P && wg.numel() == P && weights.numel() == S && weights.scalar_type() == torch::kFloat32);
  const c10::cuda::CUDAGuard guard(acc.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  accumulate_pg_kernel<<<blocks_for(P >> 2), kThreads, S * sizeof(float), stream>>>(
      acc.data_ptr<float>(), wg.data_ptr<float>(), wl.data_ptr<float>(), P, S, weights.data_ptr<float>(),
      active.has_value() ? active->data_ptr<int>() : nullptr);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flute
