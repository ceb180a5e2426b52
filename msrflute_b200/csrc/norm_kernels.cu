// GroupNorm forward/backward fused with affine, residual-add and ReLU.
//
// Reference call site: experiments/cv_resnet_fedcifar100/group_normalization.py:59-84 — GroupNorm implemented as a
// reshape + F.batch_norm with a *per-group* affine pair, followed by separate add / ReLU kernels in the block
// (model.py:44-60).  Here one kernel does  y = relu( (x-mean)*rstd*gamma + beta + residual )  and one kernel
// the whole backward (dx, d_residual, per-row d_gamma/d_beta partials).
//
// Shapes in the FL benchmarks are tiny: a row (one sample x one group) is (C/G)*H*W in {512,128,32,8,2} contiguous
// elements, and there are N*G in {640..5120} rows.  One warp per row, lanes stride the row, shuffle reductions; the
// kernels are launch-latency bound, so the win is the fusion (1 launch instead of 3-4), not bandwidth.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

constexpr int kWarpsPerBlock = 8;

// gamma/beta index for element j of row (n, g): per-group -> g ; per-channel -> g*cpg + j/HW
template <typename T, bool kPerGroup>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
group_norm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const T* __restrict__ res, T* __restrict__ y, float* __restrict__ mean_out,
                      float* __restrict__ rstd_out, int rows, int G, int cpg, int HW, float eps, bool relu,
                      int rows_per_set, long long set_stride) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kWarpsPerBlock;
  const int L = cpg * HW;
  for (int row = warp; row < rows; row += nwarps) {
    const int64_t base = static_cast<int64_t>(row) * L;
    const int g = row % G;
    const long long aoff = (row / rows_per_set) * set_stride;   // which affine set (simulated client) the row uses
    float s = 0.f;
    for (int j = lane; j < L; j += 32) s += to_f<T>(x[base + j]);
    const float mean = warp_sum(s) / L;
    float ss = 0.f;                                  // two-pass variance: rows are tiny and L1-resident, and
    for (int j = lane; j < L; j += 32) {             // E[x^2]-mean^2 cancels catastrophically for L = 2..8
      const float d = to_f<T>(x[base + j]) - mean;
      ss = fmaf(d, d, ss);
    }
    const float var = warp_sum(ss) / L;
    const float rstd = rsqrtf(var + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    for (int j = lane; j < L; j += 32) {
      const long long a = aoff + (kPerGroup ? g : g * cpg + j / HW);
      float v = (to_f<T>(x[base + j]) - mean) * rstd * gamma[a] + beta[a];
      if (res != nullptr) v += to_f<T>(res[base + j]);
      if (relu) v = fmaxf(v, 0.f);
      y[base + j] = from_f<T>(v);
    }
  }
}

// dgamma_part / dbeta_part: per-group -> [rows]; per-channel -> [rows*cpg]  (summed over N afterwards)
template <typename T, bool kPerGroup>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
group_norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const T* __restrict__ y,
                      T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dgamma_part,
                      float* __restrict__ dbeta_part, int rows, int G, int cpg, int HW, bool relu, int rows_per_set,
                      long long set_stride, float* __restrict__ dg_acc, float* __restrict__ db_acc) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kWarpsPerBlock;
  const int L = cpg * HW;
  for (int row = warp; row < rows; row += nwarps) {
    const int64_t base = static_cast<int64_t>(row) * L;
    const int g = row % G;
    const long long aoff = (row / rows_per_set) * set_stride;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float a = 0.f, b = 0.f;        // sum(dy*gamma), sum(dy*gamma*xhat)
    if (kPerGroup) {
      const float gm = gamma[aoff + g];
      float sg = 0.f, sb = 0.f;
      for (int j = lane; j < L; j += 32) {
        float d = to_f<T>(dy[base + j]);
        if (relu && to_f<T>(y[base + j]) <= 0.f) d = 0.f;
        const float xh = (to_f<T>(x[base + j]) - mean) * rstd;
        sb += d;
        sg = fmaf(d, xh, sg);
      }
      sg = warp_sum(sg);
      sb = warp_sum(sb);
      a = sb * gm;
      b = sg * gm;
      if (lane == 0) {
        if (dg_acc != nullptr) { atomicAdd(dg_acc + aoff + g, sg); atomicAdd(db_acc + aoff + g, sb); }
        else { dgamma_part[row] = sg; dbeta_part[row] = sb; }
      }
    } else {
      for (int c = 0; c < cpg; ++c) {
        const float gm = gamma[aoff + g * cpg + c];
        float sg = 0.f, sb = 0.f;
        for (int j = lane; j < HW; j += 32) {
          const int64_t k = base + static_cast<int64_t>(c) * HW + j;
          float d = to_f<T>(dy[k]);
          if (relu && to_f<T>(y[k]) <= 0.f) d = 0.f;
          const float xh = (to_f<T>(x[k]) - mean) * rstd;
          sb += d;
          sg = fmaf(d, xh, sg);
        }
        sg = warp_sum(sg);
        sb = warp_sum(sb);
        a = fmaf(sb, gm, a);
        b = fmaf(sg, gm, b);
        if (lane == 0) {
          if (dg_acc != nullptr) { atomicAdd(dg_acc + aoff + g * cpg + c, sg); atomicAdd(db_acc + aoff + g * cpg + c, sb); }
          else {
            dgamma_part[static_cast<int64_t>(row) * cpg + c] = sg;
            dbeta_part[static_cast<int64_t>(row) * cpg + c] = sb;
          }
        }
      }
    }
    const float c1 = a / L, c2 = b / L;
    for (int j = lane; j < L; j += 32) {
      float d = to_f<T>(dy[base + j]);
      if (relu && to_f<T>(y[base + j]) <= 0.f) d = 0.f;
      if (dres != nullptr) dres[base + j] = from_f<T>(d);
      const float gm = gamma[aoff + (kPerGroup ? g : g * cpg + j / HW)];
      const float xh = (to_f<T>(x[base + j]) - mean) * rstd;
      dx[base + j] = from_f<T>(rstd * (d * gm - c1 - xh * c2));
    }
  }
}

static inline int gn_blocks(int rows) {
  return std::max(1, std::min((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, 148 * 8));
}

std::vector<torch::Tensor> group_norm_fwd(torch::Tensor x, torch::Tensor weight, torch::Tensor bias,
                                          c10::optional<torch::Tensor> residual, int64_t G, double eps, bool relu,
                                          bool per_group_affine, int64_t sets) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.dim() >= 2, "x must be contiguous CUDA (N, C, ...)");
  const int N = static_cast<int>(x.size(0)), C = static_cast<int>(x.size(1));
  TORCH_CHECK(C % G == 0, "channels not divisible by groups");
  const int cpg = C / static_cast<int>(G);
  const int HW = static_cast<int>(x.numel() / (static_cast<int64_t>(N) * C));
  const int rows = N * static_cast<int>(G);
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto y = torch::empty_like(x);
  auto fopt = x.options().dtype(torch::kFloat32);
  auto mean = torch::empty({rows}, fopt), rstd = torch::empty({rows}, fopt);
  auto wf = weight.to(torch::kFloat32).contiguous(), bf = bias.to(torch::kFloat32).contiguous();
  const int A = static_cast<int>(per_group_affine ? G : C);
  TORCH_CHECK(sets >= 1 && N % sets == 0 && wf.numel() == sets * A && bf.numel() == sets * A,
              "affine parameter size mismatch");
  const int rows_per_set = rows / static_cast<int>(sets);
  if (residual.has_value()) TORCH_CHECK(residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type());
  const int blocks = gn_blocks(rows);
#define LAUNCH_FWD(T, PG)                                                                                              \
  group_norm_fwd_kernel<T, PG><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(                                           \
      reinterpret_cast<const T*>(x.data_ptr()), wf.data_ptr<float>(), bf.data_ptr<float>(),                           \
      residual.has_value() ? reinterpret_cast<const T*>(residual->data_ptr()) : nullptr,                              \
      reinterpret_cast<T*>(y.data_ptr()), mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, static_cast<int>(G),  \
      cpg, HW, static_cast<float>(eps), relu, rows_per_set, static_cast<long long>(A))
  if (x.scalar_type() == torch::kFloat32) {
    if (per_group_affine) LAUNCH_FWD(float, true); else LAUNCH_FWD(float, false);
  } else if (x.scalar_type() == torch::kBFloat16) {
    if (per_group_affine) LAUNCH_FWD(__nv_bfloat16, true); else LAUNCH_FWD(__nv_bfloat16, false);
  } else {
    TORCH_CHECK(false, "group_norm: fp32 or bf16 only");
  }
#undef LAUNCH_FWD
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {y, mean, rstd};
}

std::vector<torch::Tensor> group_norm_bwd(torch::Tensor dy, torch::Tensor x, torch::Tensor weight, torch::Tensor mean,
                                          torch::Tensor rstd, c10::optional<torch::Tensor> y, int64_t G, bool relu,
                                          bool per_group_affine, bool has_residual, int64_t sets) {
  const int N = static_cast<int>(x.size(0)), C = static_cast<int>(x.size(1));
  const int cpg = C / static_cast<int>(G);
  const int HW = static_cast<int>(x.numel() / (static_cast<int64_t>(N) * C));
  const int rows = N * static_cast<int>(G);
  TORCH_CHECK(!relu || y.has_value(), "ReLU backward needs the saved output");
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto dx = torch::empty_like(x);
  torch::Tensor dres = has_residual ? torch::empty_like(x) : torch::Tensor();
  auto fopt = x.options().dtype(torch::kFloat32);
  const int64_t parts = per_group_affine ? rows : static_cast<int64_t>(rows) * cpg;
  auto dgp = torch::empty({parts}, fopt), dbp = torch::empty({parts}, fopt);
  auto wf = weight.to(torch::kFloat32).contiguous();
  const int A_ = static_cast<int>(per_group_affine ? G : C);
  const int rows_per_set = rows / static_cast<int>(sets);
  const int blocks = gn_blocks(rows);
#define LAUNCH_BWD(T, PG)                                                                                              \
  group_norm_bwd_kernel<T, PG><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(                                           \
      reinterpret_cast<const T*>(dy.data_ptr()), reinterpret_cast<const T*>(x.data_ptr()), wf.data_ptr<float>(),      \
      mean.data_ptr<float>(), rstd.data_ptr<float>(), y.has_value() ? reinterpret_cast<const T*>(y->data_ptr()) : nullptr, \
      reinterpret_cast<T*>(dx.data_ptr()), has_residual ? reinterpret_cast<T*>(dres.data_ptr()) : nullptr,            \
      dgp.data_ptr<float>(), dbp.data_ptr<float>(), rows, static_cast<int>(G), cpg, HW, relu, rows_per_set,            \
      static_cast<long long>(A_), nullptr, nullptr)
  if (x.scalar_type() == torch::kFloat32) {
    if (per_group_affine) LAUNCH_BWD(float, true); else LAUNCH_BWD(float, false);
  } else if (x.scalar_type() == torch::kBFloat16) {
    if (per_group_affine) LAUNCH_BWD(__nv_bfloat16, true); else LAUNCH_BWD(__nv_bfloat16, false);
  } else {
    TORCH_CHECK(false, "group_norm: fp32 or bf16 only");
  }
#undef LAUNCH_BWD
  FLUTE_CUDA_CHECK(cudaGetLastError());
  const int64_t A = per_group_affine ? G : C;
  auto dw = dgp.view({sets, N / sets, A}).sum(1).reshape(weight.sizes()).to(weight.scalar_type());
  auto db = dbp.view({sets, N / sets, A}).sum(1).reshape(weight.sizes()).to(weight.scalar_type());
  return {dx, dw, db, has_residual ? dres : torch::Tensor()};
}

// ---- arena variants for the slot-batched engine: gamma/beta are read from, and their gradients accumulated into,
//      the [S, P] parameter / gradient arenas (slot s at base + s*P + offset); x is [S*B, C, H, W] fp32.
std::vector<torch::Tensor> group_norm_fwd_arena(torch::Tensor x, torch::Tensor w_arena, int64_t w_off, int64_t b_off,
                                                c10::optional<torch::Tensor> residual, int64_t G, double eps, bool relu,
                                                bool per_group_affine, int64_t sets) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.scalar_type() == torch::kFloat32 && x.dim() >= 2);
  TORCH_CHECK(w_arena.dim() == 2 && w_arena.is_contiguous() && w_arena.scalar_type() == torch::kFloat32);
  const int N = static_cast<int>(x.size(0)), C = static_cast<int>(x.size(1));
  const int cpg = C / static_cast<int>(G);
  const int HW = static_cast<int>(x.numel() / (static_cast<int64_t>(N) * C));
  const int rows = N * static_cast<int>(G);
  TORCH_CHECK(N % sets == 0 && w_arena.size(0) >= sets);
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto y = torch::empty_like(x);
  auto mean = torch::empty({rows}, x.options()), rstd = torch::empty({rows}, x.options());
  const int rows_per_set = rows / static_cast<int>(sets);
  const float* gp = w_arena.data_ptr<float>() + w_off;
  const float* bp = w_arena.data_ptr<float>() + b_off;
  const float* rp = residual.has_value() ? residual->data_ptr<float>() : nullptr;
  const long long P = w_arena.size(1);
  const int blocks = gn_blocks(rows);
  if (per_group_affine)
    group_norm_fwd_kernel<float, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
        x.data_ptr<float>(), gp, bp, rp, y.data_ptr<float>(), mean.data_ptr<float>(), rstd.data_ptr<float>(), rows,
        static_cast<int>(G), cpg, HW, static_cast<float>(eps), relu, rows_per_set, P);
  else
    group_norm_fwd_kernel<float, false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
        x.data_ptr<float>(), gp, bp, rp, y.data_ptr<float>(), mean.data_ptr<float>(), rstd.data_ptr<float>(), rows,
        static_cast<int>(G), cpg, HW, static_cast<float>(eps), relu, rows_per_set, P);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {y, mean, rstd};
}

std::vector<torch::Tensor> group_norm_bwd_arena(torch::Tensor dy, torch::Tensor x, torch::Tensor w_arena, int64_t w_off,
                                                torch::Tensor mean, torch::Tensor rstd, c10::optional<torch::Tensor> y,
                                                int64_t G, bool relu, bool per_group_affine, bool has_residual,
                                                int64_t sets, torch::Tensor g_arena, int64_t gw_off, int64_t gb_off) {
  const int N = static_cast<int>(x.size(0)), C = static_cast<int>(x.size(1));
  const int cpg = C / static_cast<int>(G);
  const int HW = static_cast<int>(x.numel() / (static_cast<int64_t>(N) * C));
  const int rows = N * static_cast<int>(G);
  TORCH_CHECK(dy.is_contiguous() && dy.scalar_type() == torch::kFloat32 && g_arena.sizes() == w_arena.sizes());
  const c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto dx = torch::empty_like(x);
  torch::Tensor dres = has_residual ? torch::empty_like(x) : torch::Tensor();
  const int rows_per_set = rows / static_cast<int>(sets);
  const long long P = w_arena.size(1);
  const float* yp = y.has_value() ? y->data_ptr<float>() : nullptr;
  float* drp = has_residual ? dres.data_ptr<float>() : nullptr;
  const int blocks = gn_blocks(rows);
  if (per_group_affine)
    group_norm_bwd_kernel<float, true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
        dy.data_ptr<float>(), x.data_ptr<float>(), w_arena.data_ptr<float>() + w_off, mean.data_ptr<float>(),
        rstd.data_ptr<float>(), yp, dx.data_ptr<float>(), drp, nullptr, nullptr, rows, static_cast<int>(G), cpg, HW, relu,
        rows_per_set, P, g_arena.data_ptr<float>() + gw_off, g_arena.data_ptr<float>() + gb_off);
  else
    group_norm_bwd_kernel<float, false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
        dy.data_ptr<float>(), x.data_ptr<float>(), w_arena.data_ptr<float>() + w_off, mean.data_ptr<float>(),
        rstd.data_ptr<float>(), yp, dx.data_ptr<float>(), drp, nullptr, nullptr, rows, static_cast<int>(G), cpg, HW, relu,
        rows_per_set, P, g_arena.data_ptr<float>() + gw_off, g_arena.data_ptr<float>() + gb_off);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {dx, has_residual ? dres : torch::Tensor()};
}

}  // namespace flute
