// Fused recurrent-cell pointwise kernels.
//
// GRU (reference: experiments/nlg_gru/model.py:19-30, a python time loop of ~8 ATen kernels per step):
//     r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)   h' = n + z * (h - n)
// LSTM (nn.LSTM gate order i, f, g, o):
//     c' = sigmoid(f) * c + sigmoid(i) * tanh(g)            h' = sigmoid(o) * tanh(c')
//
// The gate GEMMs run on the tensor cores (gemm_tcgen05.cu / cuBLAS); these kernels fuse everything after them into
// one launch forward and one backward.  The backward recomputes the gate activations from the saved
// pre-activations, so no activation tensors are stored.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/types.h>
#include "common.cuh"

namespace flute {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void gru_cell_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ h, float* __restrict__ out, int B, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const float* gib = gi + static_cast<int64_t>(b) * 3 * H;
  const float* ghb = gh + static_cast<int64_t>(b) * 3 * H;
  const float r = sigmoidf_(gib[j] + ghb[j]);
  const float z = sigmoidf_(gib[H + j] + ghb[H + j]);
  const float n = tanhf(gib[2 * H + j] + r * ghb[2 * H + j]);
  const float hp = h[idx];
  out[idx] = n + z * (hp - n);
}

__global__ void gru_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gi,
                                    const float* __restrict__ gh, const float* __restrict__ h,
                                    float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dhp, int B,
                                    int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const int64_t o = static_cast<int64_t>(b) * 3 * H;
  const float hr = gh[o + j], hz = gh[o + H + j], hn = gh[o + 2 * H + j];
  const float r = sigmoidf_(gi[o + j] + hr);
  const float z = sigmoidf_(gi[o + H + j] + hz);
  const float n = tanhf(gi[o + 2 * H + j] + r * hn);
  const float hp = h[idx], g = dh[idx];
  // h' = n + z (h - n)
  const float dn = g * (1.f - z);
  const float dz = g * (hp - n);
  const float dpre_n = dn * (1.f - n * n);
  const float dr = dpre_n * hn;
  const float dpre_r = dr * r * (1.f - r);
  const float dpre_z = dz * z * (1.f - z);
  dgi[o + j] = dpre_r;           dgh[o + j] = dpre_r;
  dgi[o + H + j] = dpre_z;       dgh[o + H + j] = dpre_z;
  dgi[o + 2 * H + j] = dpre_n;   dgh[o + 2 * H + j] = dpre_n * r;
  dhp[idx] = g * z;
}

__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c,
                                     float* __restrict__ h2, float* __restrict__ c2, int B, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const float* g = gates + static_cast<int64_t>(b) * 4 * H;
  const float i = sigmoidf_(g[j]), f = sigmoidf_(g[H + j]), gg = tanhf(g[2 * H + j]), o = sigmoidf_(g[3 * H + j]);
  const float cn = f * c[idx] + i * gg;
  c2[idx] = cn;
  h2[idx] = o * tanhf(cn);
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dc,
                                     const float* __restrict__ gates, const float* __restrict__ c,
                                     const float* __restrict__ c2, float* __restrict__ dgates,
                                     float* __restrict__ dcp, int B, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const int64_t off = static_cast<int64_t>(b) * 4 * H;
  const float i = sigmoidf_(gates[off + j]), f = sigmoidf_(gates[off + H + j]);
  const float gg = tanhf(gates[off + 2 * H + j]), o = sigmoidf_(gates[off + 3 * H + j]);
  const float tc = tanhf(c2[idx]);
  const float dho = dh[idx];
  const float dcn = dc[idx] + dho * o * (1.f - tc * tc);
  dgates[off + j] = dcn * gg * i * (1.f - i);
  dgates[off + H + j] = dcn * c[idx] * f * (1.f - f);
  dgates[off + 2 * H + j] = dcn * i * (1.f - gg * gg);
  dgates[off + 3 * H + j] = dho * tc * o * (1.f - o);
  dcp[idx] = dcn * f;
}

static void check2d(const at::Tensor& t, int64_t cols, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous() && t.dim() == 2 && t.size(1) == cols, name,
              ": expected contiguous fp32 CUDA [B, ", cols, "]");
}

at::Tensor gru_cell_fwd(at::Tensor gi, at::Tensor gh, at::Tensor h) {
  const int B = static_cast<int>(h.size(0)), H = static_cast<int>(h.size(1));
  check2d(h, H, "h"); check2d(gi, 3 * H, "gi"); check2d(gh, 3 * H, "gh");
  const c10::cuda::CUDAGuard guard(h.device());
  auto out = at::empty_like(h);
  const int n = B * H;
  gru_cell_fwd_kernel<<<(n + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      gi.data_ptr<float>(), gh.data_ptr<float>(), h.data_ptr<float>(), out.data_ptr<float>(), B, H);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

std::vector<at::Tensor> gru_cell_bwd(at::Tensor dh, at::Tensor gi, at::Tensor gh, at::Tensor h) {
  const int B = static_cast<int>(h.size(0)), H = static_cast<int>(h.size(1));
  check2d(dh, H, "dh");
  const c10::cuda::CUDAGuard guard(h.device());
  auto dgi = at::empty_like(gi), dgh = at::empty_like(gh), dhp = at::empty_like(h);
  const int n = B * H;
  gru_cell_bwd_kernel<<<(n + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      dh.data_ptr<float>(), gi.data_ptr<float>(), gh.data_ptr<float>(), h.data_ptr<float>(), dgi.data_ptr<float>(),
      dgh.data_ptr<float>(), dhp.data_ptr<float>(), B, H);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {dgi, dgh, dhp};
}

std::vector<at::Tensor> lstm_cell_fwd(at::Tensor gates, at::Tensor c) {
  const int B = static_cast<int>(c.size(0)), H = static_cast<int>(c.size(1));
  check2d(c, H, "c"); check2d(gates, 4 * H, "gates");
  const c10::cuda::CUDAGuard guard(c.device());
  auto h2 = at::empty_like(c), c2 = at::empty_like(c);
  const int n = B * H;
  lstm_cell_fwd_kernel<<<(n + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      gates.data_ptr<float>(), c.data_ptr<float>(), h2.data_ptr<float>(), c2.data_ptr<float>(), B, H);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {h2, c2};
}

std::vector<at::Tensor> lstm_cell_bwd(at::Tensor dh, at::Tensor dc, at::Tensor gates, at::Tensor c, at::Tensor c2) {
  const int B = static_cast<int>(c.size(0)), H = static_cast<int>(c.size(1));
  const c10::cuda::CUDAGuard guard(c.device());
  auto dgates = at::empty_like(gates), dcp = at::empty_like(c);
  const int n = B * H;
  lstm_cell_bwd_kernel<<<(n + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      dh.data_ptr<float>(), dc.data_ptr<float>(), gates.data_ptr<float>(), c.data_ptr<float>(), c2.data_ptr<float>(),
      dgates.data_ptr<float>(), dcp.data_ptr<float>(), B, H);
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return {dgates, dcp};
}

}  // namespace flute
