// Slot-layout gather path of the device engine.
//
// The slot arenas hold every client's parameters in the executor's own layout ([Cout, live taps, Cin] filters, dead
// taps elided); `map[j]` is element j's position in the global (PyTorch-layout) arena.  Going through the map once PER
// CLIENT makes the broadcast / gather passes sector-amplified scatter/gathers (a [.., Cin] run is strided by KH*KW in
// the global layout).  Here the permutation is paid once per round on ONE row:
//
//   slot_gather_bcast : wg_slot[j] = wg[map[j]]  and  W[s, j] = wg_slot[j] for every slot       (model distribution)
//   slot_pg_sqnorm    : n[s] = || wg_slot - W[s] ||^2                                           (local-DP clip / normalise)
//   slot_gather_fused : acc_slot[j] += sum_s coef[s] * (wg_slot[j] - W[s, j]) + sum_s sig[s] * N(seed[s], j)
//                       — pseudo-gradient, aggregation weight, local-DP scale and Philox Gaussian noise in ONE
//                       vectorised pass over the slot arenas (the reference: client.py:380-383, dga.py:142-146,
//                       privacy/__init__.py:154-201, strategies/utils.py:21-33 — three flat copies per client + D2H)
//   slot_scatter_acc  : acc[map[j]] += acc_slot[j]; acc_slot[j] = 0                              (one row through the map)
//   dead_coord_noise  : acc[idx[i]] += sig * N(seed, i)  for coordinates that exist in no slot (local-DP noise is added
//                       to EVERY coordinate of a client's update; the sum of K independent N(0, sig_s^2) is one
//                       N(0, sum sig_s^2) draw)
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {
namespace gk {

constexpr int kThreads = 256;
constexpr int kMaxSlots = 64;

static inline int blocks_for(int64_t n_items, int per_thread = 1) {
  const int64_t want = (n_items + static_cast<int64_t>(kThreads) * per_thread - 1) / (static_cast<int64_t>(kThreads) * per_thread);
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 4)));
}

__global__ void __launch_bounds__(kThreads)
gather_bcast_kernel(float* __restrict__ W, float* __restrict__ wg_slot, const float* __restrict__ wg,
                    const int* __restrict__ map, int64_t Pc, int S) {
  const int64_t n4 = Pc >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const int4 m = *reinterpret_cast<const int4*>(map + 4 * i);
    float4 v;
    v.x = m.x >= 0 ? __ldg(wg + m.x) : 0.f;
    v.y = m.y >= 0 ? __ldg(wg + m.y) : 0.f;
    v.z = m.z >= 0 ? __ldg(wg + m.z) : 0.f;
    v.w = m.w >= 0 ? __ldg(wg + m.w) : 0.f;
    reinterpret_cast<float4*>(wg_slot)[i] = v;
    for (int s = 0; s < S; ++s) st_stream(reinterpret_cast<float4*>(W + static_cast<int64_t>(s) * Pc) + i, v);
  }
}

__global__ void __launch_bounds__(kThreads)
pg_sqnorm_kernel(const float* __restrict__ W, const float* __restrict__ wg_slot, int64_t Pc, float* __restrict__ out) {
  const int s = blockIdx.y;
  const float4* wv = reinterpret_cast<const float4*>(W + static_cast<int64_t>(s) * Pc);
  const float4* gv = reinterpret_cast<const float4*>(wg_slot);
  const int64_t n4 = Pc >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  float a = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 w = ld_stream(wv + i), g = __ldg(gv + i);
    const float dx = g.x - w.x, dy = g.y - w.y, dz = g.z - w.z, dw = g.w - w.w;
    a += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float2 r = block_sum2(a, 0.f);
  if (threadIdx.x == 0) atomicAdd(out + s, r.x);
}

// coef / sig / seed are read from device memory (computed on the device from the norms: no host round trip)
__global__ void __launch_bounds__(kThreads)
gather_fused_kernel(float* __restrict__ acc_slot, const float* __restrict__ W, const float* __restrict__ wg_slot, int64_t Pc,
                    int S, const float* __restrict__ coef, const float* __restrict__ sig,
                    const long long* __restrict__ seed, int any_noise) {
  __shared__ float s_coef[kMaxSlots], s_sig[kMaxSlots];
  __shared__ unsigned long long s_seed[kMaxSlots];
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    s_coef[s] = coef[s];
    s_sig[s] = sig != nullptr ? sig[s] : 0.f;
    s_seed[s] = seed != nullptr ? static_cast<unsigned long long>(seed[s]) : 0ull;
  }
  __syncthreads();
  const int64_t n4 = Pc >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(wg_slot) + i);
    float4 r = reinterpret_cast<float4*>(acc_slot)[i];
    for (int s = 0; s < S; ++s) {
      const float c = s_coef[s];
      if (c != 0.f) {
        const float4 w = ld_stream(reinterpret_cast<const float4*>(W + static_cast<int64_t>(s) * Pc) + i);
        r.x = fmaf(c, g.x - w.x, r.x); r.y = fmaf(c, g.y - w.y, r.y);
        r.z = fmaf(c, g.z - w.z, r.z); r.w = fmaf(c, g.w - w.w, r.w);
      }
      if (any_noise && s_sig[s] != 0.f) {
        const float4 n = philox_normal4(s_seed[s], static_cast<uint64_t>(i));
        r.x = fmaf(s_sig[s], n.x, r.x); r.y = fmaf(s_sig[s], n.y, r.y);
        r.z = fmaf(s_sig[s], n.z, r.z); r.w = fmaf(s_sig[s], n.w, r.w);
      }
    }
    reinterpret_cast<float4*>(acc_slot)[i] = r;
  }
}

__global__ void __launch_bounds__(kThreads)
scatter_acc_kernel(float* __restrict__ acc, float* __restrict__ acc_slot, const int* __restrict__ map, int64_t Pc) {
  const int64_t n4 = Pc >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const int4 m = *reinterpret_cast<const int4*>(map + 4 * i);
    const float4 v = reinterpret_cast<float4*>(acc_slot)[i];
    if (m.x >= 0) acc[m.x] += v.x;              // the map is injective: no atomics
    if (m.y >= 0) acc[m.y] += v.y;
    if (m.z >= 0) acc[m.z] += v.z;
    if (m.w >= 0) acc[m.w] += v.w;
    reinterpret_cast<float4*>(acc_slot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kThreads)
dead_noise_kernel(float* __restrict__ acc, const int* __restrict__ idx, int64_t n, const float* __restrict__ sig2_sum,
                  long long seed) {
  const float sg = sqrtf(fmaxf(*sig2_sum, 0.f));
  if (sg == 0.f) return;
  const int64_t nq = (n + 3) >> 2, stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; q < nq; q += stride) {
    const float4 z = philox_normal4(static_cast<uint64_t>(seed), static_cast<uint64_t>(q));
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t i = 4 * q + e;
      if (i < n) acc[idx[i]] += sg * zz[e];
    }
  }
}

static void check_slot(const torch::Tensor& W) {
  TORCH_CHECK(W.is_cuda() && W.scalar_type() == torch::kFloat32 && W.dim() == 2 && W.is_contiguous() && W.size(1) % 4 == 0,
              "slot arena: contiguous fp32 CUDA [S, P], P % 4 == 0");
  TORCH_CHECK(W.size(0) <= kMaxSlots, "at most ", kMaxSlots, " slots");
}

}  // namespace gk

void slot_gather_bcast(torch::Tensor W, torch::Tensor wg_slot, torch::Tensor wg, torch::Tensor map) {
  using namespace gk;
  check_slot(W);
  const int64_t Pc = W.size(1);
  TORCH_CHECK(map.is_cuda() && map.scalar_type() == torch::kInt32 && map.numel() == Pc && wg_slot.numel() == Pc &&
              wg.is_cuda() && wg.scalar_type() == torch::kFloat32);
  const c10::cuda::CUDAGuard guard(W.device());
  gather_bcast_kernel<<<blocks_for(Pc >> 2), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
      W.data_ptr<float>(), wg_slot.data_ptr<float>(), wg.data_ptr<float>(), map.data_ptr<int>(), Pc,
      static_cast<int>(W.size(0)));
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// out[s] += ||wg_slot - W[s]||^2   (out must be zeroed by the caller)
void slot_pg_sqnorm(torch::Tensor W, torch::Tensor wg_slot, torch::Tensor out) {
  using namespace gk;
  check_slot(W);
  const int64_t Pc = W.size(1);
  TORCH_CHECK(wg_slot.numel() == Pc && out.numel() == W.size(0) && out.scalar_type() == torch::kFloat32);
  const c10::cuda::CUDAGuard guard(W.device());
  dim3 grid(std::min(blocks_for(Pc >> 2, 4), 148), static_cast<unsigned>(W.size(0)));
  pg_sqnorm_kernel<<<grid, kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(W.data_ptr<float>(), wg_slot.data_ptr<float>(), Pc,
                                                                           out.data_ptr<float>());
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void slot_gather_fused(torch::Tensor acc_slot, torch::Tensor W, torch::Tensor wg_slot, torch::Tensor coef,
                       c10::optional<torch::Tensor> sig, c10::optional<torch::Tensor> seed) {
  using namespace gk;
  check_slot(W);
  const int64_t Pc = W.size(1);
  const int S = static_cast<int>(W.size(0));
  TORCH_CHECK(acc_slot.numel() == Pc && wg_slot.numel() == Pc && coef.numel() == S && coef.scalar_type() == torch::kFloat32);
  const bool noise = sig.has_value() && seed.has_value();
  if (noise) TORCH_CHECK(sig->numel() == S && seed->numel() == S && seed->scalar_type() == torch::kInt64);
  const c10::cuda::CUDAGuard guard(W.device());
  gather_fused_kernel<<<blocks_for(Pc >> 2), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
      acc_slot.data_ptr<float>(), W.data_ptr<float>(), wg_slot.data_ptr<float>(), Pc, S, coef.data_ptr<float>(),
      noise ? sig->data_ptr<float>() : nullptr, noise ? reinterpret_cast<const long long*>(seed->data_ptr<int64_t>()) : nullptr,
      noise ? 1 : 0);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void slot_scatter_acc(torch::Tensor acc, torch::Tensor acc_slot, torch::Tensor map) {
  using namespace gk;
  const int64_t Pc = acc_slot.numel();
  TORCH_CHECK(map.numel() == Pc && map.scalar_type() == torch::kInt32 && Pc % 4 == 0 && acc.is_cuda());
  const c10::cuda::CUDAGuard guard(acc.device());
  scatter_acc_kernel<<<blocks_for(Pc >> 2), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
      acc.data_ptr<float>(), acc_slot.data_ptr<float>(), map.data_ptr<int>(), Pc);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

void dead_coord_noise(torch::Tensor acc, torch::Tensor idx, torch::Tensor sig2_sum, int64_t seed) {
  using namespace gk;
  TORCH_CHECK(idx.scalar_type() == torch::kInt32 && sig2_sum.numel() == 1 && sig2_sum.scalar_type() == torch::kFloat32);
  const int64_t n = idx.numel();
  if (n == 0) return;
  const c10::cuda::CUDAGuard guard(acc.device());
  dead_noise_kernel<<<blocks_for((n + 3) >> 2), kThreads, 0, at::cuda::getCurrentCUDAStream()>>>(
      acc.data_ptr<float>(), idx.data_ptr<int>(), n, sig2_sum.data_ptr<float>(), static_cast<long long>(seed));
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// ---- personalization mixing weight (/root/reference/utils/utils.py:605-617) -------------------------------
// out[0] += sum_i (wp_i - wg_i) * (alpha * gp_i + (1 - alpha) * gg_i): ONE pass over the four flat arenas instead of a
// dot product per parameter tensor plus two concatenations.
namespace gk {
__global__ void __launch_bounds__(256) alpha_dot_kernel(const float* __restrict__ wp, const float* __restrict__ wg,
                                                        const float* __restrict__ gp, const float* __restrict__ gg,
                                                        long long n, float alpha, double* __restrict__ out) {
  double acc = 0.0;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(wp)[i], b = reinterpret_cast<const float4*>(wg)[i];
    const float4 c = reinterpret_cast<const float4*>(gp)[i], d = reinterpret_cast<const float4*>(gg)[i];
    float s = (a.x - b.x) * (alpha * c.x + (1.f - alpha) * d.x);
    s += (a.y - b.y) * (alpha * c.y + (1.f - alpha) * d.y);
    s += (a.z - b.z) * (alpha * c.z + (1.f - alpha) * d.z);
    s += (a.w - b.w) * (alpha * c.w + (1.f - alpha) * d.w);
    acc += static_cast<double>(s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) acc += static_cast<double>((wp[i] - wg[i]) * (alpha * gp[i] + (1.f - alpha) * gg[i]));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}
}  // namespace gk

torch::Tensor alpha_dot(torch::Tensor wp, torch::Tensor wg, torch::Tensor gp, torch::Tensor gg, double alpha) {
  const int64_t n = wp.numel();
  for (const auto& t : {wp, wg, gp, gg})
    TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kFloat32 && t.numel() == n &&
                reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, "alpha_dot: four aligned fp32 CUDA vectors of one length");
  const c10::cuda::CUDAGuard guard(wp.device());
  auto out = torch::zeros({1}, wp.options().dtype(torch::kFloat64));
  const int blocks = static_cast<int>(std::min<int64_t>(148 * 8, std::max<int64_t>(1, (n / 4 + 255) / 256)));
  gk::alpha_dot_kernel<<<blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      wp.data_ptr<float>(), wg.data_ptr<float>(), gp.data_ptr<float>(), gg.data_ptr<float>(), n, static_cast<float>(alpha),
      out.data_ptr<double>());
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return out;
}

}  // namespace flute
