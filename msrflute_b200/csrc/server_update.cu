// Fused gather -> reduce -> normalise -> DP noise -> clip -> server optimizer -> broadcast.
//
// The reference does this with (W-1)*K*(1+3n) NCCL point-to-point messages staged through the CPU, a per-tensor
// `p.grad += g` loop, `p.grad /= weight_sum`, a flat-copy Gaussian add, and a torch.optim step, then re-sends the
// model to every worker the same way at the start of the next round (core/federated.py:112-188,330-334;
// strategies/fedavg.py:119-166; extensions/privacy/__init__.py:128-151; core/trainer.py:127-137).
//
// Here every rank keeps its weighted pseudo-gradient sum in a flat fp32 accumulator that is *peer-mapped* on the
// server GPU (CUDA IPC / symmetric memory over NVLink-5).  One kernel on the server:
//
//     g[i]  = (sum_r acc_r[i]) / sum_w              P2P 16-byte loads from every rank's accumulator
//     g[i] += sigma * N(0,1)                        Philox4x32-10 keyed by (seed, i/4): GPU-count invariant
//     w[i]  = OPT(w[i], g[i], state[i])             SGD(momentum) | Adam | AdamW | Adamax   (LAMB/LARS: 2 extra passes)
//     w_r[i] = w[i]  for every rank r               P2P 16-byte stores: the next round's broadcast, fused
//
// so the transfer overlaps the math element-tile by element-tile and no NCCL call is on the path.  With server-side
// clipping / norm logging / LAMB / LARS a grid-wide reduction is needed between "g" and "update": the kernel then runs
// as phase R (reduce+noise, emits g and per-block sum-of-squares) and phase U (update+broadcast).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {

constexpr int kT = 256;
enum Opt { SGD = 0, ADAM = 1, ADAMW = 2, ADAMAX = 3, LAMB = 4, LARS = 5 };

struct OptParams {
  int kind;
  int step;           // 1-based step index
  float lr, b1, b2, eps, wd, mom, damp;
  int nesterov, correct_bias;
  float bc1, bc2;     // 1 - b1^t, 1 - b2^t (host-computed in double)
};

static inline int grid_for(int64_t n4) {
  int64_t want = (n4 + kT * 2 - 1) / (kT * 2);
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 4)));
}

__device__ __forceinline__ float4 gather_sum(const PtrList& accs, int64_t i) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int r = 0; r < accs.n; ++r) {
    const float4 v = ld_coherent(reinterpret_cast<const float4*>(accs.p[r]) + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  return s;
}

__device__ __forceinline__ float opt_elem(const OptParams& o, float w, float g, float& m, float& v) {
  switch (o.kind) {
    case SGD: {
      float d = fmaf(o.wd, w, g);
      if (o.mom != 0.f) {
        m = (o.step == 1) ? d : fmaf(o.mom, m, (1.f - o.damp) * d);
        d = o.nesterov ? fmaf(o.mom, m, d) : m;
      }
      return fmaf(-o.lr, d, w);
    }
    case ADAM: {
      const float d = fmaf(o.wd, w, g);
      m = fmaf(o.b1, m, (1.f - o.b1) * d);
      v = fmaf(o.b2, v, (1.f - o.b2) * d * d);
      const float denom = sqrtf(v) / sqrtf(o.bc2) + o.eps;
      return w - (o.lr / o.bc1) * (m / denom);
    }
    case ADAMW: {   // utils/optimizers AdamW: optional bias correction, decoupled decay applied after the step
      m = fmaf(o.b1, m, (1.f - o.b1) * g);
      v = fmaf(o.b2, v, (1.f - o.b2) * g * g);
      const float step = o.correct_bias ? o.lr * sqrtf(o.bc2) / o.bc1 : o.lr;
      float wn = w - step * (m / (sqrtf(v) + o.eps));
      if (o.wd > 0.f) wn *= (1.f - o.lr * o.wd);
      return wn;
    }
    case ADAMAX: {
      const float d = fmaf(o.wd, w, g);
      m = fmaf(o.b1, m, (1.f - o.b1) * d);
      v = fmaxf(o.b2 * v, fabsf(d) + o.eps);
      return w - (o.lr / o.bc1) * (m / v);
    }
    default:
      return w;
  }
}

// Single-pass variant: reduce + noise + optimizer + broadcast.  Used when nothing needs the global norm of g.
__global__ void __launch_bounds__(kT)
fused_reduce_update_bcast_kernel(float* __restrict__ w, PtrList accs, const float* __restrict__ wsum,
                                 float* __restrict__ m, float* __restrict__ v, float* __restrict__ g_out, PtrList bcast,
                                 int64_t P, OptParams o, float noise_scale, uint64_t seed, int zero_local_acc) {
  const float inv = 1.f / *wsum;
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kT;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n4; i += stride) {
    float4 g = gather_sum(accs, i);
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    if (noise_scale > 0.f) {
      const float4 z = philox_normal4(seed, static_cast<uint64_t>(i));
      g.x = fmaf(noise_scale, z.x, g.x); g.y = fmaf(noise_scale, z.y, g.y);
      g.z = fmaf(noise_scale, z.z, g.z); g.w = fmaf(noise_scale, z.w, g.w);
    }
    float4 ww = ld_na(reinterpret_cast<const float4*>(w) + i);
    float4 mm = m ? ld_na(reinterpret_cast<const float4*>(m) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 vv = v ? ld_na(reinterpret_cast<const float4*>(v) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    ww.x = opt_elem(o, ww.x, g.x, mm.x, vv.x); ww.y = opt_elem(o, ww.y, g.y, mm.y, vv.y);
    ww.z = opt_elem(o, ww.z, g.z, mm.z, vv.z); ww.w = opt_elem(o, ww.w, g.w, mm.w, vv.w);
    st_stream(reinterpret_cast<float4*>(w) + i, ww);
    if (m) st_stream(reinterpret_cast<float4*>(m) + i, mm);
    if (v) st_stream(reinterpret_cast<float4*>(v) + i, vv);
    if (g_out) st_stream(reinterpret_cast<float4*>(g_out) + i, g);
    for (int r = 0; r < bcast.n; ++r) st_stream(reinterpret_cast<float4*>(bcast.p[r]) + i, ww);   // NVLink P2P stores
    if (zero_local_acc) st_stream(reinterpret_cast<float4*>(accs.p[0]) + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// Phase R: g = sum/ wsum (+noise); partial[b] = (sum g^2 before noise, sum g^2 after noise)
__global__ void __launch_bounds__(kT)
reduce_noise_kernel(PtrList accs, const float* __restrict__ wsum, float* __restrict__ g_out, int64_t P,
                    float noise_scale, uint64_t seed, float2* __restrict__ partial, int zero_local_acc) {
  const float inv = 1.f / *wsum;
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kT;
  float s_pre = 0.f, s_post = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n4; i += stride) {
    float4 g = gather_sum(accs, i);
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    s_pre += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    if (noise_scale > 0.f) {
      const float4 z = philox_normal4(seed, static_cast<uint64_t>(i));
      g.x = fmaf(noise_scale, z.x, g.x); g.y = fmaf(noise_scale, z.y, g.y);
      g.z = fmaf(noise_scale, z.z, g.z); g.w = fmaf(noise_scale, z.w, g.w);
    }
    s_post += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    st_stream(reinterpret_cast<float4*>(g_out) + i, g);
    if (zero_local_acc) st_stream(reinterpret_cast<float4*>(accs.p[0]) + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  const float2 r = block_sum2(s_pre, s_post);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// Phase U (element-wise optimizers): clip by the post-noise norm, update, broadcast.
__global__ void __launch_bounds__(kT)
update_bcast_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    PtrList bcast, int64_t P, OptParams o, const float2* __restrict__ partial, int nb_partial,
                    float max_norm, float* __restrict__ stats_out) {
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < nb_partial; i += blockDim.x) { a += partial[i].x; b += partial[i].y; }
  const float2 tot = block_sum2(a, b);
  const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (sqrtf(tot.y) + 1e-6f)) : 1.f;
  if (stats_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    stats_out[0] = sqrtf(tot.x);
    stats_out[1] = sqrtf(tot.y);
  }
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kT;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n4; i += stride) {
    float4 gg = ld_na(reinterpret_cast<const float4*>(g) + i);
    gg.x *= coef; gg.y *= coef; gg.z *= coef; gg.w *= coef;
    float4 ww = ld_na(reinterpret_cast<const float4*>(w) + i);
    float4 mm = m ? ld_na(reinterpret_cast<const float4*>(m) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 vv = v ? ld_na(reinterpret_cast<const float4*>(v) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    ww.x = opt_elem(o, ww.x, gg.x, mm.x, vv.x); ww.y = opt_elem(o, ww.y, gg.y, mm.y, vv.y);
    ww.z = opt_elem(o, ww.z, gg.z, mm.z, vv.z); ww.w = opt_elem(o, ww.w, gg.w, mm.w, vv.w);
    st_stream(reinterpret_cast<float4*>(w) + i, ww);
    if (m) st_stream(reinterpret_cast<float4*>(m) + i, mm);
    if (v) st_stream(reinterpret_cast<float4*>(v) + i, vv);
    if (coef != 1.f) st_stream(reinterpret_cast<float4*>(g) + i, gg);
    for (int r = 0; r < bcast.n; ++r) st_stream(reinterpret_cast<float4*>(bcast.p[r]) + i, ww);
  }
}

// ---- LAMB / LARS: per-tensor trust ratios.  Pass 1 builds the update direction u (in g) and per-segment norms.
__global__ void __launch_bounds__(kT)
layerwise_direction_kernel(const float* __restrict__ w, float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ v, const int64_t* __restrict__ segs, OptParams o, float coef_src_norm,
                           const float2* __restrict__ partial, int nb_partial, float max_norm,
                           float* __restrict__ seg_norms) {
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < nb_partial; i += blockDim.x) { a += partial[i].x; b += partial[i].y; }
  const float2 tot = block_sum2(a, b);
  const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (sqrtf(tot.y) + 1e-6f)) : 1.f;
  const int seg = blockIdx.y;
  const int64_t off = segs[2 * seg], n = segs[2 * seg + 1];
  float sw = 0.f, su = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kT) {
    const int64_t k = off + i;
    const float ww = w[k];
    const float gg = g[k] * coef;
    float u;
    if (o.kind == LAMB) {
      const float mm = fmaf(o.b1, m[k], (1.f - o.b1) * gg);
      const float vv = fmaf(o.b2, v[k], (1.f - o.b2) * gg * gg);
      m[k] = mm; v[k] = vv;
      u = mm / (sqrtf(vv) + o.eps) + o.wd * ww;
    } else {  // LarsSGD (arXiv:1904.00962 alg. 1): weight decay is a no-op in the reference implementation
      u = gg;
      if (o.mom != 0.f) {
        const float mm = (o.step == 1) ? gg : fmaf(o.mom, m[k], (1.f - o.mom) * gg);
        m[k] = mm;
        u = o.nesterov ? fmaf(o.mom, mm, gg) : mm;
      }
    }
    g[k] = u;
    sw = fmaf(ww, ww, sw);
    su = fmaf(u, u, su);
  }
  const float2 r = block_sum2(sw, su);
  if (threadIdx.x == 0 && (r.x != 0.f || r.y != 0.f)) {
    atomicAdd(seg_norms + 2 * seg, r.x);
    atomicAdd(seg_norms + 2 * seg + 1, r.y);
  }
}

__global__ void __launch_bounds__(kT)
layerwise_apply_kernel(float* __restrict__ w, const float* __restrict__ u, const int64_t* __restrict__ segs,
                       OptParams o, const float* __restrict__ seg_norms, PtrList bcast) {
  const int seg = blockIdx.y;
  const int64_t off = segs[2 * seg], n = segs[2 * seg + 1];
  const float wn = sqrtf(seg_norms[2 * seg]), un = sqrtf(seg_norms[2 * seg + 1]);
  float scale;
  if (o.kind == LAMB) {
    const float wc = fminf(wn, 10.f);
    scale = o.lr * ((wc == 0.f || un == 0.f) ? 1.f : wc / un);
  } else {
    scale = fminf(fmaxf(o.lr * wn / (un + 1e-8f), 0.f), 10.f);
  }
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kT) {
    const int64_t k = off + i;
    const float wnew = fmaf(-scale, u[k], w[k]);
    w[k] = wnew;
    for (int r = 0; r < bcast.n; ++r) bcast.p[r][k] = wnew;
  }
}

static PtrList to_list(const std::vector<torch::Tensor>& ts, int64_t P, const float* skip = nullptr) {
  PtrList l;
  l.n = 0;
  TORCH_CHECK(ts.size() <= static_cast<size_t>(kMaxPeers), "too many peers");
  for (const auto& t : ts) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.numel() == P, "peer buffer shape/dtype mismatch");
    if (skip != nullptr && t.data_ptr<float>() == skip) continue;
    l.p[l.n++] = t.data_ptr<float>();
  }
  return l;
}

void server_update(torch::Tensor w, std::vector<torch::Tensor> accs, torch::Tensor weight_sum,
                   c10::optional<torch::Tensor> m, c10::optional<torch::Tensor> v, c10::optional<torch::Tensor> grad_out,
                   c10::optional<torch::Tensor> segments, std::vector<torch::Tensor> bcast,
                   c10::optional<torch::Tensor> stats_out, int64_t kind, int64_t step, double lr, double b1, double b2,
                   double eps, double wd, double mom, double damp, bool nesterov, bool correct_bias, double noise_scale,
                   int64_t seed, double max_grad_norm, bool zero_accs) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == torch::kFloat32 && w.is_contiguous() && w.numel() % 4 == 0);
  TORCH_CHECK(!accs.empty(), "need at least the local accumulator");
  const int64_t P = w.numel();
  const c10::cuda::CUDAGuard guard(w.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  PtrList A = to_list(accs, P);
  PtrList B = to_list(bcast, P, w.data_ptr<float>());
  OptParams o;
  o.kind = static_cast<int>(kind); o.step = static_cast<int>(step);
  o.lr = static_cast<float>(lr); o.b1 = static_cast<float>(b1); o.b2 = static_cast<float>(b2);
  o.eps = static_cast<float>(eps); o.wd = static_cast<float>(wd); o.mom = static_cast<float>(mom);
  o.damp = static_cast<float>(damp); o.nesterov = nesterov; o.correct_bias = correct_bias;
  o.bc1 = static_cast<float>(1.0 - std::pow(b1, static_cast<double>(step)));
  o.bc2 = static_cast<float>(1.0 - std::pow(b2, static_cast<double>(step)));
  float* mp = m.has_value() ? m->data_ptr<float>() : nullptr;
  float* vp = v.has_value() ? v->data_ptr<float>() : nullptr;
  const bool layerwise = (kind == LAMB || kind == LARS);
  const bool need_norm = layerwise || max_grad_norm > 0.0 || stats_out.has_value();
  const int nb = grid_for(P >> 2);
  if (!need_norm) {
    fused_reduce_update_bcast_kernel<<<nb, kT, 0, stream>>>(
        w.data_ptr<float>(), A, weight_sum.data_ptr<float>(), mp, vp,
        grad_out.has_value() ? grad_out->data_ptr<float>() : nullptr, B, P, o, static_cast<float>(noise_scale),
        static_cast<uint64_t>(seed), zero_accs ? 1 : 0);
    FLUTE_CUDA_CHECK(cudaGetLastError());
    return;
  }
  torch::Tensor g = grad_out.has_value() ? *grad_out : torch::empty({P}, w.options());
  auto partial = torch::empty({nb, 2}, w.options());
  reduce_noise_kernel<<<nb, kT, 0, stream>>>(A, weight_sum.data_ptr<float>(), g.data_ptr<float>(), P,
                                             static_cast<float>(noise_scale), static_cast<uint64_t>(seed),
                                             reinterpret_cast<float2*>(partial.data_ptr<float>()), zero_accs ? 1 : 0);
  if (!layerwise) {
    update_bcast_kernel<<<nb, kT, 0, stream>>>(w.data_ptr<float>(), g.data_ptr<float>(), mp, vp, B, P, o,
                                               reinterpret_cast<const float2*>(partial.data_ptr<float>()), nb,
                                               static_cast<float>(max_grad_norm),
                                               stats_out.has_value() ? stats_out->data_ptr<float>() : nullptr);
  } else {
    TORCH_CHECK(segments.has_value(), "LAMB/LARS need the per-tensor segment table");
    auto segs = segments->to(torch::kInt64).contiguous();
    const int n_seg = static_cast<int>(segs.size(0));
    auto seg_norms = torch::zeros({n_seg, 2}, w.options());
    dim3 grid(64, n_seg);
    layerwise_direction_kernel<<<grid, kT, 0, stream>>>(
        w.data_ptr<float>(), g.data_ptr<float>(), mp, vp, segs.data_ptr<int64_t>(), o, 0.f,
        reinterpret_cast<const float2*>(partial.data_ptr<float>()), nb, static_cast<float>(max_grad_norm),
        seg_norms.data_ptr<float>());
    layerwise_apply_kernel<<<grid, kT, 0, stream>>>(w.data_ptr<float>(), g.data_ptr<float>(), segs.data_ptr<int64_t>(), o,
                                                    seg_norms.data_ptr<float>(), B);
    if (stats_out.has_value()) {
      auto tot = partial.sum(0).sqrt();
      stats_out->narrow(0, 0, 2).copy_(tot);
    }
  }
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ sharded update
// Transport v2: EVERY rank runs this kernel on its own 1/N slice of the arena —
//     g[i]   = sum_r acc_r[i] / sum_r wsum_r        P2P loads from every rank's accumulator slice, or ONE
//                                                   multimem.ld_reduce through the NVSwitch multicast object (NVLS)
//     w[i]   = OPT(w[i], g[i], state[i])            optimizer state exists only for the rank's slice (ZeRO-1)
//     w_r[i] = w[i] for every rank r                P2P stores, or ONE multimem.st
// so a GPU moves (N-1)/N * P floats each way instead of (N-1) * P through the server's links, sum(weights) travels as a
// symmetric scalar (no NCCL all-reduce), and the Philox counter is the GLOBAL quad index (noise is shard-invariant).
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_ptr) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc_ptr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc_ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kT)
sharded_update_kernel(float* __restrict__ w_local, PtrList accs, PtrList wsums, float* __restrict__ m, float* __restrict__ v,
                      PtrList bcast, PtrList m_mirror, PtrList v_mirror, const float* acc_mc, float* w_mc, int64_t q_lo,
                      int64_t q_hi, OptParams o, float noise_scale, uint64_t seed) {
  float tot = 0.f;
  for (int r = 0; r < wsums.n; ++r) {
    float t;
    asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(t) : "l"(wsums.p[r]) : "memory");
    tot += t;
  }
  const float inv = 1.f / tot;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kT;
  for (int64_t i = q_lo + static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < q_hi; i += stride) {
    float4 g = acc_mc != nullptr ? multimem_ld_reduce_add(acc_mc + 4 * i) : gather_sum(accs, i);
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    if (noise_scale > 0.f) {
      const float4 z = philox_normal4(seed, static_cast<uint64_t>(i));
      g.x = fmaf(noise_scale, z.x, g.x); g.y = fmaf(noise_scale, z.y, g.y);
      g.z = fmaf(noise_scale, z.z, g.z); g.w = fmaf(noise_scale, z.w, g.w);
    }
    float4 ww = ld_na(reinterpret_cast<const float4*>(w_local) + i);
    float4 mm = m ? ld_na(reinterpret_cast<const float4*>(m) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 vv = v ? ld_na(reinterpret_cast<const float4*>(v) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    ww.x = opt_elem(o, ww.x, g.x, mm.x, vv.x); ww.y = opt_elem(o, ww.y, g.y, mm.y, vv.y);
    ww.z = opt_elem(o, ww.z, g.z, mm.z, vv.z); ww.w = opt_elem(o, ww.w, g.w, mm.w, vv.w);
    if (m) st_stream(reinterpret_cast<float4*>(m) + i, mm);
    if (v) st_stream(reinterpret_cast<float4*>(v) + i, vv);
    for (int r = 0; r < m_mirror.n; ++r) st_stream(reinterpret_cast<float4*>(m_mirror.p[r]) + i, mm);   // checkpoint copy on the server
    for (int r = 0; r < v_mirror.n; ++r) st_stream(reinterpret_cast<float4*>(v_mirror.p[r]) + i, vv);
    if (w_mc != nullptr) {
      multimem_st(w_mc + 4 * i, ww);
    } else {
      for (int r = 0; r < bcast.n; ++r) st_stream(reinterpret_cast<float4*>(bcast.p[r]) + i, ww);
    }
  }
}

static PtrList to_list_any(const std::vector<torch::Tensor>& ts) {
  PtrList l;
  l.n = 0;
  TORCH_CHECK(ts.size() <= static_cast<size_t>(kMaxPeers), "too many peers");
  for (const auto& t : ts) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32, "fp32 CUDA peer buffer expected");
    l.p[l.n++] = t.data_ptr<float>();
  }
  return l;
}

// ws / accs / wsums: every rank's weight buffer / accumulator / sum(weights) scalar (peer mapped; index == rank).
// m, v: this rank's optimizer state (full-size buffers, only the slice is touched); m_mirror / v_mirror: buffers that
// additionally receive the slice (the server's state arenas, for checkpoints).  acc_mc / w_mc: multicast addresses of
// the accumulator / weight buffers (0 = P2P loops).  Returns nothing; the caller brackets it with device barriers.
void sharded_server_update(std::vector<torch::Tensor> ws, std::vector<torch::Tensor> accs, std::vector<torch::Tensor> wsums,
                           c10::optional<torch::Tensor> m, c10::optional<torch::Tensor> v,
                           std::vector<torch::Tensor> m_mirror, std::vector<torch::Tensor> v_mirror, int64_t acc_mc,
                           int64_t w_mc, int64_t rank, int64_t kind, int64_t step, double lr, double b1, double b2,
                           double eps, double wd, double mom, double damp, bool nesterov, bool correct_bias,
                           double noise_scale, int64_t seed) {
  const int N = static_cast<int>(ws.size());
  TORCH_CHECK(N >= 1 && static_cast<int>(accs.size()) == N && static_cast<int>(wsums.size()) == N && rank >= 0 && rank < N);
  TORCH_CHECK(kind != LAMB && kind != LARS, "layer-wise optimizers are not sharded");
  const int64_t P = ws[rank].numel();
  TORCH_CHECK(P % 4 == 0);
  const c10::cuda::CUDAGuard guard(ws[rank].device());
  const int64_t n4 = P >> 2;
  int64_t per = (n4 + N - 1) / N;
  per = (per + 31) / 32 * 32;                                   // slices start on 512-byte boundaries
  const int64_t lo = std::min<int64_t>(n4, rank * per), hi = std::min<int64_t>(n4, lo + per);
  if (hi <= lo) return;
  OptParams o;
  o.kind = static_cast<int>(kind); o.step = static_cast<int>(step);
  o.lr = static_cast<float>(lr); o.b1 = static_cast<float>(b1); o.b2 = static_cast<float>(b2);
  o.eps = static_cast<float>(eps); o.wd = static_cast<float>(wd); o.mom = static_cast<float>(mom);
  o.damp = static_cast<float>(damp); o.nesterov = nesterov; o.correct_bias = correct_bias;
  o.bc1 = static_cast<float>(1.0 - std::pow(b1, static_cast<double>(step)));
  o.bc2 = static_cast<float>(1.0 - std::pow(b2, static_cast<double>(step)));
  PtrList A = to_list(accs, P), B = to_list(ws, P), WS = to_list_any(wsums);
  PtrList MM = to_list_any(m_mirror), VM = to_list_any(v_mirror);
  sharded_update_kernel<<<grid_for(hi - lo), kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      ws[rank].data_ptr<float>(), A, WS, m.has_value() ? m->data_ptr<float>() : nullptr,
      v.has_value() ? v->data_ptr<float>() : nullptr, B, MM, VM, reinterpret_cast<const float*>(acc_mc),
      reinterpret_cast<float*>(w_mc), lo, hi, o, static_cast<float>(noise_scale), static_cast<uint64_t>(seed));
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

// Plain P2P copy kernel: dst_r[i] = src[i] for every peer (explicit broadcast at round 0 / after checkpoint loads).
__global__ void __launch_bounds__(kT) bcast_copy_kernel(const float* __restrict__ src, PtrList dst, int64_t P) {
  const int64_t n4 = P >> 2, stride = static_cast<int64_t>(gridDim.x) * kT;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n4; i += stride) {
    const float4 x = ld_stream(reinterpret_cast<const float4*>(src) + i);
    for (int r = 0; r < dst.n; ++r) st_stream(reinterpret_cast<float4*>(dst.p[r]) + i, x);
  }
}

void p2p_broadcast(torch::Tensor src, std::vector<torch::Tensor> dsts) {
  const int64_t P = src.numel();
  TORCH_CHECK(src.is_cuda() && src.scalar_type() == torch::kFloat32 && P % 4 == 0);
  const c10::cuda::CUDAGuard guard(src.device());
  PtrList D = to_list(dsts, P, src.data_ptr<float>());
  if (D.n == 0) return;
  bcast_copy_kernel<<<grid_for(P >> 2), kT, 0, at::cuda::getCurrentCUDAStream()>>>(src.data_ptr<float>(), D, P);
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flute
