// Gradient quantization fused into the device engine's gather (SURVEY K14 -> K21; the reference:
// extensions/quantization/quant.py:9-100 — per tensor min / max, |g| quantile via a full sort, linspace + bucketize,
// threshold sparsification, all on materialised per-client gradients).
//
// Here the per-client pseudo-gradient d = wg - W[s] is never materialised.  For every (client slot s, tensor t):
//   1. radix select of the |d| order statistics that torch.quantile interpolates between: three histogram passes over the
//      IEEE bit pattern of |d| (11 + 11 + 10 bits), all S x T segments in one launch per pass; pass 0 also reduces
//      min / max of d.  Filter taps that the slot layout does not store (they are exactly zero in the reference
//      tensor) are accounted as `n_dead` extra zeros, so the statistics equal those of the full tensor.
//   2. one "next larger key" pass when the two neighbouring order statistics differ.
//   3. slot_quant_gather: acc_slot[j] += sum_s coef[s] * Q_{s,t}(d)   — binning to 2^bits levels on [lo, hi],
//      sparsification below the threshold, aggregation weight, in the same single pass as the plain fused gather.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace flute {
namespace qg {

constexpr int kT = 256;
constexpr int kBins = 2048;

struct SegState {            // per (slot, tensor)
  unsigned int prefix;       // key bits decided so far
  unsigned int k_rem;        // rank still to be located inside the current prefix bucket
  unsigned int below;        // elements strictly below the current prefix bucket (full multiset)
  unsigned int eq;           // multiplicity of the selected key (after pass 2)
  int lo_enc, hi_enc;        // min / max of d as order-preserving ints
  unsigned int next_key;     // smallest key above the selected one (0xFFFFFFFF = none)
  unsigned int pad;
};

__device__ __forceinline__ int enc(float f) {             // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// seg table: int64 [T, 3] = (offset in the slot row, live elements, elements of the full tensor)
template <int PASS>
__global__ void __launch_bounds__(kT) qsel_hist_kernel(const float* __restrict__ W, const float* __restrict__ wg, int64_t Pc,
                                                       const long long* __restrict__ segs, int T,
                                                       SegState* __restrict__ st, unsigned int* __restrict__ hist) {
  __shared__ unsigned int h[kBins];
  const int t = blockIdx.y, s = blockIdx.z;
  const long long off = segs[3 * t], n = segs[3 * t + 1];
  for (int i = threadIdx.x; i < kBins; i += kT) h[i] = 0;
  __syncthreads();
  const SegState cur = st[s * T + t];
  const float* w = W + static_cast<long long>(s) * Pc + off;
  const float* g = wg + off;
  float mn = INFINITY, mx = -INFINITY;
  unsigned int nk = 0xFFFFFFFFu;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * kT) {
    const float d = __ldg(g + i) - w[i];
    const unsigned int key = __float_as_uint(fabsf(d));
    if (PASS == 0) {
      mn = fminf(mn, d); mx = fmaxf(mx, d);
      atomicAdd(&h[key >> 21], 1u);
    } else if (PASS == 1) {
      if ((key >> 21) == cur.prefix) atomicAdd(&h[(key >> 10) & 0x7FFu], 1u);
    } else if (PASS == 2) {
      if ((key >> 10) == cur.prefix) atomicAdd(&h[key & 0x3FFu], 1u);
    } else {                                  // PASS 3: smallest key above the selected one
      if (key > cur.prefix) nk = min(nk, key);
    }
  }
  if (PASS == 3) {
    for (int o = 16; o > 0; o >>= 1) nk = min(nk, __shfl_xor_sync(0xffffffffu, nk, o));
    if ((threadIdx.x & 31) == 0 && nk != 0xFFFFFFFFu) atomicMin(&st[s * T + t].next_key, nk);
    return;
  }
  __syncthreads();
  unsigned int* gh = hist + (static_cast<long long>(s) * T + t) * kBins;
  for (int i = threadIdx.x; i < kBins; i += kT)
    if (h[i] != 0) atomicAdd(gh + i, h[i]);
  if (PASS == 0) {
    mn = warp_min(mn); mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
      atomicMin(&st[s * T + t].lo_enc, enc(mn));
      atomicMax(&st[s * T + t].hi_enc, enc(mx));
    }
  }
}

// one block per (tensor, slot): locate the bucket that holds rank k_rem, descend
template <int PASS>
__global__ void __launch_bounds__(kT) qsel_pick_kernel(const long long* __restrict__ segs, int T, SegState* __restrict__ st,
                                                       unsigned int* __restrict__ hist) {
  __shared__ unsigned int part[kT];
  __shared__ unsigned int chosen[3];
  const int t = blockIdx.x, s = blockIdx.y;
  SegState cur = st[s * T + t];
  unsigned int* gh = hist + (static_cast<long long>(s) * T + t) * kBins;
  // zero correction: the full tensor has (total - iterated) more zeros than the iterated slot range (elided dead taps);
  // it is NEGATIVE when the range contains layout padding that is not part of the tensor (the stem's 147 -> 160 columns).
  // Zeros sit in bin 0 of every level as long as the prefix is all zero.
  const long long n_dead = segs[3 * t + 2] - segs[3 * t + 1];
  const int nb = PASS == 2 ? 1024 : kBins, per = nb / kT;
  const bool zeros_here = n_dead != 0 && cur.prefix == 0;
  unsigned int loc[8];
  unsigned int sum = 0;
  for (int i = 0; i < per; ++i) {
    const int b = threadIdx.x * per + i;
    unsigned int c = gh[b];
    if (zeros_here && b == 0) c = static_cast<unsigned int>(max(0ll, static_cast<long long>(c) + n_dead));
    loc[i] = c;
    sum += c;
    gh[b] = 0;                                   // ready for the next pass
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int cum = 0;
    int th = 0;
    for (; th < kT - 1; ++th) {
      if (cum + part[th] > cur.k_rem) break;
      cum += part[th];
    }
    chosen[0] = th; chosen[1] = cum;
  }
  __syncthreads();
  if (threadIdx.x == chosen[0]) {
    unsigned int cum = chosen[1];
    int i = 0;
    for (; i < per - 1; ++i) {
      if (cum + loc[i] > cur.k_rem) break;
      cum += loc[i];
    }
    const unsigned int bin = threadIdx.x * per + i;
    cur.below += cum;
    cur.k_rem -= cum;
    cur.prefix = PASS == 0 ? bin : PASS == 1 ? ((cur.prefix << 11) | bin) : ((cur.prefix << 10) | bin);
    if (PASS == 2) cur.eq = loc[i];
    st[s * T + t] = cur;
  }
}

__global__ void qsel_init_kernel(const long long* __restrict__ segs, int T, int S, SegState* __restrict__ st, float q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * T) return;
  const int t = i % T;
  const long long n = segs[3 * t + 2];
  SegState z;
  const double pos = static_cast<double>(q) * static_cast<double>(n - 1);
  z.prefix = 0;
  z.k_rem = static_cast<unsigned int>(n > 0 ? floor(pos) : 0);
  z.below = 0; z.eq = 0;
  z.lo_enc = 0x7FFFFFFF; z.hi_enc = static_cast<int>(0x80000000);
  z.next_key = 0xFFFFFFFFu; z.pad = 0;
  st[i] = z;
}

// params[s][t] = (lo, width, thresh, hi) of d in tensor t of slot s  (torch.quantile's linear interpolation)
__global__ void qsel_finish_kernel(const long long* __restrict__ segs, int T, int S, const SegState* __restrict__ st, float q,
                                   int levels, float4* __restrict__ params) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * T) return;
  const int t = i % T;
  const long long n = segs[3 * t + 2], n_dead = n - segs[3 * t + 1];
  const SegState z = st[i];
  float lo = dec(z.lo_enc), hi = dec(z.hi_enc);
  if (z.lo_enc == 0x7FFFFFFF) { lo = 0.f; hi = 0.f; }
  if (n_dead > 0) { lo = fminf(lo, 0.f); hi = fmaxf(hi, 0.f); }
  const double pos = static_cast<double>(q) * static_cast<double>(n - 1);
  const unsigned int k0 = static_cast<unsigned int>(floor(pos));
  const float frac = static_cast<float>(pos - floor(pos));
  const float v0 = __uint_as_float(z.prefix);
  // is rank k0 + 1 still the same value?  (below + eq = number of elements <= v0)
  float v1 = v0;
  if (static_cast<long long>(z.below) + z.eq <= static_cast<long long>(k0) + 1 && z.next_key != 0xFFFFFFFFu)
    v1 = __uint_as_float(z.next_key);
  const float thresh = v0 + frac * (v1 - v0);
  const float width = (hi - lo) / static_cast<float>(levels - 1);
  params[i] = make_float4(lo, width, thresh, hi);
}

__device__ __forceinline__ float quant_one(float d, float4 p, float lmax) {
  if (!(fabsf(d) > p.z)) return 0.f;
  if (!(p.y > 0.f)) return p.x;
  const float idx = fminf(fmaxf(ceilf((d - p.x) / p.y - 0.5f), 0.f), lmax);
  return fmaf(idx, p.y, p.x);
}

// acc_slot[j] += sum_s coef[s] * Q_{s, seg(j)}(wg[j] - W[s][j]);  seg_of_blk[j / 32] = tensor of the 32-element block
__global__ void __launch_bounds__(kT)
quant_gather_kernel(float* __restrict__ acc_slot, const float* __restrict__ W, const float* __restrict__ wg, int64_t Pc, int S,
                    int T, const float* __restrict__ coef, const float4* __restrict__ params,
                    const short* __restrict__ seg_of_blk, int levels) {
  const float lmax = static_cast<float>(levels - 1);
  const int64_t n4 = Pc >> 2, stride = static_cast<int64_t>(gridDim.x) * kT;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kT + threadIdx.x; i < n4; i += stride) {
    const int seg = seg_of_blk[i >> 3];
    if (seg < 0) continue;
    const float4 g = __ldg(reinterpret_cast<const float4*>(wg) + i);
    float4 r = reinterpret_cast<float4*>(acc_slot)[i];
    for (int s = 0; s < S; ++s) {
      const float c = __ldg(coef + s);
      if (c == 0.f) continue;
      const float4 p = __ldg(params + s * T + seg);
      const float4 w = ld_stream(reinterpret_cast<const float4*>(W + static_cast<int64_t>(s) * Pc) + i);
      r.x = fmaf(c, quant_one(g.x - w.x, p, lmax), r.x); r.y = fmaf(c, quant_one(g.y - w.y, p, lmax), r.y);
      r.z = fmaf(c, quant_one(g.z - w.z, p, lmax), r.z); r.w = fmaf(c, quant_one(g.w - w.w, p, lmax), r.w);
    }
    reinterpret_cast<float4*>(acc_slot)[i] = r;
  }
}

}  // namespace qg

// Returns params [S, T, 4] = (lo, width, thresh, hi) per (slot, tensor) of the pseudo-gradient wg - W[s].
torch::Tensor slot_quant_stats(torch::Tensor W, torch::Tensor wg, torch::Tensor segs, double q, int64_t bits) {
  using namespace qg;
  TORCH_CHECK(W.is_cuda() && W.scalar_type() == torch::kFloat32 && W.dim() == 2 && W.is_contiguous() && wg.is_cuda() &&
              wg.scalar_type() == torch::kFloat32 && wg.numel() == W.size(1));
  TORCH_CHECK(segs.is_cuda() && segs.scalar_type() == torch::kInt64 && segs.dim() == 2 && segs.size(1) == 3 && segs.is_contiguous());
  const int S = static_cast<int>(W.size(0)), T = static_cast<int>(segs.size(0));
  const int64_t Pc = W.size(1);
  const c10::cuda::CUDAGuard guard(W.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto st = torch::empty({S * T * static_cast<int64_t>(sizeof(SegState))}, W.options().dtype(torch::kUInt8));
  auto hist = torch::zeros({static_cast<int64_t>(S) * T * kBins}, W.options().dtype(torch::kInt32));
  auto params = torch::empty({S, T, 4}, W.options());
  auto* stp = reinterpret_cast<SegState*>(st.data_ptr());
  auto* hp = reinterpret_cast<unsigned int*>(hist.data_ptr<int>());
  const auto* sg = reinterpret_cast<const long long*>(segs.data_ptr<int64_t>());
  const int nst = (S * T + 127) / 128;
  qsel_init_kernel<<<nst, 128, 0, stream>>>(sg, T, S, stp, static_cast<float>(q));
  const dim3 grid(8, T, S), pick(T, S);
  qsel_hist_kernel<0><<<grid, kT, 0, stream>>>(W.data_ptr<float>(), wg.data_ptr<float>(), Pc, sg, T, stp, hp);
  qsel_pick_kernel<0><<<pick, kT, 0, stream>>>(sg, T, stp, hp);
  qsel_hist_kernel<1><<<grid, kT, 0, stream>>>(W.data_ptr<float>(), wg.data_ptr<float>(), Pc, sg, T, stp, hp);
  qsel_pick_kernel<1><<<pick, kT, 0, stream>>>(sg, T, stp, hp);
  qsel_hist_kernel<2><<<grid, kT, 0, stream>>>(W.data_ptr<float>(), wg.data_ptr<float>(), Pc, sg, T, stp, hp);
  qsel_pick_kernel<2><<<pick, kT, 0, stream>>>(sg, T, stp, hp);
  qsel_hist_kernel<3><<<grid, kT, 0, stream>>>(W.data_ptr<float>(), wg.data_ptr<float>(), Pc, sg, T, stp, hp);
  qsel_finish_kernel<<<nst, 128, 0, stream>>>(sg, T, S, stp, static_cast<float>(q), 1 << static_cast<int>(bits),
                                              reinterpret_cast<float4*>(params.data_ptr<float>()));
  FLUTE_CUDA_CHECK(cudaGetLastError());
  return params;
}

void slot_quant_gather(torch::Tensor acc_slot, torch::Tensor W, torch::Tensor wg, torch::Tensor coef, torch::Tensor params,
                       torch::Tensor seg_of_blk, int64_t bits) {
  using namespace qg;
  const int S = static_cast<int>(W.size(0)), T = static_cast<int>(params.size(1));
  const int64_t Pc = W.size(1);
  TORCH_CHECK(acc_slot.numel() == Pc && wg.numel() == Pc && coef.numel() == S && params.size(0) == S && Pc % 32 == 0 &&
              seg_of_blk.scalar_type() == torch::kInt16 && seg_of_blk.numel() == Pc / 32);
  const c10::cuda::CUDAGuard guard(W.device());
  const int64_t n4 = Pc >> 2;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n4 + kT - 1) / kT, 148 * 4)));
  quant_gather_kernel<<<grid, kT, 0, at::cuda::getCurrentCUDAStream()>>>(
      acc_slot.data_ptr<float>(), W.data_ptr<float>(), wg.data_ptr<float>(), Pc, S, T, coef.data_ptr<float>(),
      reinterpret_cast<const float4*>(params.data_ptr<float>()), seg_of_blk.data_ptr<int16_t>(), 1 << static_cast<int>(bits));
  FLUTE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flute
