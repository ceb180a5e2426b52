// Persistent LSTM layer (replaces cuDNN's RNN at
// /root/reference/experiments/nlp_rnn_fedshakespeare/model.py:18-23 and experiments/ecg_cnn/model.py:99-106).
//
// One launch runs ALL T time steps of one layer.  A thread-block cluster of 8 CTAs owns a chunk of 4 batch rows; CTA j
// owns hidden units [j*U, (j+1)*U), U = H / 8:
//
//   forward : its 4U rows of W_hh ([i|f|g|o] x U units, all H columns) live in SHARED MEMORY for the whole sequence
//             (128 KB for H = 256) — the recurrent weights are read from HBM exactly once per launch instead of once per
//             time step.  Per step every CTA computes its 4U gate pre-activations (input projection gx precomputed by
//             one GEMM for all steps), applies the cell update to its U units and writes the new h values straight into
//             the h buffers of all 8 CTAs (distributed shared memory, st.shared::cluster), then one cluster barrier.
//   backward: the mirror image — its U COLUMNS of W_hh (all 4H rows) are resident; per step it turns (dh, dc) of its
//             units into gate gradients, broadcasts those 4U x 4 values to every CTA of the cluster, barrier, and
//             multiplies the full gate-gradient vector with its weight columns to get next step's dh for its units.
//
// The time loop never leaves the SMs: no per-step launches, no per-step weight traffic, the only global traffic per
// step is gx / the saved activations.  dW_hh, dW_ih, dx are plain GEMMs over the saved [B*T, .] tensors (torch.matmul).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <vector>
#include "common.cuh"

namespace flute {
namespace lstm {

constexpr int NC = 8;            // CTAs per cluster
constexpr int BP = 4;            // batch rows per cluster (one float4)
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct FwdP {
  const float* gx;      // [B, T, 4H]  x W_ih^T + b_ih + b_hh
  const float* whh;     // [4H, H]
  const float* h0;      // [B, H] or null
  const float* c0;      // [B, H] or null
  float* hs;            // [B, T, H]
  float* gates;         // [B, T, 4H] activated i, f, g, o (saved for backward)
  float* cs;            // [B, T, H]  cell states
  int B, T, H;
};

// shared memory: Wt [H][R] (k-major slice, R = 4U rows) | hbuf [2][H][BP] | part [KS][R][BP] | gact [R][BP]
__global__ void __launch_bounds__(kThreads, 1) lstm_fwd_kernel(const FwdP p) {
  extern __shared__ float sm[];
  const int H = p.H, U = H / NC, R = 4 * U, KS = kThreads / R;      // KS k-segments of H / KS columns each
  float* Wt = sm;
  float* hbuf = Wt + static_cast<size_t>(H) * R;
  float* part = hbuf + 2 * H * BP;
  float* gact = part + KS * R * BP;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int b0 = blockIdx.y * BP;
  // resident weight slice: local row r = gate * U + u  <->  global row gate * H + rank * U + u
  for (int i = tid; i < R * H; i += kThreads) {
    const int r = i / H, k = i - r * H;
    const int gate = r / U, u = r - gate * U;
    Wt[k * R + r] = __ldg(p.whh + static_cast<long long>(gate * H + rank * U + u) * H + k);
  }
  for (int i = tid; i < H * BP; i += kThreads) {
    const int k = i / BP, b = i - k * BP;
    hbuf[i] = (p.h0 != nullptr && b0 + b < p.B) ? p.h0[static_cast<long long>(b0 + b) * H + k] : 0.f;
  }
  // cell state of (unit u, batch b) lives in thread tid = u * BP + b (tid < U * BP)
  float c_state = 0.f;
  if (tid < U * BP) {
    const int u = tid / BP, b = tid - u * BP;
    if (p.c0 != nullptr && b0 + b < p.B) c_state = p.c0[static_cast<long long>(b0 + b) * H + rank * U + u];
  }
  cluster_sync();
  const int r = tid % R, ks = tid / R;
  const int kper = H / KS;
  for (int t = 0; t < p.T; ++t) {
    const float* hcur = hbuf + (t & 1) * H * BP;
    float* hnext = hbuf + ((t + 1) & 1) * H * BP;
    // ---- gate pre-activations of my R rows: partial dot products over my k segment
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wp = Wt + static_cast<size_t>(ks) * kper * R + r;
    const float4* hp = reinterpret_cast<const float4*>(hcur) + ks * kper;
#pragma unroll 8
    for (int k = 0; k < kper; ++k) {
      const float w = wp[static_cast<size_t>(k) * R];
      const float4 hv = hp[k];
      acc.x = fmaf(w, hv.x, acc.x); acc.y = fmaf(w, hv.y, acc.y);
      acc.z = fmaf(w, hv.z, acc.z); acc.w = fmaf(w, hv.w, acc.w);
    }
    reinterpret_cast<float4*>(part)[ks * R + r] = acc;
    __syncthreads();
    if (tid < R) {
      float4 s = reinterpret_cast<float4*>(part)[tid];
      for (int q = 1; q < KS; ++q) {
        const float4 o = reinterpret_cast<float4*>(part)[q * R + tid];
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
      const int gate = tid / U, u = tid - gate * U;
      const int grow = gate * H + rank * U + u;
      float sv[BP] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int b = 0; b < BP; ++b) {
        float pre = sv[b];
        if (b0 + b < p.B) pre += __ldg(p.gx + (static_cast<long long>(b0 + b) * p.T + t) * 4 * H + grow);
        sv[b] = gate == 2 ? tanhf(pre) : sigmoidf_(pre);
        if (b0 + b < p.B) p.gates[(static_cast<long long>(b0 + b) * p.T + t) * 4 * H + grow] = sv[b];
      }
      reinterpret_cast<float4*>(gact)[tid] = make_float4(sv[0], sv[1], sv[2], sv[3]);
    }
    __syncthreads();
    // ---- cell update of my U units, new h broadcast to every CTA of the cluster
    if (tid < U * BP) {
      const int u = tid / BP, b = tid - u * BP;
      const float ig = gact[(0 * U + u) * BP + b], fg = gact[(1 * U + u) * BP + b];
      const float gg = gact[(2 * U + u) * BP + b], og = gact[(3 * U + u) * BP + b];
      c_state = fg * c_state + ig * gg;
      const float hv = og * tanhf(c_state);
      if (b0 + b < p.B) {
        const long long o = (static_cast<long long>(b0 + b) * p.T + t) * H + rank * U + u;
        p.hs[o] = hv;
        p.cs[o] = c_state;
      }
      // gather the 4 batch values of unit u in lane b == 0 and store one float4 into all 8 h buffers
      const float h1 = __shfl_down_sync(0xffffffffu, hv, 1), h2 = __shfl_down_sync(0xffffffffu, hv, 2),
                  h3 = __shfl_down_sync(0xffffffffu, hv, 3);
      if (b == 0) {
        const uint32_t local = smem_addr(hnext + (rank * U + u) * BP);
        const float4 v = make_float4(hv, h1, h2, h3);
#pragma unroll
        for (int dst = 0; dst < NC; ++dst) st_cluster_v4(mapa(local, dst), v);
      }
    }
    cluster_sync();
  }
}

struct BwdP {
  const float* dhs;     // [B, T, H] gradient wrt the layer's outputs
  const float* gates;   // [B, T, 4H] activated gates saved by forward
  const float* cs;      // [B, T, H]
  const float* c0;      // [B, H] or null
  const float* whh;     // [4H, H]
  const float* dhT;     // [B, H] gradient wrt h_T (or null)
  const float* dcT;     // [B, H] gradient wrt c_T (or null)
  float* dgx;           // [B, T, 4H] gradient wrt the gate pre-activations
  float* dh0;           // [B, H]
  float* dc0;           // [B, H]
  int B, T, H;
};

// shared memory: Wc [4H][U] (my U columns of W_hh) | dg [2][4H][BP] (double buffered by step parity) | part [SEG][U][BP]
__global__ void __launch_bounds__(kThreads, 1) lstm_bwd_kernel(const BwdP p) {
  extern __shared__ float sm[];
  const int H = p.H, U = H / NC, R4 = 4 * H, SEG = kThreads / U;     // SEG row segments of 4H / SEG rows each
  float* Wc = sm;
  float* dgbuf = Wc + static_cast<size_t>(R4) * U;
  float* part = dgbuf + 2 * R4 * BP;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int b0 = blockIdx.y * BP;
  for (int i = tid; i < R4 * U; i += kThreads) {
    const int rr = i / U, u = i - rr * U;
    Wc[i] = __ldg(p.whh + static_cast<long long>(rr) * H + rank * U + u);
  }
  // (unit u, batch b) state in thread tid = u * BP + b
  float dh_rec = 0.f, dc_next = 0.f;
  const int u_own = tid / BP, b_own = tid - u_own * BP;
  const bool owner = tid < U * BP;
  const bool live = owner && (b0 + b_own < p.B);
  if (live) {
    const long long o = static_cast<long long>(b0 + b_own) * H + rank * U + u_own;
    if (p.dhT != nullptr) dh_rec = p.dhT[o];
    if (p.dcT != nullptr) dc_next = p.dcT[o];
  }
  cluster_sync();
  const int um = tid % U, seg = tid / U, rper = R4 / SEG;
  for (int t = p.T - 1; t >= 0; --t) {
    // gate-gradient buffer of this step: a CTA that is one step ahead writes the OTHER buffer, so one cluster barrier
    // per step is enough (a buffer is rewritten two steps later, after everybody passed the barrier in between)
    float* dg = dgbuf + (t & 1) * R4 * BP;
    // ---- gate gradients of my units
    if (owner) {
      float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);      // (di, df, dg, do) pre-activation
      if (live) {
        const long long bt = static_cast<long long>(b0 + b_own) * p.T + t;
        const float* ga = p.gates + bt * 4 * H + rank * U + u_own;
        const float ig = ga[0], fg = ga[H], gg = ga[2 * H], og = ga[3 * H];
        const float c = p.cs[bt * H + rank * U + u_own];
        const float cprev = t > 0 ? p.cs[(bt - 1) * H + rank * U + u_own]
                                  : (p.c0 != nullptr ? p.c0[static_cast<long long>(b0 + b_own) * H + rank * U + u_own] : 0.f);
        const float dh = dh_rec + p.dhs[bt * H + rank * U + u_own];
        const float tc = tanhf(c);
        const float dc = dc_next + dh * og * (1.f - tc * tc);
        d4.x = dc * gg * ig * (1.f - ig);
        d4.y = dc * cprev * fg * (1.f - fg);
        d4.z = dc * ig * (1.f - gg * gg);
        d4.w = dh * tc * og * (1.f - og);
        dc_next = dc * fg;
        float* dgo = p.dgx + bt * 4 * H + rank * U + u_own;
        dgo[0] = d4.x; dgo[H] = d4.y; dgo[2 * H] = d4.z; dgo[3 * H] = d4.w;
      }
      // collect the 4 batch values per (gate, unit) in lane b == 0: one float4 per gate row to all 8 CTAs
      float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) {
        const float v0 = dv[gate];
        const float v1 = __shfl_down_sync(0xffffffffu, v0, 1), v2 = __shfl_down_sync(0xffffffffu, v0, 2),
                    v3 = __shfl_down_sync(0xffffffffu, v0, 3);
        if (b_own == 0) {
          const uint32_t local = smem_addr(dg + (gate * H + rank * U + u_own) * BP);
          const float4 v = make_float4(v0, v1, v2, v3);
#pragma unroll
          for (int dst = 0; dst < NC; ++dst) st_cluster_v4(mapa(local, dst), v);
        }
      }
    }
    cluster_sync();
    // ---- dh_{t-1} of my units = sum over all 4H gate rows of dgates * W_hh[:, my columns]
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wp = Wc + static_cast<size_t>(seg) * rper * U + um;
    const float4* dp = reinterpret_cast<const float4*>(dg) + seg * rper;
#pragma unroll 8
    for (int rr = 0; rr < rper; ++rr) {
      const float w = wp[static_cast<size_t>(rr) * U];
      const float4 d = dp[rr];
      acc.x = fmaf(w, d.x, acc.x); acc.y = fmaf(w, d.y, acc.y);
      acc.z = fmaf(w, d.z, acc.z); acc.w = fmaf(w, d.w, acc.w);
    }
    reinterpret_cast<float4*>(part)[seg * U + um] = acc;
    __syncthreads();
    if (owner) {
      float s = 0.f;
      for (int q = 0; q < SEG; ++q) s += part[(q * U + u_own) * BP + b_own];
      dh_rec = s;
    }
    __syncthreads();                       // `part` is rewritten by the next step's matvec
  }
  if (live) {
    const long long o = static_cast<long long>(b0 + b_own) * H + rank * U + u_own;
    p.dh0[o] = dh_rec;
    p.dc0[o] = dc_next;
  }
}

static size_t fwd_smem(int H) {
  const int U = H / NC, R = 4 * U, KS = kThreads / R;
  return (static_cast<size_t>(H) * R + 2 * H * BP + KS * R * BP + R * BP) * sizeof(float);
}
static size_t bwd_smem(int H) {
  const int U = H / NC, SEG = kThreads / U;
  return (static_cast<size_t>(4 * H) * U + 2 * 4 * H * BP + SEG * U * BP) * sizeof(float);
}

template <typename P>
static void launch_cluster(void (*kernel)(const P), const P& p, int chunks, size_t smem, cudaStream_t stream) {
  FLUTE_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(NC, chunks);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FLUTE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p));
}

static void check(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.is_contiguous(), name, ": contiguous fp32 CUDA tensor");
}

}  // namespace lstm

bool lstm_supported(int64_t H) { return H == 64 || H == 128 || H == 256; }

// gx [B, T, 4H], whh [4H, H], h0 / c0 [B, H] optional -> {hs [B,T,H], gates [B,T,4H], cs [B,T,H]}
std::vector<torch::Tensor> lstm_layer_fwd(torch::Tensor gx, torch::Tensor whh, c10::optional<torch::Tensor> h0,
                                          c10::optional<torch::Tensor> c0) {
  using namespace lstm;
  check(gx, "gx");
  check(whh, "whh");
  const int64_t B = gx.size(0), T = gx.size(1), H = whh.size(1);
  TORCH_CHECK(gx.dim() == 3 && gx.size(2) == 4 * H && whh.size(0) == 4 * H && lstm_supported(H),
              "lstm_layer_fwd: gx [B, T, 4H], whh [4H, H], H in {64, 128, 256}");
  const c10::cuda::CUDAGuard guard(gx.device());
  auto hs = torch::empty({B, T, H}, gx.options());
  auto gates = torch::empty({B, T, 4 * H}, gx.options());
  auto cs = torch::empty({B, T, H}, gx.options());
  FwdP p;
  p.gx = gx.data_ptr<float>(); p.whh = whh.data_ptr<float>();
  torch::Tensor h0c, c0c;
  if (h0.has_value()) { h0c = h0->contiguous(); check(h0c, "h0"); }
  if (c0.has_value()) { c0c = c0->contiguous(); check(c0c, "c0"); }
  p.h0 = h0.has_value() ? h0c.data_ptr<float>() : nullptr;
  p.c0 = c0.has_value() ? c0c.data_ptr<float>() : nullptr;
  p.hs = hs.data_ptr<float>(); p.gates = gates.data_ptr<float>(); p.cs = cs.data_ptr<float>();
  p.B = static_cast<int>(B); p.T = static_cast<int>(T); p.H = static_cast<int>(H);
  launch_cluster(lstm_fwd_kernel, p, static_cast<int>((B + BP - 1) / BP), fwd_smem(p.H), at::cuda::getCurrentCUDAStream());
  return {hs, gates, cs};
}

// -> {dgx [B,T,4H], dh0 [B,H], dc0 [B,H]}
std::vector<torch::Tensor> lstm_layer_bwd(torch::Tensor dhs, torch::Tensor gates, torch::Tensor cs, torch::Tensor whh,
                                          c10::optional<torch::Tensor> c0, c10::optional<torch::Tensor> dhT,
                                          c10::optional<torch::Tensor> dcT) {
  using namespace lstm;
  check(dhs, "dhs"); check(gates, "gates"); check(cs, "cs"); check(whh, "whh");
  const int64_t B = dhs.size(0), T = dhs.size(1), H = dhs.size(2);
  TORCH_CHECK(lstm_supported(H) && gates.size(2) == 4 * H, "lstm_layer_bwd: bad shapes");
  const c10::cuda::CUDAGuard guard(dhs.device());
  auto dgx = torch::zeros({B, T, 4 * H}, dhs.options());
  auto dh0 = torch::zeros({B, H}, dhs.options());
  auto dc0 = torch::zeros({B, H}, dhs.options());
  BwdP p;
  torch::Tensor c0c, dhc, dcc;
  if (c0.has_value()) { c0c = c0->contiguous(); check(c0c, "c0"); }
  if (dhT.has_value()) { dhc = dhT->contiguous(); check(dhc, "dhT"); }
  if (dcT.has_value()) { dcc = dcT->contiguous(); check(dcc, "dcT"); }
  p.dhs = dhs.data_ptr<float>(); p.gates = gates.data_ptr<float>(); p.cs = cs.data_ptr<float>();
  p.c0 = c0.has_value() ? c0c.data_ptr<float>() : nullptr;
  p.whh = whh.data_ptr<float>();
  p.dhT = dhT.has_value() ? dhc.data_ptr<float>() : nullptr;
  p.dcT = dcT.has_value() ? dcc.data_ptr<float>() : nullptr;
  p.dgx = dgx.data_ptr<float>(); p.dh0 = dh0.data_ptr<float>(); p.dc0 = dc0.data_ptr<float>();
  p.B = static_cast<int>(B); p.T = static_cast<int>(T); p.H = static_cast<int>(H);
  launch_cluster(lstm_bwd_kernel, p, static_cast<int>((B + BP - 1) / BP), bwd_smem(p.H), at::cuda::getCurrentCUDAStream());
  return {dgx, dh0, dc0};
}

}  // namespace flute
