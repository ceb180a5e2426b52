// Shared device helpers for the msrflute_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define FLUTE_CUDA_CHECK(expr)                                                                         \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    TORCH_CHECK(_e == cudaSuccess, "CUDA error ", cudaGetErrorString(_e), " at ", __FILE__, ":", __LINE__); \
  } while (0)

namespace flute {

constexpr int kMaxPeers = 16;

// Up to kMaxPeers raw device pointers passed by value (peer-mapped or local buffers).
struct PtrList {
  float* p[kMaxPeers];
  int n;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of up to two values; result valid in every thread.  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ float2 block_sum2(float a, float b) {
  __shared__ float sa[32], sb[32];
  __syncthreads();  // protect smem reuse across consecutive calls
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sa[wid] = a; sb[wid] = b; }
  __syncthreads();
  a = lane < nw ? sa[lane] : 0.f;
  b = lane < nw ? sb[lane] : 0.f;
  a = warp_sum(a);
  b = warp_sum(b);
  return make_float2(a, b);
}

// 16-byte global accesses.  Weights/gradients of an arena are streamed once per kernel: bypass L1.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// Same without .nc: for buffers the same kernel also writes (w, g are updated in place).
__device__ __forceinline__ float4 ld_na(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
// Plain (coherent) 16-byte load: for buffers another GPU may have just written (peer accumulators).
__device__ __forceinline__ float4 ld_coherent(const float4* p) {
  float4 r;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------ Philox4x32-10 (counter based RNG)
// Counter = element-quad index, key = seed: the noise added to element i does not depend on the launch geometry
// or on the number of GPUs (SURVEY §7.3 risk 7).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // (0,1]
  return (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
// four standard normals for quad index q
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t q) {
  uint4 c = make_uint4(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), 0x464C5554u /*"FLUT"*/, 0u);
  uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  const float u0 = u32_to_unit(r.x), u1 = u32_to_unit(r.y), u2 = u32_to_unit(r.z), u3 = u32_to_unit(r.w);
  const float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace flute
