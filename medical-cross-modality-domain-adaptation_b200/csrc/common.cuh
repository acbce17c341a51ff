// Shared device/host helpers for the PnP-AdaNet B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PNP_OK 0
#define PNP_ERR_BAD_ARG 100001
#define PNP_ERR_UNSUPPORTED 100002
#define PNP_ERR_DRIVER 100003

#define PNP_LAUNCH_CHECK()                        \
  do {                                            \
    cudaError_t _e = cudaGetLastError();          \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

#define PNP_CUDA(call)                            \
  do {                                            \
    cudaError_t _e = (call);                      \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

static inline int pnp_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG.  One call yields the 4 uniforms for the 4 consecutive elements
// [4*idx4, 4*idx4+3] of a tensor; `stream` separates dropout call sites, the seed lives in device
// memory so that a captured CUDA graph sees a fresh seed on every replay.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 pnp_philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

struct PnpDropout {
  const unsigned long long* seed_ptr;  // device scalar; nullptr => dropout disabled
  unsigned long long stream;           // call-site id
  float keep;                          // keep probability
  float inv_keep;                      // 1/keep
};

// multipliers (0 or 1/keep) for elements 4*idx4 .. 4*idx4+3
__device__ __forceinline__ float4 pnp_dropout_mult4(const PnpDropout& d, unsigned long long seed, unsigned long long idx4) {
  uint4 r = pnp_philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)d.stream, (uint32_t)(d.stream >> 32)),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 2.3283064365386963e-10f;  // 2^-32
  float4 m;
  m.x = (r.x * s < d.keep) ? d.inv_keep : 0.f;
  m.y = (r.y * s < d.keep) ? d.inv_keep : 0.f;
  m.z = (r.z * s < d.keep) ? d.inv_keep : 0.f;
  m.w = (r.w * s < d.keep) ? d.inv_keep : 0.f;
  return m;
}

__device__ __forceinline__ float pnp_dropout_mult1(const PnpDropout& d, unsigned long long seed, unsigned long long idx) {
  float4 m = pnp_dropout_mult4(d, seed, idx >> 2);
  int l = (int)(idx & 3);
  return l == 0 ? m.x : (l == 1 ? m.y : (l == 2 ? m.z : m.w));
}

__device__ __forceinline__ float pnp_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double pnp_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
