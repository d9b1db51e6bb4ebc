// Shared device helpers for libmfm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mfm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmfm_b200 targets sm_100a (B200) only"
#endif

extern unsigned long long g_mfm_launches;   // host-side counter, see abi.cu

// Per-device facts and one-time kernel attributes (abi.cu).  The opt-in shared-memory limit and the SM count belong to
// the CURRENT device, and cudaFuncSetAttribute applies per device: both are cached per (device[, kernel]) behind a
// mutex, so a process that drives several GPUs -- or calls from autograd's backward thread -- stays correct.
struct MfmDevInfo { int sms; int smem_optin; };
const MfmDevInfo& mfm_dev_info();
int mfm_func_smem(const void* func, int bytes);      // 0 or a cudaError_t
template <typename F> static inline int mfm_func_smem_t(F* f, int bytes) { return mfm_func_smem(reinterpret_cast<const void*>(f), bytes); }

#define MFM_LAUNCH_CHECK()                                   \
  do {                                                       \
    ++g_mfm_launches;                                        \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

#define MFM_REQUIRE(cond) \
  do { if (!(cond)) return MFM_ERR_ARG; } while (0)

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Gate activations of the recurrences.  The cell update is transcendental-bound (5 per hidden unit per step), so
// these use the SFU directly: ex2.approx (2 ulp) and rcp.approx (1 ulp); absolute error <= ~2e-7, far inside the
// 1e-3 parity budget, ~4x cheaper than expf/tanhf.  Arguments are clamped so 1+e never overflows.
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gate_sigmoid(float x) {
  x = fminf(fmaxf(x, -30.0f), 30.0f);
  return rcp_fast(1.0f + __expf(-x));
}
__device__ __forceinline__ float gate_tanh(float x) {
  x = fminf(fmaxf(x, -15.0f), 15.0f);
  const float e = __expf(-2.0f * x);
  return (1.0f - e) * rcp_fast(1.0f + e);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MFM_ACT_RELU: return fmaxf(v, 0.0f);
    case MFM_ACT_TANH: return tanhf(v);
    case MFM_ACT_SIGMOID: return sigmoidf_acc(v);
    default: return v;
  }
}

// ---- counter-based dropout RNG (stateless; backward never needs the mask: relu+dropout output > 0 <=> kept) ----
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// Stream key of (seed, step, site): every input goes through its own mixing round, so consecutive steps (or sites) give
// unrelated keys -- NOT seed + step*C, which made the step-s+1 stream the step-s stream shifted by one element.
__host__ __device__ __forceinline__ uint32_t site_key(uint32_t seed, uint32_t step, uint32_t site) {
  return fmix32(fmix32(seed ^ fmix32(step + 0x9E3779B9u)) + site * 0x7F4A7C15u);
}
__device__ __forceinline__ uint32_t site_seed(const long long* rng, int site) {
  return site_key((uint32_t)rng[0], (uint32_t)rng[1], (uint32_t)site);
}
// Counter -> 32 random bits under a stream key: two rounds, the key enters both (a one-round hash(idx*C + key) of two keys
// is the same sequence at two offsets).
__host__ __device__ __forceinline__ uint32_t rng_bits(uint32_t sseed, uint32_t idx) {
  return fmix32(fmix32(idx + sseed) ^ sseed);
}
__device__ __forceinline__ bool drop_keep(uint32_t sseed, uint32_t idx, float p) {
  float u = (float)(rng_bits(sseed, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p;
}

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
// block-wide sum; result valid in thread 0. `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.0f;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.0f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
