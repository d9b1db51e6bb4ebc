// tcgen05 / mbarrier / split-bf16 helpers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Wait on an mbarrier phase.  try_wait carries a suspend-time hint, so a waiting warp is parked by the hardware until
// the phase completes instead of re-polling: a polling loop (the first version, with a clock64() time-out per iteration)
// spent a third of an SM's issue slots on waiters -- slots the co-resident CTA's epilogue needed.  Bounded: a barrier
// that never completes traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int spin = 0; spin < (1 << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_NONE
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// interior fast path: 16 B-aligned, fully in bounds, no predicates
__device__ __forceinline__ void load8_fast(const float* __restrict__ p, float v[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 consecutive fp32 along the contiguous dimension at (row, col..col+7); zero outside [0,nrows) x [0,ncols)
__device__ __forceinline__ void load8(const float* __restrict__ src, long long ld, int row, int nrows, int col, int ncols,
                                      bool vec_ok, float v[8]) {
  if (row < nrows && col + 8 <= ncols) {
    const float* p = src + (long long)row * ld + col;
    if (vec_ok) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(p + i);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (row < nrows) {
      const float* p = src + (long long)row * ld + col;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < ncols) v[i] = __ldg(p + i);
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);      // .x = a (low address), .y = b
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 x8 -> 16 B of bf16 "hi" and (optionally) 16 B of bf16 "lo" = bf16(x - hi)
__device__ __forceinline__ void split_store(const float v[8], unsigned char* hi_dst, unsigned char* lo_dst, bool want_lo) {
  uint4 h;
  h.x = pack_bf16(v[0], v[1]); h.y = pack_bf16(v[2], v[3]); h.z = pack_bf16(v[4], v[5]); h.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(hi_dst) = h;
  if (want_lo) {
    float r[8];
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[2 * i] = v[2 * i] - __uint_as_float(hw[i] << 16);
      r[2 * i + 1] = v[2 * i + 1] - __uint_as_float(hw[i] & 0xFFFF0000u);
    }
    uint4 l;
    l.x = pack_bf16(r[0], r[1]); l.y = pack_bf16(r[2], r[3]); l.z = pack_bf16(r[4], r[5]); l.w = pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(lo_dst) = l;
  }
}

