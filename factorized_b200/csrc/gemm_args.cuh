// GEMM argument block and the fused epilogue shared by the CUDA-core and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

struct GemmArgs {
  int M, N, K;
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  const float* bias; const float* bias2;
  int act, accumulate;
  const float* mask; long long ldmask; float mask_scale;
  float drop_p; int drop_site; const long long* rng;
  int kchunk;      // K range per blockIdx.z
  int atomic;      // split-K: atomicAdd partial sums into C
  float* colsum_out;   // TN only (tensor-core kernel): colsum_out[m] += sum_k A[k,m]  (bias gradient fused as a ones column of B)
};

// epilogue of include/mfm_b200.h::mfm_gemm for one output element
__device__ __forceinline__ void gemm_epilogue_store(const GemmArgs& a, int m, int n, float v, bool do_drop, uint32_t sseed,
                                                    float keep_scale) {
  float* cp = a.C + (long long)m * a.ldc + n;
  if (a.atomic) { atomicAdd(cp, v); return; }
  if (a.bias) v += __ldg(a.bias + n);
  if (a.bias2) v += __ldg(a.bias2 + n);
  v = apply_act(v, a.act);
  if (do_drop) v = drop_keep(sseed, (uint32_t)m * (uint32_t)a.N + (uint32_t)n, a.drop_p) ? v * keep_scale : 0.0f;
  if (a.mask) v = (__ldg(a.mask + (long long)m * a.ldmask + n) > 0.0f) ? v * a.mask_scale : 0.0f;
  if (a.accumulate) v += *cp;
  *cp = v;
}
