// GEMM argument block and the fused epilogue shared by the CUDA-core and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

// Fused reconstruction head (include/mfm_b200.h::mfm_gemm_mse): with x set, the epilogue value v = x_hat becomes
// C = grad_scale * (v - x), *slot += loss_scale * sum (v - x)^2, and v itself is stored to xhat only when asked for.
struct GemmMse {
  const float* x; long long ldx;
  float loss_scale, grad_scale;
  float* slot;
  float* xhat; long long ldxhat;
};
// mfm_gemm_mse parks its extension here for the launcher that builds the GemmArgs of the same call (same host thread)
extern thread_local GemmMse g_pending_mse;
static inline GemmMse take_pending_mse() {
  GemmMse m = g_pending_mse;
  g_pending_mse = GemmMse{nullptr, 0, 0.0f, 0.0f, nullptr, nullptr, 0};
  return m;
}

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
  GemmMse mse;         // NT only: fused MSE head (x == nullptr: off)
};

// epilogue of include/mfm_b200.h::mfm_gemm for one output element
__device__ __forceinline__ void gemm_epilogue_store(const GemmArgs& a, int m, int n, float v, bool do_drop, uint32_t sseed,
                                                    float keep_scale, float* sq_acc = nullptr) {
  float* cp = a.C + (long long)m * a.ldc + n;
  if (a.atomic) { atomicAdd(cp, v); return; }
  if (a.bias) v += __ldg(a.bias + n);
  if (a.bias2) v += __ldg(a.bias2 + n);
  v = apply_act(v, a.act);
  if (a.mse.x) {        // the caller sums the squared residuals of its elements and adds them once per warp
    const float r = v - __ldg(a.mse.x + (long long)m * a.mse.ldx + n);
    if (sq_acc) *sq_acc = fmaf(r, r, *sq_acc);
    else atomicAdd(a.mse.slot, a.mse.loss_scale * r * r);
    if (a.mse.xhat) a.mse.xhat[(long long)m * a.mse.ldxhat + n] = v;
    v = a.mse.grad_scale * r;
  }
  if (do_drop) v = drop_keep(sseed, (uint32_t)m * (uint32_t)a.N + (uint32_t)n, a.drop_p) ? v * keep_scale : 0.0f;
  if (a.mask) v = (__ldg(a.mask + (long long)m * a.ldmask + n) > 0.0f) ? v * a.mask_scale : 0.0f;
  if (a.accumulate) v += *cp;
  *cp = v;
}
