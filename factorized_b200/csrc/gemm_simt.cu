// Generic fp32 CUDA-core GEMM with the fused epilogue of include/mfm_b200.h::mfm_gemm.
// This is the exact-fp32 path (MFM_PATH_SIMT_FP32) and the fallback for shapes the tcgen05
// path does not take (K < 16, tiny M).  64x64x16 tiles, 256 threads, 4x4 register blocking.
#include "common.cuh"

#include "gemm_args.cuh"

#define BM 64
#define BN 64
#define BK 16
#define SPAD 4

template <int MODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  __shared__ __align__(16) float As[BK][BM + SPAD];
  __shared__ __align__(16) float Bs[BK][BN + SPAD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // loader coordinates
  const int lr = t >> 2, lq = (t & 3) * 4;     // operand with contiguous K: row lr, k-quad lq
  const int kr = t >> 4, kq = (t & 15) * 4;    // operand with contiguous M/N: k row kr, col-quad kq

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (MODE == MFM_GEMM_TN) {
      const int k = k0 + kr;
      const float* src = a.A + (long long)k * a.lda + m0 + kq;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        As[kr][kq + i] = (k < kend && m0 + kq + i < a.M) ? __ldg(src + i) : 0.0f;
    } else {
      const int m = m0 + lr;
      const float* src = a.A + (long long)m * a.lda + k0 + lq;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        As[lq + i][lr] = (m < a.M && k0 + lq + i < kend) ? __ldg(src + i) : 0.0f;
    }
    // ---- B tile -> Bs[k][n]
    if (MODE == MFM_GEMM_NT) {
      const int n = n0 + lr;
      const float* src = a.B + (long long)n * a.ldb + k0 + lq;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        Bs[lq + i][lr] = (n < a.N && k0 + lq + i < kend) ? __ldg(src + i) : 0.0f;
    } else {
      const int k = k0 + kr;
      const float* src = a.B + (long long)k * a.ldb + n0 + kq;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        Bs[kr][kq + i] = (k < kend && n0 + kq + i < a.N) ? __ldg(src + i) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  uint32_t sseed = 0;
  const bool do_drop = a.drop_p > 0.0f;
  if (do_drop) sseed = site_seed(a.rng, a.drop_site);
  const float keep_scale = do_drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      gemm_epilogue_store(a, m, n, acc[i][j], do_drop, sseed, keep_scale, &sq);
    }
  }
  if (a.mse.x) {                                       // fused MSE head: one atomic per warp
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(a.mse.slot, a.mse.loss_scale * sq);
  }
}

int gemm_simt_launch(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                     float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                     const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                     const long long* rng, cudaStream_t st) {
  GemmArgs a{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
             drop_p, drop_site, rng, K, 0, nullptr};
  a.mse = g_pending_mse;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  // split K over CTAs for the weight-gradient shape (tiny M,N; K = T*B rows), plain sums only
  const bool plain = !bias && !bias2 && act == MFM_ACT_NONE && !mask && drop_p <= 0.0f && accumulate;
  if (plain && K >= 1024) {
    long long tiles = (long long)grid.x * grid.y;
    int splits = (int)((2 * mfm_dev_info().sms + tiles - 1) / tiles);
    int maxs = K / 64;       // the K loop is not software-pipelined: a CTA pays one load latency per 16 k, so short slices win
    if (splits > maxs) splits = maxs;
    if (splits > 1) {
      int kc = (K + splits - 1) / splits;
      kc = ((kc + BK - 1) / BK) * BK;
      a.kchunk = kc;
      a.atomic = 1;
      grid.z = (K + kc - 1) / kc;
    }
  }
  switch (mode) {
    case MFM_GEMM_NT: gemm_simt_kernel<MFM_GEMM_NT><<<grid, 256, 0, st>>>(a); break;
    case MFM_GEMM_NN: gemm_simt_kernel<MFM_GEMM_NN><<<grid, 256, 0, st>>>(a); break;
    case MFM_GEMM_TN: gemm_simt_kernel<MFM_GEMM_TN><<<grid, 256, 0, st>>>(a); break;
    default: return MFM_ERR_ARG;
  }
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
