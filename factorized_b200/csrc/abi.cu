// C-ABI glue: library version, launch counter, GEMM path selection and dispatch.
#include <mutex>
#include <set>
#include <utility>
#include "gemm_args.cuh"

unsigned long long g_mfm_launches = 0;

// ---- per-device caches ----------------------------------------------------------------------------------------------
static std::mutex g_dev_mu;
static MfmDevInfo g_dev[64];
static bool g_dev_ok[64];
static std::set<std::pair<int, const void*>> g_attr_done;

const MfmDevInfo& mfm_dev_info() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_ok[dev]) {
    MfmDevInfo d;
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) d.sms = 148;
    if (cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) d.smem_optin = 48 * 1024;
    (void)cudaGetLastError();
    g_dev[dev] = d;
    g_dev_ok[dev] = true;
  }
  return g_dev[dev];
}

int mfm_func_smem(const void* func, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  const auto key = std::make_pair(dev, func);
  if (g_attr_done.count(key)) return 0;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return (int)e;
  g_attr_done.insert(key);
  return 0;
}
static int g_gemm_path = MFM_PATH_TC_BF16X3;   // tensor cores, split-bf16 operands; MFM_PATH_SIMT_FP32 is the exact-fp32 alternative

int gemm_simt_launch(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                     float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                     const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                     const long long* rng, cudaStream_t st);

extern "C" int mfm_version(void) { return MFM_B200_VERSION; }
extern "C" unsigned long long mfm_launch_count(void) { return g_mfm_launches; }
int gemm_tc_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                   float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                   const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                   const long long* rng, float* colsum_out, void* ws, size_t ws_bytes, cudaStream_t st);
int gemm_tcp_launch_tn_pair(int passes, int M, int K, const float* A, long long lda, int N1, const float* B1, long long ldb1,
                            float* C1, long long ldc1, float* colsum1, int N2, const float* B2, long long ldb2, float* C2,
                            long long ldc2, cudaStream_t st);
static long long g_tc_min_work = 1LL << 20;   // M*N*K below this stays on the CUDA-core kernel

extern "C" int mfm_set_gemm_path(int path) {
  if (path != MFM_PATH_SIMT_FP32 && path != MFM_PATH_TC_BF16X3 && path != MFM_PATH_TC_BF16) return MFM_ERR_UNSUPPORTED;
  g_gemm_path = path;
  return MFM_OK;
}
extern "C" int mfm_set_gemm_tc_min_work(long long mnk) {
  if (mnk < 0) return MFM_ERR_ARG;
  g_tc_min_work = mnk;
  return MFM_OK;
}
extern "C" int mfm_get_gemm_path(void) { return g_gemm_path; }

extern "C" int mfm_gemm(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                        float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                        const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                        const long long* rng, float* colsum_out, void* stream) {
  return mfm_gemm_ws(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale, drop_p,
                     drop_site, rng, colsum_out, nullptr, 0, stream);
}

extern "C" int mfm_gemm_ws(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                           float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                           const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                           const long long* rng, float* colsum_out, void* ws, long long ws_bytes, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C);
  MFM_REQUIRE(!colsum_out || mode == MFM_GEMM_TN);
  MFM_REQUIRE(mode == MFM_GEMM_NT || mode == MFM_GEMM_NN || mode == MFM_GEMM_TN);
  MFM_REQUIRE(act >= MFM_ACT_NONE && act <= MFM_ACT_SIGMOID);
  MFM_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || rng));
  if (g_gemm_path != MFM_PATH_SIMT_FP32 && (long long)M * N * K >= g_tc_min_work)
    return gemm_tc_launch(g_gemm_path == MFM_PATH_TC_BF16X3 ? 3 : 1, mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act,
                          accumulate, mask, ldmask, mask_scale, drop_p, drop_site, rng, colsum_out, ws, ws_bytes > 0 ? (size_t)ws_bytes : 0,
                          (cudaStream_t)stream);
  if (colsum_out) {      // CUDA-core path: the column sums are a separate streaming kernel
    int rc = mfm_colsum(K, M, A, lda, colsum_out, stream);
    if (rc) return rc;
  }
  return gemm_simt_launch(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
                          drop_p, drop_site, rng, (cudaStream_t)stream);
}

thread_local GemmMse g_pending_mse = {nullptr, 0, 0.0f, 0.0f, nullptr, nullptr, 0};

extern "C" int mfm_gemm_mse(int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, const float* bias,
                            const float* x, long long ldx, float loss_scale, float grad_scale, float* slot,
                            float* dxhat, long long lddx, float* xhat, long long ldxhat, void* ws, long long ws_bytes, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && x && slot && dxhat);
  g_pending_mse = GemmMse{x, ldx, loss_scale, grad_scale, slot, xhat, ldxhat};
  const int rc = mfm_gemm_ws(MFM_GEMM_NT, M, N, K, A, lda, B, ldb, dxhat, lddx, bias, nullptr, MFM_ACT_NONE, 0, nullptr, 0, 1.0f, 0.0f, 0,
                             nullptr, nullptr, ws, ws_bytes, stream);
  (void)take_pending_mse();          // (an argument error returned before any launcher consumed it)
  return rc;
}

extern "C" int mfm_gemm_tn_pair(int M, int K, const float* A, long long lda, int N1, const float* B1, long long ldb1,
                                float* C1, long long ldc1, float* colsum1, int N2, const float* B2, long long ldb2,
                                float* C2, long long ldc2, void* stream) {
  MFM_REQUIRE(M > 0 && K > 0 && N1 > 0 && N2 > 0 && A && B1 && B2 && C1 && C2);
  if (g_gemm_path != MFM_PATH_SIMT_FP32 && (long long)M * (N1 + N2) * K >= g_tc_min_work) {
    const int rc = gemm_tcp_launch_tn_pair(g_gemm_path == MFM_PATH_TC_BF16X3 ? 3 : 1, M, K, A, lda, N1, B1, ldb1, C1, ldc1, colsum1,
                                           N2, B2, ldb2, C2, ldc2, (cudaStream_t)stream);
    if (rc != MFM_ERR_UNSUPPORTED) return rc;
  }
  int rc = mfm_gemm(MFM_GEMM_TN, M, N1, K, A, lda, B1, ldb1, C1, ldc1, nullptr, nullptr, MFM_ACT_NONE, 1, nullptr, 0, 1.0f, 0.0f, 0,
                    nullptr, colsum1, stream);
  if (rc) return rc;
  return mfm_gemm(MFM_GEMM_TN, M, N2, K, A, lda, B2, ldb2, C2, ldc2, nullptr, nullptr, MFM_ACT_NONE, 1, nullptr, 0, 1.0f, 0.0f, 0,
                  nullptr, nullptr, stream);
}
