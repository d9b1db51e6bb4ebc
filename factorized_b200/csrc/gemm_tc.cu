// tcgen05 GEMM for sm_100a: fp32 in HBM, split-bf16 (hi+lo) operands in shared memory, fp32 accumulate in TMEM.
//
//   D[128 x BN] (TMEM, fp32) += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T        (MFM_PATH_TC_BF16X3, ~2^-16 relative)
//   D[128 x BN]              += A_hi B_hi^T                                     (MFM_PATH_TC_BF16)
//
// Why this shape: the step's time-parallel GEMMs have M = T*B (tens of thousands of rows) against N,K <= 512, so
// they are bound by streaming the activation matrix once from HBM, not by the tensor pipe.  The kernel therefore
// reads fp32 activations straight from their producer's layout (any leading dimension, any column offset: no
// repack pass, no TMA alignment rule), converts in registers, and stores the bf16 planes in the tcgen05
// "no-swizzle" canonical layouts, which a SIMT store hits conflict-free:
//   K-major  operand (rows x K, K contiguous in HBM):   byte(r,k) = (k/8)*LBO + r*16 + (k%8)*2,   SBO = 128
//   MN-major operand (K x cols, cols contiguous in HBM): byte(k,c) = (k/8)*LBO + (c/8)*128 + (k%8)*16 + (c%8)*2
// so all three GEMM modes (NT forward, NN data-gradient, TN weight-gradient) run without a transpose pass.
// One elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN<=256, K=16); completion is tracked
// with tcgen05.commit -> mbarrier.  256 threads: the raw fp32 of chunk c+1 is prefetched into registers while chunk c
// is converted, fenced and multiplied; two shared-memory stages let the MMAs of chunk c overlap the conversion of c+1.
// The epilogue reads the accumulator with tcgen05.ld, transposes 32x32 blocks through shared memory so that every
// store / split-K reduction is a coalesced 128 B row segment, and applies the fused epilogue of mfm_gemm.
#include <cstdlib>
#include "gemm_args.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

#define TC_BM 128
#define TC_BK 32
#define TC_THREADS 256
#define TC_STAGES 2
#define TC_A_PLANE (4 * (TC_BM * 16 + 32))        // bytes of one A plane (K-major with 32 B slab padding; MN-major needs less)
#define TC_NA 2                                   // A items (8 fp32 each) per thread per chunk: 128 rows * 4 slabs / 256
#define TC_NB 4                                   // B items per thread per chunk (BN <= 256)


// One operand tile of a K chunk travels HBM -> registers (raw fp32) -> split-bf16 planes in shared memory.
// The two halves are separate calls so the loads of chunk c+1 are in flight while chunk c is converted,
// fenced, and multiplied (register-level software pipelining; nothing waits on a load it just issued).
//   K-major  source [rows, K] (K contiguous):  item = (row r, 8-wide k slab)
//   MN-major source [K, cols] (cols contiguous): item = (k row, 8-wide column group)
template <bool MN, int NI>
__device__ __forceinline__ void tile_fetch(float (&v)[NI][8], const float* __restrict__ src, long long ld, int mn0, int mn_max,
                                           int k0, int kend, int tile_mn, bool vec_ok, int ones_col = -1) {
  const int items = MN ? ((((tile_mn >> 3) + 3) & ~3) * TC_BK) : tile_mn * 4;
#pragma unroll
  for (int u = 0; u < NI; ++u) {
    const int idx = u * TC_THREADS + threadIdx.x;
    if (idx < items) {
      if (MN) {
        const int k = (idx & 7) + 8 * ((idx >> 5) & 3), mg = ((idx >> 3) & 3) + 4 * (idx >> 7);
        if (mg < (tile_mn >> 3)) {
          load8(src, ld, k0 + k, kend, mn0 + mg * 8, mn_max, vec_ok, v[u]);
          if (ones_col >= 0 && k0 + k < kend) {       // virtual all-ones column: column sums of the other operand
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (mn0 + mg * 8 + i == ones_col) v[u][i] = 1.0f;
          }
        }
      } else {
        const int slab = idx & 3, r = idx >> 2;
        load8(src, ld, mn0 + r, mn_max, k0 + slab * 8, kend, vec_ok, v[u]);
      }
    }
  }
}
template <bool MN, int NI>
__device__ __forceinline__ void tile_commit(const float (&v)[NI][8], unsigned char* hi, unsigned char* lo, int lbo, int tile_mn,
                                            bool want_lo) {
  const int items = MN ? ((((tile_mn >> 3) + 3) & ~3) * TC_BK) : tile_mn * 4;
#pragma unroll
  for (int u = 0; u < NI; ++u) {
    const int idx = u * TC_THREADS + threadIdx.x;
    if (idx < items) {
      if (MN) {
        const int k = (idx & 7) + 8 * ((idx >> 5) & 3), mg = ((idx >> 3) & 3) + 4 * (idx >> 7);
        if (mg < (tile_mn >> 3)) {
          const int off = (k >> 3) * lbo + mg * 128 + (k & 7) * 16;
          split_store(v[u], hi + off, lo + off, want_lo);
        }
      } else {
        const int slab = idx & 3, r = idx >> 2;
        split_store(v[u], hi + slab * lbo + r * 16, lo + slab * lbo + r * 16, want_lo);
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(TcArgs ta) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[TC_STAGES];
  __shared__ uint32_t tmem_holder;
  const GemmArgs& a = ta.g;
  const int BN = ta.BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // block indices pinned in registers: the compiler otherwise re-reads SR_CTAID (S2UR, long latency) inside the K loop
  unsigned bx, by, bz;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
  asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
  const int m0 = by * TC_BM, n0 = bx * BN;
  const int kbeg = bz * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const bool want_lo = ta.passes == 3;
  constexpr bool A_MN = (MODE == MFM_GEMM_TN);
  constexpr bool B_MN = (MODE != MFM_GEMM_NT);
  const int lboA = A_MN ? (TC_BM / 8) * 128 : (TC_BM * 16 + 32);
  const int lboB = B_MN ? (BN / 8) * 128 : (BN * 16 + 32);
  const int b_plane = 4 * (BN * 16 + 32);
  const int stage_bytes = 2 * TC_A_PLANE + 2 * b_plane;
  const bool vecA = ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) && ((a.lda & 3) == 0);
  const bool vecB = ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0) && ((a.ldb & 3) == 0);
  const int ones_col = (MODE == MFM_GEMM_TN && a.colsum_out) ? a.N : -1;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)ta.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // first chunk's loads go out before anything else waits
  float va[TC_NA][8], vb[TC_NB][8];
  const int nchunks = (kend - kbeg + TC_BK - 1) / TC_BK;
  // interior fast path (aligned, tile fully in bounds): per-thread item pointers computed once, advanced per chunk
  const bool fastA = vecA && (m0 + TC_BM <= a.M), fastB = vecB && (n0 + BN <= a.N) && ones_col < 0;
  const float* pa[TC_NA];
  const float* pb[TC_NB];
  bool okb[TC_NB];
#pragma unroll
  for (int u = 0; u < TC_NA; ++u) {
    const int idx = u * TC_THREADS + tid;
    if (A_MN) pa[u] = a.A + (long long)(kbeg + (idx & 7) + 8 * ((idx >> 5) & 3)) * a.lda + m0 + (((idx >> 3) & 3) + 4 * (idx >> 7)) * 8;
    else      pa[u] = a.A + (long long)(m0 + (idx >> 2)) * a.lda + kbeg + (idx & 3) * 8;
  }
#pragma unroll
  for (int u = 0; u < TC_NB; ++u) {
    const int idx = u * TC_THREADS + tid;
    if (B_MN) {
      const int mg = ((idx >> 3) & 3) + 4 * (idx >> 7);
      okb[u] = mg < (BN >> 3);
      pb[u] = a.B + (long long)(kbeg + (idx & 7) + 8 * ((idx >> 5) & 3)) * a.ldb + n0 + mg * 8;
    } else {
      okb[u] = idx < BN * 4;
      pb[u] = a.B + (long long)(n0 + (idx >> 2)) * a.ldb + kbeg + (idx & 3) * 8;
    }
  }
  const long long stepA = A_MN ? (long long)TC_BK * a.lda : TC_BK, stepB = B_MN ? (long long)TC_BK * a.ldb : TC_BK;
  auto fetch = [&](int c) {
    const int k1 = kbeg + c * TC_BK;
    const bool kfull = k1 + TC_BK <= kend;
    if (fastA && kfull) {
#pragma unroll
      for (int u = 0; u < TC_NA; ++u) load8_fast(pa[u] + c * stepA, va[u]);
    } else {
      tile_fetch<A_MN, TC_NA>(va, a.A, a.lda, m0, a.M, k1, kend, TC_BM, vecA);
    }
    if (fastB && kfull) {
#pragma unroll
      for (int u = 0; u < TC_NB; ++u)
        if (okb[u]) load8_fast(pb[u] + c * stepB, vb[u]);
    } else {
      tile_fetch<B_MN, TC_NB>(vb, a.B, a.ldb, n0, a.N, k1, kend, BN, vecB, ones_col);
    }
  };
  if (nchunks > 0) fetch(0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    unsigned char* st = smem + s * stage_bytes;
    unsigned char* Ahi = st;
    unsigned char* Alo = st + TC_A_PLANE;
    unsigned char* Bhi = st + 2 * TC_A_PLANE;
    unsigned char* Blo = Bhi + b_plane;
    if (c >= TC_STAGES && !(ta.dbg & 1)) mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / TC_STAGES - 1) & 1));   // MMAs that read this stage are done
    if (!(ta.dbg & 2)) {
      tile_commit<A_MN, TC_NA>(va, Ahi, Alo, lboA, TC_BM, want_lo);
      tile_commit<B_MN, TC_NB>(vb, Bhi, Blo, lboB, BN, want_lo);
    }
    // generic-proxy stores -> visible to the tensor core.  The fence lowers to MEMBAR.ALL.CTA, which waits for ALL of
    // this thread's outstanding memory operations: it must come BEFORE the prefetch loads are issued, or every chunk
    // pays the full HBM latency at the fence and nothing overlaps.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (c + 1 < nchunks && !(ta.dbg & 8)) fetch(c + 1);
    __syncthreads();
    if (tid == 0 && !(ta.dbg & 1)) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH = smem_u32(Ahi), aL = smem_u32(Alo), bH = smem_u32(Bhi), bL = smem_u32(Blo);
#pragma unroll
      for (int kk = 0; kk < TC_BK / 16; ++kk) {
        const uint32_t ao = kk * 2 * lboA, bo = kk * 2 * lboB;
        const uint64_t dAh = make_smem_desc(aH + ao, lboA, 128), dBh = make_smem_desc(bH + bo, lboB, 128);
        umma_bf16(tmem_base, dAh, dBh, idesc, (c > 0 || kk > 0) ? 1u : 0u);
        if (want_lo) {
          const uint64_t dAl = make_smem_desc(aL + ao, lboA, 128), dBl = make_smem_desc(bL + bo, lboB, 128);
          umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
          umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bars[s]));      // implies tcgen05.fence::before_thread_sync
    }
  }
  // all MMAs complete when the last commit lands (commits are ordered)
  if (nchunks > 0 && !(ta.dbg & 1)) {
    const int c = nchunks - 1;
    mbar_wait(smem_u32(&bars[c & 1]), (uint32_t)((c / TC_STAGES) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue (tc_epilogue.cuh): TMEM -> registers -> transposed through the now idle stage memory -> coalesced stores
  if (!(ta.dbg & 4)) tc_epilogue(ta, tmem_base, reinterpret_cast<float*>(smem), warp, lane, m0, n0, nchunks > 0, ones_col);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ta.tmem_cols)
                 : "memory");
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// gemm_tcp.cu: the pipelined (warp-specialised, cp.async-staged) kernel for 16 B-aligned operands
bool gemm_tcp_eligible(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb);
int gemm_tcp_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                    const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                    const long long* rng, float* colsum_out, void* ws, size_t ws_bytes, cudaStream_t st);

int gemm_tc_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                   float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                   const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                   const long long* rng, float* colsum_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (gemm_tcp_eligible(mode, M, N, K, A, lda, B, ldb)) {
    const int rc = gemm_tcp_launch(passes, mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask,
                                   mask_scale, drop_p, drop_site, rng, colsum_out, ws, ws_bytes, st);
    if (rc != MFM_ERR_UNSUPPORTED) return rc;       // unsupported operand alignment: the register-prefetch kernel below
  }
  TcArgs ta;
  ta.g = GemmArgs{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
                  drop_p, drop_site, rng, K, 0, colsum_out};
  ta.g.mse = g_pending_mse;          // (cleared by mfm_gemm_mse after the call)
  ta.passes = passes;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("MFM_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  ta.dbg = dbg;
  // tile N: the whole (padded) N when it fits 256 columns, else near-equal tiles (+1 virtual ones column for colsum_out)
  const int n16 = round_up(N + (colsum_out ? 1 : 0), 16);
  const int ntiles = (n16 + 255) / 256;
  ta.BN = round_up((n16 + ntiles - 1) / ntiles, 16);
  int cols = 32;
  while (cols < ta.BN) cols <<= 1;
  ta.tmem_cols = cols;
  dim3 grid((N + (colsum_out ? 1 : 0) + ta.BN - 1) / ta.BN, (M + TC_BM - 1) / TC_BM, 1);
  const bool plain = !bias && !bias2 && act == MFM_ACT_NONE && !mask && drop_p <= 0.0f && accumulate;
  if (plain && K >= 2048) {
    long long tiles = (long long)grid.x * grid.y;
    int splits = (int)((2 * mfm_dev_info().sms + tiles - 1) / tiles);
    int maxs = K / 256;
    if (splits > maxs) splits = maxs;
    if (splits > 1) {
      int kc = round_up((K + splits - 1) / splits, TC_BK);
      ta.g.kchunk = kc;
      ta.g.atomic = 1;
      grid.z = (K + kc - 1) / kc;
    }
  }
  const size_t smem = (size_t)TC_STAGES * (2 * TC_A_PLANE + 2 * 4 * (ta.BN * 16 + 32)) + 128;
  {
    int e = 0;
    if (mode == MFM_GEMM_NT) e = mfm_func_smem_t(gemm_tc_kernel<MFM_GEMM_NT>, 110 * 1024);
    if (mode == MFM_GEMM_NN) e = mfm_func_smem_t(gemm_tc_kernel<MFM_GEMM_NN>, 110 * 1024);
    if (mode == MFM_GEMM_TN) e = mfm_func_smem_t(gemm_tc_kernel<MFM_GEMM_TN>, 110 * 1024);
    if (e) return e;
  }
  switch (mode) {
    case MFM_GEMM_NT: gemm_tc_kernel<MFM_GEMM_NT><<<grid, TC_THREADS, smem, st>>>(ta); break;
    case MFM_GEMM_NN: gemm_tc_kernel<MFM_GEMM_NN><<<grid, TC_THREADS, smem, st>>>(ta); break;
    case MFM_GEMM_TN: gemm_tc_kernel<MFM_GEMM_TN><<<grid, TC_THREADS, smem, st>>>(ta); break;
    default: return MFM_ERR_ARG;
  }
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
