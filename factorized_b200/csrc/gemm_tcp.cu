// Pipelined tcgen05 GEMM for sm_100a (the streaming form of gemm_tc.cu; same math, same epilogue, same three layouts).
//
// The step's time-parallel GEMMs are [T*B, <=512] activations against small weights: they are bound by streaming the
// activation operand from HBM.  gemm_tc.cu keeps one K chunk of prefetch in registers and runs all its threads in
// lock step (load -> convert -> fence -> barrier -> MMA), so every chunk exposes the whole latency chain.  Here the
// links of that chain are different warps connected by mbarriers, and the transport is TMA:
//
//   warp 8 (one thread)  PRODUCER    cp.async.bulk.tensor.2d: the raw fp32 box of A and of B for chunk c -> ring stage s,
//                                    completion by complete_tx on full[s].  The tensor maps describe the operands as
//                                    they lie in HBM (any 16 B-aligned view); M/N/K tails are zero-filled by the TMA
//                                    unit.  No thread ever has a load outstanding, several stages are always in flight.
//   warps 0..7           CONVERTERS  wait full[s]; read 32 B items of the stage, split fp32 -> bf16 hi + lo, store both
//                                    planes in the tcgen05 no-swizzle canonical layout (plane set p), fence.proxy.async;
//                                    one arrive per warp on conv[p] and on empty[s].  Warps run ahead independently.
//   warp 9 (one thread)  MMA         wait conv[p]; tcgen05.mma hi*hi + lo*hi + hi*lo (M=128, N=BN, K=16) into the fp32 TMEM
//                                    accumulator; tcgen05.commit -> pfree[p] hands the plane set back to the converters.
//   warps 0..7           EPILOGUE    after the last commit: tc_epilogue.cuh (TMEM -> transposed through the idle ring ->
//                                    512 B coalesced stores with bias / activation / dropout / mask / split-K reduction).
//
// Two CTAs are resident per SM, so one CTA's epilogue stores overlap the other's loads.  Staging layouts:
//   K-major operand  [rows, K]:  one dense box [rows][BK] fp32; item (r, slab) = 32 B at r*BK*4 + slab*32
//                                -> plane byte slab*LBO + r*16, LBO = rows*16 + 128/SLABS (rotates banks: conflict-free)
//   MN-major operand [K, cols]:  boxes of [BK][32 cols] with the 128 B TMA swizzle (16 B chunk index ^= k%8), so the
//                                eight k rows of a core matrix are read from eight different bank groups
//                                -> plane byte (k/8)*LBO + (c/8)*128 + (k%8)*16
// Requirements (else the caller falls back to gemm_tc.cu): A and B 16 B aligned, leading dimensions multiples of 4.
#include <cuda.h>
#include <cstdlib>
#include "tc_epilogue.cuh"

#define P_BM 128
#define P_MAXBN 128
#define P_NCONV 256                 // converter / epilogue threads (warps 0..7)
#define P_THREADS (P_NCONV + 64)    // + producer warp + MMA warp

template <int BK> struct PCfg {
  static constexpr int SLABS = BK / 8;
  static constexpr int PAD = 128 / SLABS;
  static constexpr int STAGES = (BK == 16) ? 4 : 2;
  static constexpr int PSETS = (BK == 16) ? 2 : 1;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds8(const unsigned char* p0, const unsigned char* p1, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p0);
  const float4 b = *reinterpret_cast<const float4*>(p1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- converter: one operand tile of a stage -> split-bf16 planes ------------------------------------------------------
template <bool MN, int BK>
__device__ __forceinline__ void convert_tile(const unsigned char* st, unsigned char* hi, unsigned char* lo, int rows_or_cols,
                                             int lbo, int tid, bool want_lo, int ones_local, int kvalid) {
  constexpr int SLABS = BK / 8;
  constexpr int NI = P_MAXBN * SLABS / P_NCONV;          // items per thread for a 128-wide tile: 1 (BK 16) or 2 (BK 32)
  float v[8];
#pragma unroll
  for (int u = 0; u < NI; ++u) {
    const int idx = u * P_NCONV + tid;
    if (!MN) {
      const int r = idx / SLABS, slab = idx % SLABS;
      if (r < rows_or_cols) {
        const unsigned char* p = st + r * (BK * 4) + slab * 32;
        lds8(p, p + 16, v);
        split_store(v, hi + slab * lbo + r * 16, lo + slab * lbo + r * 16, want_lo);
      }
    } else {
      const int klo = idx & 7, g = idx >> 3, ng = rows_or_cols >> 3;
      const int mg = g % ng, khi = g / ng;
      if (khi < SLABS) {
        const int k = khi * 8 + klo;
        const unsigned char* row = st + (mg >> 2) * (BK * 128) + k * 128;      // box of 32 columns, 128 B rows, swizzled
        const int ch = (mg & 3) * 2;
        lds8(row + ((ch ^ klo) << 4), row + (((ch + 1) ^ klo) << 4), v);
        if (ones_local >= 0 && (ones_local >> 3) == mg && k < kvalid) {                          // virtual ones column
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == (ones_local & 7)) v[i] = 1.0f;
        }
        const int off = khi * lbo + mg * 128 + klo * 16;
        split_store(v, hi + off, lo + off, want_lo);
      }
    }
  }
}

template <int MODE, int BK>
__global__ void __launch_bounds__(P_THREADS, 2) gemm_tcp_kernel(const TcArgs ta, const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB) {
  constexpr int S = PCfg<BK>::STAGES, NP = PCfg<BK>::PSETS, SLABS = PCfg<BK>::SLABS, PAD = PCfg<BK>::PAD;
  constexpr bool A_MN = (MODE == MFM_GEMM_TN);
  constexpr bool B_MN = (MODE != MFM_GEMM_NT);
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[S], empty_bar[S], conv_bar[NP], pfree_bar[NP], accum_bar;
  __shared__ uint32_t tmem_holder;
  unsigned char* const smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const GemmArgs& a = ta.g;
  const int BN = ta.BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned bx, by, bz;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
  asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
  const int m0 = by * P_BM, n0 = bx * BN;
  const int kbeg = bz * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const int nchunks = (kend - kbeg + BK - 1) / BK;
  const bool want_lo = ta.passes == 3;
  const int BNb = B_MN ? ((BN + 31) & ~31) : BN;          // MN-major B arrives in boxes of 32 columns
  const int stA = P_BM * BK * 4, stB = BNb * BK * 4, stage_bytes = stA + stB;
  const int lboA = A_MN ? (P_BM / 8) * 128 : (P_BM * 16 + PAD);
  const int lboB = B_MN ? (BN / 8) * 128 : (BN * 16 + PAD);
  const int plA = SLABS * (P_BM * 16 + PAD), plB = SLABS * (BN * 16 + PAD), pset = 2 * plA + 2 * plB;
  unsigned char* const planes = smem + S * stage_bytes;
  const int ones_col = (MODE == MFM_GEMM_TN && a.colsum_out) ? a.N : -1;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), P_NCONV / 32);
    }
    for (int p = 0; p < NP; ++p) {
      mbar_init(smem_u32(&conv_bar[p]), P_NCONV / 32);
      mbar_init(smem_u32(&pfree_bar[p]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)ta.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp < P_NCONV / 32) {
    // ================================ CONVERTERS, then EPILOGUE ================================
    const int ones_local = ones_col >= 0 ? ones_col - n0 : -1;      // column of the tile that is the virtual ones column
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % S, p = c % NP;
      const unsigned char* st = smem + s * stage_bytes;
      unsigned char* Ahi = planes + p * pset;
      unsigned char* Bhi = Ahi + 2 * plA;
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((c / S) & 1));
      if (c >= NP) mbar_wait(smem_u32(&pfree_bar[p]), (uint32_t)((c / NP - 1) & 1));
      if (!(ta.dbg & 2)) {
        convert_tile<A_MN, BK>(st, Ahi, Ahi + plA, P_BM, lboA, tid, want_lo, -1, 0);
        convert_tile<B_MN, BK>(st + stA, Bhi, Bhi + plB, BN, lboB, tid, want_lo,
                               (ones_local >= 0 && ones_local < BN) ? ones_local : -1, kend - (kbeg + c * BK));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&conv_bar[p]));       // this warp's share of the planes is written and fenced
        mbar_arrive(smem_u32(&empty_bar[s]));      // ... and its share of the stage has been read
      }
    }
    if (nchunks > 0) mbar_wait(smem_u32(&accum_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every stage has been converted and every MMA has completed: the ring is idle and serves as the transpose scratch
    if (!(ta.dbg & 4)) tc_epilogue(ta, tmem_base, reinterpret_cast<float*>(smem), warp, lane, m0, n0, nchunks > 0, ones_col, 2);
  } else if (warp == 8) {
    // ================================ TMA PRODUCER (one thread) ================================
    if (lane == 0) {
      const uint32_t stage0 = smem_u32(smem);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S;
        if (c >= S) mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)((c / S - 1) & 1));
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t sa = stage0 + s * stage_bytes, sb = sa + stA;
        const int k0 = kbeg + c * BK;
        mbar_expect_tx(fb, (uint32_t)stage_bytes);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < P_BM / 32; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, m0 + 32 * j, k0, fb);
        } else {
          tma_load_2d(sa, &tmA, k0, m0, fb);
        }
        if (B_MN) {
          for (int j = 0; j < BNb / 32; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, n0 + 32 * j, k0, fb);
        } else {
          tma_load_2d(sb, &tmB, k0, n0, fb);
        }
      }
    }
  } else if (lane == 0) {
    // ================================ MMA ISSUE (one thread) ================================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(P_BM >> 4) << 24);
    const uint32_t planes0 = smem_u32(planes);
    for (int c = 0; c < nchunks; ++c) {
      const int p = c % NP;
      mbar_wait(smem_u32(&conv_bar[p]), (uint32_t)((c / NP) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH = planes0 + p * pset, aL = aH + plA, bH = aH + 2 * plA, bL = bH + plB;
#pragma unroll
      for (int kk = 0; kk < BK / 16; ++kk) {
        const uint32_t ao = kk * 2 * lboA, bo = kk * 2 * lboB;
        const uint64_t dAh = make_smem_desc(aH + ao, lboA, 128), dBh = make_smem_desc(bH + bo, lboB, 128);
        umma_bf16(tmem_base, dAh, dBh, idesc, (c > 0 || kk > 0) ? 1u : 0u);
        if (want_lo) {
          const uint64_t dAl = make_smem_desc(aL + ao, lboA, 128), dBl = make_smem_desc(bL + bo, lboB, 128);
          umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
          umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&pfree_bar[p]));               // plane set p may be rewritten once these MMAs have read it
    }
    if (nchunks > 0) umma_commit(smem_u32(&accum_bar));    // all MMAs complete: the accumulator is final
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ta.tmem_cols)
                 : "memory");
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int BK>
static size_t tcp_smem_bytes(int BN, bool b_mn) {
  const int BNb = b_mn ? round_up(BN, 32) : BN;
  const size_t stage = (size_t)P_BM * BK * 4 + (size_t)BNb * BK * 4;
  const size_t pset = 2 * (size_t)PCfg<BK>::SLABS * (P_BM * 16 + PCfg<BK>::PAD) + 2 * (size_t)PCfg<BK>::SLABS * (BN * 16 + PCfg<BK>::PAD);
  size_t tot = PCfg<BK>::STAGES * stage + PCfg<BK>::PSETS * pset;
  if (tot < TC_EPI_SCRATCH_BYTES) tot = TC_EPI_SCRATCH_BYTES;
  return tot + 1024;                                     // slack to align the ring to the 1024 B swizzle atom
}

// ---- tensor maps (driver entry point resolved at run time: the library does not link libcuda) --------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}
// fp32 matrix view [outer, inner] with row pitch ld floats; box [box_outer][box_inner]
static bool make_map(CUtensorMap* tm, const float* base, long long ld, int inner, int outer, int box_inner, int box_outer,
                     bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int MODE, int BK>
static int tcp_launch_one(const TcArgs& ta, dim3 grid, cudaStream_t st) {
  constexpr bool A_MN = (MODE == MFM_GEMM_TN), B_MN = (MODE != MFM_GEMM_NT);
  const GemmArgs& g = ta.g;
  CUtensorMap tmA, tmB;
  bool ok = A_MN ? make_map(&tmA, g.A, g.lda, g.M, g.K, 32, BK, true) : make_map(&tmA, g.A, g.lda, g.K, g.M, BK, P_BM, false);
  ok = ok && (B_MN ? make_map(&tmB, g.B, g.ldb, g.N, g.K, 32, BK, true) : make_map(&tmB, g.B, g.ldb, g.K, g.N, BK, ta.BN, false));
  if (!ok) return MFM_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcp_kernel<MODE, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tcp_smem_bytes<BK>(P_MAXBN, B_MN));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  gemm_tcp_kernel<MODE, BK><<<grid, P_THREADS, tcp_smem_bytes<BK>(ta.BN, B_MN), st>>>(ta, tmA, tmB);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

// TMA needs 16 B-aligned bases and row pitches; extents are free (tails are zero-filled by the TMA unit).  The virtual
// ones column of the fused bias gradient must start a fresh 16 B group, and a split-K range must not end inside a chunk
// (both hold for the launcher's own choices; N % 4 is checked here).
bool gemm_tcp_eligible(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MFM_TCP"); enabled = e ? atoi(e) : 1; }
  if (!enabled || !get_encode()) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 3) || (reinterpret_cast<uintptr_t>(B) & 15) || (ldb & 3)) return false;
  return true;
}

int gemm_tcp_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                    const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                    const long long* rng, float* colsum_out, cudaStream_t st) {
  TcArgs ta;
  ta.g = GemmArgs{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
                  drop_p, drop_site, rng, K, 0, colsum_out};
  ta.passes = passes;
  static int dbg = -1, bk = -1;
  if (dbg < 0) { const char* e = getenv("MFM_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  if (bk < 0) { const char* e = getenv("MFM_TCP_BK"); bk = e ? atoi(e) : 16; }
  ta.dbg = dbg;
  // tile N: near-equal tiles of at most 128 columns (+1 virtual ones column for colsum_out)
  const int n16 = round_up(N + (colsum_out ? 1 : 0), 16);
  const int ntiles = (n16 + P_MAXBN - 1) / P_MAXBN;
  ta.BN = round_up((n16 + ntiles - 1) / ntiles, 16);
  int cols = 32;
  while (cols < ta.BN) cols <<= 1;
  ta.tmem_cols = cols;
  dim3 grid((N + (colsum_out ? 1 : 0) + ta.BN - 1) / ta.BN, (M + P_BM - 1) / P_BM, 1);
  const bool plain = !bias && !bias2 && act == MFM_ACT_NONE && !mask && drop_p <= 0.0f && accumulate;
  if (plain && K >= 2048) {                       // split-K weight gradients: fill both CTA slots of every SM
    long long tiles = (long long)grid.x * grid.y;
    int splits = (int)((2 * 148 + tiles - 1) / tiles);
    int maxs = K / 256;
    if (splits > maxs) splits = maxs;
    if (splits > 1) {
      int kc = round_up((K + splits - 1) / splits, 32);
      ta.g.kchunk = kc;
      ta.g.atomic = 1;
      grid.z = (K + kc - 1) / kc;
    }
  }
  if (bk == 32) {
    switch (mode) {
      case MFM_GEMM_NT: return tcp_launch_one<MFM_GEMM_NT, 32>(ta, grid, st);
      case MFM_GEMM_NN: return tcp_launch_one<MFM_GEMM_NN, 32>(ta, grid, st);
      case MFM_GEMM_TN: return tcp_launch_one<MFM_GEMM_TN, 32>(ta, grid, st);
    }
  } else {
    switch (mode) {
      case MFM_GEMM_NT: return tcp_launch_one<MFM_GEMM_NT, 16>(ta, grid, st);
      case MFM_GEMM_NN: return tcp_launch_one<MFM_GEMM_NN, 16>(ta, grid, st);
      case MFM_GEMM_TN: return tcp_launch_one<MFM_GEMM_TN, 16>(ta, grid, st);
    }
  }
  return MFM_ERR_ARG;
}
