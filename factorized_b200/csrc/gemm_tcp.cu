// Pipelined tcgen05 GEMM for sm_100a (the streaming form of gemm_tc.cu; same math, same epilogue, same three layouts).
//
// The step's time-parallel GEMMs are [T*B, <=512] activations against small weights: they are bound by streaming the
// activation operand from HBM.  gemm_tc.cu keeps one K chunk of prefetch in registers and runs all its threads in
// lock step (load -> convert -> fence -> barrier -> MMA), so every chunk exposes the whole latency chain.  Here the
// links of that chain are different warps connected by mbarriers over ONE ring of 4..7 stages (BK = 16), the
// transport is TMA, and the A operand never touches shared memory again after it has been split:
//
//   warp 8 (one thread)  PRODUCER    waits done[s]; cp.async.bulk.tensor.2d: the raw fp32 box of A for chunk c -> stage s
//                                    (the tensor map describes the operand as it lies in HBM, any 16 B-aligned view; M/N/K
//                                    tails are zero-filled by the TMA unit), and B either the same way or -- when the
//                                    caller supplied a workspace (weights: NT / NN) -- as ONE bulk copy of the chunk's
//                                    pre-split bf16 image, already in MMA layout.  Completion: complete_tx on full[s].
//   warps 0..7           CONVERTERS  wait full[s]; thread (row, 8-k slab) reads its 32 B of the raw A box, splits fp32 ->
//                                    bf16 hi + lo and stores both with tcgen05.st into the stage's 16 tensor-memory columns
//                                    (TMEM lane = tile row); a raw B box is split into shared-memory planes in the tcgen05
//                                    no-swizzle canonical layout (+ fence.proxy.async).  One arrive per warp on conv[s];
//                                    warps run ahead independently (no CTA-wide barrier in the loop).
//   warp 9 (one thread)  MMA         wait conv[s]; tcgen05.mma in the TS form -- A from tensor memory, B from shared memory --
//                                    hi*hi + lo*hi + hi*lo (M=128, N=BN, K=16) into the fp32 TMEM accumulator; one
//                                    tcgen05.commit -> done[s] hands the stage (shared memory and TMEM columns) back.
//   warps 0..7           EPILOGUE    after the last commit: tc_epilogue.cuh (TMEM -> transposed through the idle ring ->
//                                    512 B coalesced stores with bias / activation / dropout / mask / split-K reduction).
//
// Pre-split B (gemm_prep_kernel): a weight is the same for all T*B/128 row tiles, so splitting it inside every CTA is
// T*B/128-fold redundant and costs as much as splitting A.  One small kernel writes, per (N tile, K chunk), the image
// [hi plane | lo plane] exactly as the MMA wants it in shared memory; a stage is then 8 KB of raw A + the image.
// Why TS: with both operands split into shared-memory planes the shared-memory pipe saturated (TMA write + converter read
// + plane write + 3 x MMA operand reads = 57 KB per chunk, ~119 of 128 B/cycle with two CTAs per SM).
//
// Two CTAs of 320 threads are resident per SM (<= 110 KB and 256 TMEM columns each, no static shared memory), so one
// CTA's epilogue stores overlap the other's loads.  Measured alternatives that lost: three CTAs with four converter warps,
// 16 B cp.async transport (L1-bypassing LDGSTS fetches every sector twice and writes one wavefront per lane), separate raw /
// plane rings, two converter groups on alternate chunks.  Raw layouts:
//   K-major operand  [rows, K]:  one box [rows][16] fp32, 64 B TMA swizzle (16 B chunk ^= (row/2)%4): a warp reads the
//                                first / second 16 B half of 32 consecutive rows conflict-free
//   MN-major operand [K, cols]:  boxes of [16][32 cols], 128 B TMA swizzle (chunk ^= k%8): a warp reads one k row of its
//                                32 columns (A), or the eight k rows of a core matrix from eight bank groups (B)
//   B planes: K-major byte slab*LBO + r*16 (LBO = rows*16+64), MN-major byte (k/8)*LBO + (c/8)*128 + (k%8)*16
// Requirements (else the caller falls back to gemm_tc.cu): A 16 B aligned with a leading dimension that is a multiple
// of 4; the same for a raw B (a pre-split B may have any alignment).
#include "tcp_shared.cuh"

// ---- converter: one raw operand tile of a stage -> split-bf16 planes ---------------------------------------------------
// width = rows (K-major) or columns (MN-major) of the tile, a multiple of 16 and <= 128.
template <bool MN>
__device__ __forceinline__ void convert_tile(const unsigned char* st, unsigned char* hi, unsigned char* lo, int width,
                                             int lbo, int tid, bool want_lo, int ones_local, int kvalid) {
  float v[8];
#pragma unroll
  for (int u = 0; u < 256 / P_NCONV; ++u) {                 // up to 256 items of 32 B per tile
    const int idx = u * P_NCONV + tid;
    if (!MN) {
      const int r = idx >> 1, slab = idx & 1;               // rows x 2 slabs
      if (r < width) {
        const unsigned char* row = st + r * (P_BK * 4);
        const int sw = (r >> 1) & 3;                        // 64 B swizzle: chunk ^= (byte address >> 7) & 3
        lds8(row + (((2 * slab) ^ sw) << 4), row + (((2 * slab + 1) ^ sw) << 4), v);
        split_store(v, hi + slab * lbo + r * 16, lo + slab * lbo + r * 16, want_lo);
      }
    } else {
      const int klo = idx & 7, g = idx >> 3, ng = width >> 3;   // 16 k x (width/8) column groups
      const int mg = g % ng, khi = g / ng;
      if (khi < P_SLABS) {
        const int k = khi * 8 + klo;
        const unsigned char* row = st + (mg >> 2) * (P_BK * 128) + k * 128;    // box of 32 columns, 128 B rows
        const int ch = (mg & 3) * 2;                        // 128 B swizzle: chunk ^= k % 8
        lds8(row + ((ch ^ klo) << 4), row + (((ch + 1) ^ klo) << 4), v);
        const int off = khi * lbo + mg * 128 + klo * 16;
        split_store(v, hi + off, lo + off, want_lo);
        if (ones_local >= 0 && (ones_local >> 3) == mg && k < kvalid)          // virtual ones column (fused bias gradient):
          *reinterpret_cast<unsigned short*>(hi + off + 2 * (ones_local & 7)) = 0x3F80;   // bf16(1.0); its lo part stays 0
      }
    }
  }
}

// debug trace: per CTA 4 header words (globaltimer at start, smid, nchunks, globaltimer at end) + per chunk 6 stamps
// (clock64 relative to CTA start): producer issue, converter full-wait done, converter arrive, MMA conv-wait done, MMA commit,
// producer pfree-wait done.  One writer thread per role; enabled only through mfm_debug_set_gemm_trace.
#define P_TRACE_CHUNKS 32
#define P_TRACE_WORDS (4 + 6 * P_TRACE_CHUNKS)
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct TcpArgs {
  TcArgs t;
  TcArgs t2;                     // pair form (TN): N tiles >= ntiles1 multiply the same A by a second B into a second C
  int ntiles1;                   // 0 = plain GEMM
  const unsigned char* bimg;     // BPRE: images [N tile][K chunk][hi plane | lo plane]
  int nchunks_total;             // BPRE: K chunks per N tile in bimg
  int bar_off;                   // byte offset of the mbarrier block inside the dynamic shared buffer
  int nstages;                   // ring depth
  int group;                     // chunks the MMA thread issues per mbarrier wait (<= nstages / 2)
  long long* trace;              // debug (mfm_debug_set_gemm_trace): per-CTA clock64 stamps, see scripts/gemm_trace.py
};

template <int MODE, bool BPRE>
__global__ void __launch_bounds__(P_THREADS, 2) gemm_tcp_kernel(const TcpArgs pa, const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmB2) {
  constexpr bool A_MN = (MODE == MFM_GEMM_TN);
  constexpr bool B_MN = (MODE != MFM_GEMM_NT);
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned bx, by, bz;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
  asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
  const bool second = MODE == MFM_GEMM_TN && pa.ntiles1 > 0 && (int)bx >= pa.ntiles1;      // CTA-uniform
  const TcArgs& ta = second ? pa.t2 : pa.t;
  const CUtensorMap* const tmBsel = second ? &tmB2 : &tmB;
  if (second) bx -= (unsigned)pa.ntiles1;
  const GemmArgs& a = ta.g;
  const int BN = ta.BN, S = pa.nstages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = by * P_BM, n0 = bx * BN;
  const int kbeg = bz * a.kchunk;
  const int kend = min(a.K, kbeg + a.kchunk);
  const int nchunks = (kend - kbeg + P_BK - 1) / P_BK;
  const bool want_lo = ta.passes == 3;
  // one ring, S deep.  Stage s = [raw A | pre-split B image]  or  [raw A | raw B | B hi | B lo] in shared memory (a
  // multiple of the 1024 B swizzle atom) plus 16 tensor-memory columns (A hi 8 | A lo 8) behind the accumulator.
  const int BNb = B_MN ? ((BN + 31) & ~31) : BN;          // raw MN-major B arrives in boxes of 32 columns
  const int plB = P_SLABS * (BN * 16 + P_PAD);
  const int stA = P_BM * P_BK * 4, stB = BPRE ? 0 : BNb * P_BK * 4;
  const int offB = stA + stB;                               // B planes (hi | lo): the landed image, or written by the converters
  const int stage_bytes = (offB + 2 * plB + 1023) & ~1023;
  const int lboB = B_MN ? (BN / 8) * 128 : (BN * 16 + P_PAD);
  const uint32_t acol0 = (uint32_t)((BN + 15) & ~15);
  const int ones_col = (MODE == MFM_GEMM_TN && a.colsum_out) ? a.N : -1;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem + pa.bar_off);
  unsigned long long* const full_bar = bars;                     // [S] stage landed (TMA / bulk complete_tx)
  unsigned long long* const conv_bar = bars + P_MAXRING;         // [S] stage converted by every converter warp
  unsigned long long* const done_bar = bars + 2 * P_MAXRING;     // [S] the MMAs that read stage s are complete
  unsigned long long& accum_bar = bars[3 * P_MAXRING];
  uint32_t& tmem_holder = *reinterpret_cast<uint32_t*>(bars + 3 * P_MAXRING + 1);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), P_NCONV / 32);
      mbar_init(smem_u32(&done_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P_WMMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)ta.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  long long* const tr = pa.trace ? pa.trace + (size_t)((bz * gridDim.y + by) * gridDim.x + bx) * P_TRACE_WORDS : nullptr;
  const long long t0 = clock64();
  if (tr && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[0] = gtimer(); tr[1] = smid; tr[2] = nchunks;
    (void)bx;
  }
#define TRACE(cc, slot) do { if (tr && (cc) < P_TRACE_CHUNKS) tr[4 + 6 * (cc) + (slot)] = clock64() - t0; } while (0)

  if (warp < P_NCONV / 32) {
    // ================================ CONVERTERS, then EPILOGUE ================================
    // full[s] also means "the MMAs of chunk c-S are complete" (the producer waited for done[s] before refilling the
    // stage), so the stage's tensor-memory columns and B planes may be overwritten without a further wait.
    const int ones_local = (ones_col >= 0 && ones_col - n0 >= 0 && ones_col - n0 < BN) ? ones_col - n0 : -1;
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      unsigned char* st = smem + s * stage_bytes;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) TRACE(c, 1);
      if (!(ta.dbg & 2)) {
        convert_a_tmem<A_MN>(st, tmem_base + acol0 + (uint32_t)(s * 16), warp, lane, want_lo);
        if (!BPRE) {
          convert_tile<B_MN>(st + stA, st + offB, st + offB + plB, BN, lboB, tid, want_lo, ones_local, kend - (kbeg + c * P_BK));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&conv_bar[s]));
      if (tid == 0) TRACE(c, 2);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (nchunks > 0) mbar_wait(smem_u32(&accum_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && nchunks < 31) TRACE(31, 0);
    // every chunk has been converted and multiplied: the ring is idle and serves as the transpose scratch
    if (!(ta.dbg & 4)) tc_epilogue(ta, tmem_base, reinterpret_cast<float*>(smem), warp, lane, m0, n0, nchunks > 0, ones_col, P_NCONV / 128,
                                   (tr && nchunks < 30) ? tr + 4 + 6 * 30 : nullptr);
    if (tid == 0 && nchunks < 31) TRACE(31, 1);
  } else if (warp == P_WPROD) {
    // ================================ PRODUCER (one thread) ================================
    if (lane == 0) {
      const uint32_t stage0 = smem_u32(smem);
      const unsigned char* img = BPRE ? pa.bimg + ((size_t)bx * pa.nchunks_total + kbeg / P_BK) * (size_t)(2 * plB) : nullptr;
      const uint32_t tx = (uint32_t)(stA + (BPRE ? 2 * plB : stB));
      int s = 0;
      uint32_t ph = 0;
      for (int c = 0; c < nchunks; ++c) {
        if (c >= S) mbar_wait(smem_u32(&done_bar[s]), ph ^ 1u);             // the MMAs of chunk c-S are complete
        TRACE(c, 5);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t sa = stage0 + s * stage_bytes, sb = sa + stA;
        const int k0 = kbeg + c * P_BK;
        mbar_expect_tx(fb, tx);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < P_BM / 32; ++j) tma_load_2d(sa + j * (P_BK * 128), &tmA, m0 + 32 * j, k0, fb);
        } else {
          tma_load_2d(sa, &tmA, k0, m0, fb);
        }
        if (BPRE) {
          bulk_load(sa + offB, img + (size_t)c * (2 * plB), (uint32_t)(2 * plB), fb);
        } else if (B_MN) {
          for (int j = 0; j < BNb / 32; ++j) tma_load_2d(sb + j * (P_BK * 128), tmBsel, n0 + 32 * j, k0, fb);
        } else {
          tma_load_2d(sb, tmBsel, k0, n0, fb);
        }
        TRACE(c, 0);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (lane == 0) {
    // ================================ MMA ISSUE (one thread) ================================
    // Its loop is the serial spine of the kernel: one wait, three MMAs, one commit per chunk.
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, A from TMEM (K-major), B major, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(P_BM >> 4) << 24);
    const uint32_t stage0 = smem_u32(smem);
    const int G = pa.group;                      // chunks per wait: converter warps finish chunks in order, so "chunk c+G-1
    int s = 0;                                   // converted" implies chunks c..c+G-2 are too -- one mbarrier wait per G chunks
    uint32_t ph = 0;
    for (int c0 = 0; c0 < nchunks; c0 += G) {
      const int g = min(G, nchunks - c0);
      {
        int sl = s + g - 1;
        uint32_t pl = ph;
        if (sl >= S) { sl -= S; pl ^= 1u; }
        mbar_wait(smem_u32(&conv_bar[sl]), pl);  // converters saw full[] before arriving: the B images are acquired transitively
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = c0; c < c0 + g; ++c) {
        TRACE(c, 3);
        const uint32_t bH = stage0 + s * stage_bytes + offB, bL = bH + plB;
        const uint64_t dBh = make_smem_desc(bH, lboB, 128), dBl = make_smem_desc(bL, lboB, 128);
        const uint32_t tAh = tmem_base + acol0 + (uint32_t)(s * 16), tAl = tAh + 8;
        umma_bf16_ts(tmem_base, tAh, dBh, idesc, c > 0 ? 1u : 0u);
        if (want_lo) {
          umma_bf16_ts(tmem_base, tAl, dBh, idesc, 1u);
          umma_bf16_ts(tmem_base, tAh, dBl, idesc, 1u);
        }
        umma_commit(smem_u32(&done_bar[s]));              // stage s (shared memory and its TMEM columns) may be refilled
        TRACE(c, 4);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
    if (nchunks > 0) umma_commit(smem_u32(&accum_bar));    // all MMAs complete: the accumulator is final
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && nchunks < 31) TRACE(31, 2);
#undef TRACE
  if (tr && tid == 0) tr[3] = gtimer();
  if (warp == P_WMMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ta.tmem_cols)
                 : "memory");
  }
}

static long long* g_trace = nullptr;
static long long g_trace_words = 0;
extern "C" int mfm_debug_set_gemm_trace(void* buf, long long bytes) {
  g_trace = static_cast<long long*>(buf);
  g_trace_words = buf ? bytes / 8 : 0;
  return MFM_OK;
}


#define P_RING_BUDGET (110 * 1024)   // two CTAs per SM
// Split-K weight gradients: CTAs per SM the K splits are sized for.  Two are resident, but every split pays an epilogue of
// 128 x BN atomics and the gradients run beside the critical chain: measured over the whole step (same box) 4 -> 2.327 ms,
// 2 -> 2.286, 1.5 -> 2.259, 1 -> 2.267.  env MFM_TN_FILL overrides.
static double tn_fill() {
  static double v = -1.0;
  if (v < 0.0) { const char* e = getenv("MFM_TN_FILL"); v = e ? atof(e) : 1.5; if (v <= 0.0) v = 1.5; }
  return v;
}
struct RingCfg { int S; size_t bytes; int tmem_cols; };
static RingCfg tcp_ring(int BN, bool b_mn, bool bpre) {
  const int BNb = b_mn ? round_up(BN, 32) : BN;
  const size_t plB = (size_t)P_SLABS * (BN * 16 + P_PAD);
  size_t stage = (size_t)P_BM * P_BK * 4 + (bpre ? 0 : (size_t)BNb * P_BK * 4) + 2 * plB;
  stage = (stage + 1023) & ~(size_t)1023;
  RingCfg c;
  c.S = (int)(P_RING_BUDGET / stage);
  // each stage also owns 16 tensor-memory columns behind the accumulator; 256 columns per CTA (two CTAs share the 512)
  const int by_tmem = (256 - round_up(BN, 16)) / 16;
  if (c.S > by_tmem) c.S = by_tmem;
  if (c.S > P_MAXRING) c.S = P_MAXRING;
  if (c.S < 2) c.S = 2;
  c.bytes = c.S * stage;
  const size_t scratch = (size_t)(P_NCONV / 32) * 32 * TC_EPI_LD * 4;
  if (c.bytes < scratch) c.bytes = scratch;
  c.tmem_cols = 32;
  while (c.tmem_cols < round_up(BN, 16) + 16 * c.S) c.tmem_cols <<= 1;
  return c;
}


template <int MODE, bool BPRE>
static int tcp_launch_one(const TcpArgs& pa, dim3 grid, cudaStream_t st) {
  constexpr bool A_MN = (MODE == MFM_GEMM_TN), B_MN = (MODE != MFM_GEMM_NT);
  const GemmArgs& g = pa.t.g;
  CUtensorMap tmA, tmB;
  bool ok = A_MN ? make_map(&tmA, g.A, g.lda, g.M, g.K, 32, P_BK) : make_map(&tmA, g.A, g.lda, g.K, g.M, P_BK, P_BM);
  if (BPRE) tmB = tmA;
  else ok = ok && (B_MN ? make_map(&tmB, g.B, g.ldb, g.N, g.K, 32, P_BK) : make_map(&tmB, g.B, g.ldb, g.K, g.N, P_BK, pa.t.BN));
  if (!ok) return MFM_ERR_UNSUPPORTED;
  if (int e = mfm_func_smem_t(gemm_tcp_kernel<MODE, BPRE>, (int)(P_RING_BUDGET + P_BAR_BYTES))) return e;
  TcpArgs pb = pa;
  pb.trace = ((long long)grid.x * grid.y * grid.z * P_TRACE_WORDS <= g_trace_words) ? g_trace : nullptr;
  const RingCfg rc = tcp_ring(pa.t.BN, B_MN, BPRE);
  pb.nstages = rc.S;
  static int grp = -1;
  if (grp < 0) { const char* e = getenv("MFM_TCP_G"); grp = e ? atoi(e) : 2; }
  pb.group = grp < 1 ? 1 : (grp > rc.S / 2 ? (rc.S / 2 > 0 ? rc.S / 2 : 1) : grp);
  pb.bar_off = (int)rc.bytes;
  pb.t.tmem_cols = pb.t2.tmem_cols = rc.tmem_cols;
  CUtensorMap tmB2 = tmB;
  if (MODE == MFM_GEMM_TN && pa.ntiles1 > 0 && !make_map(&tmB2, pa.t2.g.B, pa.t2.g.ldb, pa.t2.g.N, pa.t2.g.K, 32, P_BK))
    return MFM_ERR_UNSUPPORTED;
  gemm_tcp_kernel<MODE, BPRE><<<grid, P_THREADS, (size_t)pb.bar_off + P_BAR_BYTES, st>>>(pb, tmA, tmB, tmB2);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

// TMA needs 16 B-aligned bases and row pitches; extents are free (tails are zero-filled by the TMA unit).
bool gemm_tcp_eligible(int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MFM_TCP"); enabled = e ? atoi(e) : 1; }
  if (!enabled || !get_encode()) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (lda & 3)) return false;
  return true;
}

int gemm_ps_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                   float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate, const float* mask,
                   long long ldmask, float mask_scale, float drop_p, int drop_site, const long long* rng, const GemmMse& mse,
                   void* ws, size_t ws_bytes, cudaStream_t st);

int gemm_tcp_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                    float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                    const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                    const long long* rng, float* colsum_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcpArgs pa;
  TcArgs& ta = pa.t;
  ta.g = GemmArgs{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
                  drop_p, drop_site, rng, K, 0, colsum_out};
  ta.g.mse = g_pending_mse;          // (cleared by mfm_gemm_mse after the call)
  ta.passes = passes;
  static int dbg = -1, pre = -1, maxbn_pre = -1;
  if (dbg < 0) { const char* e = getenv("MFM_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  if (pre < 0) { const char* e = getenv("MFM_TCP_PRE"); pre = e ? atoi(e) : 1; }
  if (maxbn_pre < 0) { const char* e = getenv("MFM_TCP_MAXBN"); maxbn_pre = e ? atoi(e) : P_MAXBN_PRE; }
  ta.dbg = dbg;
  pa.bimg = nullptr;
  pa.nchunks_total = 0;
  pa.trace = nullptr;
  pa.ntiles1 = 0;
  pa.t2 = pa.t;
  const bool plain = !bias && !bias2 && act == MFM_ACT_NONE && !mask && drop_p <= 0.0f && accumulate;
  const bool splitk = plain && K >= 2048;
  // pre-split B: weights (NT / NN) against many row tiles, caller-provided workspace, no split-K
  const int nck = (K + P_BK - 1) / P_BK;
  bool bpre = pre && mode != MFM_GEMM_TN && !splitk && ws && M >= 4096 && ((reinterpret_cast<uintptr_t>(ws) & 127) == 0);
  const bool b_ok = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 3) == 0);
  if (!bpre && !b_ok) return MFM_ERR_UNSUPPORTED;         // raw B goes through TMA: needs the alignment
  const int maxbn = bpre ? maxbn_pre : P_MAXBN_RAW;
  // tile N: near-equal tiles of at most maxbn columns (+1 virtual ones column for colsum_out)
  const int n16 = round_up(N + (colsum_out ? 1 : 0), 16);
  const int ntiles = (n16 + maxbn - 1) / maxbn;
  ta.BN = round_up((n16 + ntiles - 1) / ntiles, 16);
  if (bpre && !colsum_out) {                                                   // persistent kernel (gemm_ps.cu) where it applies
    const int rc = gemm_ps_launch(passes, mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask,
                                  mask_scale, drop_p, drop_site, rng, g_pending_mse, ws, ws_bytes, st);
    if (rc != MFM_ERR_UNSUPPORTED) return rc;
  }
  if (bpre) {
    const size_t need = (size_t)ntiles * nck * 2 * P_SLABS * (ta.BN * 16 + P_PAD);
    if (need > ws_bytes) {
      if (!b_ok) return MFM_ERR_UNSUPPORTED;
      return gemm_tcp_launch(passes, mode, M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask,
                             mask_scale, drop_p, drop_site, rng, colsum_out, nullptr, 0, st);
    }
  }
  ta.tmem_cols = 0;                                  // set by tcp_launch_one (accumulator + A sets)
  dim3 grid((N + (colsum_out ? 1 : 0) + ta.BN - 1) / ta.BN, (M + P_BM - 1) / P_BM, 1);
  if (splitk) {                                   // split-K weight gradients: one split per resident CTA slot
    long long tiles = (long long)grid.x * grid.y;
    const double occ = tn_fill();
    int splits = (int)((long long)(occ * mfm_dev_info().sms + tiles - 1) / tiles);
    if (splits < 1) splits = 1;
    int maxs = K / 256;
    if (splits > maxs) splits = maxs;
    if (splits > 1) {
      int kc = round_up((K + splits - 1) / splits, 32);
      ta.g.kchunk = kc;
      ta.g.atomic = 1;
      grid.z = (K + kc - 1) / kc;
    }
  }
  if (bpre) {
    pa.bimg = static_cast<const unsigned char*>(ws);
    pa.nchunks_total = nck;
    dim3 pg(grid.x, nck, 1);
    if (mode == MFM_GEMM_NT) gemm_prep_kernel<false><<<pg, 256, 0, st>>>(B, ldb, N, K, ta.BN, nck, passes == 3, static_cast<unsigned char*>(ws));
    else                     gemm_prep_kernel<true><<<pg, 256, 0, st>>>(B, ldb, N, K, ta.BN, nck, passes == 3, static_cast<unsigned char*>(ws));
    MFM_LAUNCH_CHECK();
    return mode == MFM_GEMM_NT ? tcp_launch_one<MFM_GEMM_NT, true>(pa, grid, st) : tcp_launch_one<MFM_GEMM_NN, true>(pa, grid, st);
  }
  switch (mode) {
    case MFM_GEMM_NT: return tcp_launch_one<MFM_GEMM_NT, false>(pa, grid, st);
    case MFM_GEMM_NN: return tcp_launch_one<MFM_GEMM_NN, false>(pa, grid, st);
    case MFM_GEMM_TN: return tcp_launch_one<MFM_GEMM_TN, false>(pa, grid, st);
  }
  return MFM_ERR_ARG;
}

// C1[M,N1] += A^T B1 (+ colsum1[m] += sum_k A[k,m]) and C2[M,N2] += A^T B2 in ONE launch: the two weight gradients of a
// cell (dW_ih = dG^T x, dW_hh = dG^T h_prev) share their big operand, which is then streamed from HBM once -- the N
// tiles of the second product run next to those of the first and find A's chunks in L2.
int gemm_tcp_launch_tn_pair(int passes, int M, int K, const float* A, long long lda, int N1, const float* B1, long long ldb1,
                            float* C1, long long ldc1, float* colsum1, int N2, const float* B2, long long ldb2, float* C2,
                            long long ldc2, cudaStream_t st) {
  auto ok16 = [](const float* p, long long ld) { return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 3) == 0); };
  if (!gemm_tcp_eligible(MFM_GEMM_TN, M, N1, K, A, lda, B1, ldb1) || !ok16(B1, ldb1) || !ok16(B2, ldb2)) return MFM_ERR_UNSUPPORTED;
  TcpArgs pa;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("MFM_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  const int n1 = N1 + (colsum1 ? 1 : 0);
  const int nmax = n1 > N2 ? n1 : N2;
  const int n16 = round_up(nmax, 16);
  const int nt = (n16 + P_MAXBN_RAW - 1) / P_MAXBN_RAW;
  const int BN = round_up((n16 + nt - 1) / nt, 16);
  const int tiles1 = (n1 + BN - 1) / BN, tiles2 = (N2 + BN - 1) / BN;
  pa.t.g = GemmArgs{M, N1, K, A, lda, B1, ldb1, C1, ldc1, nullptr, nullptr, MFM_ACT_NONE, 1, nullptr, 0, 1.0f, 0.0f, 0, nullptr,
                    K, 0, colsum1};
  pa.t.passes = passes; pa.t.dbg = dbg; pa.t.BN = BN; pa.t.tmem_cols = 0;
  pa.t2 = pa.t;
  pa.t2.g.N = N2; pa.t2.g.B = B2; pa.t2.g.ldb = ldb2; pa.t2.g.C = C2; pa.t2.g.ldc = ldc2; pa.t2.g.colsum_out = nullptr;
  pa.bimg = nullptr; pa.nchunks_total = 0; pa.trace = nullptr;
  pa.ntiles1 = tiles1;
  dim3 grid(tiles1 + tiles2, (M + P_BM - 1) / P_BM, 1);
  if (K >= 2048) {
    long long tiles = (long long)grid.x * grid.y;
    int splits = (int)((long long)(tn_fill() * mfm_dev_info().sms + tiles - 1) / tiles);
    if (splits < 1) splits = 1;
    int maxs = K / 256;
    if (splits > maxs) splits = maxs;
    if (splits > 1) {
      int kc = round_up((K + splits - 1) / splits, 32);
      pa.t.g.kchunk = pa.t2.g.kchunk = kc;
      pa.t.g.atomic = pa.t2.g.atomic = 1;
      grid.z = (K + kc - 1) / kc;
    }
  }
  return tcp_launch_one<MFM_GEMM_TN, false>(pa, grid, st);
}
