// Persistent streaming GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] · B)  with M = T*B rows of activations (tens of
// thousands) against a small weight (N, K <= ~512) -- the shape of every time-parallel layer of the step.
//
// gemm_tcp.cu runs one CTA per output tile: barrier set-up, TMEM allocation, the load-latency ramp (1-1.5 us under load),
// a short K loop (8..25 chunks) and a 7 000-cycle epilogue follow each other, and two co-resident CTAs only partly hide one
// another's phases (scripts/gemm_trace.py: a 128x144x128 tile lives 16 000 cycles for 2 500 cycles' worth of traffic).
// Here ONE CTA per SM stays resident and walks its tiles; its roles never drain:
//
//   warp 12 (one thread)   PRODUCER    ring of S stages of K = 32: TMA box of raw fp32 A (128 rows x 32 k, 16 KB) and -- streaming
//                                      mode -- one bulk copy of the stage's two pre-split weight images; runs ahead across tile
//                                      boundaries, so the next tiles' loads are in flight while these are multiplied and stored
//   warps 0..7             CONVERTERS  fp32 -> bf16 hi/lo, tcgen05.st into the stage's 32 tensor-memory columns (TS form)
//   warps 13, 14 (1 thread each)  MMA  hi*hi + lo*hi + hi*lo.  ONE thread sustains one tcgen05.mma per ~105 cycles whatever its
//                                      N (measured), so the CTA's tiles are taken in pairs: issuer p multiplies tile 2P+p into
//                                      accumulator p, the stages of the two tiles alternate in the ring
//   warps 8..11            EPILOGUE    tile i is drained while tiles i+1, i+2 are multiplied: tcgen05.ld (lane = row, 32 columns) ->
//                                      bias / activation / dropout in registers -> swizzled shared-memory box -> TMA store
//                                      (cp.async.bulk.tensor, 4 KB per instruction; M and N tails are clipped by the TMA unit).
//                                      The accumulator is handed back as soon as its last column block sits in registers.
//
// Tiles are (128 rows) x (BN columns), BN a multiple of 32 and <= 192, the last N tile possibly narrower.
//   resident mode   (the tile's whole weight image fits next to the ring; OFF by default, MFM_PS_RES=2 lets the model choose):
//                   CTA b serves N tile b % NT for every (gridDim.x / NT)-th row tile and loads the image ONCE -- no weight
//                   traffic in the loop, 16 KB stages; measured slower than streaming, see ps_plan()
//   streaming mode  tiles (row tile, N tile), N fastest, dealt round-robin; every stage carries its weight images (they are
//                   re-read from L2 by every row tile: as many bytes as A when BN = 128)
// ps_plan() picks N tiling and mode per shape with a small cost model.
// Tensor memory: 2 x BN accumulator columns + 32 S operand columns <= 512 (one CTA per SM owns all of it).
// Scope (else the caller uses gemm_tcp.cu): NT / NN with a pre-split B image (weights), M >= 4096, 16 B-aligned A and C
// with leading dimensions that are multiples of 4, N a multiple of 4 (the TMA store clips whole 16 B units), epilogue =
// bias (+ bias2) + activation + dropout, or -- on the plain product -- ONE operand tile (TMA-loaded one block ahead, same box and
// swizzle as the output): ReLU mask, accumulate into C, or the fused MSE head (residual, its square sum, its gradient).
#include <cstdio>
#include "tcp_shared.cuh"

#define PS_MAXRING 12
#define PS_BK 32                     // k per ring stage = two MMA sub-blocks of 16
#define PS_STA (P_BM * PS_BK * 4)    // raw fp32 A box of a stage: 16 KB
#define PS_NCONV 256
#define PS_WEPI 8                    // first epilogue warp (8..11: warp % 4 = TMEM lane quarter)
#define PS_WPROD 12
#define PS_WMMA 13                   // and 14: one MMA-issuing thread per tile of a pair
#define PS_THREADS 480
#define PS_STAG_BYTES (4 * 2 * 4096) // per epilogue warp two 32x32 fp32 boxes
#define PS_MAXN 1024
#define PS_SMEM_BUDGET (225 * 1024)

struct PsArgs {
  GemmArgs g;
  const unsigned char* bimg;
  int nck;             // 16-k sub-blocks (one weight image each)
  int nst;             // ring stages per tile = ceil(K / 32)
  int BN, BN_last, NT; // N tiling
  int MT;              // row tiles
  int resident;        // 1: each CTA serves ONE N tile and keeps that tile's whole weight image in shared memory
  int S, stage_bytes;
  int passes;
  int off_res, off_stag, off_aux, off_bias, off_bar;
  int aux;             // epilogue operand tile, TMA-loaded beside the output box: 0 none, 1 ReLU mask (C = mask > 0 ? v * scale : 0),
                       // 2 accumulate (C += v; the operand is C itself), 3 fused MSE head (operand x: C = grad_scale (v - x))
  long long* trace;    // PS_DEBUG builds only
};

// The CTA's i-th tile.  Streaming: tiles (row tile, N tile), N fastest, dealt round-robin.  Resident: CTA b serves N tile
// b % NT and every (gridDim.x / NT)-th row tile.
struct PsWalk {
  int NT, MT, res, first, step;
  __device__ __forceinline__ explicit PsWalk(const PsArgs& pa) : NT(pa.NT), MT(pa.MT), res(pa.resident) {
    first = res ? (int)blockIdx.x / NT : (int)blockIdx.x;
    step = res ? (int)gridDim.x / NT : (int)gridDim.x;
  }
  __device__ __forceinline__ bool get(int i, int& mt, int& j) const {
    if (res) { mt = first + i * step; j = (int)blockIdx.x % NT; return mt < MT; }
    const int t = first + i * step;
    mt = t / NT; j = t - mt * NT;
    return t < MT * NT;
  }
  __device__ __forceinline__ int count() const {           // tiles of this CTA
    const int total = res ? MT : MT * NT;
    return first < total ? (total - first + step - 1) / step : 0;
  }
};

// Debug trace (make EXTRA=-DPS_DEBUG=1; scripts/gemm_ps_trace.py): per CTA 64 chunks x 6 clock stamps (producer issue, converter
// warp 0 woke / arrived, converter warp 7 arrived, MMA woke / committed) + 16 tiles x 4 (epilogue woke, accumulator released,
// last store issued, MMA waited for the accumulator).
#ifndef PS_DEBUG
#define PS_DEBUG 0
#endif
#define PS_TR_CHUNKS 64
#define PS_TR_TILES 16
#define PS_TR_WORDS (4 + 6 * PS_TR_CHUNKS + 4 * PS_TR_TILES)
#if PS_DEBUG
static long long* g_ps_trace = nullptr;
#define PS_STAMP_C(g, slot) do { if (pa.trace && (g) < PS_TR_CHUNKS) pa.trace[(size_t)blockIdx.x * PS_TR_WORDS + 4 + 6 * (g) + (slot)] = clock64() - pa_t0; } while (0)
#define PS_STAMP_T(i, slot) do { if (pa.trace && (i) < PS_TR_TILES) pa.trace[(size_t)blockIdx.x * PS_TR_WORDS + 4 + 6 * PS_TR_CHUNKS + 4 * (i) + (slot)] = clock64() - pa_t0; } while (0)
#else
#define PS_STAMP_C(g, slot) do { } while (0)
#define PS_STAMP_T(i, slot) do { } while (0)
#endif
extern "C" int mfm_debug_set_gemm_ps_trace(void* buf) {
#if PS_DEBUG
  g_ps_trace = static_cast<long long*>(buf);
  return MFM_OK;
#else
  (void)buf;
  return MFM_ERR_UNSUPPORTED;
#endif
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

// Converter of one K=32 stage.  Thread (warp w, lane l) owns row 32*(w%4)+l (= its TMEM lane) and sub-block w/4 (16 k): it reads
// its 64 B of the raw box (128 B rows, 128 B TMA swizzle: 16 B unit ^= row % 8 -- conflict-free), splits fp32 -> bf16 hi + lo and
// writes 8 + 8 packed registers with tcgen05.st.  TMEM image of a stage (32 columns): [hi sub 0 | hi sub 1 | lo sub 0 | lo sub 1].
__device__ __forceinline__ void ps_convert(const unsigned char* st, uint32_t tcol, int warp, int lane, bool want_lo) {
  const int sub = warp >> 2, r = 32 * (warp & 3) + lane;
  const unsigned char* row = st + r * 128;
  const int sw = r & 7;
  uint32_t h[8], l[8];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 v = *reinterpret_cast<const float4*>(row + (((4 * sub + u) ^ sw) << 4));
    h[2 * u] = pack_bf16(v.x, v.y);
    h[2 * u + 1] = pack_bf16(v.z, v.w);
    l[2 * u] = pack_bf16(v.x - __uint_as_float(h[2 * u] << 16), v.y - __uint_as_float(h[2 * u] & 0xFFFF0000u));
    l[2 * u + 1] = pack_bf16(v.z - __uint_as_float(h[2 * u + 1] << 16), v.w - __uint_as_float(h[2 * u + 1] & 0xFFFF0000u));
  }
  const uint32_t ta = tcol + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(sub * 8);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(h[0]), "r"(h[1]), "r"(h[2]),
               "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
  if (want_lo)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta + 16), "r"(l[0]), "r"(l[1]),
                 "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
}

#define PS_AUX_MASK 1
#define PS_AUX_ACC 2
#define PS_AUX_MSE 3
template <int ACT, int AUX>
__device__ __forceinline__ void ps_epilogue(const PsArgs& pa, const CUtensorMap* tmC, const CUtensorMap* tmX, uint32_t tmem_base,
                                            unsigned char* smem, unsigned long long* acc_full, unsigned long long* acc_free,
                                            unsigned long long* aux_bar_all, int ew, int lane, long long pa_t0) {
  const GemmArgs& a = pa.g;
  // AUX: the operand tile of block k+1 is TMA-loaded (same 32x32 box, same 128 B swizzle as the output box; rows / columns
  // beyond M / N are zero-filled) into one of two buffers while block k is processed
  const uint32_t auxb = AUX ? smem_u32(smem + pa.off_aux) + (uint32_t)ew * 8192u : 0u;
  unsigned long long* const aux_bar = aux_bar_all + ew * 2;
  float sq = 0.0f;
  const float* sbias = reinterpret_cast<const float*>(smem + pa.off_bias);
  const uint32_t stag = smem_u32(smem + pa.off_stag) + (uint32_t)ew * 8192u;
  const bool do_drop = a.drop_p > 0.0f;
  uint32_t sseed = 0;
  if (do_drop) sseed = site_seed(a.rng, a.drop_site);
  const float keep_scale = do_drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);
  const uint32_t swz = (uint32_t)(lane & 7);
  int blk = 0;
  const PsWalk walk(pa);
  if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmC)) : "memory");
  if (AUX && lane == 0) {
    int mt0, j0;
    if (walk.get(0, mt0, j0)) {
      const uint32_t ab = smem_u32(&aux_bar[0]);
      mbar_expect_tx(ab, 4096u);
      tma_load_2d(auxb, tmX, j0 * pa.BN, mt0 * P_BM + ew * 32, ab);
    }
  }
  for (int i = 0, mt, j; walk.get(i, mt, j); ++i) {
    const int bn = (j == pa.NT - 1) ? pa.BN_last : pa.BN;
    const int n0 = j * pa.BN, mrow = mt * P_BM + ew * 32;
    const int buf = i & 1;
    mbar_wait(smem_u32(&acc_full[buf]), (uint32_t)((i >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ew == 0 && lane == 0) PS_STAMP_T(i, 0);
    for (int c0 = 0; c0 < bn; c0 += 32, ++blk) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tlane + (uint32_t)(buf * pa.BN + c0))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (c0 + 32 >= bn) {                           // this warp holds its last block of the accumulator: hand it back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_free[buf]));
        if (ew == 0 && lane == 0) PS_STAMP_T(i, 1);
      }
      const uint32_t sb = stag + (uint32_t)(blk & 1) * 4096u;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this box is done
      __syncwarp();
      const int n = n0 + c0;
      const uint32_t e0 = (uint32_t)(mrow + lane) * (uint32_t)a.N + (uint32_t)n;
      uint32_t ax = 0u;
      if (AUX) {
        int ni = i, nc0 = c0 + 32, nmt = mt, nj = j;          // the next block of this warp: same tile, or the next tile's first
        bool have = true;
        if (nc0 >= bn) { ni = i + 1; nc0 = 0; have = walk.get(ni, nmt, nj); }
        if (have && lane == 0) {                               // (every lane finished reading that buffer two blocks ago: the
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     //  __syncwarp above orders it before this issue)
          const uint32_t ab = smem_u32(&aux_bar[(blk + 1) & 1]);
          mbar_expect_tx(ab, 4096u);
          tma_load_2d(auxb + (uint32_t)((blk + 1) & 1) * 4096u, tmX, nj * pa.BN + nc0, nmt * P_BM + ew * 32, ab);
        }
        mbar_wait(smem_u32(&aux_bar[blk & 1]), (uint32_t)((blk >> 1) & 1));
        ax = auxb + (uint32_t)(blk & 1) * 4096u + (uint32_t)lane * 128u;
      }
      const bool row_ok = mrow + lane < a.M;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + n + 4 * q);
        float4 v;
        v.x = act_t<ACT>(__uint_as_float(r[4 * q]) + b4.x);
        v.y = act_t<ACT>(__uint_as_float(r[4 * q + 1]) + b4.y);
        v.z = act_t<ACT>(__uint_as_float(r[4 * q + 2]) + b4.z);
        v.w = act_t<ACT>(__uint_as_float(r[4 * q + 3]) + b4.w);
        if (AUX) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"(ax + (((uint32_t)q ^ swz) << 4)));
          if (AUX == PS_AUX_MASK) {
            v.x = x.x > 0.0f ? v.x * a.mask_scale : 0.0f; v.y = x.y > 0.0f ? v.y * a.mask_scale : 0.0f;
            v.z = x.z > 0.0f ? v.z * a.mask_scale : 0.0f; v.w = x.w > 0.0f ? v.w * a.mask_scale : 0.0f;
          } else if (AUX == PS_AUX_ACC) {
            v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
          } else {                                             // MSE head: residual, its square (valid elements only), its gradient
            v.x -= x.x; v.y -= x.y; v.z -= x.z; v.w -= x.w;
            if (row_ok && n + 4 * q < a.N) sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
            v.x *= a.mse.grad_scale; v.y *= a.mse.grad_scale; v.z *= a.mse.grad_scale; v.w *= a.mse.grad_scale;
          }
        }
        if (do_drop) {
          v.x = drop_keep(sseed, e0 + 4 * q, a.drop_p) ? v.x * keep_scale : 0.0f;
          v.y = drop_keep(sseed, e0 + 4 * q + 1, a.drop_p) ? v.y * keep_scale : 0.0f;
          v.z = drop_keep(sseed, e0 + 4 * q + 2, a.drop_p) ? v.z * keep_scale : 0.0f;
          v.w = drop_keep(sseed, e0 + 4 * q + 3, a.drop_p) ? v.w * keep_scale : 0.0f;
        }
        // 128 B swizzle of the TMA box: 16 B chunk q of row `lane` lives at chunk q ^ (lane % 8)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (uint32_t)lane * 128u + (((uint32_t)q ^ swz) << 4)),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, n, mrow, sb);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (ew == 0 && lane == 0) PS_STAMP_T(i, 2);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
  if (AUX == PS_AUX_MSE) {                          // one atomic per warp
    sq = warp_sum(sq);
    if (lane == 0) atomicAdd(a.mse.slot, sq * a.mse.loss_scale);
  }
}

template <bool B_MN>
__global__ void __launch_bounds__(PS_THREADS, 1) gemm_ps_kernel(const PsArgs pa, const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmC,
                                                                const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const GemmArgs& a = pa.g;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int S = pa.S;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem + pa.off_bar);
  unsigned long long* const full_bar = bars;                      // [S] stage landed
  unsigned long long* const conv_bar = bars + PS_MAXRING;         // [S] stage converted (8 warps)
  unsigned long long* const done_bar = bars + 2 * PS_MAXRING;     // [S] MMAs of the stage complete
  unsigned long long* const acc_full = bars + 3 * PS_MAXRING;     // [2] accumulator final
  unsigned long long* const acc_free = bars + 3 * PS_MAXRING + 2; // [2] accumulator drained (4 epilogue warps)
  unsigned long long* const res_bar = bars + 3 * PS_MAXRING + 4;  // resident weight image landed
  uint32_t& tmem_holder = *reinterpret_cast<uint32_t*>(bars + 3 * PS_MAXRING + 5);
  unsigned long long* const aux_bar = bars + 3 * PS_MAXRING + 6;  // [4 epilogue warps][2] epilogue operand tile landed
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), PS_NCONV / 32);
      mbar_init(smem_u32(&done_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&acc_free[b]), 4);
    }
    mbar_init(smem_u32(res_bar), 1);
    for (int b = 0; b < 8; ++b) mbar_init(smem_u32(&aux_bar[b]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    float* sbias = reinterpret_cast<float*>(smem + pa.off_bias);
    const int npad = (pa.NT - 1) * pa.BN + pa.BN_last;
    for (int n = tid; n < npad; n += PS_THREADS)
      sbias[n] = n < a.N ? (a.bias ? __ldg(a.bias + n) : 0.0f) + (a.bias2 ? __ldg(a.bias2 + n) : 0.0f) : 0.0f;
  }
  if (warp == PS_WMMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const long long pa_t0 = PS_DEBUG ? clock64() : 0;
  (void)pa_t0;
#if PS_DEBUG
  if (pa.trace && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    pa.trace[(size_t)blockIdx.x * PS_TR_WORDS] = gt;
    pa.trace[(size_t)blockIdx.x * PS_TR_WORDS + 1] = smid;
  }
#endif
  const uint32_t acol0 = (uint32_t)(2 * pa.BN);
  const int stA = PS_STA;
  const bool want_lo = pa.passes == 3;

  if (warp < PS_NCONV / 32) {
    // ================================ CONVERTERS ================================
    int s = 0, g = 0;
    uint32_t ph = 0;
    (void)g;
    const PsWalk walk(pa);
    const int nstages = walk.count() * pa.nst;
    for (; g < nstages; ++g) {
      mbar_wait(smem_u32(&full_bar[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) PS_STAMP_C(g, 1);
      ps_convert(smem + s * pa.stage_bytes, tmem_base + acol0 + (uint32_t)(s * 32), warp, lane, want_lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&conv_bar[s]));
      if (tid == 0) PS_STAMP_C(g, 2);
      if (tid == 7 * 32) PS_STAMP_C(g, 3);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else if (warp < PS_WPROD) {
    // ================================ EPILOGUE ================================
    const int ew = warp - PS_WEPI;
    if (pa.aux == PS_AUX_MASK)     ps_epilogue<MFM_ACT_NONE, PS_AUX_MASK>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0);
    else if (pa.aux == PS_AUX_ACC) ps_epilogue<MFM_ACT_NONE, PS_AUX_ACC>(pa, &tmC, &tmC, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0);
    else if (pa.aux == PS_AUX_MSE) ps_epilogue<MFM_ACT_NONE, PS_AUX_MSE>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0);
    else switch (a.act) {                          // CTA-uniform
      case MFM_ACT_RELU: ps_epilogue<MFM_ACT_RELU, 0>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0); break;
      case MFM_ACT_TANH: ps_epilogue<MFM_ACT_TANH, 0>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0); break;
      case MFM_ACT_SIGMOID: ps_epilogue<MFM_ACT_SIGMOID, 0>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0); break;
      default: ps_epilogue<MFM_ACT_NONE, 0>(pa, &tmC, &tmX, tmem_base, smem, acc_full, acc_free, aux_bar, ew, lane, pa_t0); break;
    }
  } else if (warp == PS_WPROD) {
    // ================================ PRODUCER (one thread) ================================
    // Stage order: the CTA's tiles are taken in PAIRS (2P, 2P+1) whose K stages alternate -- (2P, c), (2P+1, c), (2P, c+1), ... --
    // because each tile of a pair has its own MMA-issuing thread and accumulator; an unpaired last tile runs alone.
    if (lane == 0) {
      const uint32_t stage0 = smem_u32(smem);
      const size_t img_tile = (size_t)pa.nck * (size_t)(2 * P_SLABS * (pa.BN * 16 + P_PAD));
      int s = 0, g = 0;
      uint32_t ph = 0;
      const PsWalk walk(pa);
      const int ntl = walk.count();
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      for (int i0 = 0; i0 < ntl; i0 += 2) {
        const int np = min(2, ntl - i0);
        int mt[2] = {0, 0}, jn[2] = {0, 0};
        walk.get(i0, mt[0], jn[0]);
        if (np == 2) walk.get(i0 + 1, mt[1], jn[1]);
        for (int c = 0; c < pa.nst; ++c) {
          for (int q = 0; q < np; ++q, ++g) {
            const int bn = (jn[q] == pa.NT - 1) ? pa.BN_last : pa.BN;
            const uint32_t img_bytes = pa.resident ? 0u : (uint32_t)(2 * P_SLABS * (bn * 16 + P_PAD));
            const unsigned char* img = pa.bimg + (size_t)jn[q] * img_tile;
            mbar_wait(smem_u32(&done_bar[s]), ph ^ 1u);         // the MMAs that read this stage S stages ago are complete
            const uint32_t fb = smem_u32(&full_bar[s]);
            const uint32_t sa = stage0 + s * pa.stage_bytes;
            const uint32_t bbytes = img_bytes * (uint32_t)min(2, pa.nck - 2 * c);   // the stage's one or two weight images
            mbar_expect_tx(fb, (uint32_t)stA + bbytes);
            tma_load_2d(sa, &tmA, c * PS_BK, mt[q] * P_BM, fb);
            if (bbytes) bulk_load(sa + stA, img + (size_t)(2 * c) * img_bytes, bbytes, fb);
            PS_STAMP_C(g, 0);
            if (++s == S) { s = 0; ph ^= 1u; }
            if (pa.resident && g + 1 == min(S, ntl * pa.nst)) {
              // the ring is primed (A first: the converters start on it); now this CTA's N tile, the whole weight image, once
              const int bnr = (jn[0] == pa.NT - 1) ? pa.BN_last : pa.BN;
              const uint32_t ib = (uint32_t)(2 * P_SLABS * (bnr * 16 + P_PAD));
              const unsigned char* im = pa.bimg + (size_t)jn[0] * img_tile;
              const uint32_t rb = smem_u32(res_bar);
              mbar_expect_tx(rb, ib * (uint32_t)pa.nck);
              for (int cc = 0; cc < pa.nck; ++cc) bulk_load(stage0 + pa.off_res + (uint32_t)cc * ib, im + (size_t)cc * ib, ib, rb);
            }
          }
        }
      }
    }
  } else if (lane == 0) {
    // ================================ MMA ISSUE (two threads: warps 13 and 14) ================================
    // One thread sustains one tcgen05.mma per ~105 cycles whatever its N (measured: scripts/gemm_ps_trace.py), three per 16 k.
    // Issuer p multiplies the tiles i = p, p+2, ... into accumulator p; the two run side by side on the alternating stages.
    const int p = warp - PS_WMMA;
    const uint32_t stage0 = smem_u32(smem);
    const PsWalk walk(pa);
    const int ntl = walk.count();
    if (pa.resident && p < ntl) mbar_wait(smem_u32(res_bar), 0u);
    for (int i = p, mt, j; i < ntl; i += 2) {
      walk.get(i, mt, j);
      const int bn = (j == pa.NT - 1) ? pa.BN_last : pa.BN;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(bn >> 3) << 17) |
                             ((uint32_t)(P_BM >> 4) << 24);
      const int plB = P_SLABS * (bn * 16 + P_PAD);
      const int lboB = B_MN ? (bn / 8) * 128 : (bn * 16 + P_PAD);
      const uint32_t tD = tmem_base + (uint32_t)(p * pa.BN);
      const bool paired = (i | 1) < ntl;
      int g = (i & ~1) * pa.nst + (paired ? p : 0);
      mbar_wait(smem_u32(&acc_free[p]), (uint32_t)(((i >> 1) & 1) ^ 1));       // drained by the epilogue of tile i-2
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p == 0) PS_STAMP_T(i, 3);
      for (int c = 0; c < pa.nst; ++c, g += paired ? 2 : 1) {
        const int q = g / S, s = g - q * S;
        mbar_wait(smem_u32(&conv_bar[s]), (uint32_t)(q & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        PS_STAMP_C(g, 4);
        const int nsub = min(2, pa.nck - 2 * c);
        for (int sub = 0; sub < nsub; ++sub) {
          const uint32_t bH = pa.resident ? stage0 + pa.off_res + (uint32_t)((2 * c + sub) * 2 * plB)
                                          : stage0 + s * pa.stage_bytes + stA + (uint32_t)(sub * 2 * plB);
          const uint32_t bL = bH + plB;
          const uint64_t dBh = make_smem_desc(bH, lboB, 128), dBl = make_smem_desc(bL, lboB, 128);
          const uint32_t tAh = tmem_base + acol0 + (uint32_t)(s * 32 + sub * 8), tAl = tAh + 16;
          umma_bf16_ts(tD, tAh, dBh, idesc, (c > 0 || sub > 0) ? 1u : 0u);
          if (want_lo) {
            umma_bf16_ts(tD, tAl, dBh, idesc, 1u);
            umma_bf16_ts(tD, tAh, dBl, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&done_bar[s]));
        PS_STAMP_C(g, 5);
      }
      umma_commit(smem_u32(&acc_full[p]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#if PS_DEBUG
  if (pa.trace && tid == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    pa.trace[(size_t)blockIdx.x * PS_TR_WORDS + 2] = gt;
    pa.trace[(size_t)blockIdx.x * PS_TR_WORDS + 3] = clock64() - pa_t0;
  }
#endif
  if (warp == PS_WMMA) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int g_ps_launches = 0;
extern "C" int mfm_debug_gemm_ps_count(void) { return g_ps_launches; }

// Tiling of N and residency, by a small cost model (cycles per SM): tiles are near-equal multiples of 32 columns (<= 224: two
// accumulators + >= 4 operand stages in 512 TMEM columns), the last as narrow as N allows.  Per chunk a tile costs three
// MMAs of max(~110 cycles fixed, bn/2), 8 KB of A and -- streaming -- its weight image from L2 (~40 B/cycle/SM when every SM
// pulls), and the ring's round trip (~3 000 cycles) divided by its depth.  Env MFM_PS_NT / MFM_PS_RES pin the choice.
struct PsPlan { int BN, BN_last, NT, resident, S, stage_bytes, off_res, off_stag, off_aux, off_bias, off_bar; double cost; };
static bool ps_plan_one(int MT, int N, int nck, int sms, int nt, int resident, bool aux, PsPlan* pl) {
  const int bn = round_up((N + nt - 1) / nt, 32);
  if (bn > 224 || (long long)bn * (nt - 1) >= N || nt > sms) return false;
  const int bn_last = round_up(N - bn * (nt - 1), 32);
  const int img = 2 * P_SLABS * (bn * 16 + P_PAD);
  const int stage = resident ? PS_STA : ((PS_STA + 2 * img + 1023) & ~1023);
  const int npad = (nt - 1) * bn + bn_last;
  const int res_bytes = resident ? ((nck * img + 1023) & ~1023) : 0;
  const int aux_bytes = aux ? PS_STAG_BYTES : 0;           // two operand boxes per epilogue warp
  const int fixed = res_bytes + PS_STAG_BYTES + aux_bytes + round_up(npad * 4, 128) + 512;
  int S = (PS_SMEM_BUDGET - fixed) / stage;
  const int by_tmem = (512 - 2 * bn) / 32;
  if (S > by_tmem) S = by_tmem;
  if (S > PS_MAXRING) S = PS_MAXRING;
  if (PS_SMEM_BUDGET < fixed || S < 3) return false;
  const double rounds = resident ? (double)((MT + sms / nt - 1) / (sms / nt)) : (double)(((long long)MT * nt + sms - 1) / sms);
  const double mma = rounds * nck * (3.0 * (bn / 2 > 60 ? bn / 2 : 60) + 110.0);
  const double l2 = (rounds * (nck * (8192.0 + (resident ? 0 : img)) + 128.0 * bn * 4) + (resident ? (double)nck * img : 0.0)) / 40.0;
  const double lat = rounds * (nck / 2.0) * 3000.0 / S;
  double cost = mma > l2 ? mma : l2;
  if (lat > cost) cost = lat;
  pl->BN = bn; pl->BN_last = bn_last; pl->NT = nt; pl->resident = resident; pl->S = S; pl->stage_bytes = stage;
  pl->off_res = S * stage;
  pl->off_stag = pl->off_res + res_bytes;
  pl->off_aux = pl->off_stag + PS_STAG_BYTES;
  pl->off_bias = pl->off_aux + aux_bytes;
  pl->off_bar = pl->off_bias + round_up(npad * 4, 128);
  pl->cost = cost;
  return true;
}
static int g_ps_res = -1;                 // residency: 0 streaming only (default), 1 resident only, 2 by the cost model
extern "C" int mfm_debug_gemm_ps_residency(int mode) {
  if (mode < 0 || mode > 2) return MFM_ERR_ARG;
  g_ps_res = mode;
  return MFM_OK;
}
static bool ps_plan(int M, int N, int K, int sms, bool aux, PsPlan* best) {
  static int pin_nt = -1;
  int& pin_res = g_ps_res;
  if (pin_nt < 0) { const char* e = getenv("MFM_PS_NT"); pin_nt = e ? atoi(e) : 0; }
  // default 0 = streaming only: the resident mode's model cost is lower but it measured SLOWER, alone (36.7 vs 34.2 us on
  // 40960x400x128) and in the step (all-resident 2.337 ms, model's choice 2.243, all-streaming 2.221, same box): a CTA pinned
  // to one N tile starts with a ~100 KB image load and the row tiles divide unevenly over (SMs / NT) CTAs.  2 = by the model.
  if (pin_res < 0) { const char* e = getenv("MFM_PS_RES"); pin_res = e ? atoi(e) : 0; }
  const int MT = (M + P_BM - 1) / P_BM, nck = (K + P_BK - 1) / P_BK;
  bool found = false;
  const int nt0 = (N + 223) / 224;
  for (int pass = 0; pass < 2 && !found; ++pass) {         // (resident only: shapes it does not hold are streamed)
    for (int nt = nt0; nt <= nt0 + 5; ++nt) {
      if (pin_nt > 0 && nt != pin_nt) continue;
      for (int res = 0; res < 2; ++res) {
        if (pass == 0 && pin_res < 2 && res != pin_res) continue;
        if (pass == 1 && res != 0) continue;
        PsPlan pl;
        if (!ps_plan_one(MT, N, nck, sms, nt, res, aux, &pl)) continue;
        if (!found || pl.cost < best->cost) { *best = pl; found = true; }
      }
    }
  }
  return found;
}

int gemm_ps_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                   float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate, const float* mask,
                   long long ldmask, float mask_scale, float drop_p, int drop_site, const long long* rng, const GemmMse& mse,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MFM_PS"); enabled = e ? atoi(e) : 1; }
  if (!enabled || mode == MFM_GEMM_TN || M < 4096 || N > PS_MAXN || N < 1 || (N & 3) || K < 1 || !ws) return MFM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 3) || (reinterpret_cast<uintptr_t>(ws) & 127)) return MFM_ERR_UNSUPPORTED;
  // one epilogue operand at most, and only on the plain (no activation, no dropout) product; x_hat is not written here
  const int naux = (mask ? 1 : 0) + (accumulate ? 1 : 0) + (mse.x ? 1 : 0);
  if (naux > 1 || (naux && (act != MFM_ACT_NONE || drop_p > 0.0f)) || (mse.x && mse.xhat)) return MFM_ERR_UNSUPPORTED;
  const float* auxp = mask ? mask : mse.x ? mse.x : nullptr;
  const long long auxld = mask ? ldmask : mse.x ? mse.ldx : 0;
  if (auxp && ((reinterpret_cast<uintptr_t>(auxp) & 15) || (auxld & 3))) return MFM_ERR_UNSUPPORTED;
  PsArgs pa;
  pa.g = GemmArgs{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale, drop_p, drop_site, rng,
                  K, 0, nullptr};
  pa.g.mse = mse;
  pa.aux = mask ? PS_AUX_MASK : accumulate ? PS_AUX_ACC : mse.x ? PS_AUX_MSE : 0;
  pa.passes = passes;
#if PS_DEBUG
  pa.trace = g_ps_trace;
#else
  pa.trace = nullptr;
#endif
  const int sms = mfm_dev_info().sms;
  PsPlan pl;
  if (!ps_plan(M, N, K, sms, naux > 0, &pl)) return MFM_ERR_UNSUPPORTED;
  static int verbose = -1;
  if (verbose < 0) { const char* e = getenv("MFM_PS_VERBOSE"); verbose = e ? atoi(e) : 0; }
  if (verbose) fprintf(stderr, "gemm_ps %dx%dx%d: NT %d BN %d/%d resident %d S %d cost %.0f\n", M, N, K, pl.NT, pl.BN, pl.BN_last, pl.resident, pl.S, pl.cost);
  pa.BN = pl.BN; pa.BN_last = pl.BN_last; pa.NT = pl.NT; pa.resident = pl.resident; pa.S = pl.S; pa.stage_bytes = pl.stage_bytes;
  pa.off_res = pl.off_res; pa.off_stag = pl.off_stag; pa.off_aux = pl.off_aux; pa.off_bias = pl.off_bias; pa.off_bar = pl.off_bar;
  pa.nck = (K + P_BK - 1) / P_BK;
  pa.nst = (K + PS_BK - 1) / PS_BK;
  pa.MT = (M + P_BM - 1) / P_BM;
  const size_t img_full = (size_t)pa.nck * 2 * P_SLABS * (pa.BN * 16 + P_PAD);
  const size_t img_last = (size_t)pa.nck * 2 * P_SLABS * (pa.BN_last * 16 + P_PAD);
  if (img_full * (pa.NT - 1) + img_last > ws_bytes) return MFM_ERR_UNSUPPORTED;
  pa.bimg = static_cast<const unsigned char*>(ws);
  const size_t smem = (size_t)pa.off_bar + 512;
  CUtensorMap tmA, tmC, tmX;
  if (!make_map(&tmA, A, lda, K, M, PS_BK, P_BM) || !make_map(&tmC, C, ldc, N, M, 32, 32)) return MFM_ERR_UNSUPPORTED;
  tmX = tmC;
  if (auxp && !make_map(&tmX, auxp, auxld, N, M, 32, 32)) return MFM_ERR_UNSUPPORTED;
  dim3 pg(pa.NT, pa.nck, 1);
  if (mode == MFM_GEMM_NT) {
    if (int e = mfm_func_smem_t(gemm_ps_kernel<false>, PS_SMEM_BUDGET)) return e;
    gemm_prep_kernel<false><<<pg, 256, 0, st>>>(B, ldb, N, K, pa.BN, pa.nck, passes == 3, static_cast<unsigned char*>(ws), pa.BN_last);
  } else {
    if (int e = mfm_func_smem_t(gemm_ps_kernel<true>, PS_SMEM_BUDGET)) return e;
    gemm_prep_kernel<true><<<pg, 256, 0, st>>>(B, ldb, N, K, pa.BN, pa.nck, passes == 3, static_cast<unsigned char*>(ws), pa.BN_last);
  }
  MFM_LAUNCH_CHECK();
  int grid;
  if (pa.resident) {
    const int per = sms / pa.NT < pa.MT ? sms / pa.NT : pa.MT;
    grid = per * pa.NT;
  } else {
    const long long tiles = (long long)pa.MT * pa.NT;
    grid = tiles < sms ? (int)tiles : sms;
  }
  if (mode == MFM_GEMM_NT) gemm_ps_kernel<false><<<grid, PS_THREADS, smem, st>>>(pa, tmA, tmC, tmX);
  else                     gemm_ps_kernel<true><<<grid, PS_THREADS, smem, st>>>(pa, tmA, tmC, tmX);
  MFM_LAUNCH_CHECK();
  ++g_ps_launches;
  return MFM_OK;
}
