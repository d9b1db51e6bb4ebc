// LSTM recurrences on the tensor cores, warp-specialised (mfm_model.py:56,83,85,167-169 and their adjoint).
//
// A recurrence step is a strict chain: gate GEMM -> TMEM -> activations -> h_t -> next gate GEMM.  One chain cannot keep
// an SM busy, so a CTA owns TWO chains -- batch sub-tiles A and B of the same cell, sharing the resident weight image --
// and alternates: all sixteen compute warps run the cell update of A while the tensor core runs the gate GEMM of B, then
// swap.  What the timeline of the earlier versions showed (scripts/lstm_trace.py) and this layout answers:
//   * a warp that issues tcgen05.mma blocks for the whole GEMM (the MMA queue is shallow: ~90 cycles per instruction
//     when two issuers compete), so the issue loop belongs to a warp that has no other work: warp 16, alone in a fifth
//     warpgroup that gives its registers away (setmaxnreg: 24 there, 112 for the compute warps);
//   * two chains that each own half of the warps drift into lock step and their GEMMs are never hidden; with all
//     compute warps on one chain at a time the alternation is enforced by construction;
//   * consecutive MMAs into one accumulator serialise at the pipeline latency: the four gates (forward) / four partial
//     sums over K (backward) are separate accumulators and every pass goes to all of them in turn;
//   * the cell update is issue-bound: addresses are (64-bit running base of the step) + (32-bit column offset), rows
//     beyond B and lanes beyond h are clamped onto valid elements and only their stores are predicated (no divergent
//     branch in the loop), two logistic values share one reciprocal, and the warp index is made provably uniform.
//
//   warp 16 (one thread)   per step and chain: wait ready[c] (h_{t-1} / dG_t of the chain is in shared memory);
//                          tcgen05.mma hi*hi + lo*hi + hi*lo over K; tcgen05.commit -> done[c].
//   warps 0..15            per step, chain A then chain B: wait done[c]; tcgen05.ld; cell update; stash; h_t -> shared
//                          memory (split bf16, K-major B operand); fence.proxy.async; arrive ready[c].
//
// The gate GEMM is issued TRANSPOSED, one MMA tile PER GATE:   D_g[unit j, batch b] = W_g[j, :] . h_{t-1}[b, :]
//   A = W_g (split-bf16, resident in shared memory, K-major = W's own layout), M = 128 rows = hidden units;
//   B = h_{t-1} of the chain's NB batch rows;  D_g = NB TMEM columns, lane = unit.
// A thread owns a TMEM lane, i.e. ONE hidden unit, and finds all four gates of that unit in its own lane (columns
// g*NB + b): the cell update needs no exchange between threads, and because lane = unit, the 32 lanes of a warp touch
// 32 consecutive floats of the row-major stashes (G_x, gates, c, h, dG): one 128 B line per access.
// Small cells would leave most of the 128 lanes idle, so the unit rows of the A tile are REPLICATED (h <= 32: 4 copies,
// h <= 64: 2): every lane quadrant then holds every unit and the quadrants split the batch columns instead (such cells
// take 64-row chains).  Warp w reads lane quadrant w%4 and the column part w/4.
//
// Backward:  dh^T[unit j, batch b] = sum_k' W^T[j, k'] dG[b, k'],  k' = 4*unit + gate (unit-major, so the four gate
// gradients a thread produces are 8 contiguous bytes of the B operand); same roles, same replication; the carried dc
// stays in registers.  Cells that do not fit on chip (h > ~108 forward, h > ~104 backward) run on lstm_seq.cu.
#include <cstdlib>
#include "tc_common.cuh"

// -DWS_DEBUG=1 compiles a clock-stamp trace of CTA 0 into the forward kernel (scripts/lstm_trace.py); off in production:
// its lane-divergent stores inside the step loop cost uniform registers
#ifndef WS_DEBUG
#define WS_DEBUG 0
#endif
#if WS_DEBUG
__device__ long long g_ws_trace_buf[17 * 32 * 8];          // [warp 0..16][step 32][8 stamps]
#define WS_STAMP(w, t, k) do { if (blockIdx.x == 0 && lane == 0 && (t) < 32) g_ws_trace_buf[((w) * 32 + (t)) * 8 + (k)] = clock64(); } while (0)
#else
#define WS_STAMP(w, t, k) do { } while (0)
#endif

#define WS_CW 16                          // compute warps
#define WS_THREADS ((WS_CW + 4) * 32)     // + the issuer's warpgroup (setmaxnreg works on warpgroups)
#define WS_MAXCHAIN 2

struct WsCell {
  mfm_lstm_cell c;
  int nb;        // batch rows per chain = UMMA N (16, 32 or 64)
  int nsub;      // lane quadrants one copy of the units occupies: max(2, ceil(h / 32))
  int nact;      // compute warps that own units (the rest idle): arrivals per chain and step
  int gs;        // forward: rows per gate block of the W image (128 when replicated, else h rounded up to 8)
  int kp;        // MMA K: forward h rounded up to 16; backward 4 * (h rounded up to 8)
  int lboA, lboB;
  int cta0;      // first blockIdx.x of this cell
};
struct WsBatch {
  WsCell c[MFM_MAX_CELLS];
  int n;
};

// wait on an mbarrier phase.  The suspend-time hint parks the warp until the phase completes instead of re-polling (the
// first version re-polled ~200 times per step: a third of all issued instructions); bounded: a barrier that never
// completes traps instead of hanging the GPU
#ifndef WS_SPIN
#define WS_SPIN 0
#endif
__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#if WS_SPIN
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void ws_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

static unsigned long long g_ws_counts[8];     // launches per variant, see mfm_debug_lstm_variant_count

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float v[2]) {
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r[0]);
  v[1] = __uint_as_float(r[1]);
}
__device__ __forceinline__ unsigned short bf16_bits(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<unsigned short*>(&b);
}
// 1 + 2^a with the exponent clamped from above (2^28: products of two such terms stay finite); no lower clamp is needed:
// ex2 underflows to 0 and the logistic saturates correctly.  Two logistic values then share ONE reciprocal:
//   1/(1+ea) = (1+eb) * rcp((1+ea)(1+eb)),   1/(1+eb) = (1+ea) * rcp(...)     (the SFU takes 8 cycles per warp instruction)
__device__ __forceinline__ float one_plus_ex2(float a) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(a, 28.0f)));
  return 1.0f + e;
}
// predicated 4-byte global store without a branch (a divergent `if` around the stash stores cost a BSSY region per item)
__device__ __forceinline__ void st_if(float* p, float v, int ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(ok));
}

// which units and batch columns of a chain a compute warp owns
struct WsRole {
  int q;        // TMEM lane quadrant (= warp % 4)
  int j;        // hidden unit of this lane
  bool on;      // lane has a unit
  bool warp_on; // warp has work
  int col0;     // first batch column of this warp inside the chain's NB
  int ncol;     // its columns: 2, 4 or 8
};
__device__ __forceinline__ WsRole ws_role(int w, int lane, int h, int nsub, int nb) {
  WsRole r;
  r.q = w & 3;
  const int part = w >> 2;
  if (nsub == 2) {                         // two copies: quadrant = (copy, sub-block); copies and parts split the columns
    const int sub = r.q & 1, rep = r.q >> 1;
    r.ncol = nb >> 3;
    r.j = sub * 32 + lane;
    r.col0 = (rep * 4 + part) * r.ncol;
    r.warp_on = sub * 32 < h;              // h <= 32: the second sub-block is empty
  } else {
    r.j = r.q * 32 + lane;
    r.ncol = nb >> 2;
    r.col0 = part * r.ncol;
    r.warp_on = r.q < nsub;
  }
  r.on = r.warp_on && r.j < h;
  return r;
}

__device__ __forceinline__ const WsCell& ws_find(const WsBatch& bt, int bx) {
  int ci = 0;
  while (ci + 1 < bt.n && bx >= bt.c[ci + 1].cta0) ++ci;
  return bt.c[ci];
}

#define WS_REG_COMPUTE() asm volatile("setmaxnreg.inc.sync.aligned.u32 112;")
#define WS_REG_ISSUER() asm volatile("setmaxnreg.dec.sync.aligned.u32 24;")

// ----------------------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------------------
template <int NCH, int NSG>      // chains per CTA; groups of 4 batch columns per warp and chain (2: 32-/64-row chains, 1: half as wide)
__global__ void __launch_bounds__(WS_THREADS, 1) lstm_ws_fwd_kernel(const __grid_constant__ WsBatch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_done[WS_MAXCHAIN], bar_ready[WS_MAXCHAIN];
  __shared__ uint32_t tmem_holder;
  const WsCell& wc = ws_find(bt, (int)blockIdx.x);
  const int h = wc.c.h, B = wc.c.B, T = wc.c.T, H4 = 4 * h, gx_steps = wc.c.gx_steps;
  const int NB = wc.nb, KP = wc.kp, GS = wc.gs, nsub = wc.nsub;
  const int slabs = KP >> 3;
  const int lboA = wc.lboA, lboH = wc.lboB;
  const int ldgx = wc.c.ld_gx ? (int)wc.c.ld_gx : H4, ldcs = (int)wc.c.ld_cs, ldhs = (int)wc.c.ld_hs;
  const float* __restrict__ const Wg = wc.c.W;
  const float* __restrict__ const gx_base = wc.c.gx;
  const float* __restrict__ const bias_rest = wc.c.bias_rest;
  float* __restrict__ const gates_base = wc.c.gates;
  float* __restrict__ const cs_base = wc.c.cs;
  float* __restrict__ const csd_base = wc.c.cs_dup;
  float* __restrict__ const hs_base = wc.c.hs;
  unsigned char* Whi = smem;
  unsigned char* Wlo = Whi + slabs * lboA;
  unsigned char* Hbase = Wlo + slabs * lboA;                 // per chain: [hi plane | lo plane]
  const int chainH = 2 * slabs * lboH;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches stay uniform
  const int row0 = ((int)blockIdx.x - wc.cta0) * (NCH * NB);
  int tmem_cols = 32;
  while (tmem_cols < NCH * 4 * NB) tmem_cols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < NCH; ++i) {
      mbar_init(smem_u32(&bar_done[i]), 4);                    // one commit per issuing warp
      mbar_init(smem_u32(&bar_ready[i]), wc.nact);             // compute warps that own units
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // W [4h,h] fp32 -> resident split-bf16 K-major image: gate g at rows g*GS.., unit rows replicated when nsub <= 2
  {
    const bool vecW = ((reinterpret_cast<uintptr_t>(Wg) & 15) == 0) && ((h & 3) == 0);
    const int span = nsub <= 2 ? 32 * nsub : GS;             // rows of one copy
    const int rows = 4 * GS;
    for (int idx = tid; idx < rows * slabs; idx += WS_THREADS) {
      const int slab = idx % slabs, r = idx / slabs;
      const int g = r / GS, j = (r - g * GS) % span;
      float v[8];
      load8(Wg, h, j < h ? g * h + j : H4, H4, slab * 8, h, vecW, v);
      split_store(v, Whi + slab * lboA + r * 16, Wlo + slab * lboA + r * 16, true);
    }
    for (int idx = tid * 16; idx < NCH * chainH; idx += WS_THREADS * 16)
      *reinterpret_cast<uint4*>(Hbase + idx) = make_uint4(0, 0, 0, 0);     // h_{-1} = 0, K padding = 0
  }
  // block 0 of the histories is the zero initial state
  for (int idx = tid; idx < NCH * NB * h; idx += WS_THREADS) {
    const int b = row0 + idx / h, j = idx % h;
    if (b < B) {
      hs_base[(long long)b * ldhs + j] = 0.0f;
      cs_base[(long long)b * ldcs + j] = 0.0f;
      if (csd_base) csd_base[(long long)b * ldcs + j] = 0.0f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp >= WS_CW) {
    // ================================ MMA issue (warps 16..19: one thread and one gate each) + L2 prefetch of G_x =========
    WS_REG_ISSUER();
    {
      // One issuing thread sustains only one tcgen05.mma per ~60 cycles (descriptor arithmetic + the instruction's own
      // latency; the trace showed the tensor core waiting for the issuer, not the other way round), and MMAs into one
      // accumulator serialise anyway: the four warps of this warpgroup issue ONE GATE each, concurrently.
      const int g = warp - WS_CW;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t gofs = (uint64_t)((g * GS * 16) >> 4);
      const uint64_t dAh0 = make_smem_desc(smem_u32(Whi), lboA, 128) + gofs, dAl0 = make_smem_desc(smem_u32(Wlo), lboA, 128) + gofs;
      const int ksteps = KP >> 4;
      const uint64_t astep = (uint64_t)((2 * lboA) >> 4), bstep = (uint64_t)((2 * lboH) >> 4);
      // The cell update's G_x reads are the long pole of a step (ncu: 90 % of the active warp samples waited on them, one
      // column group of register prefetch does not cover HBM latency under load).  These warps are idle between GEMMs, so
      // they pull the G_x rows of the chain's NEXT step into L2 a whole step ahead: one bulk prefetch per batch row and gate.
      const bool pf_ok = ((reinterpret_cast<uintptr_t>(gx_base) & 15) == 0) && ((ldgx & 3) == 0) && ((h & 3) == 0);
      for (int t = 0; t < T; ++t) {                            // step 0 multiplies the zero state like every other step
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          if (lane == 0) {
            if (t > 0) ws_wait(smem_u32(&bar_ready[ch]), (uint32_t)((t - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (g == 0) WS_STAMP(16, t, ch * 2);
            const uint32_t hb = smem_u32(Hbase + ch * chainH);
            const uint64_t dBh0 = make_smem_desc(hb, lboH, 128), dBl0 = make_smem_desc(hb + slabs * lboH, lboH, 128);
            const uint32_t dcol = tmem_base + (uint32_t)((ch * 4 + g) * NB);
            uint64_t ao = 0, bo = 0;
#pragma unroll 1
            for (int kk = 0; kk < ksteps; ++kk, ao += astep, bo += bstep) {
              umma_bf16(dcol, dAh0 + ao, dBh0 + bo, idesc, kk > 0 ? 1u : 0u);
              umma_bf16(dcol, dAl0 + ao, dBh0 + bo, idesc, 1u);
              umma_bf16(dcol, dAh0 + ao, dBl0 + bo, idesc, 1u);
            }
            umma_commit(smem_u32(&bar_done[ch]));
            if (g == 0) WS_STAMP(16, t, ch * 2 + 1);
          }
          __syncwarp();
          if (pf_ok && t + 1 < gx_steps) {
            for (int r = lane; r < NB; r += 32) {
              const int b = row0 + ch * NB + r;
              if (b < B) {
                const float* src = gx_base + ((long long)(t + 1) * B + b) * ldgx + g * h;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(h * 4)) : "memory");
              }
            }
          }
        }
      }
    }
  } else {
    // ================================ cell update (warps 0..15) ================================
    WS_REG_COMPUTE();
    const WsRole ro = ws_role(warp, lane, h, nsub, NB);
    if (ro.warp_on) {
      const bool on = ro.on;
      const int jc = on ? ro.j : 0;                            // lanes without a unit shadow unit 0 (stores predicated)
      const int crow0 = row0 + ro.col0;                        // global batch row of this warp's first column in chain 0
      const int bvalid = B - crow0;                            // CTA-relative columns below this are real rows
      const int rbase = bvalid > 0 ? crow0 : B - 1;            // rows beyond B shadow the last valid row
      const int cmax = bvalid > 0 ? bvalid - 1 : 0;
      const uint32_t tl = tmem_base + ((uint32_t)(ro.q * 32) << 16) + (uint32_t)ro.col0;
      unsigned char* const Hhi = Hbase + (ro.j >> 3) * lboH + (ro.j & 7) * 2 + ro.col0 * 16;
      const int planeH = slabs * lboH;
      // running bases of the current step (element (first column of chain 0, gate g, unit jc)); + constant every step
      const float* gxp[4];
      float* gtp[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        gxp[g] = gx_base + (long long)rbase * ldgx + g * h + jc;
        gtp[g] = gates_base + (long long)rbase * H4 + g * h + jc;
      }
      float* csp = cs_base + (long long)(B + rbase) * ldcs + jc;
      float* cdp = csd_base ? csd_base + (long long)(B + rbase) * ldcs + jc : nullptr;
      float* hsp = hs_base + (long long)(B + rbase) * ldhs + jc;
      int ldx = ldgx;                                          // column pitch / step stride of the gx source: switches to the
      unsigned gx_step = (unsigned)B * (unsigned)ldgx;         // constant bias row (pitch 0) after gx_steps steps (decoder)
      const unsigned gt_step = (unsigned)B * (unsigned)H4;
      const unsigned cs_step = (unsigned)B * (unsigned)ldcs, hs_step = (unsigned)B * (unsigned)ldhs;
      float cst[NCH][4 * NSG];
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < 4 * NSG; ++i) cst[a][i] = 0.0f;
      // every warp owns NSG groups of 4 columns per chain: the group count is a template parameter, so the double-buffered
      // G_x registers are indexed statically (a run-time parity put the buffers in local memory)

      float gx[2][4][4];
      for (int t = 0; t < T; ++t) {
        if (t == gx_steps) {                                   // warp-uniform, once
#pragma unroll
          for (int g = 0; g < 4; ++g) gxp[g] = bias_rest + g * h + jc;
          ldx = 0;
          gx_step = 0u;
        }
        // G_x groups are double-buffered in processing order (group k of a step lives in gx[k & 1]) and always loaded one
        // group ahead -- across the chain switch and across the step boundary: since the gate GEMM is hidden behind the
        // other chain's cell update, nothing else would cover the load latency of a step's first group
        auto load_gx = [&](int ch, int sg, unsigned step_off, float (&dst)[4][4]) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const unsigned o = (unsigned)(min(ch * NB + sg * 4 + cc, cmax) * ldx) + step_off;
#pragma unroll
            for (int g = 0; g < 4; ++g) dst[g][cc] = __ldg(gxp[g] + o);
          }
        };
        if (t == 0 || t == gx_steps) load_gx(0, 0, 0u, gx[0]);  // (the source switched: reload)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          WS_STAMP(warp, t, ch * 3);
          ws_wait(smem_u32(&bar_done[ch]), (uint32_t)(t & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          WS_STAMP(warp, t, ch * 3 + 1);
          if ((NCH * NSG) % 2 == 1 && ch == 0 && t > 0 && t != gx_steps) {
            // an odd number of column groups per step (one half-width chain): the prefetch of this step's first group went to
            // the other buffer; the parities stay compile-time constants and the registers are moved instead
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) gx[0][g][cc] = gx[1][g][cc];
          }
#pragma unroll
          for (int sg = 0; sg < NSG; ++sg) {
            {
              constexpr int nsg = NSG;
              const int cur = (ch * NSG + sg) & 1;             // parity of this group in processing order
              float acc[4][4];
              if (ch == 0 && sg == 0) WS_STAMP(warp, t, 6);
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld4(tl + (uint32_t)((ch * 4 + g) * NB + sg * 4), acc[g]);
              // the next group's G_x (same chain, or the first group of the other chain) travels while this one is computed
              if (sg + 1 < nsg) load_gx(ch, sg + 1, 0u, gx[cur ^ 1]);
              else if (ch + 1 < NCH) load_gx(ch + 1, 0, 0u, gx[cur ^ 1]);
              else if (t + 1 < T && t + 1 != gx_steps) load_gx(0, 0, gx_step, gx[cur ^ 1]);     // next step's first group
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (ch == 0 && sg == 0) WS_STAMP(warp, t, 7);
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                const int col = sg * 4 + cc, gcol = ch * NB + col;      // column inside the chain / inside the CTA
                const int ci = min(gcol, cmax);
                const int ok = (on && gcol < bvalid) ? 1 : 0;
                constexpr float kS = -1.4426950408889634f, kT = -2.8853900817779268f;
                const float di = one_plus_ex2((acc[0][cc] + gx[cur][0][cc]) * kS), df = one_plus_ex2((acc[1][cc] + gx[cur][1][cc]) * kS);
                const float dg = one_plus_ex2((acc[2][cc] + gx[cur][2][cc]) * kT), dq = one_plus_ex2((acc[3][cc] + gx[cur][3][cc]) * kS);
                const float rif = rcp_fast(di * df), rgo = rcp_fast(dg * dq);
                const float ig = rif * df, fg = rif * di, og = rgo * dg;
                const float gg = fmaf(2.0f, rgo * dq, -1.0f);  // tanh(x) = 2 * logistic(2x) - 1
                const float cn = fmaf(fg, cst[ch][col], ig * gg);
                float hn = og * fmaf(2.0f, rcp_fast(one_plus_ex2(cn * kT)), -1.0f);
                hn = (gcol < bvalid) ? hn : 0.0f;              // rows beyond B feed zeros to the next gate GEMM
                cst[ch][col] = cn;
                const unsigned o4 = (unsigned)(ci * H4), oc = (unsigned)(ci * ldcs), oh = (unsigned)(ci * ldhs);
                st_if(gtp[0] + o4, ig, ok); st_if(gtp[1] + o4, fg, ok); st_if(gtp[2] + o4, gg, ok); st_if(gtp[3] + o4, og, ok);
                st_if(csp + oc, cn, ok);
                st_if(hsp + oh, hn, ok);
                if (cdp) st_if(cdp + oc, cn, ok);
                if (on) {                                      // h_t as the next step's B operand
                  const unsigned short hb = bf16_bits(hn);
                  const float hr = hn - __uint_as_float((uint32_t)hb << 16);
                  unsigned char* hp = Hhi + ch * chainH + col * 16;
                  *reinterpret_cast<unsigned short*>(hp) = hb;
                  *reinterpret_cast<unsigned short*>(hp + planeH) = bf16_bits(hr);
                }
              }
            }
          }
          WS_STAMP(warp, t, ch * 3 + 2);
          if (t + 1 < T) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ws_arrive(smem_u32(&bar_ready[ch]));
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) { gxp[g] += gx_step; gtp[g] += gt_step; }
        csp += cs_step;
        if (cdp) cdp += cs_step;
        hsp += hs_step;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(WS_THREADS, 1) lstm_ws_bwd_kernel(const __grid_constant__ WsBatch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_done[WS_MAXCHAIN], bar_ready[WS_MAXCHAIN];
  __shared__ uint32_t tmem_holder;
  const WsCell& wc = ws_find(bt, (int)blockIdx.x);
  const int h = wc.c.h, B = wc.c.B, T = wc.c.T, H4 = 4 * h;
  const int NB = wc.nb, KP = wc.kp, nsub = wc.nsub;         // KP = 4 * hp8
  const int slabs = KP >> 3, hp8 = KP >> 2;
  const int lboA = wc.lboA, lboB = wc.lboB;
  const int ldcs = (int)wc.c.ld_cs, lddh = (int)wc.c.ld_dh_all, lddc = (int)wc.c.ld_dc_ext;
  const float* __restrict__ const Wg = wc.c.W;
  const float* __restrict__ const gates_base = wc.c.gates;
  const float* __restrict__ const cs_base = wc.c.cs;
  const float* __restrict__ const dha_base = wc.c.dh_all;
  const float* __restrict__ const dhl_base = wc.c.dh_last;
  const float* __restrict__ const dce_base = wc.c.dc_ext;
  const float* __restrict__ const dce2_base = wc.c.dc_ext2;
  float* __restrict__ const dG_base = wc.c.dG;
  const int lddhl = (int)wc.c.ld_dh_last;
  // time-split recurrence (mfm_lstm_cell): carried dc in; W^T dG_0 and dc_0 f_0 out after one more product
  const float* __restrict__ const dcl_base = wc.c.dc_last;
  float* __restrict__ const dho_base = wc.c.dh_out;
  float* __restrict__ const dco_base = wc.c.dc_out;
  const bool carry = dho_base != nullptr;
  const bool e2full = wc.c.dc_ext2_full != 0;
  unsigned char* Ahi = smem;                                // W^T: A[j][k'], k' = 4*unit + gate
  unsigned char* Alo = Ahi + slabs * lboA;
  unsigned char* Bbase = Alo + slabs * lboA;                // per chain: dG tile [hi | lo]
  const int chainB = 2 * slabs * lboB;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int row0 = ((int)blockIdx.x - wc.cta0) * (NCH * NB);
  const int ksteps = KP >> 4;
  const int nacc = ksteps < 4 ? ksteps : 4;                 // partial sums over K: independent accumulators
  int tmem_cols = 32;
  while (tmem_cols < NCH * 4 * NB) tmem_cols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < NCH; ++i) {
      mbar_init(smem_u32(&bar_done[i]), 4);                   // one commit per issuing warp
      mbar_init(smem_u32(&bar_ready[i]), wc.nact);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[r][k' = 4*jj + g] = W[g*h + jj][j(r)]; item = (row r, slab): 8 gathered values, row fastest (coalesced over j)
  {
    const int span = nsub <= 2 ? 32 * nsub : h;
    const int rows = nsub <= 2 ? 128 : h;
    for (int idx = tid; idx < rows * slabs; idx += WS_THREADS) {
      const int r = idx % rows, slab = idx / rows;
      const int j = r % span;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int jj = 2 * slab + (e >> 2), g = e & 3;
        v[e] = (jj < h && j < h) ? __ldg(Wg + (long long)(g * h + jj) * h + j) : 0.0f;
      }
      split_store(v, Ahi + slab * lboA + r * 16, Alo + slab * lboA + r * 16, true);
    }
    for (int idx = tid * 16; idx < NCH * chainB; idx += WS_THREADS * 16)
      *reinterpret_cast<uint4*>(Bbase + idx) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp >= WS_CW) {
    WS_REG_ISSUER();
    {
      // the four warps of this warpgroup issue one partial sum over K each (k-steps a, a+4, a+8, ...), concurrently, and
      // between GEMMs pull the rows the cell update of the chain's NEXT step (t-1) reads into L2: gates (warp a = gate a),
      // cell history, external gradients
      const int a = warp - WS_CW;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t dAh0 = make_smem_desc(smem_u32(Ahi), lboA, 128), dAl0 = make_smem_desc(smem_u32(Alo), lboA, 128);
      const uint64_t astep = (uint64_t)((2 * lboA) >> 4), bstep = (uint64_t)((2 * lboB) >> 4);
      const bool pf_g = ((reinterpret_cast<uintptr_t>(gates_base) & 15) == 0) && ((h & 3) == 0);
      auto pf_row = [&](const float* base, int ld, long long row) {
        if (base && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0) && ((h & 3) == 0))
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + row * ld), "r"((uint32_t)(h * 4)) : "memory");
      };
      const int nprod = T + (carry ? 1 : 0);               // (the carried-out dh is one more product)
      for (int n = 0; n < nprod; ++n) {                     // n-th product: dG of step T-n (zero for n = 0) -> dh of step T-1-n
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          if (lane == 0) {
            if (n > 0) ws_wait(smem_u32(&bar_ready[ch]), (uint32_t)((n - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bb = smem_u32(Bbase + ch * chainB);
            const uint64_t dBh0 = make_smem_desc(bb, lboB, 128), dBl0 = make_smem_desc(bb + slabs * lboB, lboB, 128);
            const uint32_t dcol = tmem_base + (uint32_t)((ch * 4 + a) * NB);
#pragma unroll 1
            for (int kk = a; kk < ksteps; kk += 4) {
              umma_bf16(dcol, dAh0 + kk * astep, dBh0 + kk * bstep, idesc, kk > a ? 1u : 0u);
              umma_bf16(dcol, dAl0 + kk * astep, dBh0 + kk * bstep, idesc, 1u);
              umma_bf16(dcol, dAh0 + kk * astep, dBl0 + kk * bstep, idesc, 1u);
            }
            umma_commit(smem_u32(&bar_done[ch]));           // (a warp with no k-step still arrives)
          }
          __syncwarp();
          const int tn = T - 2 - n;                         // the step whose inputs to fetch: the one after the step being multiplied for
          if (tn >= 0) {
            for (int r = lane; r < NB; r += 32) {
              const int b = row0 + ch * NB + r;
              if (b < B) {
                const long long tr = (long long)tn * B + b;
                if (pf_g) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gates_base + tr * H4 + a * h), "r"((uint32_t)(h * 4)) : "memory");
                if (a == 0) pf_row(cs_base, ldcs, tr);
                if (a == 1) pf_row(dha_base, lddh, tr);
                if (a == 2) pf_row(dce_base, lddc, tr);
                if (a == 3) pf_row(dce2_base, lddc, tr);
              }
            }
          }
        }
      }
    }
  } else {
    WS_REG_COMPUTE();
    const WsRole ro = ws_role(warp, lane, h, nsub, NB);
    if (ro.warp_on) {
      const bool on = ro.on;
      const int jc = on ? ro.j : 0;
      const int crow0 = row0 + ro.col0;
      const int bvalid = B - crow0;
      const int rbase = bvalid > 0 ? crow0 : B - 1;
      const int cmax = bvalid > 0 ? bvalid - 1 : 0;
      const uint32_t tl = tmem_base + ((uint32_t)(ro.q * 32) << 16) + (uint32_t)ro.col0;
      const bool jb = ro.j < hp8;                             // lanes of the K padding keep writing zeros
      unsigned char* const Bhi = Bbase + (ro.j >> 1) * lboB + (ro.j & 1) * 8 + ro.col0 * 16;
      const int planeB = slabs * lboB;
      // running bases of the current step (t = T-1 first), moved back by a constant every step
      const long long tb0 = (long long)(T - 1) * B + rbase;
      const float* gtp[4];
      float* dgp[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        gtp[g] = gates_base + tb0 * H4 + g * h + jc;
        dgp[g] = dG_base + tb0 * H4 + g * h + jc;
      }
      const float* cpp = cs_base + tb0 * ldcs + jc;           // c_{t-1}; c_t is one block further
      const float* dhp = dha_base ? dha_base + tb0 * lddh + jc : nullptr;
      const float* dcp = dce_base ? dce_base + tb0 * lddc + jc : nullptr;
      const float* dc2p = dce2_base ? dce2_base + tb0 * lddc + jc : nullptr;
      const float* dhlp = dhl_base ? dhl_base + (long long)rbase * lddhl + jc : nullptr;
      const unsigned gt_step = (unsigned)B * (unsigned)H4, cs_step = (unsigned)B * (unsigned)ldcs;
      const unsigned dh_step = (unsigned)B * (unsigned)lddh, dc_step = (unsigned)B * (unsigned)lddc;
      float dc[NCH][8];
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dc[a][i] = (dcl_base && i < ro.ncol) ? __ldg(dcl_base + (long long)(rbase + min(a * NB + i, cmax)) * h + jc) : 0.0f;
      const int nsg = ro.ncol >> 1;                           // groups of TWO columns per chain: 2 or 4 (even: static parity)

      for (int t = T - 1; t >= 0; --t) {
        struct In { float ig, fg, gg, og, cp, cn, dhx, dcx; };
        In buf[2];                                            // (single-buffered: the issuer warps keep the next step's rows in L2)
        const bool last = t == T - 1;
        auto load_in = [&](int ch, int sg, In (&d)[2]) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int ci = min(ch * NB + sg * 2 + cc, cmax);
            const unsigned o4 = (unsigned)(ci * H4), oc = (unsigned)(ci * ldcs);
            In v;
            v.ig = gtp[0][o4]; v.fg = gtp[1][o4]; v.gg = gtp[2][o4]; v.og = gtp[3][o4];
            v.cp = cpp[oc];
            v.cn = cpp[oc + cs_step];
            v.dhx = dhp ? __ldg(dhp + (unsigned)(ci * lddh)) : 0.0f;
            if (last && dhlp) v.dhx += __ldg(dhlp + (unsigned)(ci * lddhl));
            v.dcx = dcp ? __ldg(dcp + (unsigned)(ci * lddc)) : 0.0f;
            if ((!last || e2full) && dc2p) v.dcx += __ldg(dc2p + (unsigned)(ci * lddc));
            d[cc] = v;
          }
        };
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          load_in(ch, 0, buf);                                // first group: in flight across the MMA wait
          ws_wait(smem_u32(&bar_done[ch]), (uint32_t)((T - 1 - t) & 1));     // dh_rec of this step = product T-1-t
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {
            if (sg < nsg) {
              float dh[4][2];
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                if (a < nacc) tmem_ld2(tl + (uint32_t)((ch * 4 + a) * NB + sg * 2), dh[a]);
                else dh[a][0] = dh[a][1] = 0.0f;
              }
              if (sg > 0) load_in(ch, sg, buf);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const int col = sg * 2 + cc, gcol = ch * NB + col;
                const int ci = min(gcol, cmax);
                const int ok = (on && gcol < bvalid) ? 1 : 0;
                const In& v = buf[cc];
                const float dht = (dh[0][cc] + dh[1][cc]) + (dh[2][cc] + dh[3][cc]) + v.dhx;
                const float tc = fmaf(2.0f, rcp_fast(one_plus_ex2(v.cn * -2.8853900817779268f)), -1.0f);
                const float dci = dc[ch][col] + dht * v.og * (1.0f - tc * tc) + v.dcx;
                float d_i = dci * v.gg * v.ig * (1.0f - v.ig);
                float d_f = dci * v.cp * v.fg * (1.0f - v.fg);
                float d_g = dci * v.ig * (1.0f - v.gg * v.gg);
                float d_o = dht * tc * v.og * (1.0f - v.og);
                dc[ch][col] = dci * v.fg;
                const unsigned o4 = (unsigned)(ci * H4);
                st_if(dgp[0] + o4, d_i, ok); st_if(dgp[1] + o4, d_f, ok); st_if(dgp[2] + o4, d_g, ok); st_if(dgp[3] + o4, d_o, ok);
                if (!ok) d_i = d_f = d_g = d_o = 0.0f;
                if (jb && (t > 0 || carry)) {                  // B operand: row = batch column, k' = 4j..4j+3 (8 contiguous bytes)
                  const float v4[4] = {d_i, d_f, d_g, d_o};
                  unsigned short hb[4], lb[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    hb[e] = bf16_bits(v4[e]);
                    lb[e] = bf16_bits(v4[e] - __uint_as_float((uint32_t)hb[e] << 16));
                  }
                  unsigned char* bp = Bhi + ch * chainB + col * 16;
                  *reinterpret_cast<uint2*>(bp) =
                      make_uint2((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), (uint32_t)hb[2] | ((uint32_t)hb[3] << 16));
                  *reinterpret_cast<uint2*>(bp + planeB) =
                      make_uint2((uint32_t)lb[0] | ((uint32_t)lb[1] << 16), (uint32_t)lb[2] | ((uint32_t)lb[3] << 16));
                }
              }
            }
          }
          if (t > 0 || carry) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ws_arrive(smem_u32(&bar_ready[ch]));
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) { gtp[g] -= gt_step; dgp[g] -= gt_step; }
        cpp -= cs_step;
        if (dhp) dhp -= dh_step;
        if (dcp) dcp -= dc_step;
        if (dc2p) dc2p -= dc_step;
      }
      if (carry) {                                            // product T: W^T dG_0, and the carried dc, go out
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          ws_wait(smem_u32(&bar_done[ch]), (uint32_t)(T & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {
            if (sg < nsg) {
              float dh[4][2];
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                if (a < nacc) tmem_ld2(tl + (uint32_t)((ch * 4 + a) * NB + sg * 2), dh[a]);
                else dh[a][0] = dh[a][1] = 0.0f;
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const int col = sg * 2 + cc, gcol = ch * NB + col;
                const int ci = min(gcol, cmax);
                const int ok = (on && gcol < bvalid) ? 1 : 0;
                const long long o = (long long)(rbase + ci) * h + jc;
                st_if(dho_base + o, (dh[0][cc] + dh[1][cc]) + (dh[2][cc] + dh[3][cc]), ok);
                st_if(dco_base + o, dc[ch][col], ok);
              }
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
static inline int ru(int x, int m) { return (x + m - 1) / m * m; }
static int ws_smem_limit() { return mfm_dev_info().smem_optin - 1024; }     // the opt-in limit covers static + dynamic

static int g_ws_force_nb = 0;      // tests: 16 forces the narrow chains where they are legal (nsub >= 3)
static int g_ws_force_chains = 0;  // tests: 1 or 2 chains per CTA (0 = by occupancy)
extern "C" int mfm_debug_lstm_force_nb(int nb) {
  if (nb != 0 && nb != 16 && nb != 32) return MFM_ERR_ARG;
  g_ws_force_nb = nb;
  return MFM_OK;
}
extern "C" int mfm_debug_lstm_force_chains(int n) {
  if (n < 0 || n > WS_MAXCHAIN) return MFM_ERR_ARG;
  g_ws_force_chains = n;
  return MFM_OK;
}
// copies the clock-stamp trace of the last forward launch to `host` (17*32*8 int64); only in -DWS_DEBUG=1 builds
extern "C" int mfm_debug_set_lstm_trace(void* host) {
#if WS_DEBUG
  if (!host) return MFM_ERR_ARG;
  cudaError_t e = cudaMemcpyFromSymbol(host, g_ws_trace_buf, sizeof(long long) * 17 * 32 * 8);
  return e == cudaSuccess ? MFM_OK : (int)e;
#else
  (void)host;
  return MFM_ERR_UNSUPPORTED;
#endif
}
// variant ids: 0 fwd wide chains (32 / 64 rows), 1 fwd half-width chains, 2 bwd wide, 3 bwd half-width, 4 fwd CUDA-core fallback,
// 5 bwd CUDA-core fallback, 6 launches with one chain per CTA, 7 launches with two
extern "C" unsigned long long mfm_debug_lstm_variant_count(int variant) {
  return (variant >= 0 && variant < 8) ? g_ws_counts[variant] : 0ull;
}
void ws_count_fallback(bool bwd, int ncells) { g_ws_counts[bwd ? 5 : 4] += (unsigned long long)ncells; }

// shared-memory plan of one cell; returns bytes, 0 if it does not fit
static size_t ws_plan(bool bwd, int h, int nb, int nchain, int limit, WsCell& lc) {
  const int nsub0 = (h + 31) / 32;
  if (nsub0 > 4) return 0;
  const int nsub = nsub0 < 2 ? 2 : nsub0;                   // layout: h <= 32 uses the two-copy layout with an empty sub-block
  const int R = nsub == 2 ? 2 : 1;
  const int ncol = nb / (4 * R);                            // columns per warp and chain: 8 or 4
  if (ncol * 4 * R != nb || !(ncol == 8 || ncol == 4)) return 0;
  lc.nb = nb;
  lc.nsub = nsub;
  lc.nact = nsub0 == 1 ? 8 : nsub0 == 3 ? 12 : 16;
  for (int pad = 32; pad >= 0; pad -= 32) {
    int rowsA, slabs;
    if (!bwd) {
      lc.gs = nsub <= 2 ? 128 : ru(h, 8);
      lc.kp = ru(h, 16);
      rowsA = 4 * lc.gs;
    } else {
      lc.gs = 0;
      lc.kp = 4 * ru(h, 8);
      rowsA = nsub <= 2 ? 128 : h;
    }
    slabs = lc.kp / 8;
    lc.lboA = rowsA * 16 + pad;
    lc.lboB = nb * 16 + pad;
    const size_t a = (size_t)2 * slabs * lc.lboA, b = (size_t)nchain * 2 * slabs * lc.lboB;
    // an MMA tile reads 128 rows from its first one: the last tile of a plane may run past the plane into the next region
    const int over = !bwd ? (3 * lc.gs + 128 - rowsA) * 16 : (128 - rowsA) * 16;
    const size_t need = a + (b > (size_t)(over > 0 ? over : 0) ? b : (size_t)over) + 128;
    if (need <= (size_t)limit) return need;
  }
  return 0;
}

// plans every cell for `nchain` chains per CTA, wide (8 columns per warp and chain) or narrow (4); returns the CTA count,
// -1 when a cell fits on chip but not in this shape (cells that fit in no shape go to `rest`)
static int ws_plan_all(bool bwd, const mfm_lstm_cell* cells, int ncells, int nchain, bool narrow, int lim, WsBatch& bt,
                       size_t& smem, mfm_lstm_cell* rest, int* nrest) {
  bt.n = 0;
  *nrest = 0;
  smem = 0;
  for (int i = 0; i < ncells; ++i) {
    const mfm_lstm_cell& c = cells[i];
    WsCell lc;
    size_t s = 0;
    bool any = false;
    if (c.h >= 1 && c.h <= 128) {
      const int nb_wide = c.h <= 64 ? 64 : 32;              // two-copy layouts take chains twice as wide
      const int nb = narrow ? nb_wide / 2 : nb_wide;
      s = ws_plan(bwd, c.h, nb, nchain, lim, lc);
      if (!s) {                                             // would any other shape hold it?  then the caller tries that shape
        WsCell tmp;
        for (int nc = 1; nc <= 2 && !any; ++nc)
          for (int w = 0; w < 2 && !any; ++w) any = ws_plan(bwd, c.h, w ? nb_wide / 2 : nb_wide, nc, lim, tmp) != 0;
      }
    }
    if (!s) {
      if (any) return -1;
      rest[(*nrest)++] = c;
      continue;
    }
    lc.c = c;
    bt.c[bt.n++] = lc;
    if (s > smem) smem = s;
  }
  // CTAs are dispatched in blockIdx order: the long-running (wide) cells first, so the short ones fill the tail
  for (int i = 1; i < bt.n; ++i) {
    WsCell key = bt.c[i];
    int j = i - 1;
    while (j >= 0 && bt.c[j].c.h < key.c.h) { bt.c[j + 1] = bt.c[j]; --j; }
    bt.c[j + 1] = key;
  }
  int total = 0;
  for (int i = 0; i < bt.n; ++i) {
    bt.c[i].cta0 = total;
    total += (bt.c[i].c.B + nchain * bt.c[i].nb - 1) / (nchain * bt.c[i].nb);
  }
  return total;
}

static int ws_launch(bool bwd, const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  const int lim = ws_smem_limit();
  WsBatch bt;
  size_t smem = 0;
  if (const char* e = getenv("MFM_WS_CHAINS")) g_ws_force_chains = atoi(e) == 1 ? 1 : atoi(e) == 2 ? 2 : 0;
  // Shapes, in order of preference: two wide chains per CTA (each hides the other's gate GEMM); when that leaves most
  // SMs without a CTA (a single decoder cell, small batches), two chains of half the width -- twice the CTAs, the GEMMs
  // still hidden; one wide chain per CTA when shared memory holds nothing else.
  // When even that leaves two thirds of the SMs without a CTA (a decoder cell at batch <= 1024), ONE half-width chain per
  // CTA: the step is bound by the MMA issue rate (~46 cycles per instruction; two chains = twice the instructions per step and
  // CTA), so a lone decoder cell h = 104 at batch 2048 runs 149 -> 111 us backward, 111 -> 100 us forward in this shape -- but
  // the step launches its three decoder cells side by side, 3 x 128 CTAs then queue for SMs and the step got 3 % SLOWER, so at
  // that size the two-chain shape stays.
  struct Shape { int nchain; bool narrow; };
  Shape order[4] = {{2, false}, {2, true}, {1, true}, {1, false}};
  if (g_ws_force_nb == 16) { order[0] = {2, true}; order[1] = {2, false}; }
  if (g_ws_force_chains == 1) { order[0] = {1, false}; order[1] = {2, false}; order[2] = {2, true}; order[3] = {1, true}; }
  const int sms = mfm_dev_info().sms;
  int pick = -1;
  for (int k = 0; k < 4; ++k) {
    const int t = ws_plan_all(bwd, cells, ncells, order[k].nchain, order[k].narrow, lim, bt, smem, rest, nrest);
    if (t < 0) continue;                                    // a cell does not take this shape
    if (pick >= 0 && order[k].nchain == 1 && !order[k].narrow) break;   // (one wide chain: only when nothing else fits)
    pick = k;
    // keep looking for a shape with more CTAs only when this one under-fills the GPU and nothing forces the choice
    const int want = order[k].nchain == 2 && order[k].narrow ? sms / 3 : sms * 2 / 3;
    if (t >= want || g_ws_force_nb == 16 || g_ws_force_chains) break;
  }
  if (pick < 0) {                                           // (cannot happen: every on-chip cell takes the last shape or none)
    *nrest = 0;
    for (int i = 0; i < ncells; ++i) rest[(*nrest)++] = cells[i];
    return MFM_OK;
  }
  const int nchain = order[pick].nchain;
  const bool narrow = order[pick].narrow;
  const int total = ws_plan_all(bwd, cells, ncells, nchain, narrow, lim, bt, smem, rest, nrest);
  if (!bt.n) return MFM_OK;                                 // nothing fits on chip: all cells are in `rest`
  for (int i = 0; i < bt.n; ++i) g_ws_counts[(bwd ? 2 : 0) + (narrow ? 1 : 0)] += 1;
  g_ws_counts[nchain == 1 ? 6 : 7] += 1;
  int e = 0;
  if (bwd) {
    if (nchain == 1) { if (!(e = mfm_func_smem_t(lstm_ws_bwd_kernel<1>, lim))) lstm_ws_bwd_kernel<1><<<total, WS_THREADS, smem, st>>>(bt); }
    else             { if (!(e = mfm_func_smem_t(lstm_ws_bwd_kernel<2>, lim))) lstm_ws_bwd_kernel<2><<<total, WS_THREADS, smem, st>>>(bt); }
  } else if (nchain == 1 && narrow) {
    if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<1, 1>, lim))) lstm_ws_fwd_kernel<1, 1><<<total, WS_THREADS, smem, st>>>(bt);
  } else if (nchain == 1) {
    if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<1, 2>, lim))) lstm_ws_fwd_kernel<1, 2><<<total, WS_THREADS, smem, st>>>(bt);
  } else if (narrow) {
    if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<2, 1>, lim))) lstm_ws_fwd_kernel<2, 1><<<total, WS_THREADS, smem, st>>>(bt);
  } else {
    if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<2, 2>, lim))) lstm_ws_fwd_kernel<2, 2><<<total, WS_THREADS, smem, st>>>(bt);
  }
  if (e) return e;
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

int lstm_tc_fwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  return ws_launch(false, cells, ncells, rest, nrest, st);
}
int lstm_tc_bwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  return ws_launch(true, cells, ncells, rest, nrest, st);
}
