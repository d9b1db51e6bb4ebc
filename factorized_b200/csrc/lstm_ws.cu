// LSTM recurrences on the tensor cores, warp-specialised (mfm_model.py:56,83,85,167-169 and their adjoint).
//
// A recurrence step is a strict chain: gate GEMM -> TMEM -> activations -> h_t -> next gate GEMM.  One chain cannot
// keep an SM busy (the first version of these kernels ran one chain per SM and spent 3/4 of every step waiting), so a CTA
// runs up to two independent chains -- batch sub-tiles of the same cell that share the resident weight image: while
// chain A's warps run their activations, chain B's gate GEMM executes, and the loads / SFU / stores of both overlap.
//
//   warps 8c .. 8c+7       chain c.   wait done[c]; tcgen05.ld; cell update; stash; h_t -> shared memory (split bf16,
//                          K-major B operand); fence.proxy.async; count in on arrive[c] (acq_rel).  The LAST warp to count in
//                          issues the chain's next gate GEMM (one thread): tcgen05.mma hi*hi + lo*hi + hi*lo over K,
//                          tcgen05.commit -> done[c].  (A 17th, MMA-only warp would put five warps on one scheduler
//                          and cap every thread at 96 registers.)
//
// The gate GEMM is issued TRANSPOSED, one MMA tile PER GATE:   D_g[unit j, batch b] = W_g[j, :] . h_{t-1}[b, :]
//   A = W_g (split-bf16, resident in shared memory, K-major = W's own layout), M = 128 rows = hidden units;
//   B = h_{t-1} of the chain's NB batch rows;  D_g = NB TMEM columns, lane = unit.
// A thread owns a TMEM lane, i.e. ONE hidden unit, and finds all four gates of that unit in its own lane (columns
// g*NB + b): the cell update needs no exchange between threads, and because lane = unit, the 32 lanes of a warp touch
// 32 consecutive floats of the row-major stashes (G_x, gates, c, h, dG): one 128 B line per access.
// Small cells would leave most of the 128 lanes idle, so the unit rows of the A tile are REPLICATED (h <= 32: 4 copies,
// h <= 64: 2): every lane quadrant then holds every unit and the quadrants split the batch columns instead.  Eight
// warps serve a chain: warp k reads quadrant k%4 and the column half k/4.
//
// Backward:  dh^T[unit j, batch b] = sum_k' W^T[j, k'] dG[b, k'],  k' = 4*unit + gate (unit-major, so the four gate
// gradients a thread produces are 8 contiguous bytes of the B operand); same roles, same replication; the carried dc
// stays in registers.  Cells that do not fit on chip (h > ~108 forward, h > 128 backward) run on lstm_seq.cu.
#include <cstdlib>
#include "tc_common.cuh"

// Timing experiments and the clock-stamp trace are compiled in only with -DWS_DEBUG=1 (scripts/lstm_trace.py): their
// lane-divergent branches inside the step loop cost the production kernel uniform registers.
#ifndef WS_DEBUG
#define WS_DEBUG 0
#endif

#define WS_CWARPS 8
#define WS_MAXCHAIN 2

struct WsCell {
  mfm_lstm_cell c;
  int nb;        // batch rows per chain = UMMA N (32 or 16)
  int nsub;      // ceil(h / 32): lane quadrants one copy of the units occupies
  int gs;        // forward: rows per gate block of the W image (128 when replicated, else h rounded up to 8)
  int kp;        // MMA K: forward h rounded up to 16; backward 4 * (h rounded up to 8)
  int lboA, lboB;
  int cta0;      // first blockIdx.x of this cell
};
struct WsBatch {
  WsCell c[MFM_MAX_CELLS];
  int n;
  long long* trace;   // debug (mfm_debug_set_lstm_trace): clock stamps of CTA 0, [warp 16][step 32][4]
  int dbg;       // timing experiments (env MFM_WS_DBG; results are WRONG when set): 1 no stash stores, 2 no G_x loads,
                 // 4 relaxed count-in, 8 no proxy fence, 16 no MMAs (commit only)
};

// wait on an mbarrier phase.  The suspend-time hint keeps the warp parked until the phase completes instead of
// re-polling (the first version re-polled ~200 times per step: a third of all issued instructions); bounded: a barrier
// that never completes traps instead of hanging the GPU
__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24); ++spin) {
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

static unsigned long long g_ws_counts[8];     // launches per variant, see mfm_debug_lstm_variant_count

// counts a warp in; true for the last of `n` (the counter only grows: arrivals of step s are [s*n, (s+1)*n))
__device__ __forceinline__ bool ws_count_in(unsigned int* cnt, unsigned int n) {
  unsigned int old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
  return (old + 1u) % n == 0u;
}
__device__ __forceinline__ bool ws_count_in_relaxed(unsigned int* cnt, unsigned int n) {
  unsigned int old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
  return (old + 1u) % n == 0u;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ unsigned short bf16_bits(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<unsigned short*>(&b);
}
// sigmoid and tanh through one path: tanh(x) = 2*sigmoid(2x) - 1  ->  a * rcp(1 + ex2(k * x)) + b
__device__ __forceinline__ float act_sig(float x) {
  x = fminf(fmaxf(x, -30.0f), 30.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  return rcp_fast(1.0f + e);
}
__device__ __forceinline__ float act_tanh(float x) {
  x = fminf(fmaxf(x, -15.0f), 15.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.8853900817779268f));
  return fmaf(2.0f, rcp_fast(1.0f + e), -1.0f);
}

// 1 + 2^a with the exponent clamped from above (2^28: products of two such terms stay finite); no lower clamp is needed:
// ex2 underflows to 0 and the logistic saturates correctly.  Two logistic values then share ONE reciprocal:
//   1/(1+ea) = (1+eb) * rcp((1+ea)(1+eb)),   1/(1+eb) = (1+ea) * rcp(...)         (MUFU is the scarce pipe: 8 lanes/clk/SM...)
__device__ __forceinline__ float one_plus_ex2(float a) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(a, 28.0f)));
  return 1.0f + e;
}
// predicated 4-byte global store without a branch (a divergent `if` around the stash stores cost a BSSY region per item)
__device__ __forceinline__ void st_if(float* p, float v, int ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(ok));
}
// which units and batch columns of its chain a compute warp owns
struct WsRole {
  int q;        // TMEM lane quadrant (= warp % 4)
  int j;        // hidden unit of this lane
  bool on;      // lane has a unit (warp-level activity is `warp_on`)
  bool warp_on;
  int col0;     // first batch column of this warp inside the chain's NB
  int nsg;      // groups of 4 columns: 1, 2 or 4
};
__device__ __forceinline__ WsRole ws_role(int k, int lane, int h, int nsub, int nb) {
  WsRole r;
  r.q = k & 3;
  const int half = k >> 2;
  if (nsub <= 2) {                         // replicated: quadrant = (copy, sub-block); copies and halves split the columns
    const int sub = r.q % nsub, rep = r.q / nsub, R = 4 / nsub;
    const int ncol = nb / (2 * R);
    r.j = sub * 32 + lane;
    r.col0 = (rep * 2 + half) * ncol;
    r.nsg = ncol >> 2;
    r.warp_on = true;
  } else {
    r.j = r.q * 32 + lane;
    r.col0 = half * (nb >> 1);
    r.nsg = nb >> 3;
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

// ----------------------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------------------
// Addressing: every per-element address is  (running 64-bit base of the step) + (32-bit column offset) * 4 -- one
// IMAD.WIDE per access; rows beyond B and lanes beyond h are CLAMPED onto valid elements and only their stores are
// predicated, so the loop body has no divergent branches (each one cost a BSSY region that re-materialised every
// descriptor; together with recomputed 64-bit indices the first version issued ~195 instructions per (unit, column)).
template <int NCHAIN>
__global__ void __launch_bounds__(NCHAIN * WS_CWARPS * 32, 1) lstm_ws_fwd_kernel(const __grid_constant__ WsBatch bt) {
  constexpr int NTHREADS = NCHAIN * WS_CWARPS * 32;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_done[WS_MAXCHAIN];
  __shared__ unsigned int arrive_cnt[WS_MAXCHAIN];
  __shared__ uint32_t tmem_holder;
  const WsCell& wc = ws_find(bt, (int)blockIdx.x);
  const int dbg = WS_DEBUG ? bt.dbg : 0;
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
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches become uniform branches
  const int row0 = ((int)blockIdx.x - wc.cta0) * (NCHAIN * NB);
  int tmem_cols = 32;
  while (tmem_cols < NCHAIN * 4 * NB) tmem_cols <<= 1;

  const unsigned int nact = (nsub == 3) ? 6u : 8u;           // warps with work per chain
  if (tid == 0) {
    for (int i = 0; i < NCHAIN; ++i) {
      mbar_init(smem_u32(&bar_done[i]), 1);
      arrive_cnt[i] = 0u;
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
    for (int idx = tid; idx < rows * slabs; idx += NTHREADS) {
      const int slab = idx % slabs, r = idx / slabs;
      const int g = r / GS, j = (r - g * GS) % span;
      float v[8];
      load8(Wg, h, j < h ? g * h + j : H4, H4, slab * 8, h, vecW, v);
      split_store(v, Whi + slab * lboA + r * 16, Wlo + slab * lboA + r * 16, true);
    }
    for (int idx = tid * 16; idx < NCHAIN * chainH; idx += NTHREADS * 16)
      *reinterpret_cast<uint4*>(Hbase + idx) = make_uint4(0, 0, 0, 0);     // h_{-1} = 0, K padding = 0
  }
  // block 0 of the histories is the zero initial state
  for (int idx = tid; idx < NCHAIN * NB * h; idx += NTHREADS) {
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

  // the chain's gate GEMM for the next step, issued by one thread (the last warp of the chain to count in)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto issue = [&](int ch) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t dAh0 = make_smem_desc(smem_u32(Whi), lboA, 128), dAl0 = make_smem_desc(smem_u32(Wlo), lboA, 128);
    const uint32_t hb = smem_u32(Hbase + ch * chainH);
    const uint64_t dBh0 = make_smem_desc(hb, lboH, 128), dBl0 = make_smem_desc(hb + slabs * lboH, lboH, 128);
    const int ksteps = KP >> 4;
    const uint64_t astep = (uint64_t)((2 * lboA) >> 4), bstep = (uint64_t)((2 * lboH) >> 4);
    // Consecutive MMAs into the SAME accumulator serialise at the full pipeline latency (~90 cycles each, measured), so the
    // four gates -- independent accumulators -- are interleaved: every pass of every k-step goes to all gates in turn.
    const uint32_t d0 = tmem_base + (uint32_t)(ch * 4 * NB);
    const uint64_t gstep = (uint64_t)((GS * 16) >> 4);
    uint64_t ao = 0, bo = 0;
    if (!(dbg & 16)) {
#pragma unroll 1
      for (int kk = 0; kk < ksteps; ++kk, ao += astep, bo += bstep) {
        const uint32_t acc = kk > 0 ? 1u : 0u;
#pragma unroll
        for (int g = 0; g < 4; ++g) umma_bf16(d0 + (uint32_t)(g * NB), dAh0 + ao + g * gstep, dBh0 + bo, idesc, acc);
#pragma unroll
        for (int g = 0; g < 4; ++g) umma_bf16(d0 + (uint32_t)(g * NB), dAl0 + ao + g * gstep, dBh0 + bo, idesc, 1u);
#pragma unroll
        for (int g = 0; g < 4; ++g) umma_bf16(d0 + (uint32_t)(g * NB), dAh0 + ao + g * gstep, dBl0 + bo, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar_done[ch]));
  };

  const int ch = warp / WS_CWARPS;
  const WsRole ro = ws_role(warp % WS_CWARPS, lane, h, nsub, NB);
  if (ro.warp_on) {
    const bool on = ro.on;
    const int jc = on ? ro.j : 0;                              // lanes without a unit shadow unit 0 (stores predicated)
    const int crow0 = row0 + ch * NB + ro.col0;                // global batch row of this warp's first column
    const int bvalid = B - crow0;                              // columns of this warp that are real rows (may be <= 0)
    const int rbase = bvalid > 0 ? crow0 : B - 1;              // rows beyond B shadow the last valid row
    const int cmax = bvalid > 0 ? bvalid - 1 : 0;
    const uint32_t tl = tmem_base + ((uint32_t)(ro.q * 32) << 16) + (uint32_t)(ch * 4 * NB + ro.col0);
    unsigned char* const Hhi = Hbase + ch * chainH + (ro.j >> 3) * lboH + (ro.j & 7) * 2 + ro.col0 * 16;
    unsigned char* const Hlo = Hhi + slabs * lboH;
    const uint32_t bar = smem_u32(&bar_done[ch]);
    long long* const tr = (WS_DEBUG && bt.trace && blockIdx.x == 0 && lane == 0) ? bt.trace : nullptr;
    // running bases of the current step (element (first column, gate g, unit jc)); advanced by a constant every step
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
    int ldx = ldgx;                                            // column pitch / step stride of the gx source: switch to the
    unsigned gx_step = (unsigned)B * (unsigned)ldgx;           // constant bias row (pitch 0) after gx_steps steps (decoder)
    const unsigned gt_step = (unsigned)B * (unsigned)H4;
    const unsigned cs_step = (unsigned)B * (unsigned)ldcs, hs_step = (unsigned)B * (unsigned)ldhs;
    float cst[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cst[i] = 0.0f;
    const int nsg = ro.nsg;
    // step 0 multiplies the zero initial state like every other step (D = 0 exactly): no special case in the loop
    if (warp % WS_CWARPS == 0 && lane == 0) issue(ch);

    for (int t = 0; t < T; ++t) {
      if (t == gx_steps) {                                     // warp-uniform, once
#pragma unroll
        for (int g = 0; g < 4; ++g) gxp[g] = bias_rest + g * h + jc;
        ldx = 0;
        gx_step = 0u;
      }
      float gx[2][4][4];
      auto load_gx = [&](int sg, float (&dst)[4][4]) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const unsigned o = (unsigned)(min(sg * 4 + cc, cmax) * ldx);
#pragma unroll
          for (int g = 0; g < 4; ++g) dst[g][cc] = (dbg & 2) ? 0.1f : __ldg(gxp[g] + o);
        }
      };
      load_gx(0, gx[0]);                                       // first column group: in flight across the MMA wait
      if (tr && t < 32) tr[(warp * 32 + t) * 4 + 0] = clock64();
      if (dbg & 32) mbar_wait(bar, (uint32_t)(t & 1)); else ws_wait(bar, (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tr && t < 32) tr[(warp * 32 + t) * 4 + 1] = clock64();
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        if (sg < nsg) {                                        // warp-uniform
          float acc[4][4];
          if (dbg & 64) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) acc[g][cc] = 0.0f;
          } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld4(tl + (uint32_t)(g * NB + sg * 4), acc[g]);
          }
          if (sg + 1 < nsg) load_gx(sg + 1, gx[(sg + 1) & 1]);  // next group's G_x travels while this one is computed
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int col = sg * 4 + cc;
            const int ci = min(col, cmax);
            const int ok = (on && col < bvalid && !(dbg & 1)) ? 1 : 0;
            constexpr float kS = -1.4426950408889634f, kT = -2.8853900817779268f;
            const float di = one_plus_ex2((acc[0][cc] + gx[sg & 1][0][cc]) * kS), df = one_plus_ex2((acc[1][cc] + gx[sg & 1][1][cc]) * kS);
            const float dg = one_plus_ex2((acc[2][cc] + gx[sg & 1][2][cc]) * kT), dq = one_plus_ex2((acc[3][cc] + gx[sg & 1][3][cc]) * kS);
            const float rif = rcp_fast(di * df), rgo = rcp_fast(dg * dq);
            const float ig = rif * df, fg = rif * di, og = rgo * dg;
            const float gg = fmaf(2.0f, rgo * dq, -1.0f);      // tanh(x) = 2 * logistic(2x) - 1
            const float cn = fmaf(fg, cst[col], ig * gg);
            float hn = og * fmaf(2.0f, rcp_fast(one_plus_ex2(cn * kT)), -1.0f);
            hn = (col < bvalid) ? hn : 0.0f;                   // rows beyond B feed zeros to the next gate GEMM
            cst[col] = cn;
            const unsigned o4 = (unsigned)(ci * H4), oc = (unsigned)(ci * ldcs), oh = (unsigned)(ci * ldhs);
            st_if(gtp[0] + o4, ig, ok); st_if(gtp[1] + o4, fg, ok); st_if(gtp[2] + o4, gg, ok); st_if(gtp[3] + o4, og, ok);
            st_if(csp + oc, cn, ok);
            st_if(hsp + oh, hn, ok);
            if (cdp) st_if(cdp + oc, cn, ok);
            if (on && !(dbg & 128)) {                          // h_t as the next step's B operand
              const unsigned short hb = bf16_bits(hn);
              const float hr = hn - __uint_as_float((uint32_t)hb << 16);
              *reinterpret_cast<unsigned short*>(Hhi + col * 16) = hb;
              *reinterpret_cast<unsigned short*>(Hlo + col * 16) = bf16_bits(hr);
            }
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) { gxp[g] += gx_step; gtp[g] += gt_step; }
      csp += cs_step;
      if (cdp) cdp += cs_step;
      hsp += hs_step;
      if (t + 1 < T) {
        if (!(dbg & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (tr && t < 32) tr[(warp * 32 + t) * 4 + 2] = clock64();
        if (lane == 0 && ((dbg & 4) ? ws_count_in_relaxed(&arrive_cnt[ch], nact) : ws_count_in(&arrive_cnt[ch], nact))) {
          issue(ch);
          if (tr && t < 32) tr[(warp * 32 + t) * 4 + 3] = clock64();
        }
        __syncwarp();
      }
    }
  } else if (T > 0 && (warp % WS_CWARPS) == 0 && lane == 0) {
    issue(ch);                                                 // (never: warp 0 of a chain always has work)
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
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float v[2]) {
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r[0]);
  v[1] = __uint_as_float(r[1]);
}

template <int NCHAIN>
__global__ void __launch_bounds__(NCHAIN * WS_CWARPS * 32, 1) lstm_ws_bwd_kernel(const __grid_constant__ WsBatch bt) {
  constexpr int NTHREADS = NCHAIN * WS_CWARPS * 32;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_done[WS_MAXCHAIN];
  __shared__ unsigned int arrive_cnt[WS_MAXCHAIN];
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
  unsigned char* Ahi = smem;                                // W^T: A[j][k'], k' = 4*unit + gate
  unsigned char* Alo = Ahi + slabs * lboA;
  unsigned char* Bbase = Alo + slabs * lboA;                // per chain: dG tile [hi | lo]
  const int chainB = 2 * slabs * lboB;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform: role branches become uniform branches
  const int row0 = ((int)blockIdx.x - wc.cta0) * (NCHAIN * NB);
  const int tmem_cols = NCHAIN * NB < 32 ? 32 : NCHAIN * NB;

  const unsigned int nact = (nsub == 3) ? 6u : 8u;
  if (tid == 0) {
    for (int i = 0; i < NCHAIN; ++i) {
      mbar_init(smem_u32(&bar_done[i]), 1);
      arrive_cnt[i] = 0u;
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
    for (int idx = tid; idx < rows * slabs; idx += NTHREADS) {
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
    for (int idx = tid * 16; idx < NCHAIN * chainB; idx += NTHREADS * 16)
      *reinterpret_cast<uint4*>(Bbase + idx) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto issue = [&](int ch) {                                // dG of the step just finished -> dh of the step before
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t dAh0 = make_smem_desc(smem_u32(Ahi), lboA, 128), dAl0 = make_smem_desc(smem_u32(Alo), lboA, 128);
    const uint32_t bb = smem_u32(Bbase + ch * chainB);
    const uint64_t dBh0 = make_smem_desc(bb, lboB, 128), dBl0 = make_smem_desc(bb + slabs * lboB, lboB, 128);
    const uint32_t dcol = tmem_base + (uint32_t)(ch * NB);
    const int ksteps = KP >> 4;
    const uint64_t astep = (uint64_t)((2 * lboA) >> 4), bstep = (uint64_t)((2 * lboB) >> 4);
    uint64_t ao = 0, bo = 0;
#pragma unroll 1
    for (int kk = 0; kk < ksteps; ++kk, ao += astep, bo += bstep) {
      umma_bf16(dcol, dAh0 + ao, dBh0 + bo, idesc, kk > 0 ? 1u : 0u);
      umma_bf16(dcol, dAl0 + ao, dBh0 + bo, idesc, 1u);
      umma_bf16(dcol, dAh0 + ao, dBl0 + bo, idesc, 1u);
    }
    umma_commit(smem_u32(&bar_done[ch]));
  };

  const int ch = warp / WS_CWARPS;
  const WsRole ro = ws_role(warp % WS_CWARPS, lane, h, nsub, NB);
  if (ro.warp_on) {
    const bool on = ro.on;
    const int jc = on ? ro.j : 0;
    const int crow0 = row0 + ch * NB + ro.col0;
    const int bvalid = B - crow0;
    const int rbase = bvalid > 0 ? crow0 : B - 1;
    const int cmax = bvalid > 0 ? bvalid - 1 : 0;
    const uint32_t tl = tmem_base + ((uint32_t)(ro.q * 32) << 16) + (uint32_t)(ch * NB + ro.col0);
    const bool jb = ro.j < hp8;                               // lanes of the K padding keep writing zeros
    unsigned char* const Bhi = Bbase + ch * chainB + (ro.j >> 1) * lboB + (ro.j & 1) * 8 + ro.col0 * 16;
    unsigned char* const Blo = Bhi + slabs * lboB;
    const uint32_t bar = smem_u32(&bar_done[ch]);
    // running bases of the current step (t = T-1 first), moved back by a constant every step
    const long long tb0 = (long long)(T - 1) * B + rbase;
    const float* gtp[4];
    float* dgp[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      gtp[g] = gates_base + tb0 * H4 + g * h + jc;
      dgp[g] = dG_base + tb0 * H4 + g * h + jc;
    }
    const float* cpp = cs_base + tb0 * ldcs + jc;             // c_{t-1}; c_t is one block further
    const float* dhp = dha_base ? dha_base + tb0 * lddh + jc : nullptr;
    const float* dcp = dce_base ? dce_base + tb0 * lddc + jc : nullptr;
    const float* dc2p = dce2_base ? dce2_base + tb0 * lddc + jc : nullptr;
    const float* dhlp = dhl_base ? dhl_base + (long long)rbase * lddhl + jc : nullptr;
    const unsigned gt_step = (unsigned)B * (unsigned)H4, cs_step = (unsigned)B * (unsigned)ldcs;
    const unsigned dh_step = (unsigned)B * (unsigned)lddh, dc_step = (unsigned)B * (unsigned)lddc;
    float dc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dc[i] = 0.0f;
    const int nsg = 2 * ro.nsg;                               // groups of TWO columns here (register budget)
    // the last step multiplies a zero dG like every other step (dh_rec = 0 exactly): no special case in the loop
    if (warp % WS_CWARPS == 0 && lane == 0) issue(ch);

    for (int t = T - 1; t >= 0; --t) {
      struct In { float ig, fg, gg, og, cp, cn, dhx, dcx; };
      In buf[2][2];
      const bool last = t == T - 1;
      auto load_in = [&](int sg, In (&d)[2]) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int ci = min(sg * 2 + cc, cmax);
          const unsigned o4 = (unsigned)(ci * H4), oc = (unsigned)(ci * ldcs);
          In v;
          v.ig = gtp[0][o4]; v.fg = gtp[1][o4]; v.gg = gtp[2][o4]; v.og = gtp[3][o4];
          v.cp = cpp[oc];
          v.cn = cpp[oc + cs_step];
          v.dhx = dhp ? __ldg(dhp + (unsigned)(ci * lddh)) : 0.0f;
          if (last && dhlp) v.dhx += __ldg(dhlp + (unsigned)(ci * lddhl));
          v.dcx = dcp ? __ldg(dcp + (unsigned)(ci * lddc)) : 0.0f;
          if (!last && dc2p) v.dcx += __ldg(dc2p + (unsigned)(ci * lddc));
          d[cc] = v;
        }
      };
      load_in(0, buf[0]);
      ws_wait(bar, (uint32_t)((T - 1 - t) & 1));              // dh_rec of this step = product T-1-t
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int sg = 0; sg < 8; ++sg) {
        if (sg < nsg) {
          float dh[2];
          tmem_ld2(tl + (uint32_t)(sg * 2), dh);
          if (sg + 1 < nsg) load_in(sg + 1, buf[(sg + 1) & 1]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int col = sg * 2 + cc;
            const int ci = min(col, cmax);
            const bool ok = on && col < bvalid;
            const In& v = buf[sg & 1][cc];
            const float dht = dh[cc] + v.dhx;
            const float tc = act_tanh(v.cn);
            const float dci = dc[col] + dht * v.og * (1.0f - tc * tc) + v.dcx;
            float d_i = dci * v.gg * v.ig * (1.0f - v.ig);
            float d_f = dci * v.cp * v.fg * (1.0f - v.fg);
            float d_g = dci * v.ig * (1.0f - v.gg * v.gg);
            float d_o = dht * tc * v.og * (1.0f - v.og);
            dc[col] = dci * v.fg;
            const unsigned o4 = (unsigned)(ci * H4);
            if (ok) {
              dgp[0][o4] = d_i; dgp[1][o4] = d_f; dgp[2][o4] = d_g; dgp[3][o4] = d_o;
            } else {
              d_i = d_f = d_g = d_o = 0.0f;
            }
            if (jb && t > 0) {                                 // B operand: row = batch column, k' = 4j..4j+3 (8 contiguous bytes)
              const float v4[4] = {d_i, d_f, d_g, d_o};
              unsigned short hb[4], lb[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hb[e] = bf16_bits(v4[e]);
                lb[e] = bf16_bits(v4[e] - __uint_as_float((uint32_t)hb[e] << 16));
              }
              *reinterpret_cast<uint2*>(Bhi + col * 16) =
                  make_uint2((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), (uint32_t)hb[2] | ((uint32_t)hb[3] << 16));
              *reinterpret_cast<uint2*>(Blo + col * 16) =
                  make_uint2((uint32_t)lb[0] | ((uint32_t)lb[1] << 16), (uint32_t)lb[2] | ((uint32_t)lb[3] << 16));
            }
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) { gtp[g] -= gt_step; dgp[g] -= gt_step; }
      cpp -= cs_step;
      if (dhp) dhp -= dh_step;
      if (dcp) dcp -= dc_step;
      if (dc2p) dc2p -= dc_step;
      if (t > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0 && ws_count_in(&arrive_cnt[ch], nact)) issue(ch);
        __syncwarp();
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

static long long* g_ws_trace = nullptr;
extern "C" int mfm_debug_set_lstm_trace(void* buf) {
  g_ws_trace = static_cast<long long*>(buf);
  return MFM_OK;
}
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
// variant ids: 0 fwd NB=32, 1 fwd NB=16, 2 bwd NB=32, 3 bwd NB=16, 4 fwd CUDA-core fallback, 5 bwd CUDA-core fallback,
// 6 launches with one chain per CTA, 7 launches with two
extern "C" unsigned long long mfm_debug_lstm_variant_count(int variant) {
  return (variant >= 0 && variant < 8) ? g_ws_counts[variant] : 0ull;
}
void ws_count_fallback(bool bwd, int ncells) { g_ws_counts[bwd ? 5 : 4] += (unsigned long long)ncells; }

// shared-memory plan of one cell; returns bytes, 0 if it does not fit
static size_t ws_plan(bool bwd, int h, int nb, int nchain, int limit, WsCell& lc) {
  const int nsub = (h + 31) / 32;
  if (nsub > 4) return 0;
  if (nb == 16 && nsub <= 2) return 0;                      // the replicated layouts split 32 columns over 8 warps
  lc.nb = nb;
  lc.nsub = nsub;
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

// plans every cell for `nchain` chains per CTA; returns the CTA count (cells that do not fit are left out)
static int ws_plan_all(bool bwd, const mfm_lstm_cell* cells, int ncells, int nchain, int lim, WsBatch& bt, size_t& smem,
                       mfm_lstm_cell* rest, int* nrest) {
  bt.n = 0;
  *nrest = 0;
  smem = 0;
  for (int i = 0; i < ncells; ++i) {
    const mfm_lstm_cell& c = cells[i];
    WsCell lc;
    size_t s = 0;
    if (c.h >= 1 && c.h <= 128) {
      if (g_ws_force_nb != 16) s = ws_plan(bwd, c.h, 32, nchain, lim, lc);
      if (!s) s = ws_plan(bwd, c.h, 16, nchain, lim, lc);
      if (!s && g_ws_force_nb == 16) s = ws_plan(bwd, c.h, 32, nchain, lim, lc);
    }
    if (!s) { rest[(*nrest)++] = c; continue; }
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
  {
    const char* e = getenv("MFM_WS_DBG");
    bt.dbg = e ? atoi(e) : 0;
    bt.trace = g_ws_trace;
  }
  size_t smem = 0;
  // two chains per CTA overlap each other's latency; when that leaves most SMs without a CTA (a single decoder cell,
  // small batches), one chain per CTA spreads the chains over twice as many SMs instead
  int nchain = g_ws_force_chains ? g_ws_force_chains : 2;
  int total = ws_plan_all(bwd, cells, ncells, nchain, lim, bt, smem, rest, nrest);
  if (!g_ws_force_chains && bt.n && total * 3 < mfm_dev_info().sms * 2) {
    nchain = 1;
    total = ws_plan_all(bwd, cells, ncells, nchain, lim, bt, smem, rest, nrest);
  }
  if (!bt.n) return MFM_OK;
  for (int i = 0; i < bt.n; ++i) g_ws_counts[(bwd ? 2 : 0) + (bt.c[i].nb == 16 ? 1 : 0)] += 1;
  g_ws_counts[nchain == 1 ? 6 : 7] += 1;
  int e = 0;
  if (bwd) {
    if (nchain == 1) { if (!(e = mfm_func_smem_t(lstm_ws_bwd_kernel<1>, lim))) lstm_ws_bwd_kernel<1><<<total, WS_CWARPS * 32, smem, st>>>(bt); }
    else             { if (!(e = mfm_func_smem_t(lstm_ws_bwd_kernel<2>, lim))) lstm_ws_bwd_kernel<2><<<total, 2 * WS_CWARPS * 32, smem, st>>>(bt); }
  } else {
    if (nchain == 1) { if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<1>, lim))) lstm_ws_fwd_kernel<1><<<total, WS_CWARPS * 32, smem, st>>>(bt); }
    else             { if (!(e = mfm_func_smem_t(lstm_ws_fwd_kernel<2>, lim))) lstm_ws_fwd_kernel<2><<<total, 2 * WS_CWARPS * 32, smem, st>>>(bt); }
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
