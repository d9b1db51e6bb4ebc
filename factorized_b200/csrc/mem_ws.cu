// The MFN memory recurrence on the tensor cores (mfm_model.py:177-180 and its adjoint), in the transposed tcgen05 form of
// lstm_ws.cu.  mfn.cu holds the CUDA-core kernels this replaces where it applies (mem <= 64, g1, g2 <= 128).
//
// A step is a chain of TWO small GEMMs with an elementwise stage after each:
//   forward    u_k = dropout(relu(Gkpre[t] + mem W_km^T))   (K = mem)    k = 1, 2
//              gamma_k = sigmoid(u_k W_k2^T + b_k2)         (K = g_k)
//              mem' = gamma_1 mem + gamma_2 cHat[t]
//   backward   dp_k, dPc, dmem*gamma_1 from (dmem, mem_{t-1}, gamma_k, cHat)           (elementwise)
//              du_k = (dp_k W_k2) masked by u_k > 0         (K = mem)
//              dmem_{t-1} = dmem*gamma_1 + du_1 W_1m + du_2 W_2m   (K = g_1 + g_2)
// Both GEMMs are issued transposed, D[unit, batch] = W[unit, :] . x[batch, :]: the weights are the resident A operand
// (split bf16 hi/lo, K-major, built once in the prologue), the CTA's 16 batch rows are the N dimension, and a thread owns a
// TMEM lane = one unit, so every elementwise stage finds all it needs in its own lane and touches the row-major stashes with
// lane = consecutive floats.  Two layouts of the 8 compute warps:
//   "wide" stage  (units of gamma*_fc1, <= 128 per k): warps 0..3 serve k = 1, warps 4..7 serve k = 2; lane quadrant w % 4 holds
//                 units 32 (w % 4) .., all 16 batch columns
//   "mem" stage   (memory units, <= 64): the unit rows of the A tile are replicated (rows 64.. = rows 0..), lane quadrant
//                 q = w % 4 holds units 32 (q & 1) .., and the copies (q >> 1) and warp halves (w >> 2) split the 16 batch
//                 columns four ways; gamma_1 and gamma_2 (forward) are two accumulators in the SAME lane
// Four more warps issue the MMAs, one thread each: a single thread sustains one tcgen05.mma per ~100 cycles and MMAs into one
// accumulator serialise, so issuer q takes tile k = q & 1 and the K steps of parity q >> 1 into its OWN partial accumulator
// (4 x 16 TMEM columns per GEMM); the elementwise stage adds the two partials of its tile.  hi*hi + lo*hi + hi*lo as everywhere.
// The state operands (mem / u_k, dp_k / du_k) are written by the elementwise stages as split bf16 straight into the K-major
// B-operand layout; four mbarriers (ready / done per GEMM) carry the chain.
#include <cstdlib>
#include "tc_common.cuh"

// -DMW_DEBUG=1 compiles clock stamps of CTA 0 into the forward kernel (scripts/mem_trace.py)
#ifndef MW_DEBUG
#define MW_DEBUG 0
#endif
#if MW_DEBUG
__device__ long long g_mw_trace[32 * 16];                 // [step][slot]
#define MW_STAMP(t, k) do { if (blockIdx.x == 0 && (t) < 32) g_mw_trace[(t) * 16 + (k)] = clock64(); } while (0)
#else
#define MW_STAMP(t, k) do { } while (0)
#endif
#define MW_NB 16                      // batch rows per CTA = UMMA N
#define MW_CW 8                       // compute warps
#define MW_THREADS ((MW_CW + 4) * 32)
#define MW_LBO (128 * 16)             // A images: 128 rows per K slab
#define MW_LBOB (MW_NB * 16)          // B operands: 16 rows per K slab

// The chain is strictly serial and nothing else runs on the SM: the waiters poll (test_wait).  (Parking on the barrier -- the
// try_wait form of tc_common.cuh -- was measured equal in lstm_ws.cu.)  Bounded: a barrier that never completes traps.
__device__ __forceinline__ void mw_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int spin = 0; spin < (1 << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void mw_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mw_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mw_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ unsigned short mw_bf16(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<unsigned short*>(&b);
}
// element (batch column b, k) of a K-major B operand: hi plane at `op`, lo plane `plane` bytes further
__device__ __forceinline__ void mw_put(unsigned char* op, int plane, int b, int k, float v) {
  unsigned char* p = op + (k >> 3) * MW_LBOB + b * 16 + (k & 7) * 2;
  const unsigned short hb = mw_bf16(v);
  *reinterpret_cast<unsigned short*>(p) = hb;
  *reinterpret_cast<unsigned short*>(p + plane) = mw_bf16(v - __uint_as_float((uint32_t)hb << 16));
}
__device__ __forceinline__ void mw_st_if(float* p, float v, int ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(ok));
}

// A image of one tile: 128 rows x K (a multiple of 16), element (r, k) = src[(r % span) * sr + k * sk] inside [nr, nk), else 0.
// The lanes of a warp run along the CONTIGUOUS dimension of the source (sk == 1: K slabs, 32 B per lane as two float4 when
// aligned; sr == 1: rows) -- with the lanes across a strided dimension every load instruction touched 32 lines and the
// prologue took 30 us.  Four items per pass keep 32 loads of a thread in flight.
__device__ __forceinline__ void mw_build(unsigned char* hi, int K, const float* __restrict__ src, long long sr, long long sk, int nr,
                                         int nk, int span, int tid) {
  const int slabs = K >> 3;
  unsigned char* lo = hi + slabs * MW_LBO;
  const bool kmajor = sk == 1;
  const bool vec = kmajor && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((sr & 3) == 0);
  for (int idx0 = tid; idx0 < 128 * slabs; idx0 += 4 * MW_THREADS) {
    float v[4][8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int idx = idx0 + m * MW_THREADS;
      const int r = kmajor ? idx / slabs : idx & 127, slab = kmajor ? idx - r * slabs : idx >> 7;
      const int rs = r % span;
      const bool in = idx < 128 * slabs && rs < nr;
      if (kmajor) {
        load8(src, sr, in ? rs : nr, nr, slab * 8, nk, vec, v[m]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = slab * 8 + i;
          v[m][i] = (in && k < nk) ? __ldg(src + (long long)rs * sr + (long long)k * sk) : 0.0f;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int idx = idx0 + m * MW_THREADS;
      const int r = kmajor ? idx / slabs : idx & 127, slab = kmajor ? idx - r * slabs : idx >> 7;
      if (idx < 128 * slabs) split_store(v[m], hi + slab * MW_LBO + r * 16, lo + slab * MW_LBO + r * 16, true);
    }
  }
}

// one issuer's share of a GEMM: K steps kk0, kk0 + 2, ... of tile image `img` against operand `op` into its partial accumulator
__device__ __forceinline__ void mw_issue(uint32_t tD, uint32_t img, int K, uint32_t op, int kk0, uint32_t idesc) {
  const int slabs = K >> 3, ksteps = K >> 4;
  const uint64_t dAh = make_smem_desc(img, MW_LBO, 128), dAl = make_smem_desc(img + slabs * MW_LBO, MW_LBO, 128);
  const uint64_t dBh = make_smem_desc(op, MW_LBOB, 128), dBl = make_smem_desc(op + slabs * MW_LBOB, MW_LBOB, 128);
  for (int kk = kk0; kk < ksteps; kk += 2) {
    const uint64_t ao = (uint64_t)((kk * 2 * MW_LBO) >> 4), bo = (uint64_t)((kk * 2 * MW_LBOB) >> 4);
    umma_bf16(tD, dAh + ao, dBh + bo, idesc, kk > kk0 ? 1u : 0u);
    umma_bf16(tD, dAl + ao, dBh + bo, idesc, 1u);
    umma_bf16(tD, dAh + ao, dBl + bo, idesc, 1u);
  }
}

struct MwDims { int KM, K1, K2; int offW[4]; int offOp[3]; int total; };   // K of the mem / g1 / g2 contractions (multiples of 16)
static __host__ __device__ __forceinline__ MwDims mw_dims(int mem, int g1, int g2) {
  MwDims d;
  d.KM = (mem + 15) & ~15; d.K1 = (g1 + 15) & ~15; d.K2 = (g2 + 15) & ~15;
  int o = 0;
  // images: [0], [1] contract over mem (rows = gamma*_fc1 units of k = 1, 2); [2], [3] contract over g_k (rows = memory units)
  d.offW[0] = o; o += 2 * (d.KM >> 3) * MW_LBO;
  d.offW[1] = o; o += 2 * (d.KM >> 3) * MW_LBO;
  d.offW[2] = o; o += 2 * (d.K1 >> 3) * MW_LBO;
  d.offW[3] = o; o += 2 * (d.K2 >> 3) * MW_LBO;
  // operands: [0] K = mem (two of them in backward: dp_1, dp_2), [1] K = g1, [2] K = g2
  d.offOp[0] = o; o += 2 * 2 * (d.KM >> 3) * MW_LBOB;
  d.offOp[1] = o; o += 2 * (d.K1 >> 3) * MW_LBOB;
  d.offOp[2] = o; o += 2 * (d.K2 >> 3) * MW_LBOB;
  d.total = o;
  return d;
}

#define MW_PROLOGUE()                                                                                                         \
  extern __shared__ __align__(128) unsigned char smem[];                                                                      \
  __shared__ __align__(8) unsigned long long bars[4]; /* ready / done of the mem-K GEMM, ready / done of the g-K GEMM */      \
  __shared__ uint32_t tmem_holder;                                                                                            \
  const int T = a.T, B = a.B, mem = a.mem, g1 = a.g1, g2 = a.g2;                                                              \
  const MwDims d = mw_dims(mem, g1, g2);                                                                                      \
  const int tid = threadIdx.x, lane = tid & 31;                                                                               \
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                                                                     \
  const int row0 = (int)blockIdx.x * MW_NB;                                                                                   \
  if (tid == 0) {                                                                                                             \
    mbar_init(smem_u32(&bars[0]), MW_CW); mbar_init(smem_u32(&bars[1]), 4);                                                   \
    mbar_init(smem_u32(&bars[2]), MW_CW); mbar_init(smem_u32(&bars[3]), 4);                                                   \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                                        \
  }                                                                                                                           \
  if (warp == 0) {                                                                                                            \
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(128u) \
                 : "memory");                                                                                                 \
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");                                  \
  }                                                                                                                           \
  for (int idx = tid * 16; idx < d.total - d.offOp[0]; idx += MW_THREADS * 16)                                                \
    *reinterpret_cast<uint4*>(smem + d.offOp[0] + idx) = make_uint4(0, 0, 0, 0); /* zero state, zero K padding */

#define MW_START()                                                   \
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      \
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  \
  __syncthreads();                                                   \
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   \
  const uint32_t tmem_base = tmem_holder;                            \
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MW_NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

#define MW_END()                                                                                                       \
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                                     \
  __syncthreads();                                                                                                     \
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");

// publish the operand a compute warp has just written and count it in on `bar`
#define MW_PUBLISH(bar)                                                \
  do {                                                                 \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  \
    __syncwarp();                                                      \
    if (lane == 0) mw_arrive(smem_u32(&(bar)));                        \
  } while (0)

// ----------------------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MW_THREADS, 1) mem_ws_fwd_kernel(const __grid_constant__ mfm_mem_args a) {
  MW_PROLOGUE();
  mw_build(smem + d.offW[0], d.KM, a.W1m, a.ld_w1m, 1, g1, mem, 128, tid);       // rows = units of gamma1_fc1, K = mem
  mw_build(smem + d.offW[1], d.KM, a.W2m, a.ld_w2m, 1, g2, mem, 128, tid);
  mw_build(smem + d.offW[2], d.K1, a.W12, g1, 1, mem, g1, 64, tid);              // rows = memory units (two copies), K = g1
  mw_build(smem + d.offW[3], d.K2, a.W22, g2, 1, mem, g2, 64, tid);
  for (int idx = tid; idx < MW_NB * mem; idx += MW_THREADS) {                    // block 0 of the history: the zero state
    const int r = idx / mem, j = idx - r * mem;
    if (row0 + r < B) a.mems[(long long)(row0 + r) * mem + j] = 0.0f;
  }
  MW_START();
  unsigned char* const opM = smem + d.offOp[0];
  unsigned char* const opU1 = smem + d.offOp[1];
  unsigned char* const opU2 = smem + d.offOp[2];
  const int planeM = (d.KM >> 3) * MW_LBOB;

  if (warp >= MW_CW) {
    // ================================ MMA issue (one thread per warp) ================================
    if (lane == 0) {
      const int q = warp - MW_CW, k = q & 1, kk0 = q >> 1;
      const uint32_t tA = tmem_base + (uint32_t)(q * MW_NB), tB = tmem_base + (uint32_t)(64 + q * MW_NB);
      for (int t = 0; t < T; ++t) {
        if (t > 0) mw_wait(smem_u32(&bars[0]), (uint32_t)((t - 1) & 1));         // mem_{t-1} is in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (q == 3) MW_STAMP(t, 0);
        mw_issue(tA, smem_u32(smem + d.offW[k]), d.KM, smem_u32(opM), kk0, idesc);
        umma_commit(smem_u32(&bars[1]));
        if (q == 3) MW_STAMP(t, 1);
        mw_wait(smem_u32(&bars[2]), (uint32_t)(t & 1));                          // u_1, u_2 are in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (q == 3) MW_STAMP(t, 2);
        mw_issue(tB, smem_u32(smem + d.offW[2 + k]), k ? d.K2 : d.K1, smem_u32(k ? opU2 : opU1), kk0, idesc);
        umma_commit(smem_u32(&bars[3]));
        if (q == 3) MW_STAMP(t, 3);
      }
    }
  } else {
    // ================================ elementwise stages (warps 0..7) ================================
    const int q = warp & 3, half = warp >> 2;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    // wide stage: tile k = half, unit u
    const int gk = half ? g2 : g1, u = q * 32 + lane;
    const bool u_on = u < gk;
    const int uc = u_on ? u : 0;
    const float* const gpre = half ? a.G2pre : a.G1pre;
    float* const Uo = half ? a.U2 : a.U1;
    unsigned char* const opU = half ? opU2 : opU1;
    const int planeU = ((half ? d.K2 : d.K1) >> 3) * MW_LBOB;
    const bool two_a = d.KM >= 32;                                               // a second partial exists
    const float dp = half ? a.drop_p2 : a.drop_p1;
    const bool dd = dp > 0.0f;
    const uint32_t ss = dd ? site_seed(a.rng, half ? a.site2 : a.site1) : 0u;
    const float ks = dd ? 1.0f / (1.0f - dp) : 1.0f;
    // mem stage: unit j, 4 columns from c0
    const int j = (q & 1) * 32 + lane, c0 = ((q >> 1) * 2 + half) * 4;
    const bool j_on = j < mem;
    const int jc = j_on ? j : 0;
    const float b1 = __ldg(a.b12 + jc), b2 = __ldg(a.b22 + jc);
    const bool two_1 = d.K1 >= 32, two_2 = d.K2 >= 32;
    float mprev[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int nvalid = B - row0;                                                 // batch columns below this are real rows

    // the stash reads of a step are issued one stage ahead of their use: the chain never waits for HBM
    float gp[16], ch[4];
#pragma unroll
    for (int b = 0; b < 16; ++b) gp[b] = __ldg(gpre + (long long)(row0 + min(b, nvalid - 1)) * gk + uc);
    for (int t = 0; t < T; ++t) {
      const long long tb = (long long)t * B;
      // ---- stage 1: u_k = dropout(relu(Gkpre[t] + mem W_km^T)) ----
#pragma unroll
      for (int c = 0; c < 4; ++c) ch[c] = __ldg(a.cHat + (tb + row0 + min(c0 + c, nvalid - 1)) * mem + jc);
      if (tid == 0) MW_STAMP(t, 4);
      mw_wait(smem_u32(&bars[1]), (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) MW_STAMP(t, 5);
      {
        float acc[16], acc2[16];
        mw_ld16(tl + (uint32_t)(half * MW_NB), acc);
        if (two_a) mw_ld16(tl + (uint32_t)((2 + half) * MW_NB), acc2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (tid == 0) MW_STAMP(t, 6);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          float v = gp[b] + acc[b];
          if (two_a) v += acc2[b];
          v = fmaxf(v, 0.0f);
          const long long tr = tb + row0 + b;
          if (dd) v = drop_keep(ss, (uint32_t)tr * (uint32_t)gk + (uint32_t)u, dp) ? v * ks : 0.0f;
          const int ok = (u_on && b < nvalid) ? 1 : 0;
          mw_st_if(Uo + (tb + row0 + min(b, nvalid - 1)) * gk + uc, v, ok);
          if (u_on) mw_put(opU, planeU, b, u, b < nvalid ? v : 0.0f);
        }
      }
      if (tid == 0) MW_STAMP(t, 7);
      MW_PUBLISH(bars[2]);
      if (tid == 0) MW_STAMP(t, 8);
      // ---- stage 2: gamma_k = sigmoid(u_k W_k2^T + b_k2);  mem' = gamma_1 mem + gamma_2 cHat[t] ----
      if (t + 1 < T) {
#pragma unroll
        for (int b = 0; b < 16; ++b) gp[b] = __ldg(gpre + (tb + B + row0 + min(b, nvalid - 1)) * gk + uc);
      }
      if (tid == 0) MW_STAMP(t, 9);
      mw_wait(smem_u32(&bars[3]), (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) MW_STAMP(t, 10);
      {
        float p1[4], p1b[4], p2[4], p2b[4];
        mw_ld4(tl + (uint32_t)(64 + c0), p1);
        mw_ld4(tl + (uint32_t)(64 + MW_NB + c0), p2);
        if (two_1) mw_ld4(tl + (uint32_t)(64 + 2 * MW_NB + c0), p1b);
        if (two_2) mw_ld4(tl + (uint32_t)(64 + 3 * MW_NB + c0), p2b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float x1 = b1 + p1[c], x2 = b2 + p2[c];
          if (two_1) x1 += p1b[c];
          if (two_2) x2 += p2b[c];
          const float ga1 = gate_sigmoid(x1), ga2 = gate_sigmoid(x2);
          const int col = c0 + c;
          const int ok = (j_on && col < nvalid) ? 1 : 0;
          const float nm = col < nvalid ? ga1 * mprev[c] + ga2 * ch[c] : 0.0f;
          mprev[c] = nm;
          const long long o = (tb + row0 + min(col, nvalid - 1)) * mem + jc;
          mw_st_if(a.Gam1 + o, ga1, ok);
          mw_st_if(a.Gam2 + o, ga2, ok);
          mw_st_if(a.mems + o + (long long)B * mem, nm, ok);
          if (j_on) mw_put(opM, planeM, col, j, nm);
        }
      }
      if (tid == 0) MW_STAMP(t, 11);
      if (t + 1 < T) MW_PUBLISH(bars[0]);
      if (tid == 0) MW_STAMP(t, 12);
    }
  }
  MW_END();
}

// ----------------------------------------------------------------------------------------------------------------
// backward (reverse time)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MW_THREADS, 1) mem_ws_bwd_kernel(const __grid_constant__ mfm_mem_args a) {
  MW_PROLOGUE();
  mw_build(smem + d.offW[0], d.KM, a.W12, 1, g1, g1, mem, 128, tid);             // rows = units of gamma1_fc1: W12^T, K = mem
  mw_build(smem + d.offW[1], d.KM, a.W22, 1, g2, g2, mem, 128, tid);
  mw_build(smem + d.offW[2], d.K1, a.W1m, 1, a.ld_w1m, mem, g1, 64, tid);        // rows = memory units (two copies): W1m^T, K = g1
  mw_build(smem + d.offW[3], d.K2, a.W2m, 1, a.ld_w2m, mem, g2, 64, tid);
  MW_START();
  const int planeM = (d.KM >> 3) * MW_LBOB;
  unsigned char* const opP1 = smem + d.offOp[0];
  unsigned char* const opP2 = opP1 + 2 * planeM;
  unsigned char* const opU1 = smem + d.offOp[1];
  unsigned char* const opU2 = smem + d.offOp[2];

  if (warp >= MW_CW) {
    if (lane == 0) {
      const int q = warp - MW_CW, k = q & 1, kk0 = q >> 1;
      const uint32_t tA = tmem_base + (uint32_t)(q * MW_NB), tB = tmem_base + (uint32_t)(64 + q * MW_NB);
      for (int s = 0; s < T; ++s) {
        mw_wait(smem_u32(&bars[0]), (uint32_t)(s & 1));                          // dp_1, dp_2 of this step are in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mw_issue(tA, smem_u32(smem + d.offW[k]), d.KM, smem_u32(k ? opP2 : opP1), kk0, idesc);
        umma_commit(smem_u32(&bars[1]));
        if (s + 1 < T) {                                                         // (the first time step hands nothing further back)
          mw_wait(smem_u32(&bars[2]), (uint32_t)(s & 1));                        // du_1, du_2
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          mw_issue(tB, smem_u32(smem + d.offW[2 + k]), k ? d.K2 : d.K1, smem_u32(k ? opU2 : opU1), kk0, idesc);
          umma_commit(smem_u32(&bars[3]));
        }
      }
    }
  } else {
    const int q = warp & 3, half = warp >> 2;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const int gk = half ? g2 : g1, u = q * 32 + lane;
    const bool u_on = u < gk;
    const int uc = u_on ? u : 0;
    const float* const Us = half ? a.U2 : a.U1;
    float* const dUo = half ? a.dU2 : a.dU1;
    const long long ldu = half ? (a.ld_dU2 ? a.ld_dU2 : g2) : (a.ld_dU1 ? a.ld_dU1 : g1);
    const float sc = half ? a.scale2 : a.scale1;
    unsigned char* const opU = half ? opU2 : opU1;
    const int planeU = ((half ? d.K2 : d.K1) >> 3) * MW_LBOB;
    const bool two_a = d.KM >= 32;
    const int j = (q & 1) * 32 + lane, c0 = ((q >> 1) * 2 + half) * 4;
    const bool j_on = j < mem;
    const int jc = j_on ? j : 0;
    const bool two_1 = d.K1 >= 32, two_2 = d.K2 >= 32;
    const int nvalid = B - row0;
    float dm[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      dm[c] = (j_on && c0 + c < nvalid) ? __ldg(a.dmem_last + (long long)(row0 + c0 + c) * a.ld_dmem_last + j) : 0.0f;

    // the stash reads of a step are issued about a step ahead of their use: the chain never waits for HBM
    float in0[4][4], uv[16];                                   // (mem_{t-1}, gamma_1, gamma_2, cHat) of 4 columns; u_k of 16
    auto load_in0 = [&](int t) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const long long o = ((long long)t * B + row0 + min(c0 + c, nvalid - 1)) * mem + jc;
        in0[c][0] = a.mems[o]; in0[c][1] = a.Gam1[o]; in0[c][2] = a.Gam2[o]; in0[c][3] = __ldg(a.cHat + o);
      }
    };
    auto load_uv = [&](int t) {
#pragma unroll
      for (int b = 0; b < 16; ++b) uv[b] = Us[((long long)t * B + row0 + min(b, nvalid - 1)) * gk + uc];
    };
    load_in0(T - 1);
    load_uv(T - 1);
    for (int s = 0; s < T; ++s) {
      const int t = T - 1 - s;
      const long long tb = (long long)t * B;
      // ---- stage 0: gate gradients of this step; dm <- dm * gamma_1 (the direct path) ----
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = c0 + c;
        const long long o = (tb + row0 + min(col, nvalid - 1)) * mem + jc;
        const float mp = in0[c][0], ga1 = in0[c][1], ga2 = in0[c][2], chv = in0[c][3];
        const float dmv = dm[c];
        const float dp1 = dmv * mp * ga1 * (1.0f - ga1), dp2 = dmv * chv * ga2 * (1.0f - ga2);
        const int ok = (j_on && col < nvalid) ? 1 : 0;
        mw_st_if(a.dP1 + o, dp1, ok);
        mw_st_if(a.dP2 + o, dp2, ok);
        mw_st_if(a.dPc + o, dmv * ga2 * (1.0f - chv * chv), ok);
        dm[c] = dmv * ga1;
        if (j_on) {
          mw_put(opP1, planeM, col, j, dp1);
          mw_put(opP2, planeM, col, j, dp2);
        }
      }
      MW_PUBLISH(bars[0]);
      if (t > 0) load_in0(t - 1);
      // ---- stage 1: du_k = (dp_k W_k2) * relu/dropout mask ----
      mw_wait(smem_u32(&bars[1]), (uint32_t)(s & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        float acc[16], acc2[16];
        mw_ld16(tl + (uint32_t)(half * MW_NB), acc);
        if (two_a) mw_ld16(tl + (uint32_t)((2 + half) * MW_NB), acc2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          float v = acc[b];
          if (two_a) v += acc2[b];
          v = (uv[b] > 0.0f && b < nvalid) ? v * sc : 0.0f;
          mw_st_if(dUo + (tb + row0 + min(b, nvalid - 1)) * ldu + uc, v, (u_on && b < nvalid) ? 1 : 0);
          if (u_on) mw_put(opU, planeU, b, u, v);
        }
      }
      if (s + 1 < T) {
        MW_PUBLISH(bars[2]);
        load_uv(t - 1);
        // ---- stage 2: dmem_{t-1} = dm * gamma_1 + du_1 W_1m + du_2 W_2m ----
        mw_wait(smem_u32(&bars[3]), (uint32_t)(s & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float p1[4], p1b[4], p2[4], p2b[4];
        mw_ld4(tl + (uint32_t)(64 + c0), p1);
        mw_ld4(tl + (uint32_t)(64 + MW_NB + c0), p2);
        if (two_1) mw_ld4(tl + (uint32_t)(64 + 2 * MW_NB + c0), p1b);
        if (two_2) mw_ld4(tl + (uint32_t)(64 + 3 * MW_NB + c0), p2b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v = dm[c] + p1[c] + p2[c];
          if (two_1) v += p1b[c];
          if (two_2) v += p2b[c];
          dm[c] = v;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
  }
  MW_END();
}

static unsigned long long g_mw_counts[2];
static int g_mw_off = 0;
extern "C" int mfm_debug_mem_ws_trace(long long* host32x16) {
#if MW_DEBUG
  return (int)cudaMemcpyFromSymbol(host32x16, g_mw_trace, sizeof(long long) * 32 * 16);
#else
  (void)host32x16;
  return MFM_ERR_UNSUPPORTED;
#endif
}
extern "C" int mfm_debug_mem_force_simt(int on) { g_mw_off = on; return MFM_OK; }
extern "C" unsigned long long mfm_debug_mem_ws_count(int bwd) { return g_mw_counts[bwd ? 1 : 0]; }

// MFM_ERR_UNSUPPORTED: the caller (mfn.cu) runs the CUDA-core kernel
int mem_ws_launch(const mfm_mem_args* a, bool bwd, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MFM_MEM_WS"); enabled = e ? atoi(e) : 1; }
  if (!enabled || g_mw_off) return MFM_ERR_UNSUPPORTED;
  if (a->mem > 64 || a->g1 > 128 || a->g2 > 128) return MFM_ERR_UNSUPPORTED;
  const MwDims d = mw_dims(a->mem, a->g1, a->g2);
  if (d.total > mfm_dev_info().smem_optin - 1024) return MFM_ERR_UNSUPPORTED;
  const int grid = (a->B + MW_NB - 1) / MW_NB;
  if (bwd) {
    if (int e = mfm_func_smem_t(mem_ws_bwd_kernel, mfm_dev_info().smem_optin - 1024)) return e;
    mem_ws_bwd_kernel<<<grid, MW_THREADS, d.total, st>>>(*a);
  } else {
    if (int e = mfm_func_smem_t(mem_ws_fwd_kernel, mfm_dev_info().smem_optin - 1024)) return e;
    mem_ws_fwd_kernel<<<grid, MW_THREADS, d.total, st>>>(*a);
  }
  MFM_LAUNCH_CHECK();
  ++g_mw_counts[bwd ? 1 : 0];
  return MFM_OK;
}
