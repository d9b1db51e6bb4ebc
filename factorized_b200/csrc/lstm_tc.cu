// LSTM recurrences on the tensor cores, forward and backward (mfm_model.py:56,83,85,167-169 and their adjoint).
//
// The gate GEMM is issued TRANSPOSED:   D^T[gate-unit n, batch b] = W[n, :] . h_{t-1}[b, :]
//   A operand = W (M = 128 gate-units per MMA tile, K-major = W's own row-major layout), resident in shared memory as
//               split-bf16 (hi + lo) for the whole sequence;
//   B operand = h_{t-1} for the CTA's NB batch rows (N = NB), written by the previous step's epilogue straight into
//               shared memory in the canonical K-major layout (the recurrence never leaves the SM);
//   D         = fp32 in TMEM: lane = gate-unit, column = batch row; 3 MMAs per k-step (hi*hi + lo*hi + hi*lo).
// Why transposed: an epilogue thread owns a TMEM lane.  With lane = gate-unit, the 32 lanes of a warp touch 32
// CONSECUTIVE floats of one row of the row-major stashes (G_x, gates, dG, c, h), i.e. one 128 B line per access.
// The first version had lane = batch row: every access hit 32 different lines and the kernels were bound by L1
// wavefronts (64 cycles per 512 B), 8x slower than the tensor-core work they wrapped.
// A thread sees ONE gate of a unit, so the four gates meet through a small shared-memory exchange (8 batch columns at
// a time), after which (unit, batch) items finish the cell update, again unit-fastest (coalesced).
//
// Backward:  dh^T[unit j, batch b] = sum_k' W^T[j, k'] dG[b, k'],  k' = 4*unit + gate (unit-major so the four gate
// gradients of a unit are 8 contiguous bytes of the B operand); thread (unit j, 8 or 4 batch columns) keeps the
// carried dc in registers, reads dh from TMEM, and all its global traffic is unit-fastest as well.
// Cells with h > 128 (TMEM lanes / shared memory) run on lstm_seq.cu instead.
#include "tc_common.cuh"

#define L2_THREADS 512

struct Lstm2Cell {
  mfm_lstm_cell c;
  int nb;          // batch rows per CTA = UMMA N (32 or 16)
  int tiles;       // ceil(B / nb)
  int kp;          // forward: h rounded up to 16 (MMA K).  backward: 4 * (h rounded up to 8)
  int mtiles;      // forward: ceil(4h / 128)
  int tmem_cols;
};
struct Lstm2Batch {
  Lstm2Cell c[MFM_MAX_CELLS];
  int n;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
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

// ----------------------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------------------
// CC = batch columns per gate-exchange chunk: 32 (one exchange per step) when the buffer fits next to W, else 8
template <int CC>
__global__ void __launch_bounds__(L2_THREADS, 1) lstm_tc_fwd_kernel(Lstm2Batch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[4];       // one per MMA tile: a tile's warps start as soon as it is done
  __shared__ uint32_t tmem_holder;
  const Lstm2Cell& lc = bt.c[blockIdx.y];
  if ((int)blockIdx.x >= lc.tiles) return;
  const mfm_lstm_cell& c = lc.c;
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  const int NB = lc.nb, KP = lc.kp, mtiles = lc.mtiles;
  const int slabs = KP >> 3;
  const int lboA = H4 * 16 + 32, lboH = NB * 16 + 32;
  unsigned char* Whi = smem;
  unsigned char* Wlo = Whi + slabs * lboA;
  unsigned char* Hhi = Wlo + slabs * lboA;                 // the last MMA tile may read up to 127 rows past W: lands here, ignored lanes
  unsigned char* Hlo = Hhi + slabs * lboH;
  float* Gs = reinterpret_cast<float*>(Hlo + slabs * lboH + 2048);     // [4h][CC + 1] activated gates of one column chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * NB;

  if (tid == 0) {
    for (int m = 0; m < 4; ++m) mbar_init(smem_u32(&bars[m]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)lc.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // W [4h,h] fp32 -> resident split-bf16 K-major A operand (k >= h zero)
  {
    const bool vecW = ((reinterpret_cast<uintptr_t>(c.W) & 15) == 0) && ((h & 3) == 0);
    for (int idx = tid; idx < H4 * slabs; idx += L2_THREADS) {
      const int slab = idx % slabs, n = idx / slabs;
      float v[8];
      load8(c.W, h, n, H4, slab * 8, h, vecW, v);
      split_store(v, Whi + slab * lboA + n * 16, Wlo + slab * lboA + n * 16, true);
    }
    for (int idx = tid * 16; idx < 2 * slabs * lboH + 2048; idx += L2_THREADS * 16)
      *reinterpret_cast<uint4*>(Hhi + idx) = make_uint4(0, 0, 0, 0);       // h_{-1} = 0, K padding = 0
  }
  // block 0 of the histories is the zero initial state
  for (int idx = tid; idx < NB * h; idx += L2_THREADS) {
    const int b = row0 + idx / h, j = idx % h;
    if (b < B) {
      c.hs[(long long)b * c.ld_hs + j] = 0.0f;
      c.cs[(long long)b * c.ld_cs + j] = 0.0f;
      if (c.cs_dup) c.cs_dup[(long long)b * c.ld_cs + j] = 0.0f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  // phase-1 identity: MMA tile mt, TMEM lane quadrant q, gate-unit n
  const int mt = warp >> 2, q = warp & 3;
  const int n = mt * 128 + q * 32 + lane;
  const bool tile_on = mt < mtiles;              // warp-uniform
  const bool n_on = tile_on && n < H4;
  const bool is_tanh = n_on && (n / h == 2);
  // sigmoid and tanh share one code path: tanh(x) = 2*sigmoid(2x) - 1  ->  act = aa * rcp(1 + ex2(kk * x)) + bb
  const float act_k = (is_tanh ? -2.0f : -1.0f) * 1.4426950408889634f, act_a = is_tanh ? 2.0f : 1.0f, act_b = is_tanh ? -1.0f : 0.0f;
  const float act_clamp = is_tanh ? 15.0f : 30.0f;
  const bool full = (NB == 32) && (row0 + 32 <= B);     // CTA-uniform
  const float bias_n = (n_on && c.bias_rest) ? __ldg(c.bias_rest + n) : 0.0f;
  const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NB);
  // phase-2 identity: items (cc, j) of a column chunk, unit j fastest; at most ITEMS per thread (CC*h <= CC*128)
  constexpr int ITEMS = CC * 128 / L2_THREADS, NCH = 32 / CC;
  int it_cc[ITEMS], it_j[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int item = tid + k * L2_THREADS;
    it_cc[k] = item < CC * h ? item / h : -1;
    it_j[k] = item % h;
  }
  float cprev[NCH][ITEMS];                        // c_{t-1} of this thread's items, per column chunk
#pragma unroll
  for (int a = 0; a < NCH; ++a)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) cprev[a][k] = 0.0f;
  const int nchunk = NB / CC;

  for (int t = 0; t < T; ++t) {
    // G_x[t] (or the decoder's constant bias) for this thread's gate-unit and all NB columns: issued before the MMA wait
    float gxv[32];
    if (n_on) {
      if (t < c.gx_steps) {
        const float* gp = c.gx + ((long long)t * B + row0) * H4 + n;
        if (full) {                                  // whole tile in range: no per-element predicates
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) gxv[cc] = __ldg(gp + (long long)cc * H4);
        } else {
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) gxv[cc] = (cc < NB && row0 + cc < B) ? __ldg(gp + (long long)cc * H4) : 0.0f;
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) gxv[cc] = bias_n;
      }
    }
    if (t > 0) {
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_u32(Whi), aL = smem_u32(Wlo), bH = smem_u32(Hhi), bL = smem_u32(Hlo);
        for (int m = 0; m < mtiles; ++m) {
          const uint32_t dcol = tmem_base + (uint32_t)(m * NB);
          for (int kk = 0; kk < (KP >> 4); ++kk) {
            const uint32_t ao = m * 128 * 16 + kk * 2 * lboA, bo = kk * 2 * lboH;
            const uint64_t dAh = make_smem_desc(aH + ao, lboA, 128), dAl = make_smem_desc(aL + ao, lboA, 128);
            const uint64_t dBh = make_smem_desc(bH + bo, lboH, 128), dBl = make_smem_desc(bL + bo, lboH, 128);
            umma_bf16(dcol, dAh, dBh, idesc, kk > 0 ? 1u : 0u);
            umma_bf16(dcol, dAl, dBh, idesc, 1u);
            umma_bf16(dcol, dAh, dBl, idesc, 1u);
          }
          umma_commit(smem_u32(&bars[m]));          // per-tile completion: tile m's warps need not wait for tiles m+1..
        }
      }
      if (tile_on) {
        mbar_wait(smem_u32(&bars[mt]), (uint32_t)((t - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (ch >= nchunk) break;
      const int bc0 = ch * CC;
      // phase 1: activate this thread's gate for CC batch columns; stash it; hand it to the exchange buffer
      if (tile_on) {
#pragma unroll
        for (int sub = 0; sub < CC / 8; ++sub) {
          float acc[8];
          if (t > 0) {
            tmem_ld8(tlane + (uint32_t)(bc0 + 8 * sub), acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
          }
          if (n_on) {
            float* gout = c.gates + ((long long)t * B + row0 + bc0 + 8 * sub) * H4 + n;
            float* gs = Gs + n * (CC + 1) + 8 * sub;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const int col = bc0 + 8 * sub + cc;
              const float pre = fminf(fmaxf(acc[cc] + gxv[col], -act_clamp), act_clamp);
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(pre * act_k));
              const float av = fmaf(act_a, rcp_fast(1.0f + e), act_b);
              if (full || row0 + col < B) gout[(long long)cc * H4] = av;
              gs[cc] = av;
            }
          }
        }
      }
      __syncthreads();
      // phase 2: (column, unit) items: c_t, h_t, histories, and h_t as next step's B operand
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int cc = it_cc[k], j = it_j[k];
        if (cc < 0) continue;
        const int bl = bc0 + cc, b = row0 + bl;
        float hn = 0.0f;
        if (b < B) {
          const float ig = Gs[j * (CC + 1) + cc], fg = Gs[(h + j) * (CC + 1) + cc];
          const float gg = Gs[(2 * h + j) * (CC + 1) + cc], og = Gs[(3 * h + j) * (CC + 1) + cc];
          const float cn = fg * cprev[ch][k] + ig * gg;
          hn = og * gate_tanh(cn);
          cprev[ch][k] = cn;
          c.cs[((long long)(t + 1) * B + b) * c.ld_cs + j] = cn;
          if (c.cs_dup) c.cs_dup[((long long)(t + 1) * B + b) * c.ld_cs + j] = cn;
          c.hs[((long long)(t + 1) * B + b) * c.ld_hs + j] = hn;
        }
        const unsigned short hb = bf16_bits(hn);
        const float hr = hn - __uint_as_float((uint32_t)hb << 16);
        const int off = (j >> 3) * lboH + bl * 16 + (j & 7) * 2;
        *reinterpret_cast<unsigned short*>(Hhi + off) = hb;
        *reinterpret_cast<unsigned short*>(Hlo + off) = bf16_bits(hr);
      }
      if (ch + 1 < nchunk) __syncthreads();        // Gs is rewritten by the next chunk
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)lc.tmem_cols)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------------------------
template <int CPT>      // batch columns per thread: 8 (NB = 32) or 4 (NB = 16)
__global__ void __launch_bounds__(L2_THREADS, 1) lstm_tc_bwd_kernel(Lstm2Batch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_holder;
  const Lstm2Cell& lc = bt.c[blockIdx.y];
  if ((int)blockIdx.x >= lc.tiles) return;
  const mfm_lstm_cell& c = lc.c;
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  constexpr int NB = 4 * CPT;
  const int KP = lc.kp;                               // 4 * hp8, a multiple of 32
  const int slabs = KP >> 3;
  const int lboA = h * 16 + 32, lboB = NB * 16 + 32;
  unsigned char* Ahi = smem;                          // W^T: A[j][k'], k' = 4*unit + gate
  unsigned char* Alo = Ahi + slabs * lboA;
  unsigned char* Bhi = Alo + slabs * lboA + 2048;     // the MMA tile reads rows j in [h,128): lands in valid memory, lanes ignored
  unsigned char* Blo = Bhi + slabs * lboB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * NB;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)lc.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[j][k' = 4*jj + g] = W[g*h + jj][j]; item = (j, slab): 8 gathered values, j fastest (coalesced)
  for (int idx = tid; idx < h * slabs; idx += L2_THREADS) {
    const int j = idx % h, slab = idx / h;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int jj = 2 * slab + (e >> 2), g = e & 3;
      v[e] = jj < h ? __ldg(c.W + (long long)(g * h + jj) * h + j) : 0.0f;
    }
    split_store(v, Ahi + slab * lboA + j * 16, Alo + slab * lboA + j * 16, true);
  }
  for (int idx = tid * 16; idx < 2 * slabs * lboB; idx += L2_THREADS * 16)
    *reinterpret_cast<uint4*>(Bhi + idx) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  // thread = (unit j = TMEM lane, CPT batch columns)
  const int q = warp & 3, cg = warp >> 2;
  const int j = q * 32 + lane;
  const bool j_on = j < h;
  const int hp8 = KP >> 2;
  const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CPT);
  float dc[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) dc[i] = 0.0f;

  for (int t = T - 1; t >= 0; --t) {
    // everything that does not depend on the recurrence is loaded before the MMA wait
    float ig[CPT], fg[CPT], gg[CPT], og[CPT], cp[CPT], cn[CPT], dhx[CPT], dcx[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int b = row0 + cg * CPT + i;
      ig[i] = fg[i] = gg[i] = og[i] = cp[i] = cn[i] = dhx[i] = dcx[i] = 0.0f;
      if (j_on && b < B) {
        const long long tr = (long long)t * B + b;
        const float* gp = c.gates + tr * H4 + j;
        ig[i] = gp[0]; fg[i] = gp[h]; gg[i] = gp[2 * h]; og[i] = gp[3 * h];
        cp[i] = c.cs[tr * c.ld_cs + j];
        cn[i] = c.cs[(tr + B) * c.ld_cs + j];
        if (c.dh_all) dhx[i] = __ldg(c.dh_all + tr * c.ld_dh_all + j);
        if (c.dh_last && t == T - 1) dhx[i] += __ldg(c.dh_last + (long long)b * c.ld_dh_last + j);
        if (c.dc_ext) dcx[i] = __ldg(c.dc_ext + tr * c.ld_dc_ext + j);
        if (c.dc_ext2 && t < T - 1) dcx[i] += __ldg(c.dc_ext2 + tr * c.ld_dc_ext + j);
      }
    }
    float dh[CPT];
    if (t < T - 1) {                                  // dh_rec of this step = the previous step's MMAs
      mbar_wait(smem_u32(&bar), (uint32_t)((T - 2 - t) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dprev = tlane + (uint32_t)(((t + 1) & 1) * NB);
      if (CPT == 8) tmem_ld8(dprev, dh); else tmem_ld4(dprev, dh);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int i = 0; i < CPT; ++i) dh[i] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int bl = cg * CPT + i, b = row0 + bl;
      float d_i = 0.0f, d_f = 0.0f, d_g = 0.0f, d_o = 0.0f;
      if (j_on && b < B) {
        const float dht = dh[i] + dhx[i];
        const float tc = gate_tanh(cn[i]);
        const float dci = dc[i] + dht * og[i] * (1.0f - tc * tc) + dcx[i];
        d_i = dci * gg[i] * ig[i] * (1.0f - ig[i]);
        d_f = dci * cp[i] * fg[i] * (1.0f - fg[i]);
        d_g = dci * ig[i] * (1.0f - gg[i] * gg[i]);
        d_o = dht * tc * og[i] * (1.0f - og[i]);
        dc[i] = dci * fg[i];
        float* op = c.dG + ((long long)t * B + b) * H4 + j;
        op[0] = d_i; op[h] = d_f; op[2 * h] = d_g; op[3 * h] = d_o;
      }
      if (j < hp8 && t > 0) {                         // B operand: row = batch column, k' = 4j..4j+3  (8 contiguous bytes)
        const float v4[4] = {d_i, d_f, d_g, d_o};
        unsigned short hb[4], lb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hb[e] = bf16_bits(v4[e]);
          lb[e] = bf16_bits(v4[e] - __uint_as_float((uint32_t)hb[e] << 16));
        }
        const int off = (j >> 1) * lboB + bl * 16 + (j & 1) * 8;
        *reinterpret_cast<uint2*>(Bhi + off) = make_uint2((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), (uint32_t)hb[2] | ((uint32_t)hb[3] << 16));
        *reinterpret_cast<uint2*>(Blo + off) = make_uint2((uint32_t)lb[0] | ((uint32_t)lb[1] << 16), (uint32_t)lb[2] | ((uint32_t)lb[3] << 16));
      }
    }
    if (t > 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcur = tmem_base + (uint32_t)((t & 1) * NB);
        const uint32_t aH = smem_u32(Ahi), aL = smem_u32(Alo), bH = smem_u32(Bhi), bL = smem_u32(Blo);
        for (int kk = 0; kk < (KP >> 4); ++kk) {
          const uint32_t ao = kk * 2 * lboA, bo = kk * 2 * lboB;
          const uint64_t dAh = make_smem_desc(aH + ao, lboA, 128), dAl = make_smem_desc(aL + ao, lboA, 128);
          const uint64_t dBh = make_smem_desc(bH + bo, lboB, 128), dBl = make_smem_desc(bL + bo, lboB, 128);
          umma_bf16(dcur, dAh, dBh, idesc, kk > 0 ? 1u : 0u);
          umma_bf16(dcur, dAl, dBh, idesc, 1u);
          umma_bf16(dcur, dAh, dBl, idesc, 1u);
        }
        umma_commit(smem_u32(&bar));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)lc.tmem_cols)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
static inline int ru(int x, int m) { return (x + m - 1) / m * m; }

static size_t fwd_smem(int h, int nb, int cc) {
  const int slabs = ru(h, 16) / 8;
  return (size_t)2 * slabs * (4 * h * 16 + 32) + (size_t)2 * slabs * (nb * 16 + 32) + 2048 + (size_t)4 * h * (cc + 1) * 4 + 128;
}
static size_t bwd_smem(int h, int nb) {
  const int slabs = 4 * ru(h, 8) / 8;
  return (size_t)2 * slabs * (h * 16 + 32) + 2048 + (size_t)2 * slabs * (nb * 16 + 32) + 128;
}
static int smem_limit() {
  static int lim = -1;
  if (lim < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) lim = 48 * 1024;
    lim -= 1024;      // the opt-in limit covers static + dynamic shared memory
  }
  return lim;
}

// CTAs are dispatched in blockIdx order (cell = blockIdx.y slowest): put the long-running cells first so the short
// ones fill the tail instead of the other way round.
static void sort_heavy_first(Lstm2Batch& bt) {
  for (int i = 1; i < bt.n; ++i) {
    Lstm2Cell key = bt.c[i];
    int j = i - 1;
    while (j >= 0 && bt.c[j].c.h < key.c.h) { bt.c[j + 1] = bt.c[j]; --j; }
    bt.c[j + 1] = key;
  }
}

// Launches the tensor-core forward for every cell that fits; cells that do not are returned in `rest`.
int lstm_tc_fwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  const int lim = smem_limit();
  Lstm2Batch b32, b8;          // by exchange-chunk width
  b32.n = b8.n = 0;
  *nrest = 0;
  size_t smem32 = 0, smem8 = 0;
  int gx32 = 0, gx8 = 0;
  // one CTA per SM: when 32-row tiles of the whole call would leave more than half of the SMs idle (a single decoder
  // cell at batch 2048 is 64 tiles), 16-row tiles double the CTAs and shorten every step of the recurrence
  long long tiles32 = 0;
  for (int i = 0; i < ncells; ++i) tiles32 += (cells[i].B + 31) / 32;
  const bool narrow = tiles32 <= 74;
  for (int i = 0; i < ncells; ++i) {
    const mfm_lstm_cell& c = cells[i];
    int nb = 0, cc = 0;
    if (c.h >= 1 && c.h <= 128) {
      if (narrow && fwd_smem(c.h, 16, 8) <= (size_t)lim) { nb = 16; cc = 8; }
      else if (fwd_smem(c.h, 32, 32) <= (size_t)lim) { nb = 32; cc = 32; }
      else if (fwd_smem(c.h, 32, 8) <= (size_t)lim) { nb = 32; cc = 8; }
      else if (fwd_smem(c.h, 16, 8) <= (size_t)lim) { nb = 16; cc = 8; }
    }
    if (!nb) { rest[(*nrest)++] = c; continue; }
    Lstm2Batch& bt = cc == 32 ? b32 : b8;
    Lstm2Cell& lc = bt.c[bt.n++];
    lc.c = c;
    lc.nb = nb;
    lc.tiles = (c.B + nb - 1) / nb;
    lc.kp = ru(c.h, 16);
    lc.mtiles = (4 * c.h + 127) / 128;
    int cols = 32;
    while (cols < lc.mtiles * nb) cols <<= 1;
    lc.tmem_cols = cols;
    const size_t sm = fwd_smem(c.h, nb, cc);
    if (cc == 32) { smem32 = smem32 > sm ? smem32 : sm; gx32 = gx32 > lc.tiles ? gx32 : lc.tiles; }
    else          { smem8 = smem8 > sm ? smem8 : sm; gx8 = gx8 > lc.tiles ? gx8 : lc.tiles; }
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_tc_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  sort_heavy_first(b32);
  sort_heavy_first(b8);
  if (b8.n) {        // the widest cells first: they are the long pole
    lstm_tc_fwd_kernel<8><<<dim3(gx8, b8.n), L2_THREADS, smem8, st>>>(b8);
    MFM_LAUNCH_CHECK();
  }
  if (b32.n) {
    lstm_tc_fwd_kernel<32><<<dim3(gx32, b32.n), L2_THREADS, smem32, st>>>(b32);
    MFM_LAUNCH_CHECK();
  }
  return MFM_OK;
}

int lstm_tc_bwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  const int lim = smem_limit();
  Lstm2Batch b32, b16;
  b32.n = b16.n = 0;
  *nrest = 0;
  size_t smem32 = 0, smem16 = 0;
  int gx32 = 0, gx16 = 0;
  for (int i = 0; i < ncells; ++i) {
    const mfm_lstm_cell& c = cells[i];
    int nb = 0;
    if (c.h >= 1 && c.h <= 128) {             // (16-row tiles for under-occupied launches, as in forward, measured slower here)
      if (bwd_smem(c.h, 32) <= (size_t)lim) nb = 32;
      else if (bwd_smem(c.h, 16) <= (size_t)lim) nb = 16;
    }
    if (!nb) { rest[(*nrest)++] = c; continue; }
    Lstm2Batch& bt = nb == 32 ? b32 : b16;
    Lstm2Cell& lc = bt.c[bt.n++];
    lc.c = c;
    lc.nb = nb;
    lc.tiles = (c.B + nb - 1) / nb;
    lc.kp = 4 * ru(c.h, 8);
    lc.mtiles = 1;
    lc.tmem_cols = nb == 32 ? 64 : 32;
    if (nb == 32) { smem32 = smem32 > bwd_smem(c.h, 32) ? smem32 : bwd_smem(c.h, 32); gx32 = gx32 > lc.tiles ? gx32 : lc.tiles; }
    else          { smem16 = smem16 > bwd_smem(c.h, 16) ? smem16 : bwd_smem(c.h, 16); gx16 = gx16 > lc.tiles ? gx16 : lc.tiles; }
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_tc_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  sort_heavy_first(b32);
  sort_heavy_first(b16);
  if (b32.n) {
    lstm_tc_bwd_kernel<8><<<dim3(gx32, b32.n), L2_THREADS, smem32, st>>>(b32);
    MFM_LAUNCH_CHECK();
  }
  if (b16.n) {
    lstm_tc_bwd_kernel<4><<<dim3(gx16, b16.n), L2_THREADS, smem16, st>>>(b16);
    MFM_LAUNCH_CHECK();
  }
  return MFM_OK;
}
