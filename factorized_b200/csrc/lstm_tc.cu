// LSTM recurrence on the tensor cores: the gate GEMM h_{t-1} W^T is a tcgen05.mma per timestep, fused with the
// sigmoid/tanh/Hadamard cell update, T steps inside the kernel  (mfm_model.py:56,83,85,167-169).
//
// A CTA owns MT (128, or 64 when shared memory is tight) batch rows of one cell for the whole sequence:
//   * W (the recurrent weight, [4h,h]) is converted ONCE to split-bf16 (hi+lo) and stays resident in shared memory
//     as the K-major B operand (no-swizzle canonical layout, see gemm_tc.cu);
//   * h_{t-1} is the A operand: the epilogue of step t-1 writes it (hi+lo bf16) straight into shared memory in the
//     canonical K-major layout, so the recurrence never round-trips through HBM;
//   * per step: one elected thread issues ceil(4h/256) x (hp/16) x 3 MMAs (bf16x3: hi*hi + lo*hi + hi*lo) into a
//     [MT x 4h] fp32 accumulator in TMEM and commits to an mbarrier; every thread (one batch row each) then reads
//     its row's i,f,g,o pre-activations with tcgen05.ld, adds the hoisted x-projection G_x[t] (HBM), applies the
//     gates, updates c, and stores gates / c_t / h_t (the stash the backward pass needs) and the next A operand.
// Cells whose weights do not fit (4h > 512 TMEM columns or > 227 KB shared memory) run on lstm_seq.cu instead.
#include "tc_common.cuh"

#define LT_THREADS 512       // 16 warps: TMEM lane quadrant = warp%4, unit-chunk group = warp/4

struct LstmTcCell {
  mfm_lstm_cell c;
  int mt;        // rows per CTA: 128 or 64
  int tiles;     // ceil(B / mt)
  int hp8;       // h rounded up to 8
  int hp16;      // h rounded up to 16 (MMA K)
  int n4;        // 4h rounded up to 16 (MMA N)
  int bn;        // N per MMA block (<= 256, multiple of 16)
  int nblk;
  int tmem_cols;
};
struct LstmTcBatch {
  LstmTcCell c[MFM_MAX_CELLS];
  int n;
};

__device__ __forceinline__ void ld8_global(const float* __restrict__ p, bool vec, int nvalid, float v[8]) {
  if (vec && nvalid >= 8) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < nvalid ? p[i] : 0.0f;
  }
}
__device__ __forceinline__ void st8_global(float* __restrict__ p, bool vec, int nvalid, const float v[8]) {
  if (vec && nvalid >= 8) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nvalid) p[i] = v[i];
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(LT_THREADS) lstm_tc_fwd_kernel(LstmTcBatch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_holder;
  const LstmTcCell& lc = bt.c[blockIdx.y];
  if ((int)blockIdx.x >= lc.tiles) return;
  const mfm_lstm_cell& c = lc.c;
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  const int MT = lc.mt, hp8 = lc.hp8, hp16 = lc.hp16, n4 = lc.n4;
  const int slabs = hp16 >> 3;
  const int lboW = n4 * 16 + 32, lboH = MT * 16 + 32;
  unsigned char* Whi = smem;
  unsigned char* Wlo = Whi + slabs * lboW;
  unsigned char* Hhi = Wlo + slabs * lboW;
  unsigned char* Hlo = Hhi + slabs * lboH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * MT;

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
  // W [4h,h] fp32 -> resident split-bf16 K-major B operand (rows >= 4h and k >= h are zero)
  {
    const bool vecW = aligned16(c.W) && ((h & 3) == 0);
    const int items = n4 * slabs;
    for (int idx = tid; idx < items; idx += LT_THREADS) {
      const int slab = idx % slabs, n = idx / slabs;
      float v[8];
      load8(c.W, h, n, H4, slab * 8, h, vecW, v);
      split_store(v, Whi + slab * lboW + n * 16, Wlo + slab * lboW + n * 16, true);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT >> 4) << 24);

  // this thread's batch row: TMEM lane 32*warp+lane holds tile row 32*warp+lane (MT=128) or 16*warp+lane (MT=64, lane<16)
  const int quad = warp & 3, cg = warp >> 2;
  const bool owns = (MT == 128) || (lane < 16);
  const int r = (MT == 128) ? (quad * 32 + lane) : (quad * 16 + lane);
  const int row = row0 + r;
  const bool valid = owns && row < B;
  const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const bool vec_gx = aligned16(c.gx) && ((h & 3) == 0);
  const bool vec_gt = aligned16(c.gates) && ((h & 3) == 0);
  const bool vec_hs = aligned16(c.hs) && ((c.ld_hs & 3) == 0);
  const bool vec_cs = aligned16(c.cs) && ((c.ld_cs & 3) == 0);
  const bool vec_b = c.bias_rest && aligned16(c.bias_rest) && ((h & 3) == 0);

  if (valid && cg == 0) {   // block 0 of the histories is the zero initial state
    for (int j = 0; j < h; ++j) {
      c.hs[(long long)row * c.ld_hs + j] = 0.0f;
      c.cs[(long long)row * c.ld_cs + j] = 0.0f;
    }
  }

  for (int t = 0; t < T; ++t) {
    if (t > 0) {
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_u32(Hhi), aL = smem_u32(Hlo), bH = smem_u32(Whi), bL = smem_u32(Wlo);
        for (int b = 0; b < lc.nblk; ++b) {
          const int nb0 = b * lc.bn;
          const int nsz = min(lc.bn, n4 - nb0);
          const uint32_t idesc = idesc_base | ((uint32_t)(nsz >> 3) << 17);
          const uint32_t dcol = tmem_base + (uint32_t)nb0;
          for (int kk = 0; kk < (hp16 >> 4); ++kk) {
            const uint32_t ao = kk * 2 * lboH, bo = kk * 2 * lboW + nb0 * 16;
            const uint64_t dAh = make_smem_desc(aH + ao, lboH, 128), dAl = make_smem_desc(aL + ao, lboH, 128);
            const uint64_t dBh = make_smem_desc(bH + bo, lboW, 128), dBl = make_smem_desc(bL + bo, lboW, 128);
            umma_bf16(dcol, dAh, dBh, idesc, kk > 0 ? 1u : 0u);
            umma_bf16(dcol, dAl, dBh, idesc, 1u);
            umma_bf16(dcol, dAh, dBl, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bar));
      }
      mbar_wait(smem_u32(&bar), (uint32_t)((t - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long tr = (long long)t * B + row;
    for (int j0 = 8 * cg; j0 < hp8; j0 += 32) {     // the four warps of a lane quadrant interleave the unit chunks
      float a[4][8];
      if (t > 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld8(tlane + (uint32_t)(g * h + j0), a[g]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) a[g][i] = 0.0f;
      }
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hv[i] = 0.0f;
      if (valid) {
        const int nv = min(8, h - j0);
        float x[4][8], cp[8], cn[8];
        if (t < c.gx_steps) {
#pragma unroll
          for (int g = 0; g < 4; ++g) ld8_global(c.gx + tr * H4 + g * h + j0, vec_gx, nv, x[g]);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) ld8_global(c.bias_rest + g * h + j0, vec_b, nv, x[g]);
        }
        if (t > 0) ld8_global(c.cs + tr * c.ld_cs + j0, vec_cs, nv, cp);     // written by this same thread at step t-1
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cp[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ig = gate_sigmoid(a[0][i] + x[0][i]);
          const float fg = gate_sigmoid(a[1][i] + x[1][i]);
          const float gg = gate_tanh(a[2][i] + x[2][i]);
          const float og = gate_sigmoid(a[3][i] + x[3][i]);
          cn[i] = fg * cp[i] + ig * gg;
          hv[i] = i < nv ? og * gate_tanh(cn[i]) : 0.0f;
          x[0][i] = ig; x[1][i] = fg; x[2][i] = gg; x[3][i] = og;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) st8_global(c.gates + tr * H4 + g * h + j0, vec_gt, nv, x[g]);
        st8_global(c.cs + (tr + B) * c.ld_cs + j0, vec_cs, nv, cn);
        st8_global(c.hs + (tr + B) * c.ld_hs + j0, vec_hs, nv, hv);
      }
      if (owns) split_store(hv, Hhi + (j0 >> 3) * lboH + r * 16, Hlo + (j0 >> 3) * lboH + r * 16, true);
    }
    if (hp16 > hp8 && owns && cg == 0) {
      const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      split_store(z, Hhi + (hp8 >> 3) * lboH + r * 16, Hlo + (hp8 >> 3) * lboH + r * 16, true);
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

// ------------------------------------------------------------------------------------------------------------------
// backward recurrence on the tensor cores.  Per step (t = T-1 .. 0), per CTA (MT batch rows of one cell):
//   elementwise: dh = dh_rec (TMEM, from the previous step's MMAs) + external dh;  dc += dh o (1-tanh^2 c) + external dc;
//                dG_t = (d_i, d_f, d_g, d_o) pre-activation gradients -> HBM (stash for the weight-gradient GEMMs) and,
//                split to bf16 hi/lo, into shared memory as the next A operand;  dc carry (dc*f) -> a [B,h] scratch row;
//   MMA:         dh_rec[MT, h] = dG_t[MT, 4h] W[4h, h]   (K = 4h) into the other half of a double-buffered TMEM tile.
// K is ordered unit-major (k' = 4*j + gate) so the gradients of a group of units form a contiguous K range: the A
// operand is produced and consumed in slots of 32 units while later units are still being computed.
// W^T stays resident in shared memory as the K-major B operand in the same k' order.
// ------------------------------------------------------------------------------------------------------------------
#define LB_SLOT_UNITS 32

struct LstmTcBwdCell {
  mfm_lstm_cell c;
  int mt, tiles, hp8, npad, nslot, tmem_cols;
};
struct LstmTcBwdBatch {
  LstmTcBwdCell c[MFM_MAX_CELLS];
  int n;
};

__global__ void __launch_bounds__(LT_THREADS) lstm_tc_bwd_kernel(LstmTcBwdBatch bt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long slotbar[2];
  __shared__ __align__(8) unsigned long long stepbar;
  __shared__ uint32_t tmem_holder;
  const LstmTcBwdCell& lc = bt.c[blockIdx.y];
  if ((int)blockIdx.x >= lc.tiles) return;
  const mfm_lstm_cell& c = lc.c;
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  const int MT = lc.mt, hp8 = lc.hp8, npad = lc.npad, nslot = lc.nslot;
  const int kslabs = hp8 >> 1;                        // K' = 4*hp8, 8 per slab
  const int lboB = npad * 16 + 32, lboA = MT * 16 + 32;
  const int slot_bytes = (LB_SLOT_UNITS / 2) * lboA;  // one plane of one slot: 16 slabs
  unsigned char* Bhi = smem;
  unsigned char* Blo = Bhi + kslabs * lboB;
  unsigned char* Aring = Blo + kslabs * lboB;         // [nslot][hi|lo][16 slabs]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * MT;

  if (tid == 0) {
    mbar_init(smem_u32(&slotbar[0]), 1);
    mbar_init(smem_u32(&slotbar[1]), 1);
    mbar_init(smem_u32(&stepbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)lc.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B operand: B[n, k'=4j+g] = W[g*h + j][n]  (zero for j >= h or n >= h); item = (n, slab): 8 gathered values
  for (int idx = tid; idx < npad * kslabs; idx += LT_THREADS) {
    const int n = idx % npad, slab = idx / npad;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = 2 * slab + (q >> 2), g = q & 3;
      v[q] = (j < h && n < h) ? __ldg(c.W + (long long)(g * h + j) * h + n) : 0.0f;
    }
    split_store(v, Bhi + slab * lboB + n * 16, Blo + slab * lboB + n * 16, true);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);

  const int quad = warp & 3, cg = warp >> 2;
  const bool owns = (MT == 128) || (lane < 16);
  const int r = (MT == 128) ? (quad * 32 + lane) : (quad * 16 + lane);
  const int row = row0 + r;
  const bool valid = owns && row < B;
  const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const bool hv4 = (h & 3) == 0;
  const bool vec_gt = aligned16(c.gates) && hv4, vec_dg = aligned16(c.dG) && hv4;
  const bool vec_cs = aligned16(c.cs) && ((c.ld_cs & 3) == 0);
  const bool vec_sc = aligned16(c.dc_scratch) && hv4;
  const bool vec_dha = c.dh_all && aligned16(c.dh_all) && ((c.ld_dh_all & 3) == 0);
  const bool vec_dhl = c.dh_last && aligned16(c.dh_last) && ((c.ld_dh_last & 3) == 0);
  const bool vec_dce = c.dc_ext && aligned16(c.dc_ext) && ((c.ld_dc_ext & 3) == 0);
  const int nslots_per_step = (hp8 + LB_SLOT_UNITS - 1) / LB_SLOT_UNITS;
  int slot_uses[2] = {0, 0};
  int step_commits = 0;

  for (int t = T - 1; t >= 0; --t) {
    const uint32_t dprev = tlane + (uint32_t)(((t + 1) & 1) * npad);
    const uint32_t dcur = tmem_base + (uint32_t)((t & 1) * npad);
    if (t < T - 1) {                       // dh_rec of this step = result of the previous step's MMAs
      mbar_wait(smem_u32(&stepbar), (uint32_t)((step_commits - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long tr = (long long)t * B + row;
    for (int u = 0; u < nslots_per_step; ++u) {
      const int sl = (nslot == 2) ? (u & 1) : 0;
      unsigned char* Ahi = Aring + sl * 2 * slot_bytes;
      unsigned char* Alo = Ahi + slot_bytes;
      if (slot_uses[sl] > 0) mbar_wait(smem_u32(&slotbar[sl]), (uint32_t)((slot_uses[sl] - 1) & 1));   // MMAs that read it are done
      const int j0 = u * LB_SLOT_UNITS + 8 * cg;
      if (j0 < hp8) {
        float dh[8];
        if (t < T - 1) {
          tmem_ld8(dprev + (uint32_t)j0, dh);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[i] = 0.0f;
        }
        float dg[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) dg[g][i] = 0.0f;
        if (valid) {
          const int nv = min(8, h - j0);
          float x[4][8], cp[8], cn[8], e[8], dc[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) ld8_global(c.gates + tr * H4 + g * h + j0, vec_gt, nv, x[g]);
          ld8_global(c.cs + tr * c.ld_cs + j0, vec_cs, nv, cp);
          ld8_global(c.cs + (tr + B) * c.ld_cs + j0, vec_cs, nv, cn);
          if (c.dh_all) {
            ld8_global(c.dh_all + tr * c.ld_dh_all + j0, vec_dha, nv, e);
#pragma unroll
            for (int i = 0; i < 8; ++i) dh[i] += e[i];
          }
          if (c.dh_last && t == T - 1) {
            ld8_global(c.dh_last + (long long)row * c.ld_dh_last + j0, vec_dhl, nv, e);
#pragma unroll
            for (int i = 0; i < 8; ++i) dh[i] += e[i];
          }
          if (t < T - 1) ld8_global(c.dc_scratch + (long long)row * h + j0, vec_sc, nv, dc);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dc[i] = 0.0f;
          }
          if (c.dc_ext) {
            ld8_global(c.dc_ext + tr * c.ld_dc_ext + j0, vec_dce, nv, e);
#pragma unroll
            for (int i = 0; i < 8; ++i) dc[i] += e[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ig = x[0][i], fg = x[1][i], gg = x[2][i], og = x[3][i];
            const float tc = gate_tanh(cn[i]);
            const float dci = dc[i] + dh[i] * og * (1.0f - tc * tc);
            const bool ok = i < nv;
            dg[0][i] = ok ? dci * gg * ig * (1.0f - ig) : 0.0f;
            dg[1][i] = ok ? dci * cp[i] * fg * (1.0f - fg) : 0.0f;
            dg[2][i] = ok ? dci * ig * (1.0f - gg * gg) : 0.0f;
            dg[3][i] = ok ? dh[i] * tc * og * (1.0f - og) : 0.0f;
            dc[i] = dci * fg;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) st8_global(c.dG + tr * H4 + g * h + j0, vec_dg, nv, dg[g]);
          if (t > 0) st8_global(c.dc_scratch + (long long)row * h + j0, vec_sc, nv, dc);
        }
        if (owns && t > 0) {               // A operand: slab q of this chunk = units (j0+2q, j0+2q+1) x gates (i,f,g,o)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v[8] = {dg[0][2 * q], dg[1][2 * q], dg[2][2 * q], dg[3][2 * q],
                                dg[0][2 * q + 1], dg[1][2 * q + 1], dg[2][2 * q + 1], dg[3][2 * q + 1]};
            const int off = (4 * cg + q) * lboA + r * 16;
            split_store(v, Ahi + off, Alo + off, true);
          }
        }
      }
      if (t > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int units = min(LB_SLOT_UNITS, hp8 - u * LB_SLOT_UNITS);
          const int ksteps = units >> 2;                                  // 4 units = 16 k'
          const uint32_t aH = smem_u32(Ahi), aL = smem_u32(Alo);
          const uint32_t bH = smem_u32(Bhi) + (uint32_t)(u * (LB_SLOT_UNITS / 2) * lboB), bL = bH + (uint32_t)(kslabs * lboB);
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint32_t ao = kk * 2 * lboA, bo = kk * 2 * lboB;
            const uint64_t dAh = make_smem_desc(aH + ao, lboA, 128), dAl = make_smem_desc(aL + ao, lboA, 128);
            const uint64_t dBh = make_smem_desc(bH + bo, lboB, 128), dBl = make_smem_desc(bL + bo, lboB, 128);
            umma_bf16(dcur, dAh, dBh, idesc, (u > 0 || kk > 0) ? 1u : 0u);
            umma_bf16(dcur, dAl, dBh, idesc, 1u);
            umma_bf16(dcur, dAh, dBl, idesc, 1u);
          }
          umma_commit(smem_u32(&slotbar[sl]));
          if (u == nslots_per_step - 1) umma_commit(smem_u32(&stepbar));
        }
        ++slot_uses[sl];
        if (u == nslots_per_step - 1) ++step_commits;
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

static inline int ru(int x, int m) { return (x + m - 1) / m * m; }

static size_t lstm_tc_bwd_smem(int h, int mt, int nslot) {
  const int hp8 = ru(h, 8), npad = ru(h, 16);
  return (size_t)2 * (hp8 / 2) * (npad * 16 + 32) + (size_t)nslot * 2 * (LB_SLOT_UNITS / 2) * (mt * 16 + 32) + 128;
}

static bool lstm_tc_bwd_plan(const mfm_lstm_cell& c, int smem_limit, LstmTcBwdCell& out) {
  if (c.h > 128 || c.h < 1 || !c.dc_scratch) return false;
  int mt = 0, nslot = 0;
  const int cand_mt[2] = {128, 64};
  for (int a = 0; a < 2 && !mt; ++a) {
    if (cand_mt[a] == 128 && c.B <= 64) continue;
    for (int ns = 2; ns >= 1 && !mt; --ns)
      if (lstm_tc_bwd_smem(c.h, cand_mt[a], ns) <= (size_t)smem_limit) { mt = cand_mt[a]; nslot = ns; }
  }
  if (!mt) return false;
  out.c = c;
  out.mt = mt;
  out.nslot = nslot;
  out.tiles = (c.B + mt - 1) / mt;
  out.hp8 = ru(c.h, 8);
  out.npad = ru(c.h, 16);
  int cols = 32;
  while (cols < 2 * out.npad) cols <<= 1;
  out.tmem_cols = cols;
  return true;
}

int lstm_tc_bwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  static int lim = -1;
  if (lim < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) lim = 48 * 1024;
    lim -= 1024;
  }
  LstmTcBwdBatch bt;
  bt.n = 0;
  *nrest = 0;
  size_t smem = 0;
  int gx = 0;
  for (int i = 0; i < ncells; ++i) {
    LstmTcBwdCell lc;
    if (lstm_tc_bwd_plan(cells[i], lim, lc)) {
      bt.c[bt.n++] = lc;
      const size_t s = lstm_tc_bwd_smem(lc.c.h, lc.mt, lc.nslot);
      if (s > smem) smem = s;
      if (lc.tiles > gx) gx = lc.tiles;
    } else {
      rest[(*nrest)++] = cells[i];
    }
  }
  if (bt.n == 0) return MFM_OK;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  lstm_tc_bwd_kernel<<<dim3(gx, bt.n), LT_THREADS, smem, st>>>(bt);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

// shared-memory bytes of the forward kernel for (h, mt)
static size_t lstm_tc_fwd_smem(int h, int mt) {
  const int hp16 = ru(h, 16), n4 = ru(4 * h, 16), slabs = hp16 / 8;
  return (size_t)2 * slabs * (n4 * 16 + 32) + (size_t)2 * slabs * (mt * 16 + 32) + 128;
}

// plan one cell; returns false if it must run on the CUDA-core kernel
static bool lstm_tc_plan(const mfm_lstm_cell& c, int smem_limit, LstmTcCell& out) {
  if (4 * c.h > 512 || c.h < 1) return false;
  int mt = 0;
  if (c.B > 64 && lstm_tc_fwd_smem(c.h, 128) <= (size_t)smem_limit) mt = 128;
  else if (lstm_tc_fwd_smem(c.h, 64) <= (size_t)smem_limit) mt = 64;
  if (!mt) return false;
  out.c = c;
  out.mt = mt;
  out.tiles = (c.B + mt - 1) / mt;
  out.hp8 = ru(c.h, 8);
  out.hp16 = ru(c.h, 16);
  out.n4 = ru(4 * c.h, 16);
  out.nblk = (out.n4 + 255) / 256;
  out.bn = ru((out.n4 + out.nblk - 1) / out.nblk, 16);
  int cols = 32;
  while (cols < out.n4) cols <<= 1;
  out.tmem_cols = cols;
  return true;
}

// Launches the tensor-core forward for every cell that fits; cells that do not are returned in `rest`.
int lstm_tc_fwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st) {
  static int lim = -1;
  if (lim < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) lim = 48 * 1024;
    lim -= 1024;      // the opt-in limit covers static + dynamic shared memory; the kernel has a few static bytes
  }
  LstmTcBatch bt;
  bt.n = 0;
  *nrest = 0;
  size_t smem = 0;
  int gx = 0;
  for (int i = 0; i < ncells; ++i) {
    LstmTcCell lc;
    if (lstm_tc_plan(cells[i], lim, lc)) {
      bt.c[bt.n++] = lc;
      const size_t s = lstm_tc_fwd_smem(lc.c.h, lc.mt);
      if (s > smem) smem = s;
      if (lc.tiles > gx) gx = lc.tiles;
    } else {
      rest[(*nrest)++] = cells[i];
    }
  }
  if (bt.n == 0) return MFM_OK;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  lstm_tc_fwd_kernel<<<dim3(gx, bt.n), LT_THREADS, smem, st>>>(bt);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
