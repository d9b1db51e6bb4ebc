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

#define LT_THREADS 128

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
  const bool owns = (MT == 128) || (lane < 16);
  const int r = (MT == 128) ? (warp * 32 + lane) : (warp * 16 + lane);
  const int row = row0 + r;
  const bool valid = owns && row < B;
  const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
  const bool vec_gx = aligned16(c.gx) && ((h & 3) == 0);
  const bool vec_gt = aligned16(c.gates) && ((h & 3) == 0);
  const bool vec_hs = aligned16(c.hs) && ((c.ld_hs & 3) == 0);
  const bool vec_cs = aligned16(c.cs) && ((c.ld_cs & 3) == 0);
  const bool vec_b = c.bias_rest && aligned16(c.bias_rest) && ((h & 3) == 0);

  if (valid) {   // block 0 of the histories is the zero initial state
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
    for (int j0 = 0; j0 < hp8; j0 += 8) {
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
        ld8_global(c.cs + tr * c.ld_cs + j0, vec_cs, nv, cp);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ig = sigmoidf_acc(a[0][i] + x[0][i]);
          const float fg = sigmoidf_acc(a[1][i] + x[1][i]);
          const float gg = tanhf(a[2][i] + x[2][i]);
          const float og = sigmoidf_acc(a[3][i] + x[3][i]);
          cn[i] = fg * cp[i] + ig * gg;
          hv[i] = i < nv ? og * tanhf(cn[i]) : 0.0f;
          x[0][i] = ig; x[1][i] = fg; x[2][i] = gg; x[3][i] = og;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) st8_global(c.gates + tr * H4 + g * h + j0, vec_gt, nv, x[g]);
        st8_global(c.cs + (tr + B) * c.ld_cs + j0, vec_cs, nv, cn);
        st8_global(c.hs + (tr + B) * c.ld_hs + j0, vec_hs, nv, hv);
      }
      if (owns) split_store(hv, Hhi + (j0 >> 3) * lboH + r * 16, Hlo + (j0 >> 3) * lboH + r * 16, true);
    }
    if (hp16 > hp8 && owns) {
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

static inline int ru(int x, int m) { return (x + m - 1) / m * m; }

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
