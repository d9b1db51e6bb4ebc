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
// with tcgen05.commit -> mbarrier; two shared-memory stages overlap the loads of chunk c+1 with the MMAs of chunk c;
// the epilogue reads the accumulator with tcgen05.ld (32x32b.x16) and applies the fused epilogue of mfm_gemm.
#include <cuda_bf16.h>
#include "gemm_args.cuh"

#define TC_BM 128
#define TC_BK 32
#define TC_THREADS 128
#define TC_STAGES 2
#define TC_A_PLANE (4 * (TC_BM * 16 + 32))        // bytes of one A plane (K-major with 32 B slab padding; MN-major needs less)

struct TcArgs {
  GemmArgs g;
  int BN;          // tile N (multiple of 16, <= 256)
  int passes;      // 3 = hi/lo split, 1 = plain bf16
  int tmem_cols;   // power of two >= max(32, BN)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded wait: a barrier that never completes traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_NONE
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 8 consecutive fp32 along the contiguous dimension at (row, col..col+7); zero outside [0,nrows) x [0,ncols)
__device__ __forceinline__ void load8(const float* __restrict__ src, long long ld, int row, int nrows, int col, int ncols,
                                      bool vec_ok, float v[8]) {
  if (row < nrows && col + 8 <= ncols) {
    const float* p = src + (long long)row * ld + col;
    if (vec_ok) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(p + i);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (row < nrows) {
      const float* p = src + (long long)row * ld + col;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < ncols) v[i] = __ldg(p + i);
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);      // .x = a (low address), .y = b
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 x8 -> 16 B of bf16 "hi" and (optionally) 16 B of bf16 "lo" = bf16(x - hi)
__device__ __forceinline__ void split_store(const float v[8], unsigned char* hi_dst, unsigned char* lo_dst, bool want_lo) {
  uint4 h;
  h.x = pack_bf16(v[0], v[1]); h.y = pack_bf16(v[2], v[3]); h.z = pack_bf16(v[4], v[5]); h.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(hi_dst) = h;
  if (want_lo) {
    float r[8];
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[2 * i] = v[2 * i] - __uint_as_float(hw[i] << 16);
      r[2 * i + 1] = v[2 * i + 1] - __uint_as_float(hw[i] & 0xFFFF0000u);
    }
    uint4 l;
    l.x = pack_bf16(r[0], r[1]); l.y = pack_bf16(r[2], r[3]); l.z = pack_bf16(r[4], r[5]); l.w = pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(lo_dst) = l;
  }
}

// source [rows, K] with K contiguous -> K-major planes.  item = (row r, 8-wide k slab)
__device__ __forceinline__ void load_kmajor(const float* __restrict__ src, long long ld, int row0, int nrows, int k0, int kend,
                                            int tile_rows, unsigned char* hi, unsigned char* lo, int lbo, bool vec_ok,
                                            bool want_lo) {
  const int items = tile_rows * 4;
  for (int base = 0; base < items; base += TC_THREADS * 4) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * TC_THREADS + threadIdx.x;
      const int slab = idx & 3, r = idx >> 2;
      if (idx < items) load8(src, ld, row0 + r, nrows, k0 + slab * 8, kend, vec_ok, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * TC_THREADS + threadIdx.x;
      const int slab = idx & 3, r = idx >> 2;
      if (idx < items) split_store(v[u], hi + slab * lbo + r * 16, lo + slab * lbo + r * 16, want_lo);
    }
  }
}

// source [K, cols] with cols contiguous -> MN-major planes.  item = (k row, 8-wide column group)
__device__ __forceinline__ void load_mnmajor(const float* __restrict__ src, long long ld, int k0, int kend, int col0, int ncols,
                                             int tile_cols, unsigned char* hi, unsigned char* lo, int lbo, bool vec_ok,
                                             bool want_lo) {
  const int groups = tile_cols >> 3;
  const int items = ((groups + 3) & ~3) * TC_BK;
  for (int base = 0; base < items; base += TC_THREADS * 4) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * TC_THREADS + threadIdx.x;
      const int k = (idx & 7) + 8 * ((idx >> 5) & 3), mg = ((idx >> 3) & 3) + 4 * (idx >> 7);
      if (idx < items && mg < groups) load8(src, ld, k0 + k, kend, col0 + mg * 8, ncols, vec_ok, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * TC_THREADS + threadIdx.x;
      const int k = (idx & 7) + 8 * ((idx >> 5) & 3), mg = ((idx >> 3) & 3) + 4 * (idx >> 7);
      if (idx < items && mg < groups) {
        const int off = (k >> 3) * lbo + mg * 128 + (k & 7) * 16;
        split_store(v[u], hi + off, lo + off, want_lo);
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
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * a.kchunk;
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  const int nchunks = (kend - kbeg + TC_BK - 1) / TC_BK;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    unsigned char* st = smem + s * stage_bytes;
    unsigned char* Ahi = st;
    unsigned char* Alo = st + TC_A_PLANE;
    unsigned char* Bhi = st + 2 * TC_A_PLANE;
    unsigned char* Blo = Bhi + b_plane;
    if (c >= TC_STAGES) mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / TC_STAGES - 1) & 1));   // MMAs that read this stage are done
    const int k0 = kbeg + c * TC_BK;
    if (A_MN) load_mnmajor(a.A, a.lda, k0, kend, m0, a.M, TC_BM, Ahi, Alo, lboA, vecA, want_lo);
    else      load_kmajor(a.A, a.lda, m0, a.M, k0, kend, TC_BM, Ahi, Alo, lboA, vecA, want_lo);
    if (B_MN) load_mnmajor(a.B, a.ldb, k0, kend, n0, a.N, BN, Bhi, Blo, lboB, vecB, want_lo);
    else      load_kmajor(a.B, a.ldb, n0, a.N, k0, kend, BN, Bhi, Blo, lboB, vecB, want_lo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
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
  if (nchunks > 0) {
    const int c = nchunks - 1;
    mbar_wait(smem_u32(&bars[c & 1]), (uint32_t)((c / TC_STAGES) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 == tile rows
  uint32_t sseed = 0;
  const bool do_drop = a.drop_p > 0.0f;
  if (do_drop) sseed = site_seed(a.rng, a.drop_site);
  const float keep_scale = do_drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  const int m = m0 + warp * 32 + lane;
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (m < a.M && nchunks > 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        if (n < a.N) gemm_epilogue_store(a, m, n, __uint_as_float(r[j]), do_drop, sseed, keep_scale);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ta.tmem_cols)
                 : "memory");
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

int gemm_tc_launch(int passes, int mode, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                   float* C, long long ldc, const float* bias, const float* bias2, int act, int accumulate,
                   const float* mask, long long ldmask, float mask_scale, float drop_p, int drop_site,
                   const long long* rng, cudaStream_t st) {
  TcArgs ta;
  ta.g = GemmArgs{M, N, K, A, lda, B, ldb, C, ldc, bias, bias2, act, accumulate, mask, ldmask, mask_scale,
                  drop_p, drop_site, rng, K, 0};
  ta.passes = passes;
  // tile N: the whole (padded) N when it fits 256 columns, else near-equal tiles
  const int n16 = round_up(N, 16);
  const int ntiles = (n16 + 255) / 256;
  ta.BN = round_up((n16 + ntiles - 1) / ntiles, 16);
  int cols = 32;
  while (cols < ta.BN) cols <<= 1;
  ta.tmem_cols = cols;
  dim3 grid((N + ta.BN - 1) / ta.BN, (M + TC_BM - 1) / TC_BM, 1);
  const bool plain = !bias && !bias2 && act == MFM_ACT_NONE && !mask && drop_p <= 0.0f && accumulate;
  if (plain && K >= 2048) {
    long long tiles = (long long)grid.x * grid.y;
    int splits = (int)((2 * 148 + tiles - 1) / tiles);
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
  static bool attr[3] = {false, false, false};
  if (!attr[mode]) {
    cudaError_t e = cudaSuccess;
    if (mode == MFM_GEMM_NT) e = cudaFuncSetAttribute(gemm_tc_kernel<MFM_GEMM_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (mode == MFM_GEMM_NN) e = cudaFuncSetAttribute(gemm_tc_kernel<MFM_GEMM_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (mode == MFM_GEMM_TN) e = cudaFuncSetAttribute(gemm_tc_kernel<MFM_GEMM_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr[mode] = true;
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
