// Device and host helpers shared by the pipelined tcgen05 GEMM kernels (gemm_tcp.cu, gemm_ps.cu): chunk geometry, TMA /
// bulk-copy / mbarrier wrappers, the fp32 -> split-bf16 converter that feeds tensor memory, tensor-map construction.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "tc_epilogue.cuh"

#define P_BM 128
#define P_BK 16
#define P_SLABS 2
#define P_PAD 64
#define P_MAXRING 8                 // upper bound of either ring depth
#define P_NCONV 256                 // converter / epilogue threads (warps 0..7)
#define P_THREADS (P_NCONV + 64)    // + producer warp + MMA warp
#define P_WPROD (P_NCONV / 32)
#define P_WMMA (P_NCONV / 32 + 1)
#define P_BAR_BYTES 384             // mbarriers + TMEM address live at the tail of the dynamic buffer (no static smem)
#define P_MAXBN_RAW 128             // B staged raw and split in the kernel
#define P_MAXBN_PRE 160             // B pre-split: 160 accumulator + 6 x 16 A columns = 256 TMEM columns

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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds8(const unsigned char* p0, const unsigned char* p1, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p0);
  const float4 b = *reinterpret_cast<const float4*>(p1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- TS mode: the A operand lives in tensor memory -------------------------------------------------------------------
// Thread (warp w, lane l) owns row 32*(w%4)+l of the tile (= its TMEM lane) and the 8-k slab w/4 of the chunk: it reads
// its 8 fp32 from the raw stage (conflict-free through the TMA swizzle), splits them, and writes 4 packed hi registers and
// 4 packed lo registers with tcgen05.st.  No shared-memory store, no proxy fence, and the MMA reads only B from shared
// memory: per chunk this removes 8.4 KB of plane stores and 12 KB of operand reads from the shared-memory pipe, which
// is what bounds the SS form.  TMEM image of one plane: lane = row, column j = (k = 2j, 2j+1) packed.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
template <bool MN>
__device__ __forceinline__ void convert_a_tmem(const unsigned char* st, uint32_t taddr_set, int warp, int lane, bool want_lo) {
  const int slab = warp >> 2;
  float v[8];
  if (!MN) {
    const int r = 32 * (warp & 3) + lane;
    const unsigned char* row = st + r * (P_BK * 4);
    const int sw = (r >> 1) & 3;                            // 64 B swizzle
    lds8(row + (((2 * slab) ^ sw) << 4), row + (((2 * slab + 1) ^ sw) << 4), v);
  } else {
    const unsigned char* box = st + (warp & 3) * (P_BK * 128) + slab * (8 * 128) + (lane & 3) * 4;   // box = 32 columns
#pragma unroll
    for (int i = 0; i < 8; ++i)                             // k = 8*slab + i; 128 B swizzle: chunk ^= k % 8 = i
      v[i] = *reinterpret_cast<const float*>(box + i * 128 + (((lane >> 2) ^ i) << 4));
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    l[j] = pack_bf16(v[2 * j] - __uint_as_float(h[j] << 16), v[2 * j + 1] - __uint_as_float(h[j] & 0xFFFF0000u));
  }
  const uint32_t ta = taddr_set + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(slab * 4);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  if (want_lo)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta + 8), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

// ---- B pre-split: block (N tile, K chunk) writes that chunk's MMA-ready image [hi plane | lo plane] -------------------
template <bool MN>
__global__ void __launch_bounds__(256) gemm_prep_kernel(const float* __restrict__ B, long long ldb, int N, int K, int BN,
                                                        int nchunks_total, int want_lo, unsigned char* __restrict__ img,
                                                        int BN_last = 0) {
  // BN_last != 0: the last N tile is narrower (gemm_ps.cu); its images follow those of the full-width tiles
  const int bx = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  const int n0 = bx * BN, k0 = c * P_BK;
  const size_t tile_base = (size_t)bx * nchunks_total * (size_t)(2 * P_SLABS * (BN * 16 + P_PAD));
  if (BN_last && bx == (int)gridDim.x - 1) BN = BN_last;
  const int plB = P_SLABS * (BN * 16 + P_PAD);
  const int lbo = MN ? (BN / 8) * 128 : (BN * 16 + P_PAD);
  unsigned char* hi = img + tile_base + (size_t)c * (size_t)(2 * plB);
  unsigned char* lo = hi + plB;
  const bool vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 3) == 0);
  float v[8];
  if (!MN) {                                   // B[N, K]: item (row r, slab)
    for (int idx = tid; idx < BN * P_SLABS; idx += 256) {
      const int r = idx >> 1, slab = idx & 1;
      load8(B, ldb, n0 + r, N, k0 + slab * 8, K, vec, v);
      split_store(v, hi + slab * lbo + r * 16, lo + slab * lbo + r * 16, want_lo != 0);
    }
  } else {                                     // B[K, N]: item (k, column group mg)
    const int ng = BN >> 3;
    for (int idx = tid; idx < P_BK * ng; idx += 256) {
      const int klo = idx & 7, g = idx >> 3, mg = g % ng, khi = g / ng;
      load8(B, ldb, k0 + khi * 8 + klo, K, n0 + mg * 8, N, vec, v);
      const int off = khi * lbo + mg * 128 + klo * 16;
      split_store(v, hi + off, lo + off, want_lo != 0);
    }
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

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
// A K chunk touches only 64 B of each operand row; promoting the L2 fill to 256 B makes DRAM see four chunks' worth of a
// row at once (one activate instead of four) and the next three chunks hit in L2.
static CUtensorMapL2promotion l2promo() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MFM_TCP_L2"); v = e ? atoi(e) : 256; }
  return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
       : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}
// fp32 matrix view [outer, inner] with row pitch ld floats; box [box_outer][box_inner]; swizzle span = the box row
static bool make_map(CUtensorMap* tm, const float* base, long long ld, int inner, int outer, int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_inner * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
