// Epilogue shared by the tcgen05 GEMM kernels: TMEM accumulator -> registers -> per-warp shared-memory transpose ->
// coalesced global stores with the fused epilogue of include/mfm_b200.h::mfm_gemm.
//
// 4*nhalves warps take part.  Warp w owns TMEM lanes 32*(w%4).. (tile rows) and the column range w/4 of nhalves.  A 32x32 block is read
// with tcgen05.ld (lane = row), written to the warp's scratch as float4 (row stride 36 floats: 16 B aligned and
// conflict-free for both phases), and read back transposed:
//   * fast path (16 B-aligned C and mask; bias / activation / dropout / ReLU mask / accumulate): each lane moves a float4,
//     one store instruction covers four 128 B row segments;
//   * general path (split-K reduction, ones column, unaligned C or mask): lane = column, 128 B per store.
#pragma once
#include "gemm_args.cuh"
#include "tc_common.cuh"

#define TC_EPI_LD 36                              // scratch row stride in floats
#define TC_EPI_SCRATCH_BYTES (8 * 32 * TC_EPI_LD * 4)

struct TcArgs {
  GemmArgs g;
  int BN;          // tile N (multiple of 16, <= 256)
  int passes;      // 3 = hi/lo split, 1 = plain bf16
  int tmem_cols;   // power of two >= max(32, BN)
  int dbg;         // experiment switches (env MFM_TC_DEBUG): 1 skip MMA, 2 skip convert/store, 4 skip epilogue, 8 skip loads
};

// The activation is chosen by a COMPILE-TIME constant inside the store loops (tc_epilogue dispatches once per call): with a
// run-time `act` the compiler if-converted the switch and evaluated tanhf AND expf for every output element of every GEMM
// -- ~60 instructions per element, 6 000 cycles per 32x32 block, two thirds of a small-K GEMM's CTA lifetime (measured
// with the epilogue clock stamps of scripts/gemm_trace.py).
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == MFM_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == MFM_ACT_TANH) return tanhf(v);
  if (ACT == MFM_ACT_SIGMOID) return sigmoidf_acc(v);
  return v;
}
template <int ACT>
__device__ __forceinline__ float4 act4(float4 v, float b0, float b1, float b2, float b3) {
  v.x = act_t<ACT>(v.x + b0); v.y = act_t<ACT>(v.y + b1);
  v.z = act_t<ACT>(v.z + b2); v.w = act_t<ACT>(v.w + b3);
  return v;
}

// warp, lane: indices within the epilogue warps.  scratch_all: 32*TC_EPI_LD floats of idle shared memory per warp.
template <int ACT>
__device__ __forceinline__ void tc_epilogue_t(const TcArgs& ta, uint32_t tmem_base, float* scratch_all, int warp, int lane,
                                              int m0, int n0, bool have_acc, int ones_col, int nhalves, long long* trs) {
  const GemmArgs& a = ta.g;
  const int BN = ta.BN;
  uint32_t sseed = 0;
  const bool do_drop = a.drop_p > 0.0f;
  if (do_drop) sseed = site_seed(a.rng, a.drop_site);
  const float keep_scale = do_drop ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  float* scratch = scratch_all + warp * (32 * TC_EPI_LD);
  // epilogue parameters pinned in registers (the fully generic per-element form cost ~30 instructions per output)
  float* const e_C = a.C;
  const long long e_ldc = a.ldc;
  const int e_M = a.M, e_N = a.N;
  const bool e_atomic = a.atomic != 0, e_acc = a.accumulate != 0, e_simple = !a.mask && !do_drop;
  const bool mask_vec = !a.mask || (((reinterpret_cast<uintptr_t>(a.mask) & 15) == 0) && ((a.ldmask & 3) == 0));
  const GemmMse& ms = a.mse;
  const bool e_mse = ms.x != nullptr;
  const bool mse_vec = !e_mse || (((reinterpret_cast<uintptr_t>(ms.x) & 15) == 0) && ((ms.ldx & 3) == 0) &&
                                  (!ms.xhat || (((reinterpret_cast<uintptr_t>(ms.xhat) & 15) == 0) && ((ms.ldxhat & 3) == 0))));
  float sq = 0.0f;                                  // this lane's share of sum (x_hat - x)^2
  const bool e_vec = !e_atomic && ones_col < 0 && mask_vec && mse_vec && ((reinterpret_cast<uintptr_t>(e_C) & 15) == 0) &&
                     ((e_ldc & 3) == 0) && ((n0 & 3) == 0);
  const int quad = warp & 3, half = warp >> 2;
  const int cbeg = half * (BN / nhalves), cend = cbeg + (BN / nhalves);
  const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int mbase = m0 + quad * 32;
  const int nrows = max(0, min(32, e_M - mbase));
  int trk = 0;
#define EPI_STAMP() do { if (trs && warp == 0 && lane == 0 && trk < 6) trs[trk++] = clock64(); } while (0)
  EPI_STAMP();
  for (int c0 = cbeg; c0 < cend; c0 += 32) {
    float v[32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (c0 + 8 * q < cend) {                            // warp-uniform
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tlane + (uint32_t)(c0 + 8 * q))
                     : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * q + i] = __uint_as_float(r[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * q + i] = 0.0f;
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (c0 == cbeg) EPI_STAMP();
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(scratch + lane * TC_EPI_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
    if (c0 == cbeg) EPI_STAMP();
    if (!have_acc) { __syncwarp(); continue; }
    if (e_vec) {
      // lane -> (row within a group of 4, 4-column group); 8 passes cover the 32 rows
      const int cg = lane & 7, rsub = lane >> 3;
      const int cl = c0 + cg * 4;                   // column within the tile
      const int n = n0 + cl;
      if (cl < cend && n < e_N) {
        const bool full4 = n + 3 < e_N;
        float b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (n + i < e_N) b[i] = (a.bias ? __ldg(a.bias + n + i) : 0.0f) + (a.bias2 ? __ldg(a.bias2 + n + i) : 0.0f);
        // NOT unrolled: the body carries every fused variant (MSE head, dropout, ReLU mask, accumulate, ragged tail); eight
        // copies of it made the epilogue ~60 KB of straight-line code per activation, far beyond the instruction cache,
        // and the epilogue -- two thirds of a small-K GEMM's CTA lifetime -- was bound by instruction FETCH (ncu: `no_inst`
        // on every line of it).  The plain case (bias + activation into an aligned C) has its own short loop.
        if (e_simple && !e_mse && !e_acc && full4) {
          const float* sp4 = scratch + rsub * TC_EPI_LD + cg * 4;
          float* cp4 = e_C + (long long)(mbase + rsub) * e_ldc + n;
#pragma unroll 2
          for (int p = 0; p < 8; ++p) {
            if (p * 4 + rsub < nrows) {
              float4 t = *reinterpret_cast<const float4*>(sp4 + p * 4 * TC_EPI_LD);
              t = act4<ACT>(t, b[0], b[1], b[2], b[3]);
              *reinterpret_cast<float4*>(cp4 + (long long)(p * 4) * e_ldc) = t;
            }
          }
        } else
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
          const int rr = p * 4 + rsub;
          if (rr < nrows) {
            float4 t = *reinterpret_cast<const float4*>(scratch + rr * TC_EPI_LD + cg * 4);
            t = act4<ACT>(t, b[0], b[1], b[2], b[3]);
            if (e_mse) {                                      // fused reconstruction head: residual, its square, its gradient
              const float* xp = ms.x + (long long)(mbase + rr) * ms.ldx + n;
              float4 xv;
              if (full4) xv = __ldg(reinterpret_cast<const float4*>(xp));
              else { xv.x = __ldg(xp); xv.y = n + 1 < e_N ? __ldg(xp + 1) : t.y; xv.z = n + 2 < e_N ? __ldg(xp + 2) : t.z; xv.w = t.w; }
              if (ms.xhat) {
                float* hp = ms.xhat + (long long)(mbase + rr) * ms.ldxhat + n;
                if (full4) *reinterpret_cast<float4*>(hp) = t;
                else { hp[0] = t.x; if (n + 1 < e_N) hp[1] = t.y; if (n + 2 < e_N) hp[2] = t.z; }
              }
              t.x -= xv.x; t.y -= xv.y; t.z -= xv.z; t.w -= xv.w;        // columns beyond N: x := x_hat, residual 0
              sq = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, sq))));
              t.x *= ms.grad_scale; t.y *= ms.grad_scale; t.z *= ms.grad_scale; t.w *= ms.grad_scale;
            }
            if (!e_simple) {                                  // dropout and / or ReLU mask, four columns at a time
              const int m = mbase + rr;
              if (do_drop) {
                const uint32_t e0 = (uint32_t)m * (uint32_t)e_N + (uint32_t)n;
                t.x = drop_keep(sseed, e0, a.drop_p) ? t.x * keep_scale : 0.0f;
                t.y = drop_keep(sseed, e0 + 1, a.drop_p) ? t.y * keep_scale : 0.0f;
                t.z = drop_keep(sseed, e0 + 2, a.drop_p) ? t.z * keep_scale : 0.0f;
                t.w = drop_keep(sseed, e0 + 3, a.drop_p) ? t.w * keep_scale : 0.0f;
              }
              if (a.mask) {
                const float* mp = a.mask + (long long)m * a.ldmask + n;
                float4 mv;
                if (full4) mv = __ldg(reinterpret_cast<const float4*>(mp));
                else { mv.x = __ldg(mp); mv.y = n + 1 < e_N ? __ldg(mp + 1) : 0.0f; mv.z = n + 2 < e_N ? __ldg(mp + 2) : 0.0f; mv.w = 0.0f; }
                t.x = mv.x > 0.0f ? t.x * a.mask_scale : 0.0f;
                t.y = mv.y > 0.0f ? t.y * a.mask_scale : 0.0f;
                t.z = mv.z > 0.0f ? t.z * a.mask_scale : 0.0f;
                t.w = mv.w > 0.0f ? t.w * a.mask_scale : 0.0f;
              }
            }
            float* cp = e_C + (long long)(mbase + rr) * e_ldc + n;
            if (full4) {
              if (e_acc) {
                const float4 o = *reinterpret_cast<const float4*>(cp);
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
              }
              *reinterpret_cast<float4*>(cp) = t;
            } else {
              const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (n + i < e_N) cp[i] = e_acc ? tv[i] + cp[i] : tv[i];
            }
          }
        }
      }
    } else {
      const int n = n0 + c0 + lane;
      const float* sp = scratch + lane;
      if (c0 + lane < cend) {
        if (ones_col >= 0 && n == ones_col) {                 // the ones column: bias gradient
          for (int rr = 0; rr < nrows; ++rr) atomicAdd(a.colsum_out + mbase + rr, sp[rr * TC_EPI_LD]);
        } else if (n < e_N) {
          float* cp = e_C + (long long)mbase * e_ldc + n;
          if (e_atomic) {
            for (int rr = 0; rr < nrows; ++rr, cp += e_ldc) atomicAdd(cp, sp[rr * TC_EPI_LD]);
          } else if (e_simple) {                              // bias + activation (+ C) on an unaligned C
            const float bsum = (a.bias ? __ldg(a.bias + n) : 0.0f) + (a.bias2 ? __ldg(a.bias2 + n) : 0.0f);
#pragma unroll 1
            for (int rr = 0; rr < nrows; ++rr, cp += e_ldc) {
              float t = act_t<ACT>(sp[rr * TC_EPI_LD] + bsum);
              if (e_mse) {
                if (ms.xhat) ms.xhat[(long long)(mbase + rr) * ms.ldxhat + n] = t;
                t -= __ldg(ms.x + (long long)(mbase + rr) * ms.ldx + n);
                sq = fmaf(t, t, sq);
                t *= ms.grad_scale;
              }
              cp[0] = e_acc ? t + cp[0] : t;
            }
          } else {                                            // dropout and/or ReLU-mask epilogues
            const float bsum = (a.bias ? __ldg(a.bias + n) : 0.0f) + (a.bias2 ? __ldg(a.bias2 + n) : 0.0f);
            const float* mp = a.mask ? a.mask + (long long)mbase * a.ldmask + n : nullptr;
#pragma unroll 1
            for (int rr = 0; rr < nrows; ++rr, cp += e_ldc) {
              float t = act_t<ACT>(sp[rr * TC_EPI_LD] + bsum);
              const int m = mbase + rr;
              if (do_drop) t = drop_keep(sseed, (uint32_t)m * (uint32_t)e_N + (uint32_t)n, a.drop_p) ? t * keep_scale : 0.0f;
              if (mp) t = __ldg(mp + (long long)rr * a.ldmask) > 0.0f ? t * a.mask_scale : 0.0f;
              cp[0] = e_acc ? t + cp[0] : t;
            }
          }
        }
      }
    }
    __syncwarp();
    EPI_STAMP();
  }
#undef EPI_STAMP
  if (e_mse) {                                      // one atomic per warp
    sq = warp_sum(sq);
    if (lane == 0) atomicAdd(ms.slot, sq * ms.loss_scale);
  }
}

__device__ __forceinline__ void tc_epilogue(const TcArgs& ta, uint32_t tmem_base, float* scratch_all, int warp, int lane,
                                            int m0, int n0, bool have_acc, int ones_col, int nhalves = 2,
                                            long long* trs = nullptr) {
  switch (ta.g.act) {                              // CTA-uniform
    case MFM_ACT_RELU: tc_epilogue_t<MFM_ACT_RELU>(ta, tmem_base, scratch_all, warp, lane, m0, n0, have_acc, ones_col, nhalves, trs); break;
    case MFM_ACT_TANH: tc_epilogue_t<MFM_ACT_TANH>(ta, tmem_base, scratch_all, warp, lane, m0, n0, have_acc, ones_col, nhalves, trs); break;
    case MFM_ACT_SIGMOID: tc_epilogue_t<MFM_ACT_SIGMOID>(ta, tmem_base, scratch_all, warp, lane, m0, n0, have_acc, ones_col, nhalves, trs); break;
    default: tc_epilogue_t<MFM_ACT_NONE>(ta, tmem_base, scratch_all, warp, lane, m0, n0, have_acc, ones_col, nhalves, trs); break;
  }
}
