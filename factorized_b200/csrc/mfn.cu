// MFN (memory fusion network) pieces that are not plain GEMMs:
//  * the delta-memory recurrence  mem' = gamma1*mem + gamma2*cHat  (mfm_model.py:177-180), fwd + bwd,
//    T steps inside one kernel; only the 64 memory columns of gamma*_fc1 are sequential, the 2H
//    "attended" columns were hoisted into GEMMs over all T*B rows (engine.py step 4);
//  * the softmax attention gate  attended = softmax(L) * cStar  (mfm_model.py:174-175), fwd + bwd.
#include "common.cuh"

#define MEM_THREADS 256
#define MEM_RT 8        // batch rows per CTA

// ------------------------------------------------------------------------------------------------
// forward.  smem: WA[mem][G] (= [W1m^T | W2m^T]), WB1[g1][mem] (= W12^T), WB2[g2][mem] (= W22^T),
// mem_s[RT][mem], u_s[RT][G].  If the weights do not fit, they are read from global (L2) instead.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MEM_THREADS) mfn_mem_fwd_kernel(mfm_mem_args a, int wsm) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, mem = a.mem, g1 = a.g1, g2 = a.g2, G = g1 + g2;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * MEM_RT;
  float* mem_s = smem;                    // [RT][mem]
  float* u_s = mem_s + MEM_RT * mem;      // [RT][G]
  float* WA = u_s + MEM_RT * G;           // [mem][G]
  float* WB1 = WA + mem * G;              // [g1][mem]
  float* WB2 = WB1 + g1 * mem;            // [g2][mem]
  if (wsm) {
    for (int idx = tid; idx < g1 * mem; idx += MEM_THREADS) {
      int u = idx / mem, k = idx - u * mem;
      WA[k * G + u] = __ldg(a.W1m + (long long)u * a.ld_w1m + k);
    }
    for (int idx = tid; idx < g2 * mem; idx += MEM_THREADS) {
      int u = idx / mem, k = idx - u * mem;
      WA[k * G + g1 + u] = __ldg(a.W2m + (long long)u * a.ld_w2m + k);
    }
    for (int idx = tid; idx < mem * g1; idx += MEM_THREADS) {
      int j = idx / g1, k = idx - j * g1;
      WB1[k * mem + j] = __ldg(a.W12 + idx);
    }
    for (int idx = tid; idx < mem * g2; idx += MEM_THREADS) {
      int j = idx / g2, k = idx - j * g2;
      WB2[k * mem + j] = __ldg(a.W22 + idx);
    }
  }
  for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
    mem_s[idx] = 0.0f;
    int r = idx / mem, j = idx - r * mem;
    if (row0 + r < B) a.mems[(long long)(row0 + r) * mem + j] = 0.0f;
  }
  __syncthreads();
  uint32_t ss1 = 0, ss2 = 0;
  const bool d1 = a.drop_p1 > 0.0f, d2 = a.drop_p2 > 0.0f;
  if (d1) ss1 = site_seed(a.rng, a.site1);
  if (d2) ss2 = site_seed(a.rng, a.site2);
  const float ks1 = d1 ? 1.0f / (1.0f - a.drop_p1) : 1.0f, ks2 = d2 ? 1.0f / (1.0f - a.drop_p2) : 1.0f;

  for (int t = 0; t < T; ++t) {
    // phase A: u = relu(Gpre[t] + mem W_m^T), one unit per thread, all RT rows
    for (int u = tid; u < G; u += MEM_THREADS) {
      const bool first = u < g1;
      const int uu = first ? u : u - g1;
      const int gw = first ? g1 : g2;
      const float* gpre = first ? a.G1pre : a.G2pre;
      float acc[MEM_RT];
#pragma unroll
      for (int r = 0; r < MEM_RT; ++r) {
        const int row = row0 + r;
        acc[r] = row < B ? __ldg(gpre + ((long long)t * B + row) * gw + uu) : 0.0f;
      }
      if (t > 0) {
        const float* wg = first ? a.W1m + (long long)uu * a.ld_w1m : a.W2m + (long long)uu * a.ld_w2m;
        for (int k = 0; k < mem; ++k) {
          const float w = wsm ? WA[k * G + u] : __ldg(wg + k);
#pragma unroll
          for (int r = 0; r < MEM_RT; ++r) acc[r] = fmaf(mem_s[r * mem + k], w, acc[r]);
        }
      }
      float* uo = first ? a.U1 : a.U2;
#pragma unroll
      for (int r = 0; r < MEM_RT; ++r) {
        const int row = row0 + r;
        float v = fmaxf(acc[r], 0.0f);
        const long long tr = (long long)t * B + row;
        if (first ? d1 : d2) {
          const uint32_t idx = (uint32_t)tr * (uint32_t)gw + (uint32_t)uu;
          v = drop_keep(first ? ss1 : ss2, idx, first ? a.drop_p1 : a.drop_p2) ? v * (first ? ks1 : ks2) : 0.0f;
        }
        u_s[r * G + u] = v;
        if (row < B) uo[tr * gw + uu] = v;
      }
    }
    __syncthreads();
    // phase B: gamma_k = sig(u_k W_k2^T + b), mem' = gamma1*mem + gamma2*cHat; thread owns (row, j)
    float newm[(MEM_RT * 512 + MEM_THREADS - 1) / MEM_THREADS];   // supports mem <= 512
    int cnt = 0;
    for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS, ++cnt) {
      const int r = idx / mem, j = idx - r * mem;
      const int row = row0 + r;
      float s1 = __ldg(a.b12 + j), s2 = __ldg(a.b22 + j);
      const float* us = u_s + r * G;
      if (wsm) {
        for (int k = 0; k < g1; ++k) s1 = fmaf(us[k], WB1[k * mem + j], s1);
        for (int k = 0; k < g2; ++k) s2 = fmaf(us[g1 + k], WB2[k * mem + j], s2);
      } else {
        for (int k = 0; k < g1; ++k) s1 = fmaf(us[k], __ldg(a.W12 + (long long)j * g1 + k), s1);
        for (int k = 0; k < g2; ++k) s2 = fmaf(us[g1 + k], __ldg(a.W22 + (long long)j * g2 + k), s2);
      }
      const float ga1 = gate_sigmoid(s1), ga2 = gate_sigmoid(s2);
      float nm = 0.0f;
      if (row < B) {
        const long long tr = (long long)t * B + row;
        const float ch = __ldg(a.cHat + tr * mem + j);
        nm = ga1 * mem_s[idx] + ga2 * ch;
        a.Gam1[tr * mem + j] = ga1;
        a.Gam2[tr * mem + j] = ga2;
        a.mems[(tr + B) * mem + j] = nm;
      }
      newm[cnt] = nm;
    }
    __syncthreads();     // everyone finished reading mem_s / u_s
    cnt = 0;
    for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS, ++cnt) mem_s[idx] = newm[cnt];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// backward (reverse time).  smem: natural layouts W12[mem][g1], W22[mem][g2], W1m[g1][mem], W2m[g2][mem];
// dmem_s[RT][mem], dp_s[RT][2*mem], du_s[RT][G].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MEM_THREADS) mfn_mem_bwd_kernel(mfm_mem_args a, int wsm) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, mem = a.mem, g1 = a.g1, g2 = a.g2, G = g1 + g2;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * MEM_RT;
  float* dmem_s = smem;                       // [RT][mem]
  float* dp_s = dmem_s + MEM_RT * mem;        // [RT][2*mem]  (dp1 | dp2)
  float* du_s = dp_s + MEM_RT * 2 * mem;      // [RT][G]
  float* W12s = du_s + MEM_RT * G;            // [mem][g1]
  float* W22s = W12s + mem * g1;              // [mem][g2]
  float* W1ms = W22s + mem * g2;              // [g1][mem]
  float* W2ms = W1ms + g1 * mem;              // [g2][mem]
  if (wsm) {
    for (int idx = tid; idx < mem * g1; idx += MEM_THREADS) W12s[idx] = __ldg(a.W12 + idx);
    for (int idx = tid; idx < mem * g2; idx += MEM_THREADS) W22s[idx] = __ldg(a.W22 + idx);
    for (int idx = tid; idx < g1 * mem; idx += MEM_THREADS) {
      int u = idx / mem, j = idx - u * mem;
      W1ms[idx] = __ldg(a.W1m + (long long)u * a.ld_w1m + j);
    }
    for (int idx = tid; idx < g2 * mem; idx += MEM_THREADS) {
      int u = idx / mem, j = idx - u * mem;
      W2ms[idx] = __ldg(a.W2m + (long long)u * a.ld_w2m + j);
    }
  }
  for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
    int r = idx / mem, j = idx - r * mem;
    dmem_s[idx] = (row0 + r < B) ? __ldg(a.dmem_last + (long long)(row0 + r) * a.ld_dmem_last + j) : 0.0f;
  }
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    // phase 1: per (row, j): gate gradients; dmem_s <- dmem*gamma1 (the direct path)
    for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
      const int r = idx / mem, j = idx - r * mem;
      const int row = row0 + r;
      float dp1 = 0.0f, dp2 = 0.0f;
      if (row < B) {
        const long long tr = (long long)t * B + row;
        const float dm = dmem_s[idx];
        const float mp = a.mems[tr * mem + j];
        const float ga1 = a.Gam1[tr * mem + j], ga2 = a.Gam2[tr * mem + j];
        const float ch = __ldg(a.cHat + tr * mem + j);
        dp1 = dm * mp * ga1 * (1.0f - ga1);
        dp2 = dm * ch * ga2 * (1.0f - ga2);
        a.dP1[tr * mem + j] = dp1;
        a.dP2[tr * mem + j] = dp2;
        a.dPc[tr * mem + j] = dm * ga2 * (1.0f - ch * ch);
        dmem_s[idx] = dm * ga1;
      }
      dp_s[r * 2 * mem + j] = dp1;
      dp_s[r * 2 * mem + mem + j] = dp2;
    }
    __syncthreads();
    // phase 2: du_k = (dp_k W_k2) * relu/dropout mask, one unit per thread, all rows
    for (int u = tid; u < G; u += MEM_THREADS) {
      const bool first = u < g1;
      const int uu = first ? u : u - g1;
      const int gw = first ? g1 : g2;
      const float* Wg = first ? a.W12 : a.W22;
      const float* Wsm = first ? W12s : W22s;
      const int po = first ? 0 : mem;
      float acc[MEM_RT];
#pragma unroll
      for (int r = 0; r < MEM_RT; ++r) acc[r] = 0.0f;
      for (int j = 0; j < mem; ++j) {
        const float w = wsm ? Wsm[j * gw + uu] : __ldg(Wg + (long long)j * gw + uu);
#pragma unroll
        for (int r = 0; r < MEM_RT; ++r) acc[r] = fmaf(dp_s[r * 2 * mem + po + j], w, acc[r]);
      }
      const float* U = first ? a.U1 : a.U2;
      float* dU = first ? a.dU1 : a.dU2;
      const float sc = first ? a.scale1 : a.scale2;
#pragma unroll
      for (int r = 0; r < MEM_RT; ++r) {
        const int row = row0 + r;
        float v = 0.0f;
        if (row < B) {
          const long long tr = (long long)t * B + row;
          v = (U[tr * gw + uu] > 0.0f) ? acc[r] * sc : 0.0f;
          dU[tr * gw + uu] = v;
        }
        du_s[r * G + u] = v;
      }
    }
    __syncthreads();
    // phase 3: dmem_{t-1} += du1 W1m + du2 W2m
    if (t > 0) {
      for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
        const int r = idx / mem, j = idx - r * mem;
        float s = dmem_s[idx];
        const float* du = du_s + r * G;
        if (wsm) {
          for (int u = 0; u < g1; ++u) s = fmaf(du[u], W1ms[u * mem + j], s);
          for (int u = 0; u < g2; ++u) s = fmaf(du[g1 + u], W2ms[u * mem + j], s);
        } else {
          for (int u = 0; u < g1; ++u) s = fmaf(du[u], __ldg(a.W1m + (long long)u * a.ld_w1m + j), s);
          for (int u = 0; u < g2; ++u) s = fmaf(du[g1 + u], __ldg(a.W2m + (long long)u * a.ld_w2m + j), s);
        }
        dmem_s[idx] = s;
      }
    }
    __syncthreads();
  }
}

static int mem_smem_limit() {
  static int lim = -1;
  if (lim < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) lim = 48 * 1024;
  }
  return lim;
}

static int mem_validate(const mfm_mem_args* a, bool bwd) {
  if (!a || a->T <= 0 || a->B <= 0 || a->mem <= 0 || a->mem > 512 || a->g1 <= 0 || a->g2 <= 0) return MFM_ERR_ARG;
  if (!a->cHat || !a->W1m || !a->W2m || !a->W12 || !a->W22 || !a->mems || !a->U1 || !a->U2 || !a->Gam1 || !a->Gam2)
    return MFM_ERR_ARG;
  if (!bwd && (!a->G1pre || !a->G2pre || !a->b12 || !a->b22)) return MFM_ERR_ARG;
  if (!bwd && (a->drop_p1 > 0.0f || a->drop_p2 > 0.0f) && !a->rng) return MFM_ERR_ARG;
  if (bwd && (!a->dmem_last || !a->dU1 || !a->dU2 || !a->dP1 || !a->dP2 || !a->dPc)) return MFM_ERR_ARG;
  return MFM_OK;
}

extern "C" int mfm_mfn_mem_fwd(const mfm_mem_args* a, void* stream) {
  int rc = mem_validate(a, false);
  if (rc) return rc;
  const int G = a->g1 + a->g2;
  size_t base = (size_t)(MEM_RT * a->mem + MEM_RT * G) * 4;
  size_t full = base + (size_t)(a->mem * G + a->g1 * a->mem + a->g2 * a->mem) * 4;
  const int lim = mem_smem_limit();
  const int wsm = full <= (size_t)lim;
  const size_t smem = wsm ? full : base;
  if (smem > (size_t)lim) return MFM_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mfn_mem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  mfn_mem_fwd_kernel<<<(a->B + MEM_RT - 1) / MEM_RT, MEM_THREADS, smem, (cudaStream_t)stream>>>(*a, wsm);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

extern "C" int mfm_mfn_mem_bwd(const mfm_mem_args* a, void* stream) {
  int rc = mem_validate(a, true);
  if (rc) return rc;
  const int G = a->g1 + a->g2;
  size_t base = (size_t)(MEM_RT * a->mem * 3 + MEM_RT * G) * 4;
  size_t full = base + (size_t)(2 * a->mem * G) * 4;
  const int lim = mem_smem_limit();
  const int wsm = full <= (size_t)lim;
  const size_t smem = wsm ? full : base;
  if (smem > (size_t)lim) return MFM_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mfn_mem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  mfn_mem_bwd_kernel<<<(a->B + MEM_RT - 1) / MEM_RT, MEM_THREADS, smem, (cudaStream_t)stream>>>(*a, wsm);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

// ------------------------------------------------------------------------------------------------
// softmax attention gate, one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_gate_fwd_kernel(int M, int N, float* __restrict__ L,
                                                                 const float* __restrict__ cstar,
                                                                 float* __restrict__ attended) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float* l = L + (long long)row * N;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, l[n]);
  mx = warp_max(mx);
  float s = 0.0f;
  for (int n = lane; n < N; n += 32) s += expf(l[n] - mx);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int n = lane; n < N; n += 32) {
    const float p = expf(l[n] - mx) * inv;
    l[n] = p;
    attended[(long long)row * N + n] = p * __ldg(cstar + (long long)row * N + n);
  }
}

__global__ void __launch_bounds__(256) softmax_gate_bwd_kernel(int M, int N, const float* __restrict__ dAtt,
                                                                 const float* __restrict__ att,
                                                                 const float* __restrict__ cstar,
                                                                 float* __restrict__ dL, float* __restrict__ dcs) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long o = (long long)row * N;
  float s = 0.0f;
  for (int n = lane; n < N; n += 32) s += dAtt[o + n] * cstar[o + n] * att[o + n];
  s = warp_sum(s);
  for (int n = lane; n < N; n += 32) {
    const float da = dAtt[o + n], p = att[o + n];
    dL[o + n] = p * (da * cstar[o + n] - s);
    dcs[o + n] = da * p;
  }
}

extern "C" int mfm_softmax_gate_fwd(int M, int N, float* L, const float* cstar, float* attended, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && L && cstar && attended);
  softmax_gate_fwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(M, N, L, cstar, attended);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

extern "C" int mfm_softmax_gate_bwd(int M, int N, const float* dAttended, const float* att, const float* cstar,
                                    float* dL, float* dcstar, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && dAttended && att && cstar && dL && dcstar);
  softmax_gate_bwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(M, N, dAttended, att, cstar, dL, dcstar);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
