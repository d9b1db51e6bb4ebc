// MFN (memory fusion network) pieces that are not plain GEMMs:
//  * the delta-memory recurrence  mem' = gamma1*mem + gamma2*cHat  (mfm_model.py:177-180), fwd + bwd,
//    T steps inside one kernel; only the 64 memory columns of gamma*_fc1 are sequential, the 2H
//    "attended" columns were hoisted into GEMMs over all T*B rows (engine.py step 4);
//  * the softmax attention gate  attended = softmax(L) * cStar  (mfm_model.py:174-175), fwd + bwd.
#include "common.cuh"

#define MEM_THREADS 512
#define MEM_RT 16       // batch rows per CTA (two register-blocked groups of 8)

static __host__ __device__ __forceinline__ int ru4(int x) { return (x + 3) & ~3; }

// acc[r] += sum_k act[r][k] * W(k) for 8 rows; act rows are 16 B aligned in shared memory (row pitch ldact, a multiple
// of 4), read as float4 broadcasts; W(k) = wcol[k*ldw] (k-major: shared memory, or a global matrix whose rows are k)
// or wrow[k] (global matrix whose row is this output unit).  32 FMAs per 4 weight loads + 8 broadcast loads.
template <bool COL>
__device__ __forceinline__ void dot8(float acc[8], const float* __restrict__ act, int ldact, int K,
                                     const float* __restrict__ wcol, int ldw, const float* __restrict__ wrow) {
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = COL ? wcol[(long long)(k + i) * ldw] : __ldg(wrow + k + i);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(act + r * ldact + k);
      acc[r] = fmaf(a.x, w[0], acc[r]);
      acc[r] = fmaf(a.y, w[1], acc[r]);
      acc[r] = fmaf(a.z, w[2], acc[r]);
      acc[r] = fmaf(a.w, w[3], acc[r]);
    }
  }
  for (; k < K; ++k) {
    const float w = COL ? wcol[(long long)k * ldw] : __ldg(wrow + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = fmaf(act[r * ldact + k], w, acc[r]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward.  smem: mem_s[RT][memP], u1_s[RT][g1P], u2_s[RT][g2P], gam_s[2][RT][mem], and (if they fit) the weights
// k-major: WA1[mem][g1] = W1m^T, WA2[mem][g2] = W2m^T, WB1[g1][mem] = W12^T, WB2[g2][mem] = W22^T.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MEM_THREADS, 1) mfn_mem_fwd_kernel(mfm_mem_args a, int wsm) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, mem = a.mem, g1 = a.g1, g2 = a.g2, G = g1 + g2;
  const int memP = ru4(mem), g1P = ru4(g1), g2P = ru4(g2);
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * MEM_RT;
  float* mem_s = smem;                       // [RT][memP]
  float* u1_s = mem_s + MEM_RT * memP;       // [RT][g1P]
  float* u2_s = u1_s + MEM_RT * g1P;         // [RT][g2P]
  float* gam_s = u2_s + MEM_RT * g2P;        // [2][RT][mem]
  float* WA1 = gam_s + 2 * MEM_RT * mem;     // [mem][g1]
  float* WA2 = WA1 + mem * g1;               // [mem][g2]
  float* WB1 = WA2 + mem * g2;               // [g1][mem]
  float* WB2 = WB1 + g1 * mem;               // [g2][mem]
  if (wsm) {
    for (int idx = tid; idx < g1 * mem; idx += MEM_THREADS) {
      int u = idx / mem, k = idx - u * mem;
      WA1[k * g1 + u] = __ldg(a.W1m + (long long)u * a.ld_w1m + k);
    }
    for (int idx = tid; idx < g2 * mem; idx += MEM_THREADS) {
      int u = idx / mem, k = idx - u * mem;
      WA2[k * g2 + u] = __ldg(a.W2m + (long long)u * a.ld_w2m + k);
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
  for (int idx = tid; idx < MEM_RT * (memP + g1P + g2P); idx += MEM_THREADS) mem_s[idx] = 0.0f;   // state + padding
  for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
    int r = idx / mem, j = idx - r * mem;
    if (row0 + r < B) a.mems[(long long)(row0 + r) * mem + j] = 0.0f;
  }
  __syncthreads();
  uint32_t ss1 = 0, ss2 = 0;
  const bool d1 = a.drop_p1 > 0.0f, d2 = a.drop_p2 > 0.0f;
  if (d1) ss1 = site_seed(a.rng, a.site1);
  if (d2) ss2 = site_seed(a.rng, a.site2);
  const float ks1 = d1 ? 1.0f / (1.0f - a.drop_p1) : 1.0f, ks2 = d2 ? 1.0f / (1.0f - a.drop_p2) : 1.0f;
  constexpr int RG = MEM_RT / 8;

  for (int t = 0; t < T; ++t) {
    // phase A: u_k = dropout(relu(Gkpre[t] + mem W_km^T)); item = (unit, row group of 8)
    for (int item = tid; item < G * RG; item += MEM_THREADS) {
      const int u = item % G, rg = item / G;
      const bool first = u < g1;
      const int uu = first ? u : u - g1, gw = first ? g1 : g2;
      const float* gpre = first ? a.G1pre : a.G2pre;
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = row0 + rg * 8 + r;
        acc[r] = row < B ? __ldg(gpre + ((long long)t * B + row) * gw + uu) : 0.0f;
      }
      if (t > 0) {
        const float* act = mem_s + rg * 8 * memP;
        if (wsm) dot8<true>(acc, act, memP, mem, (first ? WA1 : WA2) + uu, gw, nullptr);
        else     dot8<false>(acc, act, memP, mem, nullptr, 0,
                             first ? a.W1m + (long long)uu * a.ld_w1m : a.W2m + (long long)uu * a.ld_w2m);
      }
      float* uo = first ? a.U1 : a.U2;
      float* us = first ? u1_s + rg * 8 * g1P : u2_s + rg * 8 * g2P;
      const int up = first ? g1P : g2P;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = row0 + rg * 8 + r;
        float v = fmaxf(acc[r], 0.0f);
        const long long tr = (long long)t * B + row;
        if (first ? d1 : d2) {
          const uint32_t idx = (uint32_t)tr * (uint32_t)gw + (uint32_t)uu;
          v = drop_keep(first ? ss1 : ss2, idx, first ? a.drop_p1 : a.drop_p2) ? v * (first ? ks1 : ks2) : 0.0f;
        }
        us[r * up + uu] = v;
        if (row < B) uo[tr * gw + uu] = v;
      }
    }
    __syncthreads();
    // phase B: gamma_k = sig(u_k W_k2^T + b_k2); item = (memory unit j, which gamma, row group)
    for (int item = tid; item < mem * 2 * RG; item += MEM_THREADS) {
      const int j = item % mem, rest = item / mem;
      const int gsel = rest % 2, rg = rest / 2;
      const float bj = __ldg((gsel ? a.b22 : a.b12) + j);
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = bj;
      const int gw = gsel ? g2 : g1, up = gsel ? g2P : g1P;
      const float* act = (gsel ? u2_s : u1_s) + rg * 8 * up;
      if (wsm) dot8<true>(acc, act, up, gw, (gsel ? WB2 : WB1) + j, mem, nullptr);
      else     dot8<false>(acc, act, up, gw, nullptr, 0, (gsel ? a.W22 : a.W12) + (long long)j * gw);
      float* go = gsel ? a.Gam2 : a.Gam1;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = row0 + rg * 8 + r;
        const float ga = gate_sigmoid(acc[r]);
        gam_s[(gsel * MEM_RT + rg * 8 + r) * mem + j] = ga;
        if (row < B) go[((long long)t * B + row) * mem + j] = ga;
      }
    }
    __syncthreads();
    // phase C: mem' = gamma1 * mem + gamma2 * cHat[t]
    for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
      const int r = idx / mem, j = idx - r * mem;
      const int row = row0 + r;
      float nm = 0.0f;
      if (row < B) {
        const long long tr = (long long)t * B + row;
        nm = gam_s[r * mem + j] * mem_s[r * memP + j] + gam_s[(MEM_RT + r) * mem + j] * __ldg(a.cHat + tr * mem + j);
        a.mems[(tr + B) * mem + j] = nm;
      }
      mem_s[r * memP + j] = nm;       // only this thread touches (r, j) in this phase
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// backward (reverse time).  smem: dmem_s[RT][memP], dp1_s/dp2_s[RT][memP], du1_s[RT][g1P], du2_s[RT][g2P], and
// the weights in their natural (already k-major for these products) layouts W12[mem][g1], W22[mem][g2],
// W1m[g1][mem], W2m[g2][mem].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MEM_THREADS, 1) mfn_mem_bwd_kernel(mfm_mem_args a, int wsm) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, mem = a.mem, g1 = a.g1, g2 = a.g2, G = g1 + g2;
  const int memP = ru4(mem), g1P = ru4(g1), g2P = ru4(g2);
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * MEM_RT;
  float* dmem_s = smem;                        // [RT][memP]
  float* dp1_s = dmem_s + MEM_RT * memP;       // [RT][memP]
  float* dp2_s = dp1_s + MEM_RT * memP;        // [RT][memP]
  float* du1_s = dp2_s + MEM_RT * memP;        // [RT][g1P]
  float* du2_s = du1_s + MEM_RT * g1P;         // [RT][g2P]
  float* W12s = du2_s + MEM_RT * g2P;          // [mem][g1]
  float* W22s = W12s + mem * g1;               // [mem][g2]
  float* W1ms = W22s + mem * g2;               // [g1][mem]
  float* W2ms = W1ms + g1 * mem;               // [g2][mem]
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
  for (int idx = tid; idx < MEM_RT * (3 * memP + g1P + g2P); idx += MEM_THREADS) dmem_s[idx] = 0.0f;
  __syncthreads();
  for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
    int r = idx / mem, j = idx - r * mem;
    if (row0 + r < B) dmem_s[r * memP + j] = __ldg(a.dmem_last + (long long)(row0 + r) * a.ld_dmem_last + j);
  }
  __syncthreads();
  constexpr int RG = MEM_RT / 8;
  for (int t = T - 1; t >= 0; --t) {
    // phase 1: per (row, j): gate gradients; dmem_s <- dmem*gamma1 (the direct path)
    for (int idx = tid; idx < MEM_RT * mem; idx += MEM_THREADS) {
      const int r = idx / mem, j = idx - r * mem;
      const int row = row0 + r;
      float dp1 = 0.0f, dp2 = 0.0f;
      if (row < B) {
        const long long tr = (long long)t * B + row;
        const float dm = dmem_s[r * memP + j];
        const float mp = a.mems[tr * mem + j];
        const float ga1 = a.Gam1[tr * mem + j], ga2 = a.Gam2[tr * mem + j];
        const float ch = __ldg(a.cHat + tr * mem + j);
        dp1 = dm * mp * ga1 * (1.0f - ga1);
        dp2 = dm * ch * ga2 * (1.0f - ga2);
        a.dP1[tr * mem + j] = dp1;
        a.dP2[tr * mem + j] = dp2;
        a.dPc[tr * mem + j] = dm * ga2 * (1.0f - ch * ch);
        dmem_s[r * memP + j] = dm * ga1;
      }
      dp1_s[r * memP + j] = dp1;
      dp2_s[r * memP + j] = dp2;
    }
    __syncthreads();
    // phase 2: du_k = (dp_k W_k2) * relu/dropout mask; item = (unit, row group)
    for (int item = tid; item < G * RG; item += MEM_THREADS) {
      const int u = item % G, rg = item / G;
      const bool first = u < g1;
      const int uu = first ? u : u - g1, gw = first ? g1 : g2;
      float acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = 0.0f;
      dot8<true>(acc, (first ? dp1_s : dp2_s) + rg * 8 * memP, memP, mem,
                 (wsm ? (first ? W12s : W22s) : (first ? a.W12 : a.W22)) + uu, gw, nullptr);
      const float* U = first ? a.U1 : a.U2;
      float* dU = first ? a.dU1 : a.dU2;
      const float sc = first ? a.scale1 : a.scale2;
      float* ds = first ? du1_s + rg * 8 * g1P : du2_s + rg * 8 * g2P;
      const int up = first ? g1P : g2P;
      const long long ldu = first ? (a.ld_dU1 ? a.ld_dU1 : g1) : (a.ld_dU2 ? a.ld_dU2 : g2);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = row0 + rg * 8 + r;
        float v = 0.0f;
        if (row < B) {
          const long long tr = (long long)t * B + row;
          v = (U[tr * gw + uu] > 0.0f) ? acc[r] * sc : 0.0f;
          dU[tr * ldu + uu] = v;
        }
        ds[r * up + uu] = v;
      }
    }
    __syncthreads();
    // phase 3: dmem_{t-1} += du1 W1m + du2 W2m; item = (memory unit j, row group)
    if (t > 0) {
      for (int item = tid; item < mem * RG; item += MEM_THREADS) {
        const int j = item % mem, rg = item / mem;
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = dmem_s[(rg * 8 + r) * memP + j];
        if (wsm) {
          dot8<true>(acc, du1_s + rg * 8 * g1P, g1P, g1, W1ms + j, mem, nullptr);
          dot8<true>(acc, du2_s + rg * 8 * g2P, g2P, g2, W2ms + j, mem, nullptr);
        } else {
          dot8<true>(acc, du1_s + rg * 8 * g1P, g1P, g1, a.W1m + j, (int)a.ld_w1m, nullptr);
          dot8<true>(acc, du2_s + rg * 8 * g2P, g2P, g2, a.W2m + j, (int)a.ld_w2m, nullptr);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) dmem_s[(rg * 8 + r) * memP + j] = acc[r];
      }
    }
    __syncthreads();
  }
}

static int mem_smem_limit() { return mfm_dev_info().smem_optin; }

static int mem_validate(const mfm_mem_args* a, bool bwd) {
  if (!a || a->T <= 0 || a->B <= 0 || a->mem <= 0 || a->mem > 512 || a->g1 <= 0 || a->g2 <= 0) return MFM_ERR_ARG;
  if (!a->cHat || !a->W1m || !a->W2m || !a->W12 || !a->W22 || !a->mems || !a->U1 || !a->U2 || !a->Gam1 || !a->Gam2)
    return MFM_ERR_ARG;
  if (!bwd && (!a->G1pre || !a->G2pre || !a->b12 || !a->b22)) return MFM_ERR_ARG;
  if (!bwd && (a->drop_p1 > 0.0f || a->drop_p2 > 0.0f) && !a->rng) return MFM_ERR_ARG;
  if (bwd && (!a->dmem_last || !a->dU1 || !a->dU2 || !a->dP1 || !a->dP2 || !a->dPc)) return MFM_ERR_ARG;
  return MFM_OK;
}

int mem_ws_launch(const mfm_mem_args* a, bool bwd, cudaStream_t st);     // mem_ws.cu: the tensor-core form, where it applies

extern "C" int mfm_mfn_mem_fwd(const mfm_mem_args* a, void* stream) {
  int rc = mem_validate(a, false);
  if (rc) return rc;
  rc = mem_ws_launch(a, false, (cudaStream_t)stream);
  if (rc != MFM_ERR_UNSUPPORTED) return rc;
  const int G = a->g1 + a->g2;
  size_t base = (size_t)(MEM_RT * (ru4(a->mem) + ru4(a->g1) + ru4(a->g2)) + 2 * MEM_RT * a->mem) * 4;
  size_t full = base + (size_t)(2 * a->mem * G) * 4;
  const int lim = mem_smem_limit();
  const int wsm = full <= (size_t)lim;
  const size_t smem = wsm ? full : base;
  if (smem > (size_t)lim) return MFM_ERR_UNSUPPORTED;
  if (int e = mfm_func_smem_t(mfn_mem_fwd_kernel, lim)) return e;
  mfn_mem_fwd_kernel<<<(a->B + MEM_RT - 1) / MEM_RT, MEM_THREADS, smem, (cudaStream_t)stream>>>(*a, wsm);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

extern "C" int mfm_mfn_mem_bwd(const mfm_mem_args* a, void* stream) {
  int rc = mem_validate(a, true);
  if (rc) return rc;
  rc = mem_ws_launch(a, true, (cudaStream_t)stream);
  if (rc != MFM_ERR_UNSUPPORTED) return rc;
  const int G = a->g1 + a->g2;
  size_t base = (size_t)(MEM_RT * (3 * ru4(a->mem) + ru4(a->g1) + ru4(a->g2))) * 4;
  size_t full = base + (size_t)(2 * a->mem * G) * 4;
  const int lim = mem_smem_limit();
  const int wsm = full <= (size_t)lim;
  const size_t smem = wsm ? full : base;
  if (smem > (size_t)lim) return MFM_ERR_UNSUPPORTED;
  if (int e = mfm_func_smem_t(mfn_mem_bwd_kernel, lim)) return e;
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
