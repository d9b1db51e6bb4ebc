// Small data-movement, reduction, loss-head and optimizer kernels of the MFM step.
#include "common.cuh"

// one warp per row (blockDim = 32 x 8 rows); a lane moves 4 consecutive columns per iteration (16 B, vector access on
// whichever side is 16 B aligned: the x split reads rows of pitch 1300 B and writes aligned ones), no index division
__global__ void __launch_bounds__(256) copy2d_kernel(int M, int N, const float* __restrict__ src, long long lds,
                                                      float* __restrict__ dst, long long ldd, int accumulate) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool vs = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((lds & 3) == 0);
  const bool vd = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((ldd & 3) == 0);
  const int N4 = N & ~3;
  for (long long m = (long long)blockIdx.x * 8 + ty; m < M; m += (long long)gridDim.x * 8) {
    const float* sp = src + m * lds;
    float* dp = dst + m * ldd;
#pragma unroll 2
    for (int n = 4 * tx; n < N4; n += 128) {
      float4 v;
      if (vs) v = __ldg(reinterpret_cast<const float4*>(sp + n));
      else { v.x = __ldg(sp + n); v.y = __ldg(sp + n + 1); v.z = __ldg(sp + n + 2); v.w = __ldg(sp + n + 3); }
      if (vd) {
        float4* d4 = reinterpret_cast<float4*>(dp + n);
        if (accumulate) { const float4 o = *d4; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *d4 = v;
      } else if (accumulate) {
        dp[n] += v.x; dp[n + 1] += v.y; dp[n + 2] += v.z; dp[n + 3] += v.w;
      } else {
        dp[n] = v.x; dp[n + 1] = v.y; dp[n + 2] = v.z; dp[n + 3] = v.w;
      }
    }
    for (int n = N4 + tx; n < N; n += 32) dp[n] = accumulate ? dp[n] + __ldg(sp + n) : __ldg(sp + n);
  }
}

__global__ void add_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] + b[i];
}

__global__ void zero_kernel(long long n, float* __restrict__ p) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0.0f;
}

// out[n] += sum_m A[m,n]: blockIdx.x = 32-column strip, blockIdx.y = row chunk; 32x8 threads
__global__ void __launch_bounds__(256) colsum_kernel(int M, int N, const float* __restrict__ A, long long lda,
                                                      float* __restrict__ out, int rows_per_block) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int mbeg = blockIdx.y * rows_per_block, mend = min(M, mbeg + rows_per_block);
  float s = 0.0f;
  if (n < N)
    for (int m = mbeg + ty; m < mend; m += 8) s += __ldg(A + (long long)m * lda + n);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int q = 1; q < 8; ++q) s += part[q][tx];
    atomicAdd(out + n, s);
  }
}

__global__ void relu_bwd_kernel(int M, int N, const float* __restrict__ dy, long long lddy, const float* __restrict__ y,
                                long long ldy, float* __restrict__ out, long long ldo) {
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    out[(long long)m * ldo + n] = y[(long long)m * ldy + n] > 0.0f ? dy[(long long)m * lddy + n] : 0.0f;
  }
}

// slot += loss_scale * sum (xhat-x)^2 ; dxhat = grad_scale * (xhat - x)      (nn.MSELoss, mfm_mosi.py:437)
__global__ void __launch_bounds__(256) mse_kernel(int M, int N, const float* __restrict__ xh, long long ldxh,
                                                   const float* __restrict__ x, long long ldx, float loss_scale,
                                                   float grad_scale, float* __restrict__ slot,
                                                   float* __restrict__ dxh, long long lddx) {
  __shared__ float red[32];
  float s = 0.0f;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // one warp per row: coalesced, no index division
  for (long long m = (long long)blockIdx.x * 8 + ty; m < M; m += (long long)gridDim.x * 8) {
    const float* xr = xh + m * ldxh;
    const float* tr = x + m * ldx;
    float* dr = dxh ? dxh + m * lddx : nullptr;
#pragma unroll 4
    for (int n = tx; n < N; n += 32) {
      const float r = xr[n] - __ldg(tr + n);
      s = fmaf(r, r, s);
      if (dr) dr[n] = grad_scale * r;
    }
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, tot * loss_scale);
}

// nn.L1Loss (mfm_mosi.py:438): slot += scale * sum |yhat - y| ; dy = scale * sign(yhat - y)
__global__ void __launch_bounds__(256) l1_kernel(long long n, const float* __restrict__ yh, const float* __restrict__ y,
                                                  float scale, float* __restrict__ slot, float* __restrict__ dy) {
  __shared__ float red[32];
  float s = 0.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = yh[i] - y[i];
    s += fabsf(r);
    dy[i] = r > 0.0f ? scale : (r < 0.0f ? -scale : 0.0f);
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, tot * scale);
}

// loss_KLD (mfm_model.py:36-38): slot += -0.5 * sum(1 + logvar - mu^2 - exp(logvar))  (a sum over all elements)
__global__ void __launch_bounds__(256) kld_fwd_kernel(int M, int N, const float* __restrict__ mu, long long ldm,
                                                       const float* __restrict__ lv, long long ldl, float* __restrict__ slot) {
  __shared__ float red[32];
  float s = 0.0f;
  const long long n = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / N), c = (int)(i - (long long)r * N);
    const float m = mu[(long long)r * ldm + c], l = lv[(long long)r * ldl + c];
    s += 1.0f + l - m * m - expf(l);
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, -0.5f * tot);
}
// its gradient: dmu += s * mu;  dlogvar = s * 0.5 * (exp(logvar) - 1);  s = scale * (scale_dev ? *scale_dev : 1)
__global__ void __launch_bounds__(256) kld_bwd_kernel(int M, int N, const float* __restrict__ mu, long long ldm,
                                                       const float* __restrict__ lv, long long ldl, float scale,
                                                       const float* __restrict__ scale_dev, float* __restrict__ dmu, long long lddm,
                                                       float* __restrict__ dlv, long long lddl) {
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.0f);
  const long long n = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / N), c = (int)(i - (long long)r * N);
    dmu[(long long)r * lddm + c] += sc * mu[(long long)r * ldm + c];
    dlv[(long long)r * lddl + c] = sc * 0.5f * (expf(lv[(long long)r * ldl + c]) - 1.0f);
  }
}

// nn.CrossEntropyLoss (mfm_mosi_acc.py:450): one thread per sample, C is tiny (2..4 classes)
__global__ void __launch_bounds__(256) ce_kernel(int B, int C, const float* __restrict__ yh, const long long* __restrict__ y,
                                                  float scale, float* __restrict__ slot, float* __restrict__ dy) {
  __shared__ float red[32];
  float s = 0.0f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const float* l = yh + (long long)b * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, l[c]);
    float se = 0.0f;
    for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
    const float lse = mx + logf(se);
    const int tgt = (int)y[b];
    s += lse - l[tgt];
    for (int c = 0; c < C; ++c) dy[(long long)b * C + c] = scale * (expf(l[c] - lse) - (c == tgt ? 1.0f : 0.0f));
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, tot * scale);
}

__global__ void loss_total_kernel(float* lb, float l0, float l1, float l2, float lmmd) {
  lb[8] = lb[0] + l0 * lb[1] + l1 * lb[2] + l2 * lb[3] + lmmd * (lb[4] + lb[5] + lb[6] + lb[7]);
}

// torch.optim.Adam, defaults (mfm_mosi.py:403).  tick: step += 1 and the two bias-correction scalars.
__global__ void adam_tick_kernel(float* state, double beta1, double beta2) {
  const float t = state[1] + 1.0f;
  state[1] = t;
  state[2] = (float)((double)state[0] / (1.0 - pow(beta1, (double)t)));   // step size (torch does this on the host in double)
  state[3] = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)t)));            // 1/sqrt(bias_correction2)
}
__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ state, float gs, float b1, float omb1,
                                                    float b2, float omb2, float eps) {
  const float step_size = state[2], inv_bc2 = state[3];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2 + eps));
  }
}

__global__ void rng_tick_kernel(long long* rng) { rng[1] += 1; }

// Box-Muller over two hashed uniforms per element
__global__ void randn_kernel(long long n, float* __restrict__ out, const long long* __restrict__ rng, int site) {
  const uint32_t ss = site_seed(rng, site);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t h1 = rng_bits(ss, (uint32_t)(2 * i));
    const uint32_t h2 = rng_bits(ss, (uint32_t)(2 * i + 1));
    const float u1 = ((float)(h1 >> 8) + 1.0f) * (1.0f / 16777216.0f);     // (0,1]
    const float u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);              // [0,1)
    out[i] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
}

static inline int grid_for(long long n, int threads = 256, int cap = 0) {
  if (cap <= 0) cap = mfm_dev_info().sms * 8;
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

extern "C" int mfm_copy2d(int M, int N, const float* src, long long lds, float* dst, long long ldd, int accumulate,
                          void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && src && dst);
  const int blocks = (M + 7) / 8 < mfm_dev_info().sms * 16 ? (M + 7) / 8 : mfm_dev_info().sms * 16;
  copy2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(M, N, src, lds, dst, ldd, accumulate);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_add(long long n, const float* a, const float* b, float* out, void* stream) {
  MFM_REQUIRE(n > 0 && a && b && out);
  add_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, a, b, out);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_zero(long long n, float* p, void* stream) {
  MFM_REQUIRE(n > 0 && p);
  zero_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, p);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_colsum(int M, int N, const float* A, long long lda, float* out, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && A && out);
  const int strips = (N + 31) / 32;
  int chunks = (2 * mfm_dev_info().sms + strips - 1) / strips;
  int maxc = (M + 63) / 64;
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  const int rpb = (M + chunks - 1) / chunks;
  colsum_kernel<<<dim3(strips, (M + rpb - 1) / rpb), 256, 0, (cudaStream_t)stream>>>(M, N, A, lda, out, rpb);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_relu_bwd(int M, int N, const float* dy, long long lddy, const float* y, long long ldy, float* out,
                            long long ldo, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && dy && y && out);
  relu_bwd_kernel<<<grid_for((long long)M * N), 256, 0, (cudaStream_t)stream>>>(M, N, dy, lddy, y, ldy, out, ldo);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_mse_fwd_bwd(int M, int N, const float* xhat, long long ldxh, const float* x, long long ldx,
                               float loss_scale, float grad_scale, float* slot, float* dxhat, long long lddx,
                               void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && xhat && x && slot);
  mse_kernel<<<((M + 7) / 8 < mfm_dev_info().sms * 8 ? (M + 7) / 8 : mfm_dev_info().sms * 8), 256, 0, (cudaStream_t)stream>>>(M, N, xhat, ldxh, x, ldx, loss_scale,
                                                                                       grad_scale, slot, dxhat, lddx);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_kld_fwd(int M, int N, const float* mu, long long ldmu, const float* logvar, long long ldlv, float* slot,
                           void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && mu && logvar && slot);
  kld_fwd_kernel<<<grid_for((long long)M * N, 256, 64), 256, 0, (cudaStream_t)stream>>>(M, N, mu, ldmu, logvar, ldlv, slot);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_kld_bwd(int M, int N, const float* mu, long long ldmu, const float* logvar, long long ldlv, float scale,
                           const float* scale_dev, float* dmu, long long lddmu, float* dlogvar, long long lddlv, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && mu && logvar && dmu && dlogvar);
  kld_bwd_kernel<<<grid_for((long long)M * N, 256, 64), 256, 0, (cudaStream_t)stream>>>(M, N, mu, ldmu, logvar, ldlv, scale, scale_dev,
                                                                                      dmu, lddmu, dlogvar, lddlv);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_l1_fwd_bwd(long long n, const float* yhat, const float* y, float scale, float* slot, float* dy,
                              void* stream) {
  MFM_REQUIRE(n > 0 && yhat && y && slot && dy);
  l1_kernel<<<grid_for(n, 256, 64), 256, 0, (cudaStream_t)stream>>>(n, yhat, y, scale, slot, dy);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_ce_fwd_bwd(int B, int C, const float* yhat, const long long* y, float scale, float* slot, float* dy,
                              void* stream) {
  MFM_REQUIRE(B > 0 && C > 0 && yhat && y && slot && dy);
  ce_kernel<<<grid_for(B, 256, 64), 256, 0, (cudaStream_t)stream>>>(B, C, yhat, y, scale, slot, dy);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_loss_total(float* lb, float l0, float l1, float l2, float lmmd, void* stream) {
  MFM_REQUIRE(lb);
  loss_total_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(lb, l0, l1, l2, lmmd);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_adam_step(long long n, float* p, const float* g, float* m, float* v, float* state, float grad_scale,
                             double beta1, double beta2, double eps, void* stream) {
  MFM_REQUIRE(n > 0 && p && g && m && v && state);
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
  MFM_LAUNCH_CHECK();
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, p, g, m, v, state, grad_scale, (float)beta1, (float)(1.0 - beta1),
                                                             (float)beta2, (float)(1.0 - beta2), (float)eps);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_randn(long long n, float* out, const long long* rng, int site, void* stream) {
  MFM_REQUIRE(n > 0 && out && rng);
  randn_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, out, rng, site);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_rng_tick(long long* rng, void* stream) {
  MFM_REQUIRE(rng);
  rng_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

// ---- debug: device-side timeline marks (scripts/step_timeline.py; there is no nsys in the image) -----------------------
__global__ void stamp_kernel(long long* buf, int slot) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[slot] = t;
}
extern "C" int mfm_debug_stamp(long long* buf, int slot, void* stream) {
  MFM_REQUIRE(buf && slot >= 0);
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(buf, slot);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
