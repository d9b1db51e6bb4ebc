// Tiled MMD (mfm_model.py:14-34): K(x,y)[i,j] = exp(-|x_i - y_j|^2 / dim^2) over all B*B pairs without
// materialising the reference's [B,B,dim] tensor (1.34 GB for z_v at B=2048).
//   fwd :  *out = mean K(g,g) + mean K(z,z) - 2 mean K(g,z)
//   bwd :  dz_i += scale * (2c/B^2) * [ sum_j Kzz_ij (z_i - z_j) - sum_j Kgz_ji (z_i - g_j) ],  c = -2/dim^2
// A CTA owns MMD_RI rows i and sweeps all j in tiles of MMD_TJ rows staged in shared memory.
#include "common.cuh"

#define MMD_THREADS 256
#define MMD_RI 16
#define MMD_TJ 64
#define MMD_MAXDIM 256

// distance phase: thread computes pairs (i = p / TJ, j = p % TJ); K written to Ks[i][j]
__device__ __forceinline__ float mmd_pair_phase(const float* xi, const float* yj, float* Ks, int dimp, int dim,
                                                int ni, int nj, float inv_d2) {
  float local = 0.0f;
  for (int p = threadIdx.x; p < MMD_RI * MMD_TJ; p += MMD_THREADS) {
    const int i = p / MMD_TJ, j = p - i * MMD_TJ;
    float k = 0.0f;
    if (i < ni && j < nj) {
      float d2 = 0.0f;
      const float* a = xi + i * dimp;
      const float* b = yj + j * dimp;
      for (int d = 0; d < dim; ++d) {
        const float df = a[d] - b[d];
        d2 = fmaf(df, df, d2);
      }
      k = expf(-d2 * inv_d2);
    }
    if (Ks) Ks[p] = k;
    local += k;
  }
  return local;
}

__device__ __forceinline__ void mmd_load_tile(float* dst, const float* src, long long ld, int row0, int nrows_total,
                                              int tile_rows, int dim, int dimp) {
  for (int idx = threadIdx.x; idx < tile_rows * dim; idx += MMD_THREADS) {
    const int r = idx / dim, d = idx - r * dim;
    dst[r * dimp + d] = (row0 + r < nrows_total) ? __ldg(src + (long long)(row0 + r) * ld + d) : 0.0f;
  }
}

__global__ void __launch_bounds__(MMD_THREADS) mmd_fwd_kernel(int B, int dim, const float* __restrict__ z, long long ldz,
                                                              const float* __restrict__ g, long long ldg,
                                                              float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int dimp = dim | 1;                          // odd row pitch: conflict-free column walks
  float* zi = smem;                                  // [RI][dimp]
  float* gi = zi + MMD_RI * dimp;                    // [RI][dimp]
  float* tj = gi + MMD_RI * dimp;                    // [TJ][dimp]
  const int i0 = blockIdx.x * MMD_RI;
  const int ni = min(MMD_RI, B - i0);
  const float inv_d2 = 1.0f / ((float)dim * (float)dim);
  mmd_load_tile(zi, z, ldz, i0, B, MMD_RI, dim, dimp);
  mmd_load_tile(gi, g, ldg, i0, B, MMD_RI, dim, dimp);
  float acc = 0.0f;
  for (int j0 = 0; j0 < B; j0 += MMD_TJ) {
    const int nj = min(MMD_TJ, B - j0);
    __syncthreads();
    mmd_load_tile(tj, z, ldz, j0, B, MMD_TJ, dim, dimp);
    __syncthreads();
    acc += mmd_pair_phase(zi, tj, nullptr, dimp, dim, ni, nj, inv_d2);          // K(z,z)
    acc -= 2.0f * mmd_pair_phase(gi, tj, nullptr, dimp, dim, ni, nj, inv_d2);   // K(g,z)
    __syncthreads();
    mmd_load_tile(tj, g, ldg, j0, B, MMD_TJ, dim, dimp);
    __syncthreads();
    acc += mmd_pair_phase(gi, tj, nullptr, dimp, dim, ni, nj, inv_d2);          // K(g,g)
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, tot / ((float)B * (float)B));
}

__global__ void __launch_bounds__(MMD_THREADS) mmd_bwd_kernel(int B, int dim, const float* __restrict__ z, long long ldz,
                                                              const float* __restrict__ g, long long ldg, float coef,
                                                              const float* __restrict__ scale_dev,
                                                              float* __restrict__ dz, long long lddz) {
  extern __shared__ __align__(16) float smem[];
  const int dimp = dim | 1;
  float* zi = smem;                                  // [RI][dimp]
  float* tj = zi + MMD_RI * dimp;                    // [TJ][dimp]
  float* Ks = tj + MMD_TJ * dimp;                    // [RI][TJ]
  const int i0 = blockIdx.x * MMD_RI;
  const int ni = min(MMD_RI, B - i0);
  const float inv_d2 = 1.0f / ((float)dim * (float)dim);
  mmd_load_tile(zi, z, ldz, i0, B, MMD_RI, dim, dimp);
  const float sdev = scale_dev ? __ldg(scale_dev) : 1.0f;
  constexpr int MAXI = (MMD_RI * MMD_MAXDIM + MMD_THREADS - 1) / MMD_THREADS;
  float acc[MAXI];
#pragma unroll
  for (int q = 0; q < MAXI; ++q) acc[q] = 0.0f;
  for (int pass = 0; pass < 2; ++pass) {            // pass 0: j over z (+), pass 1: j over g (-)
    const float* src = pass == 0 ? z : g;
    const long long ld = pass == 0 ? ldz : ldg;
    const float sgn = pass == 0 ? 1.0f : -1.0f;
    for (int j0 = 0; j0 < B; j0 += MMD_TJ) {
      const int nj = min(MMD_TJ, B - j0);
      __syncthreads();
      mmd_load_tile(tj, src, ld, j0, B, MMD_TJ, dim, dimp);
      __syncthreads();
      mmd_pair_phase(zi, tj, Ks, dimp, dim, ni, nj, inv_d2);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < MAXI; ++q) {
        const int item = threadIdx.x + q * MMD_THREADS;
        if (item < MMD_RI * dim) {
          const int i = item / dim, d = item - i * dim;
          const float zv = zi[i * dimp + d];
          float s = 0.0f;
          for (int j = 0; j < nj; ++j) s = fmaf(Ks[i * MMD_TJ + j], zv - tj[j * dimp + d], s);
          acc[q] = fmaf(sgn, s, acc[q]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MAXI; ++q) {
    const int item = threadIdx.x + q * MMD_THREADS;
    if (item < MMD_RI * dim) {
      const int i = item / dim, d = item - i * dim;
      if (i < ni) dz[(long long)(i0 + i) * lddz + d] += coef * sdev * acc[q];
    }
  }
}

extern "C" int mfm_mmd_fwd(int B, int dim, const float* z, long long ldz, const float* g, long long ldg, float* out,
                           void* stream) {
  MFM_REQUIRE(B > 0 && dim > 0 && dim <= MMD_MAXDIM && z && g && out);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int dimp = dim | 1;
  const size_t smem = (size_t)(2 * MMD_RI + MMD_TJ) * dimp * sizeof(float);
  if (int ea = mfm_func_smem_t(mmd_fwd_kernel, 128 * 1024)) return ea;
  mmd_fwd_kernel<<<(B + MMD_RI - 1) / MMD_RI, MMD_THREADS, smem, st>>>(B, dim, z, ldz, g, ldg, out);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

extern "C" int mfm_mmd_bwd(int B, int dim, const float* z, long long ldz, const float* g, long long ldg, float scale,
                           const float* scale_dev, float* dz, long long lddz, void* stream) {
  MFM_REQUIRE(B > 0 && dim > 0 && dim <= MMD_MAXDIM && z && g && dz);
  const int dimp = dim | 1;
  const size_t smem = ((size_t)(MMD_RI + MMD_TJ) * dimp + MMD_RI * MMD_TJ) * sizeof(float);
  if (int ea = mfm_func_smem_t(mmd_bwd_kernel, 128 * 1024)) return ea;
  const float c = -2.0f / ((float)dim * (float)dim);
  const float coef = scale * 2.0f * c / ((float)B * (float)B);
  mmd_bwd_kernel<<<(B + MMD_RI - 1) / MMD_RI, MMD_THREADS, smem, (cudaStream_t)stream>>>(B, dim, z, ldz, g, ldg, coef,
                                                                                         scale_dev, dz, lddz);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}


// ------------------------------------------------------------------------------------------------
// GEMM formulation used by the training schedule (engine.py step 7): the pair matrix S = X Y^T comes from the
// tensor-core GEMM, then  K_ij = exp(-(|x_i|^2 + |y_j|^2 - 2 S_ij)/dim^2)  in place, with the weighted mean
// accumulated into the loss slot.  The K matrices stay in HBM/L2 for the backward pass, which is two more GEMMs
// (K_zz Z and K_gz^T G) and mmd_combine.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rownorm2_kernel(int B, int dim, const float* __restrict__ x, long long ld,
                                                        float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.0f;
  for (int d = lane; d < dim; d += 32) {
    const float v = __ldg(x + (long long)row * ld + d);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// The MMD is a difference of three O(1) means that cancel to O(1/B): the means need ~1e-7 absolute accuracy for the
// loss term to hold 1e-3 relative at batch 2048.  Thousands of fp32 atomicAdds into one O(1) slot lose that (each rounds
// at 3e-8), so the sums are carried in double: per thread, per block, and in the accumulator (slot64); a float slot is
// supported for the stand-alone API and gets one add per block of the already weighted double sum.
__global__ void __launch_bounds__(256) mmd_kexp_kernel(int M, int N, float* __restrict__ S, const float* __restrict__ nx,
                                                        const float* __restrict__ ny, float inv_d2, double weight,
                                                        float* __restrict__ slot, double* __restrict__ slot64) {
  __shared__ double red[8];
  // one warp per row at a time, lanes along the row (float4 when the row allows): no 64-bit index division per element
  // (the first version spent most of its 26 us per 2048 x 2048 matrix on exactly that)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && ((reinterpret_cast<uintptr_t>(ny) & 15) == 0);
  double acc = 0.0;
  for (int m = blockIdx.x * 8 + wib; m < M; m += gridDim.x * 8) {
    const float nxm = __ldg(nx + m);
    float* row = S + (long long)m * N;
    float racc = 0.0f;                                        // one row's share of a lane: <= N/32 values in (0, 1]
    if (vec) {
      for (int n = lane * 4; n < N; n += 128) {
        float4 s4 = *reinterpret_cast<const float4*>(row + n);
        const float4 y4 = __ldg(reinterpret_cast<const float4*>(ny + n));
        s4.x = expf(-fmaxf(nxm + y4.x - 2.0f * s4.x, 0.0f) * inv_d2);
        s4.y = expf(-fmaxf(nxm + y4.y - 2.0f * s4.y, 0.0f) * inv_d2);
        s4.z = expf(-fmaxf(nxm + y4.z - 2.0f * s4.z, 0.0f) * inv_d2);
        s4.w = expf(-fmaxf(nxm + y4.w - 2.0f * s4.w, 0.0f) * inv_d2);
        *reinterpret_cast<float4*>(row + n) = s4;
        racc += (s4.x + s4.y) + (s4.z + s4.w);
      }
    } else {
      for (int n = lane; n < N; n += 32) {
        const float k = expf(-fmaxf(nxm + __ldg(ny + n) - 2.0f * row[n], 0.0f) * inv_d2);
        row[n] = k;
        racc += k;
      }
    }
    acc += (double)racc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    if (slot64) atomicAdd(slot64, tot * weight);
    else atomicAdd(slot, (float)(tot * weight));
  }
}
// slots[i] = (float)acc[i]: the double accumulators of the training schedule folded into the fp32 loss buffer
__global__ void mmd_fold_kernel(int n, const double* __restrict__ acc, float* __restrict__ slots) {
  const int i = threadIdx.x;
  if (i < n) slots[i] = (float)acc[i];
}

// dz += coef * sdev * ((rs - cs) * z - t1 + t2)
__global__ void __launch_bounds__(256) mmd_combine_kernel(int B, int dim, const float* __restrict__ z, long long ldz,
                                                           const float* __restrict__ rs, const float* __restrict__ cs,
                                                           const float* __restrict__ t1, const float* __restrict__ t2,
                                                           float coef, const float* __restrict__ scale_dev,
                                                           float* __restrict__ dz, long long lddz) {
  const float c = coef * (scale_dev ? __ldg(scale_dev) : 1.0f);
  const long long total = (long long)B * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / dim), d = (int)(i - (long long)b * dim);
    const float v = (rs[b] - cs[b]) * __ldg(z + (long long)b * ldz + d) - t1[i] + t2[i];
    dz[(long long)b * lddz + d] += c * v;
  }
}

extern "C" int mfm_rownorm2(int B, int dim, const float* x, long long ld, float* out, void* stream) {
  MFM_REQUIRE(B > 0 && dim > 0 && x && out);
  rownorm2_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(B, dim, x, ld, out);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
static int mmd_kexp_launch(int M, int N, float* S, const float* nx, const float* ny, int dim, double weight, float* slot,
                           double* slot64, void* stream) {
  MFM_REQUIRE(M > 0 && N > 0 && dim > 0 && S && nx && ny && (slot || slot64));
  long long blocks = (M + 7) / 8;
  if (blocks > mfm_dev_info().sms * 8) blocks = mfm_dev_info().sms * 8;
  if (blocks < 1) blocks = 1;
  mmd_kexp_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(M, N, S, nx, ny, 1.0f / ((float)dim * (float)dim), weight, slot,
                                                                  slot64);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_mmd_kexp(int M, int N, float* S, const float* nx, const float* ny, int dim, float weight, float* slot,
                            void* stream) {
  return mmd_kexp_launch(M, N, S, nx, ny, dim, (double)weight, slot, nullptr, stream);
}
extern "C" int mfm_mmd_kexp64(int M, int N, float* S, const float* nx, const float* ny, int dim, double weight, double* slot64,
                              void* stream) {
  return mmd_kexp_launch(M, N, S, nx, ny, dim, weight, nullptr, slot64, stream);
}
extern "C" int mfm_mmd_fold(int n, const double* acc, float* slots, void* stream) {
  MFM_REQUIRE(n > 0 && n <= 32 && acc && slots);
  mmd_fold_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(n, acc, slots);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
extern "C" int mfm_mmd_combine(int B, int dim, const float* z, long long ldz, const float* rs, const float* cs,
                               const float* t1, const float* t2, float scale, const float* scale_dev, float* dz,
                               long long lddz, void* stream) {
  MFM_REQUIRE(B > 0 && dim > 0 && z && rs && cs && t1 && t2 && dz);
  const float c = -2.0f / ((float)dim * (float)dim);
  const float coef = scale * 2.0f * c / ((float)B * (float)B);
  long long blocks = ((long long)B * dim + 255) / 256;
  if (blocks > mfm_dev_info().sms * 8) blocks = mfm_dev_info().sms * 8;
  mmd_combine_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(B, dim, z, ldz, rs, cs, t1, t2, coef, scale_dev, dz, lddz);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
