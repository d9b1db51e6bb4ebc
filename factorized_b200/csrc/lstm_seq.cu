// LSTM recurrences with the timestep loop inside the kernel (fp32 CUDA-core path).
//
// forward : pre_t = h_{t-1} W^T + gx[t]  ->  i,f,g,o  ->  c_t, h_t        (mfm_model.py:56,83,85,167-169)
// backward: reverse loop, dG_t (grad of pre-activation gates) written to HBM, dh_{t-1} = dG_t W.
// The x-side projection (x_t W_ih^T + b) is NOT here: it is hoisted into one GEMM over all T*B rows
// (engine.py step 1); weight gradients are likewise time-parallel GEMMs over the stashed dG.
//
// Work split: a CTA owns RT = groups*R batch rows of one cell (blockIdx.y = cell) for all T steps; a
// thread owns hidden unit j for R rows, so the four gates of (row, j) meet in one thread and the cell
// update needs no exchange.  W lives in shared memory (transposed for the forward pass so that
// consecutive threads read consecutive banks); h_{t-1} / dG_t tiles are double-buffered in shared memory,
// one __syncthreads per step.
#include "common.cuh"

#define LSTM_THREADS 256
#define LSTM_R 8

struct LstmBatch {
  mfm_lstm_cell c[MFM_MAX_CELLS];
  int n;
  int smem_limit;
};

__device__ __forceinline__ void lstm_geometry(int h, int& groups, int& jstride) {
  if (h >= LSTM_THREADS) { groups = 1; jstride = LSTM_THREADS; }
  else { groups = LSTM_THREADS / h; jstride = h; }
}

static inline int lstm_groups_host(int h) { return h >= LSTM_THREADS ? 1 : LSTM_THREADS / h; }

static inline size_t lstm_fwd_smem(int h, bool with_w) {
  int hp = (h + 3) & ~3;
  int RT = lstm_groups_host(h) * LSTM_R;
  size_t s = (size_t)2 * RT * hp * sizeof(float);
  if (with_w) s += (size_t)hp * 4 * h * sizeof(float);
  return s;
}
static inline size_t lstm_bwd_smem(int h, bool with_w) {
  int RT = lstm_groups_host(h) * LSTM_R;
  size_t s = (size_t)2 * RT * 4 * h * sizeof(float) + (size_t)2 * RT * h * sizeof(float);
  if (with_w) s += (size_t)4 * h * h * sizeof(float);
  return s;
}

__global__ void __launch_bounds__(LSTM_THREADS) lstm_seq_fwd_kernel(LstmBatch bt) {
  extern __shared__ __align__(16) float smem[];
  const mfm_lstm_cell& c = bt.c[blockIdx.y];
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  const int hp = (h + 3) & ~3;
  int groups, jstride;
  lstm_geometry(h, groups, jstride);
  const int RT = groups * LSTM_R;
  const int row0 = blockIdx.x * RT;
  if (row0 >= B) return;
  const int tid = threadIdx.x;
  const bool wsm = (size_t)2 * RT * hp * 4 + (size_t)hp * H4 * 4 <= (size_t)bt.smem_limit;
  float* hbuf = smem;                      // [2][RT][hp]
  float* Wt = smem + 2 * RT * hp;          // [hp][H4]  Wt[k][n] = W[n][k]
  if (wsm) {
    for (int idx = tid; idx < hp * H4; idx += LSTM_THREADS) {
      int n = idx / hp, k = idx - n * hp;  // read W row-wise (coalesced), scatter transposed
      Wt[k * H4 + n] = (k < h) ? __ldg(c.W + (long long)n * h + k) : 0.0f;
    }
  }
  for (int idx = tid; idx < 2 * RT * hp; idx += LSTM_THREADS) hbuf[idx] = 0.0f;
  // zero block 0 of the histories
  for (int idx = tid; idx < RT * h; idx += LSTM_THREADS) {
    int r = idx / h, j = idx - r * h;
    if (row0 + r < B) {
      c.hs[(long long)(row0 + r) * c.ld_hs + j] = 0.0f;
      c.cs[(long long)(row0 + r) * c.ld_cs + j] = 0.0f;
      if (c.cs_dup) c.cs_dup[(long long)(row0 + r) * c.ld_cs + j] = 0.0f;
    }
  }
  __syncthreads();
  const int gid = tid / jstride;
  const int j0 = tid - gid * jstride;
  const bool active = gid < groups;
  const int rbase = gid * LSTM_R;          // first tile row of this thread

  for (int t = 0; t < T; ++t) {
    const float* hcur = hbuf + (t & 1) * RT * hp;
    float* hnxt = hbuf + ((t + 1) & 1) * RT * hp;
    if (active) {
      for (int j = j0; j < h; j += jstride) {
        float acc[LSTM_R][4];
        if (t < c.gx_steps) {
#pragma unroll
          for (int r = 0; r < LSTM_R; ++r) {
            const int row = row0 + rbase + r;
            if (row < B) {
              const float* gp = c.gx + ((long long)t * B + row) * (c.ld_gx ? c.ld_gx : (long long)H4) + j;
#pragma unroll
              for (int g = 0; g < 4; ++g) acc[r][g] = __ldg(gp + g * h);
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g) acc[r][g] = 0.0f;
            }
          }
        } else {
          float bb[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) bb[g] = c.bias_rest ? __ldg(c.bias_rest + g * h + j) : 0.0f;
#pragma unroll
          for (int r = 0; r < LSTM_R; ++r)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[r][g] = bb[g];
        }
        if (t > 0) {
          if (wsm) {
            for (int k = 0; k < hp; k += 4) {
              float w[4][4];
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int g = 0; g < 4; ++g) w[kk][g] = Wt[(k + kk) * H4 + g * h + j];
#pragma unroll
              for (int r = 0; r < LSTM_R; ++r) {
                const float4 hv = *reinterpret_cast<const float4*>(hcur + (rbase + r) * hp + k);
                const float hr[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                  for (int g = 0; g < 4; ++g) acc[r][g] = fmaf(hr[kk], w[kk][g], acc[r][g]);
              }
            }
          } else {
            for (int k = 0; k < h; ++k) {
              float w[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) w[g] = __ldg(c.W + (long long)(g * h + j) * h + k);
#pragma unroll
              for (int r = 0; r < LSTM_R; ++r) {
                const float hv = hcur[(rbase + r) * hp + k];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[r][g] = fmaf(hv, w[g], acc[r][g]);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < LSTM_R; ++r) {
          const int row = row0 + rbase + r;
          if (row >= B) continue;
          const float ig = gate_sigmoid(acc[r][0]);
          const float fg = gate_sigmoid(acc[r][1]);
          const float gg = gate_tanh(acc[r][2]);
          const float og = gate_sigmoid(acc[r][3]);
          const float cp = c.cs[((long long)t * B + row) * c.ld_cs + j];
          const float cn = fg * cp + ig * gg;
          const float hn = og * gate_tanh(cn);
          float* gp = c.gates + ((long long)t * B + row) * H4 + j;
          gp[0] = ig; gp[h] = fg; gp[2 * h] = gg; gp[3 * h] = og;
          c.cs[((long long)(t + 1) * B + row) * c.ld_cs + j] = cn;
          if (c.cs_dup) c.cs_dup[((long long)(t + 1) * B + row) * c.ld_cs + j] = cn;
          c.hs[((long long)(t + 1) * B + row) * c.ld_hs + j] = hn;
          hnxt[(rbase + r) * hp + j] = hn;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(LSTM_THREADS) lstm_seq_bwd_kernel(LstmBatch bt) {
  extern __shared__ __align__(16) float smem[];
  const mfm_lstm_cell& c = bt.c[blockIdx.y];
  const int h = c.h, B = c.B, T = c.T, H4 = 4 * c.h;
  int groups, jstride;
  lstm_geometry(h, groups, jstride);
  const int RT = groups * LSTM_R;
  const int row0 = blockIdx.x * RT;
  if (row0 >= B) return;
  const int tid = threadIdx.x;
  const bool wsm = ((size_t)2 * RT * H4 + (size_t)2 * RT * h + (size_t)H4 * h) * 4 <= (size_t)bt.smem_limit;
  float* dGs = smem;                        // [2][RT][H4]
  float* dhs = smem + 2 * RT * H4;          // [RT][h]   recurrent dh
  float* dcs = dhs + RT * h;                // [RT][h]   carried dc
  float* Ws = dcs + RT * h;                 // [H4][h]   natural layout
  if (wsm)
    for (int idx = tid; idx < H4 * h; idx += LSTM_THREADS) Ws[idx] = __ldg(c.W + idx);
  for (int idx = tid; idx < 2 * RT * h; idx += LSTM_THREADS) dhs[idx] = 0.0f;   // dhs and dcs
  if (c.dc_last)                                                                // time-split recurrence: carried dc comes in
    for (int idx = tid; idx < RT * h; idx += LSTM_THREADS) {
      const int r = idx / h, j = idx - r * h;
      if (row0 + r < B) dcs[idx] = __ldg(c.dc_last + (long long)(row0 + r) * h + j);
    }
  for (int idx = tid; idx < 2 * RT * H4; idx += LSTM_THREADS) dGs[idx] = 0.0f;
  __syncthreads();
  const int gid = tid / jstride;
  const int j0 = tid - gid * jstride;
  const bool active = gid < groups;
  const int rbase = gid * LSTM_R;

  for (int t = T - 1; t >= 0; --t) {
    float* dGc = dGs + (t & 1) * RT * H4;
    if (active) {
      for (int j = j0; j < h; j += jstride) {
#pragma unroll
        for (int r = 0; r < LSTM_R; ++r) {
          const int row = row0 + rbase + r;
          if (row >= B) continue;
          const long long tr = (long long)t * B + row;
          float dh = dhs[(rbase + r) * h + j];
          if (c.dh_all) dh += __ldg(c.dh_all + tr * c.ld_dh_all + j);
          if (c.dh_last && t == T - 1) dh += __ldg(c.dh_last + (long long)row * c.ld_dh_last + j);
          const float* gp = c.gates + tr * H4 + j;
          const float ig = gp[0], fg = gp[h], gg = gp[2 * h], og = gp[3 * h];
          const float cp = c.cs[tr * c.ld_cs + j];
          const float cn = c.cs[(tr + B) * c.ld_cs + j];
          const float tc = gate_tanh(cn);
          float dc = dcs[(rbase + r) * h + j] + dh * og * (1.0f - tc * tc);
          if (c.dc_ext) dc += __ldg(c.dc_ext + tr * c.ld_dc_ext + j);
          if (c.dc_ext2 && (t < c.T - 1 || c.dc_ext2_full)) dc += __ldg(c.dc_ext2 + tr * c.ld_dc_ext + j);
          const float d_o = dh * tc * og * (1.0f - og);
          const float d_i = dc * gg * ig * (1.0f - ig);
          const float d_f = dc * cp * fg * (1.0f - fg);
          const float d_g = dc * ig * (1.0f - gg * gg);
          float* op = c.dG + tr * H4 + j;
          op[0] = d_i; op[h] = d_f; op[2 * h] = d_g; op[3 * h] = d_o;
          float* sp = dGc + (rbase + r) * H4 + j;
          sp[0] = d_i; sp[h] = d_f; sp[2 * h] = d_g; sp[3 * h] = d_o;
          dcs[(rbase + r) * h + j] = dc * fg;
        }
      }
    }
    __syncthreads();
    if (active && (t > 0 || c.dh_out)) {
      for (int j = j0; j < h; j += jstride) {
        float acc[LSTM_R];
#pragma unroll
        for (int r = 0; r < LSTM_R; ++r) acc[r] = 0.0f;
        for (int n = 0; n < H4; n += 4) {
          float w[4];
          if (wsm) {
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = Ws[(n + q) * h + j];
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = __ldg(c.W + (long long)(n + q) * h + j);
          }
#pragma unroll
          for (int r = 0; r < LSTM_R; ++r) {
            const float4 dv = *reinterpret_cast<const float4*>(dGc + (rbase + r) * H4 + n);
            acc[r] = fmaf(dv.x, w[0], acc[r]);
            acc[r] = fmaf(dv.y, w[1], acc[r]);
            acc[r] = fmaf(dv.z, w[2], acc[r]);
            acc[r] = fmaf(dv.w, w[3], acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < LSTM_R; ++r) dhs[(rbase + r) * h + j] = acc[r];
      }
    }
    // no second barrier: the next step writes the other dG buffer, and dhs/dcs entries are thread-private
  }
  if (active && c.dh_out) {                                   // time-split recurrence: the carried state goes out
    for (int j = j0; j < h; j += jstride)
#pragma unroll
      for (int r = 0; r < LSTM_R; ++r) {
        const int row = row0 + rbase + r;
        if (row >= B) continue;
        c.dh_out[(long long)row * h + j] = dhs[(rbase + r) * h + j];
        c.dc_out[(long long)row * h + j] = dcs[(rbase + r) * h + j];
      }
  }
}

static int lstm_validate(const mfm_lstm_cell* cells, int n, bool bwd) {
  if (!cells || n <= 0 || n > MFM_MAX_CELLS) return MFM_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    const mfm_lstm_cell& c = cells[i];
    if (c.T <= 0 || c.B <= 0 || c.h <= 0 || !c.W || !c.cs || !c.gates) return MFM_ERR_ARG;
    if (!bwd && (!c.hs || !c.gx || c.gx_steps <= 0 || c.gx_steps > c.T)) return MFM_ERR_ARG;
    if (bwd && !c.dG) return MFM_ERR_ARG;
    if (bwd && ((c.dh_out == nullptr) != (c.dc_out == nullptr))) return MFM_ERR_ARG;
  }
  return MFM_OK;
}

static int smem_optin_limit() { return mfm_dev_info().smem_optin; }

int lstm_tc_fwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st);
void ws_count_fallback(bool bwd, int ncells);

extern "C" int mfm_lstm_seq_fwd(const mfm_lstm_cell* cells, int ncells, void* stream) {
  int rc = lstm_validate(cells, ncells, false);
  if (rc) return rc;
  mfm_lstm_cell rest[MFM_MAX_CELLS];
  if (mfm_get_gemm_path() != MFM_PATH_SIMT_FP32) {      // tensor-core recurrence for every cell that fits on chip
    int nrest = 0;
    rc = lstm_tc_fwd_launch(cells, ncells, rest, &nrest, (cudaStream_t)stream);
    if (rc) return rc;
    if (nrest == 0) return MFM_OK;
    cells = rest;
    ncells = nrest;
  }
  LstmBatch bt;
  bt.n = ncells;
  bt.smem_limit = smem_optin_limit();
  size_t smem = 0;
  int gx = 0;
  for (int i = 0; i < ncells; ++i) {
    bt.c[i] = cells[i];
    size_t s = lstm_fwd_smem(cells[i].h, true);
    if (s > (size_t)bt.smem_limit) s = lstm_fwd_smem(cells[i].h, false);
    if (s > (size_t)bt.smem_limit) return MFM_ERR_UNSUPPORTED;
    if (s > smem) smem = s;
    int RT = lstm_groups_host(cells[i].h) * LSTM_R;
    int tiles = (cells[i].B + RT - 1) / RT;
    if (tiles > gx) gx = tiles;
  }
  if (int e = mfm_func_smem_t(lstm_seq_fwd_kernel, bt.smem_limit)) return e;
  ws_count_fallback(false, ncells);
  lstm_seq_fwd_kernel<<<dim3(gx, ncells), LSTM_THREADS, smem, (cudaStream_t)stream>>>(bt);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}

int lstm_tc_bwd_launch(const mfm_lstm_cell* cells, int ncells, mfm_lstm_cell* rest, int* nrest, cudaStream_t st);

extern "C" int mfm_lstm_seq_bwd(const mfm_lstm_cell* cells, int ncells, void* stream) {
  int rc = lstm_validate(cells, ncells, true);
  if (rc) return rc;
  mfm_lstm_cell rest[MFM_MAX_CELLS];
  if (mfm_get_gemm_path() != MFM_PATH_SIMT_FP32) {
    int nrest = 0;
    rc = lstm_tc_bwd_launch(cells, ncells, rest, &nrest, (cudaStream_t)stream);
    if (rc) return rc;
    if (nrest == 0) return MFM_OK;
    cells = rest;
    ncells = nrest;
  }
  LstmBatch bt;
  bt.n = ncells;
  bt.smem_limit = smem_optin_limit();
  size_t smem = 0;
  int gx = 0;
  for (int i = 0; i < ncells; ++i) {
    bt.c[i] = cells[i];
    size_t s = lstm_bwd_smem(cells[i].h, true);
    if (s > (size_t)bt.smem_limit) s = lstm_bwd_smem(cells[i].h, false);
    if (s > (size_t)bt.smem_limit) return MFM_ERR_UNSUPPORTED;
    if (s > smem) smem = s;
    int RT = lstm_groups_host(cells[i].h) * LSTM_R;
    int tiles = (cells[i].B + RT - 1) / RT;
    if (tiles > gx) gx = tiles;
  }
  if (int e = mfm_func_smem_t(lstm_seq_bwd_kernel, bt.smem_limit)) return e;
  ws_count_fallback(true, ncells);
  lstm_seq_bwd_kernel<<<dim3(gx, ncells), LSTM_THREADS, smem, (cudaStream_t)stream>>>(bt);
  MFM_LAUNCH_CHECK();
  return MFM_OK;
}
