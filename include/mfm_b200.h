/*
 * mfm_b200.h -- C ABI of libmfm_b200.so: the sm_100a kernels behind the MFM training step.
 *
 * The reference (pliang279/factorized) has no FFI or operator registry: its hot path is
 * Python calling torch ops.  The drop-in boundary a user sees is
 * therefore the Python nn.Module API (factorized_b200/mfm_model.py mirrors
 * /root/reference/mfm_model.py), and this header is what that host code binds through
 * ctypes.  Each entry point names the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers, sizes, leading dimensions (in elements), a cudaStream_t
 *    passed as void*.  No torch types, no allocation, no host synchronisation: every call
 *    only enqueues kernels on `stream`, so a whole step is CUDA-graph capturable.
 *  - matrices are row-major fp32 with unit column stride and leading dimension `ld`.
 *  - "[T*B, n]" means time-major row blocks: row t*B+b.  State histories have T+1 blocks,
 *    block 0 being the zero initial state (written by the kernel).
 *  - return value: 0 on success, MFM_ERR_* (<0) for bad arguments, or a positive
 *    cudaError_t from the launch.  Nothing throws.
 */
#ifndef MFM_B200_H
#define MFM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MFM_B200_VERSION 100

#define MFM_OK 0
#define MFM_ERR_ARG (-1)       /* null pointer / non-positive size / unsupported combination */
#define MFM_ERR_UNSUPPORTED (-2)

/* GEMM operand layouts */
#define MFM_GEMM_NT 0   /* C[M,N] = A[M,K] * B[N,K]^T   (y = x W^T: forward Linear / gate projections) */
#define MFM_GEMM_NN 1   /* C[M,N] = A[M,K] * B[K,N]     (dx = dy W: data gradients)                     */
#define MFM_GEMM_TN 2   /* C[M,N] = A[K,M]^T * B[K,N]   (dW = dy^T x: weight gradients, K = T*B rows)   */

#define MFM_ACT_NONE 0
#define MFM_ACT_RELU 1
#define MFM_ACT_TANH 2
#define MFM_ACT_SIGMOID 3

/* GEMM math paths (mfm_set_gemm_path) */
#define MFM_PATH_SIMT_FP32 0    /* CUDA-core fp32 FMA */
#define MFM_PATH_TC_BF16X3 1    /* tcgen05 kind::f16, operands split hi+lo bf16, 3 MMAs, fp32 accumulate in TMEM */
#define MFM_PATH_TC_BF16 2      /* tcgen05 kind::f16, single bf16 pass */

int mfm_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
unsigned long long mfm_launch_count(void);
int mfm_set_gemm_path(int path);
int mfm_get_gemm_path(void);
/* GEMMs with M*N*K below this stay on the CUDA-core kernel even when a tcgen05 path is selected (default 2^20) */
int mfm_set_gemm_tc_min_work(long long mnk);

/* C = epilogue(op(A) op(B)).  Replaces every nn.Linear / LSTMCell input projection on the path:
 * mfm_model.py:56,61,83,85,90,167-169,174,176,178-179,535,539-542,552 and their autograd adjoints.
 * epilogue, in order: + bias[n] + bias2[n]; activation; counter-based dropout (drop_p > 0:
 * keep iff u(rng, drop_site, m*N+n) >= drop_p, scaled 1/(1-p));  * (mask[m,n] > 0) * mask_scale;
 * + C if accumulate.  MFM_GEMM_TN may split K across CTAs and then needs accumulate=1.
 * colsum_out (MFM_GEMM_TN only, may be NULL): colsum_out[m] += sum_k A[k,m] -- the bias gradient that always
 * accompanies a weight gradient dW = dY^T X, fused as a virtual all-ones column of X. */
int mfm_gemm(int mode, int M, int N, int K,
             const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
             const float* bias, const float* bias2, int act, int accumulate,
             const float* mask, long long ldmask, float mask_scale,
             float drop_p, int drop_site, const long long* rng, float* colsum_out, void* stream);
/* The same GEMM with a caller-owned device workspace (128 B aligned, ws_bytes >= 8*(N+64)*(K+16) is always enough; the
 * library never allocates).  With it, NT / NN GEMMs over many rows (M >= 4096) split the small B operand -- the weight --
 * into its bf16 hi/lo MMA image once per call instead of once per 128-row tile.  The workspace is only read by kernels
 * enqueued by this call on `stream`; reuse it for the next call on the same stream, not across streams. */
/* Debug aid (scripts/gemm_trace.py): while a device buffer is registered, every pipelined-GEMM CTA records clock stamps of
 * its producer / converter / MMA roles into it (4 + 6*32 int64 words per CTA).  NULL disables.  Not for production use. */
int mfm_debug_set_gemm_trace(void* device_buf, long long bytes);
/* Debug aid (scripts/step_timeline.py): a one-thread kernel that writes %globaltimer (ns) to buf[slot] when the stream
 * reaches it -- milestones of a CUDA-graph replay, with all cross-stream overlap included. */
int mfm_debug_stamp(long long* buf, int slot, void* stream);
int mfm_gemm_ws(int mode, int M, int N, int K,
                const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                const float* bias, const float* bias2, int act, int accumulate,
                const float* mask, long long ldmask, float mask_scale,
                float drop_p, int drop_site, const long long* rng, float* colsum_out,
                void* ws, long long ws_bytes, void* stream);

/* The reconstruction head fused into one GEMM (decoderLSTM.forward's fc1, mfm_model.py:88-90, with the MSE term of the
 * train step, mfm_mosi.py:437):   x_hat = A B^T + bias;   *slot += loss_scale * sum (x_hat - x)^2;
 * dxhat = grad_scale * (x_hat - x)  (the gradient of lambda * mean((x_hat - x)^2) when grad_scale = 2 lambda / numel).
 * x_hat itself is written only when `xhat` is not NULL (predict()); in training it never touches HBM. */
int mfm_gemm_mse(int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, const float* bias,
                 const float* x, long long ldx, float loss_scale, float grad_scale, float* slot,
                 float* dxhat, long long lddx, float* xhat, long long ldxhat, void* ws, long long ws_bytes, void* stream);

/* Two weight gradients that share dY -- dW_ih = dG^T x and dW_hh = dG^T h_prev of one LSTM cell (autograd adjoint of
 * mfm_model.py:56,167-169), or the two column blocks of gamma*_fc1 (:178-179) -- in one launch:
 *   C1[M,N1] += A^T B1,  colsum1[m] += sum_k A[k,m] (may be NULL),  C2[M,N2] += A^T B2,   A = dY [K, M].
 * dY is streamed from HBM once instead of twice. */
int mfm_gemm_tn_pair(int M, int K, const float* A, long long lda,
                     int N1, const float* B1, long long ldb1, float* C1, long long ldc1, float* colsum1,
                     int N2, const float* B2, long long ldb2, float* C2, long long ldc2, void* stream);

/* One LSTM cell unrolled over T steps inside the kernel (encoderLSTM.forward mfm_model.py:47-62,
 * decoderLSTM.forward :72-91, the three cells of MFN.forward :167-169).
 * pre_t = h_{t-1} W^T + (t < gx_steps ? gx[t] : bias_rest);  gates i,f,g,o (torch LSTMCell order);
 * c_t = sig(f) c_{t-1} + sig(i) tanh(g);  h_t = sig(o) tanh(c_t). */
typedef struct mfm_lstm_cell {
  int T, B, h, gx_steps;
  const float* gx;          /* [gx_steps*B, 4h] contiguous: hoisted x W_ih^T + b_ih + b_hh          */
  const float* bias_rest;   /* [4h] or NULL: added for t >= gx_steps (decoder)                      */
  const float* W;           /* [4h, h] contiguous recurrent weight (W_hh, or W_ih+W_hh for decoder) */
  float* hs; long long ld_hs;      /* [(T+1)*B, h]  hidden history, block 0 = 0                     */
  float* cs; long long ld_cs;      /* [(T+1)*B, h]  cell history,   block 0 = 0                     */
  float* gates;             /* [T*B, 4h] contiguous post-activation i,f,g,o (stash for backward)    */
  /* backward only */
  const float* dh_all; long long ld_dh_all;   /* [T*B, h] dL/dh_t from outside, or NULL             */
  const float* dh_last; long long ld_dh_last; /* [B, h]   dL/dh_{T-1} from outside, or NULL         */
  const float* dc_ext; long long ld_dc_ext;   /* [T*B, h] dL/dc_t from outside, or NULL             */
  float* dG;                /* [T*B, 4h] contiguous: dL/d(pre-activation gates)                      */
  float* dc_scratch;        /* [B, h] contiguous workspace (carried dc) for the tensor-core backward;
                               NULL selects the CUDA-core kernel                                      */
  /* The MFN attention reads cat(c_{t-1}, c_t) (mfm_model.py:171-173).  Instead of copying the cell history into
   * that layout, the forward kernel writes every c block a second time (cs_dup, leading dimension ld_cs) and the
   * backward kernel adds a second external cell gradient: dc_t += dc_ext2[t] for t < T-1 (ld_dc_ext), i.e. the
   * "previous c" half of the next step's concatenation.  Both may be NULL. */
  float* cs_dup;            /* forward only:  [(T+1)*B, h], receives the same blocks as cs           */
  const float* dc_ext2;     /* backward only: [(T-1)*B, h]                                            */
  long long ld_gx;          /* row pitch of gx in floats (0 = 4h): lets the hoisted projections of the two cells that
                               read the same modality (encoder + MFN cell) be column blocks of ONE GEMM's output  */
  /* Backward only: a recurrence split in TIME into two launches -- steps [t0, T) first, then [0, t0) -- so that the
   * weight-gradient GEMM of the late half can start while the early half is still running.  The first launch hands the
   * recurrent state on (dh_out, dc_out), the second takes it (dh_last = that dh_out, dc_last = that dc_out).  All NULL / 0
   * for an unsplit recurrence. */
  const float* dc_last;     /* [B, h] contiguous: dL/dc carried INTO the last step of this launch, or NULL       */
  float* dh_out;            /* [B, h] contiguous: receives W^T dG_0, the gradient w.r.t. h before step 0, or NULL */
  float* dc_out;            /* [B, h] contiguous: receives dc_0 * f_0 (w.r.t. c before step 0); set with dh_out  */
  int dc_ext2_full;         /* dc_ext2 has T blocks and also applies at the last step of this launch             */
} mfm_lstm_cell;
#define MFM_MAX_CELLS 8
/* all cells of one call run concurrently (blockIdx.y = cell) */
int mfm_lstm_seq_fwd(const mfm_lstm_cell* cells, int ncells, void* stream);
int mfm_lstm_seq_bwd(const mfm_lstm_cell* cells, int ncells, void* stream);
/* Test hooks for the recurrence kernels (tests/test_gpu_primitives.py): force the narrow (16-row) chains where they are
 * legal (nb = 16; 0 = automatic), and read how many cells each variant has served since load --
 * 0 fwd 32-row chains, 1 fwd 16-row chains, 2 bwd 32-row, 3 bwd 16-row, 4 / 5 fwd / bwd CUDA-core kernel (h > 128, exact-fp32 path). */
int mfm_debug_lstm_force_nb(int nb);
int mfm_debug_lstm_force_chains(int n);        /* chains (batch sub-tiles) per CTA: 1, 2, or 0 = by occupancy */
/* Debug aid (scripts/lstm_trace.py): while a device buffer of 16*32*4 int64 is registered, CTA 0 of every forward
 * recurrence launch records clock stamps per warp and step (wait start, wait done, epilogue done, MMA issue done). */
int mfm_debug_set_lstm_trace(void* device_buf);
unsigned long long mfm_debug_lstm_variant_count(int variant);
/* Launches of the persistent streaming GEMM (csrc/gemm_ps.cu) since load: the tests assert that the large-M layers run on it. */
int mfm_debug_gemm_ps_count(void);
/* Residency of the persistent GEMM's weight image: 0 = stream it with every ring stage (default), 1 = keep the N tile's whole image
 * in shared memory where it fits (each CTA then serves one N tile), 2 = let the cost model choose.  (env MFM_PS_RES sets the start-up value.) */
int mfm_debug_gemm_ps_residency(int mode);
/* Debug aid (scripts/gemm_ps_trace.py; library built with -DPS_DEBUG=1, else MFM_ERR_UNSUPPORTED): while a device buffer of
 * 148 * (4 + 6*64 + 4*16) int64 is registered, every CTA of the persistent GEMM records clock stamps of its roles. */
int mfm_debug_set_gemm_ps_trace(void* device_buf);

/* The MFN memory recurrence, mfm_model.py:177-180, T steps in one kernel:
 * u_k = dropout(relu(Gkpre[t] + mem W_km^T));  gamma_k = sig(u_k W_k2^T + b_k2);
 * mem' = gamma_1 mem + gamma_2 cHat[t].  Gkpre already holds attended W_k1[:, :2H]^T + b. */
typedef struct mfm_mem_args {
  int T, B, mem, g1, g2;
  const float* G1pre; const float* G2pre;     /* [T*B, g1], [T*B, g2] contiguous */
  const float* cHat;                          /* [T*B, mem] contiguous           */
  const float* W1m; long long ld_w1m;         /* [g1, mem] = gamma1_fc1.weight[:, 2H:] */
  const float* W2m; long long ld_w2m;
  const float* W12; const float* b12;         /* [mem, g1], [mem] */
  const float* W22; const float* b22;
  float* mems;                                /* [(T+1)*B, mem] contiguous, block 0 = 0 */
  float* U1; float* U2;                       /* [T*B, g*]  post relu/dropout            */
  float* Gam1; float* Gam2;                   /* [T*B, mem] */
  float drop_p1, drop_p2; int site1, site2; const long long* rng;
  /* backward only */
  float scale1, scale2;                       /* 1/(1-p) when dropout was active, else 1 */
  const float* dmem_last; long long ld_dmem_last;   /* [B, mem] */
  float* dU1; float* dU2;                     /* [T*B, g*]  grad wrt pre-relu gamma*_fc1 output */
  float* dP1; float* dP2;                     /* [T*B, mem] grad wrt pre-sigmoid gamma*_fc2 output */
  float* dPc;                                 /* [T*B, mem] grad wrt pre-tanh att2_fc2 output */
  long long ld_dU1, ld_dU2;                   /* row pitch of dU1 / dU2 (0 = contiguous): lets them be column blocks of one
                                                 matrix so the data gradient of `attended` is a single GEMM */
} mfm_mem_args;
int mfm_mfn_mem_fwd(const mfm_mem_args* a, void* stream);
int mfm_mfn_mem_bwd(const mfm_mem_args* a, void* stream);
/* Test hooks: the recurrence runs on the tensor cores (csrc/mem_ws.cu) when mem <= 64 and g1, g2 <= 128, else on CUDA cores
 * (csrc/mfn.cu); force the latter (on != 0), and read how many launches the tensor-core kernels have served since load. */
int mfm_debug_mem_force_simt(int on);
unsigned long long mfm_debug_mem_ws_count(int bwd);
/* Debug aid (scripts/mem_trace.py; library built with -DMW_DEBUG=1, else MFM_ERR_UNSUPPORTED): copies the clock stamps CTA 0 of the
 * last forward launch recorded (32 steps x 16 slots, int64) to the HOST buffer. */
int mfm_debug_mem_ws_trace(long long* host32x16);

/* attention = softmax(L, dim=1) (in place); attended = attention * cstar  (mfm_model.py:174-175) */
int mfm_softmax_gate_fwd(int M, int N, float* L, const float* cstar, float* attended, void* stream);
int mfm_softmax_gate_bwd(int M, int N, const float* dAttended, const float* att, const float* cstar,
                         float* dL, float* dcstar, void* stream);

/* loss_MMD (mfm_model.py:14-34) with the Gaussian sample passed in:  *out = mean K(g,g) + mean K(z,z)
 * - 2 mean K(g,z),  K(x,y) = exp(-|x-y|^2/dim^2).  No [B,B,dim] tensor is materialised. dim <= 256. */
int mfm_mmd_fwd(int B, int dim, const float* z, long long ldz, const float* g, long long ldg, float* out, void* stream);
/* dz += scale * (scale_dev ? *scale_dev : 1) * d(MMD)/dz ; scale_dev is an optional DEVICE scalar (autograd's
 * upstream gradient) so the drop-in path needs no host synchronisation */
int mfm_mmd_bwd(int B, int dim, const float* z, long long ldz, const float* g, long long ldg, float scale,
                const float* scale_dev, float* dz, long long lddz, void* stream);
/* GEMM formulation of the same loss for the training schedule: S = X Y^T is produced by mfm_gemm, then
 *   mfm_mmd_kexp:    S[i,j] <- K_ij = exp(-(nx[i] + ny[j] - 2 S[i,j]) / dim^2);  *slot += weight * sum_ij K_ij
 *   mfm_rownorm2:    out[i] = |x_i|^2
 *   mfm_mmd_combine: dz += scale * (scale_dev ? *scale_dev : 1) * (2c/B^2) * ((rs - cs) * z - t1 + t2),  c = -2/dim^2,
 *                    with rs = row sums of K(z,z), cs = column sums of K(g,z), t1 = K(z,z) Z, t2 = K(g,z)^T G */
int mfm_rownorm2(int B, int dim, const float* x, long long ld, float* out, void* stream);
int mfm_mmd_kexp(int M, int N, float* S, const float* nx, const float* ny, int dim, float weight, float* slot, void* stream);
/* The same with a DOUBLE accumulator, and the fold of n such accumulators into fp32 slots.  The MMD is a difference of
 * three O(1) means that cancel to O(1/B); thousands of fp32 atomic adds into one slot do not hold the loss term to 1e-3
 * relative at batch 2048, the double accumulator does (the training schedule uses this pair). */
int mfm_mmd_kexp64(int M, int N, float* S, const float* nx, const float* ny, int dim, double weight, double* slot64, void* stream);
int mfm_mmd_fold(int n, const double* acc, float* slots, void* stream);
int mfm_mmd_combine(int B, int dim, const float* z, long long ldz, const float* rs, const float* cs, const float* t1,
                    const float* t2, float scale, const float* scale_dev, float* dz, long long lddz, void* stream);
/* out[i] ~ N(0,1), counter-based (Box-Muller over the library's hash RNG keyed by rng=[seed,step] and site):
 * the Gaussian sample of loss_MMD (mfm_model.py:26) generated on the device for the fused training path */
int mfm_randn(long long n, float* out, const long long* rng, int site, void* stream);

/* small data movement / reductions */
int mfm_copy2d(int M, int N, const float* src, long long lds, float* dst, long long ldd, int accumulate, void* stream);
int mfm_add(long long n, const float* a, const float* b, float* out, void* stream);
int mfm_zero(long long n, float* p, void* stream);
int mfm_colsum(int M, int N, const float* A, long long lda, float* out, void* stream);         /* out[n] += sum_m A[m,n] */
int mfm_relu_bwd(int M, int N, const float* dy, long long lddy, const float* y, long long ldy, float* out, long long ldo, void* stream);

/* loss heads (mfm_mosi.py:437-439; mfm_mosi_acc.py:450). slot is a device float that is ACCUMULATED. */
int mfm_mse_fwd_bwd(int M, int N, const float* xhat, long long ldxh, const float* x, long long ldx,
                    float loss_scale, float grad_scale, float* slot, float* dxhat, long long lddx, void* stream);
int mfm_l1_fwd_bwd(long long n, const float* yhat, const float* y, float scale, float* slot, float* dy, void* stream);
int mfm_ce_fwd_bwd(int B, int C, const float* yhat, const long long* y, float scale, float* slot, float* dy, void* stream);
/* loss_KLD of MFM_KL / MFM_KL_EF (mfm_model.py:36-38): *slot += -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) over [M,N];
 * its gradient: dmu += s * mu, dlogvar = s * 0.5 * (exp(logvar) - 1), s = scale * (scale_dev ? *scale_dev : 1). */
int mfm_kld_fwd(int M, int N, const float* mu, long long ldmu, const float* logvar, long long ldlv, float* slot, void* stream);
int mfm_kld_bwd(int M, int N, const float* mu, long long ldmu, const float* logvar, long long ldlv, float scale,
                const float* scale_dev, float* dmu, long long lddmu, float* dlogvar, long long lddlv, void* stream);
/* lb[8] = lb[0] + l0 lb[1] + l1 lb[2] + l2 lb[3] + lmmd (lb[4]+lb[5]+lb[6]+lb[7]) */
int mfm_loss_total(float* lb, float l0, float l1, float l2, float lmmd, void* stream);

/* torch.optim.Adam defaults (mfm_mosi.py:403) over one flat buffer.  state (device floats):
 * [0]=lr (host-written), [1]=step (incremented here), [2],[3] scratch.  g is scaled by grad_scale first. */
int mfm_adam_step(long long n, float* p, const float* g, float* m, float* v, float* state,
                  float grad_scale, double beta1, double beta2, double eps, void* stream);
/* rng[1] += 1 (per-step dropout stream) */
int mfm_rng_tick(long long* rng, void* stream);

#ifdef __cplusplus
}
#endif
#endif
