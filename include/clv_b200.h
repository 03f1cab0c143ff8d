/*
 * clv_b200.h -- C-ABI of the B200-native CL-VAE / CL-VRNN hot path (libclv_b200.so).
 *
 * The reference (mobeets/classifying-vae-lstm) has NO plugin / FFI interface: its hot path is the
 * Keras graph built by get_model (code/cl_vrnn/model.py:164-267, code/cl_vae/model.py:130-224),
 * executed by model.fit / model.predict inside TensorFlow, the optimizer update of
 * code/utils/weightnorm.py:75-178 and the Python sampling loops generate_sample
 * (code/cl_vrnn/model.py:9-60, code/cl_vae/model.py:9-42).  This header is the seam a maintainer
 * binds instead (ctypes stub in INTEGRATION.md); each entry point names the reference lines it
 * replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *     the library never allocates or frees device memory and never synchronises (clv_fp32_peak_probe, a
 *     diagnostic, excepted).  Process-wide state is limited to: the auxiliary streams / events that
 *     clv_runtime_init creates per device, a per-device "kernel attributes set" flag per kernel, the
 *     diagnostic launch counter (atomic) and a thread-local launch-attribute switch -- nothing that depends
 *     on a model or a call;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered (graph-capturable);
 *   - return value: 0 = CLV_OK, negative = CLV_E_*; clv_error_string() names it;
 *   - matrices are row-major fp32 in Keras [in,out] layout; piano-rolls are uint8 {0,1}
 *     [n_frames, D] with per-sequence first-frame offsets (int32) -- windows are gathered, never
 *     materialised;
 *   - there is no CPU fallback anywhere: without a CUDA device every compute call fails.
 */
#ifndef CLV_B200_H
#define CLV_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CLV_OK 0
#define CLV_E_INVALID (-1)      /* bad argument (null pointer, size out of range)            */
#define CLV_E_UNSUPPORTED (-2)  /* shape outside what the kernels are built for              */
#define CLV_E_CUDA (-3)         /* a CUDA runtime call / launch failed (see cudaGetLastError) */
#define CLV_E_WORKSPACE (-4)    /* workspace too small                                       */

int clv_version(void);
int64_t clv_launch_count(void); /* diagnostic: kernels launched by this library so far */
/* Once per process and device, OUTSIDE stream capture: creates the auxiliary stream + events used
 * when clv_cfg.overlap_wgrad is set.  Optional: without it everything runs on `stream`. */
int clv_runtime_init(void);
const char* clv_error_string(int code);

/* ---------------------------------------------------------------- model description ------- */
/* Mirrors the arguments of get_model (cl_vrnn/model.py:164; cl_vae/model.py:130-134). */
typedef struct clv_cfg {
  int32_t model;        /* 0 = CL-VRNN, 1 = CL-VAE                                            */
  int32_t B;            /* sequences (VRNN) / frames (VAE) in THIS call (local batch)         */
  int32_t B_global;     /* batch the means are taken over (== B on one GPU; sum over ranks)    */
  int32_t L;            /* seq_length (VRNN); 1 for VAE                                       */
  int32_t D;            /* original_dim (88)                                                  */
  int32_t H;            /* intermediate_dim: LSTM units (VRNN) / latent_dim_0 (VAE)           */
  int32_t Hc;           /* VAE only: class_dim_0 (intermediate_class_dim)                     */
  int32_t Z;            /* latent_dim                                                         */
  int32_t C;            /* n_classes                                                          */
  int32_t use_x_prev;   /* --use_x_prev                                                       */
  float class_weight, kl_weight, w_kl_weight, w_log_var_prior;
  int32_t gen_noise;    /* 1: draw eps_w/eps_z in-kernel (Philox) ; 0: read the eps buffers   */
  int32_t do_backward;  /* 0: forward + losses only (validation pass)                         */
  int32_t accumulate;   /* 1: add into grads/losses (micro-batching) instead of zeroing first */
  int32_t gemm_algo;    /* 0: exact fp32 SIMT GEMMs ; 1: tcgen05 tensor-core GEMMs where built */
  int32_t x_shift;      /* frame of `current` inside a window; 0 = default (1 with use_x_prev
                           [history = frames 0..L-1, current = 1..L], else 0).  L for windows
                           stored as [history | current] when the two inputs do not overlap     */
  int32_t gemm_algo_tc_lstm_min; /* batch from which the tcgen05 recurrence is used (0 = default 16384: 128 CTAs of 128 rows; measured: at 12 288 the register-resident FFMA kernels still win, 5.06 vs 5.28 ms/step) */
  int32_t overlap_wgrad;/* 1: run the weight-gradient GEMMs on the library's auxiliary stream
                           (needs clv_runtime_init()); forked from / joined into `stream`       */
  int32_t y_shift;      /* frame of the reconstruction TARGET inside a window; 0 = the target is
                           `current` (y == x, cl_vrnn/train.py:51-57).  --predict_next: windows of L+1
                           frames, current = frames 0..L-1 (x_shift 0), target = frames 1..L (y_shift 1) */
  int32_t pair_bwd;     /* 1: both BPTTs as ONE wavefront launch (clv_lstm_pair_bwd).  Off by default: at B = 200
                           it is 35.5 us against 18.7 + 19.3 us for the two launches, but the decoder-side weight
                           gradients and Adam-WN range then no longer overlap the encoder BPTT (0.147 vs 0.137 ms/step) */
  uint64_t seed;        /* Philox key for gen_noise (make it rank-dependent)                  */
} clv_cfg;

/* Flat parameter buffer layout = Keras weighted-layer order, so RUN.h5 I/O is one memcpy per
 * tensor (cl_vrnn: hW, Wargs, encoder_h{kernel,recurrent_kernel,bias}, Z_mean, Z_log_var,
 * decoder_h{..}, X_decoded_mean -- 16 tensors; cl_vae: h_w, w_mean, w_log_var, h, z_mean,
 * z_log_var, decoder_h, x_decoded_mean -- 16 tensors).  offs/rows/cols have 16 entries; a bias
 * has rows == 0.  Returns the parameter count P (or a negative error). */
#define CLV_N_TENSORS 16
int64_t clv_param_layout(const clv_cfg* cfg, int64_t* offs, int32_t* rows, int32_t* cols);

/* ---------------------------------------------------------------- K1: GEMM ---------------- */
/* C[M,N] = act( (accumulate? C : 0) + op(A)[M,K] @ op(B)[K,N] + bias[N] + rowadd[m/ra_grp, N] ) * mask
 * Replaces every Dense / TimeDistributed(Dense) / LSTM input projection MatMul and, through the
 * transposed forms, their TF-autodiff dgrad / wgrad (cl_vrnn/model.py:174-175,196-209,225-234;
 * cl_vae/model.py:141-143,162-167,182-186).  A may be fp32 or a uint8 piano-roll gathered by
 * per-sequence frame offsets. */
typedef struct clv_gemm_args {
  int32_t M, N, K;
  const void* A; int64_t lda;   /* A(m,k): a_kmajor ? A[row(m)*lda + k] : A[row(k)*lda + m]   */
  int32_t a_u8;                 /* A elements are uint8                                        */
  int32_t a_kmajor;             /* 1: rows indexed by m (NN / NT); 0: rows indexed by k (TN)    */
  const int32_t* a_off;         /* optional gather: row(r) = a_off[r / a_grp] + a_shift + r % a_grp */
  int32_t a_grp, a_shift;
  int32_t a_row_delta, a_skip_grp; /* TN only: use row r + delta; rows with r % skip_grp == 0 are zero */
  const float* Bm; int64_t ldb; /* B(k,n): b_nmajor ? B[k*ldb + n] : B[n*ldb + k]              */
  int32_t b_nmajor;
  float* C; int64_t ldc;
  const float* bias;            /* [N] or null                                                 */
  const float* rowadd; int64_t ldra; int32_t ra_grp; /* [M/ra_grp, N] or null                  */
  const float* relu_mask; int64_t ldmask; /* multiply by (mask[m,n] > 0) (ReLU backward) or null */
  int32_t relu;                 /* apply max(.,0)                                              */
  int32_t accumulate;           /* C += ...                                                    */
  int32_t split_k;              /* >1: split the K range over grid.z, atomicAdd into C (C must
                                   already hold the value to add to; no bias/act allowed)       */
} clv_gemm_args;
int clv_gemm(const clv_gemm_args* args, void* stream);
/* Tensor-core form of the hoisted LSTM input projection (tcgen05.mma + TMEM + bulk-TMA weight loads):
 * C[M,N] = roll_u8 rows @ W[D,N] (+ rowadd[m/ra_grp,:]), N = 4H = 352, D <= 96, D % 8 == 0.
 * fp32-exact: the roll is {0,1} (exact in bf16) and W is split into bf16 hi+mid+lo, accumulated in
 * fp32 in TMEM.  `scratch` (clv_inproj_tc_scratch_bytes(), 16-byte aligned) receives the split weight
 * image, rebuilt on every call because the weights change every step. */
int64_t clv_inproj_tc_scratch_bytes(void);
int clv_inproj_tc(const uint8_t* roll, const int32_t* win_off, int32_t grp, int32_t shift, int32_t D,
                  const float* W, int64_t ldw, int32_t N, void* scratch, float* C, int64_t ldc,
                  int64_t M, const float* rowadd, int64_t ldra, int32_t ra_grp, void* stream);
/* Tensor-core form of the LSTM weight gradients that reduce over all R = B*L rows (one launch):
 *   gKx[D,4H] += X^T @ dA,  gU[H,4H] += Hprev^T @ dA (Hprev[b,t] = h[b,t-1], 0 at t = 0),
 *   gKz[Z,4H] += Zs^T @ dA  (gKx / Zs+gKz optional).  tcgen05.mma with MN-major bf16 operands built
 * in shared memory (roll exact; dA, h, Zs split hi+mid), accumulators in TMEM, rows split over CTAs,
 * partials added with red.add (the gradient buffer must hold the value to add to).  H = 88, Z <= 8. */
int clv_lstm_wgrad_tc(const float* dA, const uint8_t* roll, const int32_t* win_off, int32_t L,
                      int32_t shift, int32_t D, const float* h, const float* Zs, int32_t Z, float* gKx,
                      float* gU, float* gKz, int64_t R, int32_t H, void* stream);
/* C[M,N] = act(C + bias[N])  -- epilogue of a split-K forward GEMM; relu: 0 = none, 1 = ReLU,
 * 2 = sigmoid (the X_decoded_mean head of the sampler sub-models' predict()) */
int clv_bias_act(float* C, int64_t ldc, int32_t M, int32_t N, const float* bias, int32_t relu,
                 void* stream);
/* out[N] (+)= sum_m A[m,N]  (bias gradients) */
int clv_colsum(const float* A, int64_t lda, int32_t M, int32_t N, float* out, int32_t accumulate,
               void* stream);

/* ---------------------------------------------------------------- K2: logistic-normal ----- */
/* sampling_w + W2 + w_kl_loss + w_rec_loss + accuracy (cl_vrnn/model.py:183-191,244-255,264;
 * cl_vae/model.py:146-157,198-208).  Wargs[B,2(C-1)] = [W_mean | W_log_var] (ld = ldwa; the VAE
 * keeps the two heads in one buffer too), eps_w[B,C-1], labels[B] (argmax of the one-hot key).
 * Writes W[B,C]; adds scale_b * sum_b {w_kl, w_rec, correct} into loss_acc[1], [2], [4]. */
int clv_logitnormal_fwd(const float* Wargs, int64_t ldwa, float* eps_w, const int32_t* labels,
                        float* W, float* loss_acc, int32_t B, int32_t C, float w_log_var_prior,
                        float scale_b, int32_t gen_noise, uint64_t seed, const uint64_t* ctr,
                        void* stream);
/* dWargs[B,2(C-1)] from dW_ext[B,C] (gradient reaching W from its consumers) plus the w_rec and
 * w_kl terms; cw_over_B = class_weight/B_global, wkl_over_B = w_kl_weight/B_global. */
int clv_logitnormal_bwd(const float* Wargs, int64_t ldwa, const float* eps_w, const int32_t* labels,
                        const float* W, const float* dW_ext, float* dWargs, int32_t B, int32_t C,
                        float w_log_var_prior, float cw_over_B, float wkl_over_B, void* stream);

/* ---------------------------------------------------------------- K2b: Gaussian heads ----- */
/* Z_mean / Z_log_var Dense heads + sampling + kl_loss in one pass over h (cl_vrnn/model.py:200-216,
 * 236-239; cl_vae/model.py:165-174,193-196).  h[R,H]; Km,Kv[H,Z]; bm,bv[Z]; eps[R,Z].
 * Writes Zargs[R,2Z] = [mu | log_var], Zs[R,Z]; adds scale * sum_r kl into loss_acc[3]. */
int clv_gauss_heads_fwd(const float* h, const float* Km, const float* bm, const float* Kv,
                        const float* bv, float* eps, float* Zargs, float* Zs, float* loss_acc,
                        int64_t R, int32_t H, int32_t Z, float scale, int32_t gen_noise,
                        uint64_t seed, const uint64_t* ctr, void* stream);
/* Backward: dZ[R,Z] -> dh[R,H] (overwritten; null = weight gradients only; multiplied by [h>0] when relu_input, the CL-VAE case
 * where h is a ReLU output) and atomically accumulated dKm,dbm,dKv,dbv.
 * klw_scale = kl_weight / (B_global*L). */
int clv_gauss_heads_bwd(const float* h, const float* Km, const float* Kv, const float* eps,
                        const float* Zargs, const float* dZ, float* dh, float* dKm, float* dbm,
                        float* dKv, float* dbv, int64_t R, int32_t H, int32_t Z, float klw_scale,
                        int32_t relu_input, void* stream);

/* ---------------------------------------------------------------- K3: persistent LSTM ----- */
/* Keras-2.0.0 LSTM recurrence (cl_vrnn/model.py:196-199,225-228): hard-sigmoid gates i,f,o, tanh
 * candidate, gate order i,f,c,o.  `gates` holds the hoisted input projection x@kernel+bias
 * [B,L,4H] on entry and the ACTIVATED gates on exit (the stash for BPTT).  U = recurrent_kernel
 * [H,4H] stays in registers for all L steps.  h0/c0 may be null (zeros). */
int clv_lstm_fwd(float* gates, const float* U, float* h, float* c, const float* h0,
                 const float* c0, int32_t B, int32_t L, int32_t H, void* stream);
/* BPTT: dh_out[B,L,H] is dLoss/dh_t from the layers above.  `gates` holds the activated gates on
 * entry and dLoss/d(pre-activation) [B,L,4H] on exit; dAsum[B,4H] = its sum over t. */
int clv_lstm_bwd(float* gates, const float* U, const float* h, const float* c, const float* dh_out,
                 float* dAsum, int32_t B, int32_t L, int32_t H, void* stream);

/* Fused forms used by clv_train_step.  Forward: the input terms that are not a GEMM over the roll are
 * folded into the recurrence: a_t = (has_xproj ? gates[b,t,:] : 0) + bias + Wv[b,:] @ Ww (the
 * RepeatVector(W) columns; Ww = [C,4H] rows of the kernel) + Zs[b,t,:] @ Kz (the Z columns; Kz =
 * [Z,4H]) + h_{t-1} @ U.  Any of bias / Wv / Zs may be null.  Backward additionally emits
 * dZ[B,L,Z] = dA @ Kz^T and dW_ext[B,C] (+)= dAsum @ Ww^T (either may be null). */
int clv_lstm_fwd_fused(float* gates, int32_t has_xproj, const float* U, const float* bias,
                       const float* Wv, const float* Ww, int32_t C, const float* Zs, const float* Kz,
                       int32_t Z, float* h, float* c, int32_t B, int32_t L, int32_t H, void* stream);
int clv_lstm_bwd_fused(float* gates, const float* U, const float* c, const float* dh_out, float* dAsum,
                       const float* Ww, int32_t C, float* dW_ext, int32_t dW_accumulate,
                       const float* Kz, int32_t Z, float* dZ, int32_t B, int32_t L, int32_t H,
                       void* stream);

/* clv_lstm_bwd_fused with the Z-head exchange of the CL-VRNN (cl_vrnn/model.py:200-216,236-239) folded
 * into the two BPTT kernels so that no kernel sits between them:
 *   decoder call (dZargs_out, Zargs, eps_z non-null): besides dZ writes dZargs_out[B,L,2Z] =
 *     dLoss/d(Z_mean | Z_log_var) = backward of Z = mu + exp(lv/2) eps plus the kl term
 *     (klw_scale = kl_weight / (B_global L)), i.e. the first half of clv_gauss_heads_bwd;
 *   encoder call (dZargs_in, Kzm, Kzv non-null, Zh <= 2): dh_out may be null;
 *     dLoss/dh[b,t,:] (+)= dZargs_in[b,t,:] @ [Kzm | Kzv]^T per cell, i.e. its second half.
 * The head weight gradients stay with clv_gauss_heads_bwd (dh = null: weight gradients only). */
int clv_lstm_bwd_heads(float* gates, const float* U, const float* c, const float* dh_out, float* dAsum,
                       const float* Ww, int32_t C, float* dW_ext, int32_t dW_accumulate,
                       const float* Kz, int32_t Z, float* dZ, const float* Zargs, const float* eps_z,
                       float klw_scale, float* dZargs_out, const float* dZargs_in, const float* Kzm,
                       const float* Kzv, int32_t Zh, int32_t B, int32_t L, int32_t H, void* stream);

/* Encoder LSTM + Z heads / sampling / z-KL + decoder LSTM of one CL-VRNN forward pass
 * (cl_vrnn/model.py:193-228,236-239) as ONE wavefront launch: CTA 2p = encoder of row group p (4 sequences),
 * CTA 2p+1 = its decoder one or two steps behind; the decoder's helper warp computes Z_t from the encoder's
 * h rows, which it polls through L2 until they differ from the 0xFFFFFFFF pattern the CALLER fills h_e with
 * before the launch (no flags, no fences).  Arguments as clv_lstm_fwd_fused for both LSTMs (gates_*: hoisted
 * projection in, activated gates out; K*_w: [C,4H] rows that multiply W; Kd_z: [Z,4H]) and as
 * clv_gauss_heads_fwd for the heads (kl_scale = 1/(B_global L)).  H = 88, Z <= 2, C <= 16.
 * The recurrence uses tanh through ex2.approx/rcp.approx (abs. error < 5e-7). */
int clv_lstm_pair_fwd(float* gates_e, const float* Ue, const float* be, const float* Ke_w, float* h_e,
                      float* c_e, float* gates_d, int32_t has_xproj_d, const float* Ud, const float* bd,
                      const float* Kd_w, const float* Kd_z, float* h_d, float* c_d, const float* Wv, int32_t C,
                      const float* Kzm, const float* bzm, const float* Kzv, const float* bzv, float* eps_z,
                      float* Zargs, float* Zs, float* loss_acc, float kl_scale, int32_t gen_noise,
                      uint64_t seed, const uint64_t* ctr, int32_t B, int32_t L, int32_t H, int32_t Z,
                      void* stream);

/* Decoder BPTT + Z-head exchange + encoder BPTT of one CL-VRNN backward pass as one wavefront launch (the
 * mirror of clv_lstm_pair_fwd; replaces the two clv_lstm_bwd_heads calls): CTA 2p = decoder BPTT of row group
 * p, CTA 2p+1 = encoder BPTT one or two steps behind, fed with dLoss/d(Z_mean | Z_log_var) through dZargs,
 * which its helper warp polls until the rows differ from the 0xFFFFFFFF pattern the CALLER fills dZargs with.
 * dW_ext[B,C] must be ZEROED by the caller: both BPTTs add dAsum @ K_w^T to it atomically.  Arguments as
 * clv_lstm_bwd_heads (gates_*: activated gates in, dLoss/d(pre-activation) out).  H = 88, Z <= 2, C <= 16. */
int clv_lstm_pair_bwd(float* gates_d, const float* Ud, const float* c_d, const float* dh_d, float* dAsum_d,
                      const float* Kd_w, const float* Kd_z, float* dZ, const float* Zargs, const float* eps_z,
                      float klw_scale, float* dZargs, float* gates_e, const float* Ue, const float* c_e,
                      float* dAsum_e, const float* Ke_w, const float* Kzm, const float* Kzv, float* dW_ext,
                      int32_t C, int32_t B, int32_t L, int32_t H, int32_t Z, void* stream);

/* Tensor-core form of the forward recurrence for large batches: 128 rows per CTA, h_{t-1} @ U on
 * tcgen05 (fp16 hi+lo splits of both operands, 3 products, fp32 accumulate in TMEM), cell state in
 * TMEM, U resident in shared memory.  `gates` must already hold x@kernel + bias + W term (use
 * clv_inproj_tc with rowadd); the optional Z term is added in the epilogue.  scratch:
 * clv_lstm_fwd_tc_scratch_bytes(), 16-byte aligned. */
int64_t clv_lstm_fwd_tc_scratch_bytes(void);
int clv_lstm_fwd_tc(float* gates, const float* U, const float* Zs, const float* Kz, int32_t Z, float* h,
                    float* c, void* scratch, int32_t B, int32_t L, int32_t H, void* stream);

/* ---------------------------------------------------------------- K4: Bernoulli loss ------ */
/* vae_loss = 88*mean_k Keras-BCE with clip->logit semantics, fused with its backward
 * (cl_vrnn/model.py:241-242; cl_vae/model.py:190-191).  logits[R,D] are overwritten with
 * dLoss/dlogits = scale*(p-x)*[eps<=p<=1-eps] when do_backward; x is row (x_off[r/x_grp] +
 * x_shift + r%x_grp) of the uint8 roll.  Adds scale * sum_r loss_r into loss_acc[0]. */
int clv_bernoulli_ce_fwd_bwd(float* logits, const uint8_t* roll, const int32_t* x_off,
                             int32_t x_grp, int32_t x_shift, float* loss_acc, int64_t R, int32_t D,
                             float scale, int32_t do_backward, void* stream);

/* K1+K4 fused for the sigmoid head (H = D = 88): logits = h @ Kx + bx, the Bernoulli loss above,
 * dlogits[R,D] and dh[R,H] = dlogits @ Kx^T in ONE pass over h with Kx resident in shared memory
 * (cl_vrnn/model.py:229-234,241-242 and their backward). */
int clv_xhead_fwd_bwd(const float* h, const float* Kx, const float* bx, const uint8_t* roll,
                      const int32_t* x_off, int32_t x_grp, int32_t x_shift, float* loss_acc,
                      float* dlogits, float* dh, int64_t R, int32_t H, int32_t D, float scale,
                      int32_t do_backward, void* stream);

/* The same head on 5th-gen tensor cores for large batches (R >= 128 rows per SM in clv_train_step): two
 * chained tcgen05 GEMMs per 128-row tile (logits = h @ Kx in TMEM -> loss / dlogits epilogue -> dh = dlogits @
 * Kx^T in TMEM), fp32 operands as bf16 hi + mid splits (three products, ~2^-16 relative).  Always computes the
 * backward.  gKx[H,D] / gbx[D] (both or neither; pre-zeroed or accumulating): the head's weight and bias gradients
 * h^T @ dlogits and colsum(dlogits), a third GEMM over the two tiles already in shared memory, added with red.add.  scratch: clv_xhead_tc_scratch_bytes() bytes, 16-byte aligned (split weight images, rebuilt per
 * call).  (cl_vrnn/model.py:229-234,241-242 and their backward) */
int64_t clv_xhead_tc_scratch_bytes(void);
int clv_xhead_tc(const float* h, const float* Kx, const float* bx, const uint8_t* roll, const int32_t* x_off,
                 int32_t x_grp, int32_t x_shift, float* loss_acc, float* dlogits, float* dh, float* gKx, float* gbx,
                 void* scratch, int64_t R, int32_t H, int32_t D, float scale, void* stream);

/* ---------------------------------------------------------------- key encoder (fused) ------ */
/* hW = relu(flat(window) @ Khw + b) as a gather-sum over the SET keys of the binary window (exact
 * fp32), Wargs = hW @ Kwa + b, then K2 forward -- one CTA per sequence (cl_vrnn/model.py:174-191,
 * 244-255,264).  Needs L*D <= 65535 and D % 4 == 0 (clv_train_step falls back to K1+K2 otherwise). */
int clv_keyenc_fwd(const uint8_t* roll, const int32_t* win_off, int32_t shift, int32_t L, int32_t D,
                   const float* Khw, const float* bhw, const float* Kwa, const float* bwa, float* eps_w,
                   const int32_t* labels, float* hW, float* Wargs, float* W, float* loss_acc, int32_t B,
                   int32_t C, float w_log_var_prior, float scale_b, int32_t gen_noise, uint64_t seed,
                   const uint64_t* ctr, void* stream);
/* K2 backward + dgrad through Wargs and the ReLU of hW: dWargs[B,2(C-1)], dhW[B,D]. */
int clv_keyenc_bwd(const float* Wargs, const float* eps_w, const int32_t* labels, const float* W,
                   const float* dW_ext, const float* Kwa, const float* hW, float* dWargs, float* dhW,
                   int32_t B, int32_t C, int32_t D, float w_log_var_prior, float cw_over_B,
                   float wkl_over_B, void* stream);

/* clv_keyenc_bwd plus ALL key-encoder weight gradients in the same kernel (one CTA per sequence):
 * gKwa += hW^T (x) dWargs, gbwa, gbhw, and gKhw[p,:] += dhW for every SET key p of the window -- the
 * sparse-scatter transpose of the forward gather-sum (red.add into pre-zeroed gradients). */
int clv_keyenc_bwd_full(const uint8_t* roll, const int32_t* win_off, int32_t shift, int32_t L, int32_t D,
                        const float* Wargs, const float* eps_w, const int32_t* labels, const float* W,
                        const float* dW_ext, const float* Kwa, const float* hW, float* dWargs, float* dhW,
                        float* gKhw, float* gbhw, float* gKwa, float* gbwa, int32_t B, int32_t C,
                        float w_log_var_prior, float cw_over_B, float wkl_over_B, void* stream);

/* ---------------------------------------------------------------- K6: Adam + weight-norm -- */
/* AdamWithWeightnorm.get_updates (utils/weightnorm.py:75-143,146-178) on the flat buffers.
 * state = [ m (P) | v (P) | V_scaler (ncols) | m_g (ncols) | v_g (ncols) | cached bias-correction
 * factor (2) | iterations (1) | retired-block counter (1) ], ncols =
 * total number of matrix columns; clv_adamwn_state_floats gives its size, clv_adamwn_init fills
 * V_scaler with ones and the rest with zeros.  grad_scale multiplies the gradient (1/N after a
 * sum-allreduce is already folded into B_global, so normally 1). */
int64_t clv_adamwn_state_floats(const clv_cfg* cfg);
int clv_adamwn_init(const clv_cfg* cfg, float* state, void* stream);
int clv_adamwn_step(const clv_cfg* cfg, float* params, const float* grads, float* state, double lr,
                    double beta_1, double beta_2, double epsilon, double grad_scale, int32_t weightnorm,
                    void* stream);
/* The same update restricted to the tensors [t_first, t_last) of clv_param_layout's order (weight-norm
 * is per tensor column, so ranges are independent).  `advance` != 0 on exactly one -- the last -- call
 * of a step: it increments `iterations` when its last block retires; all ranges of a step must be
 * issued before that call completes.  loss_mirror (nullable, used by the advancing call): the 8 loss
 * scalars stored behind the gradients (grads[P..P+8), the [grads | losses] buffer) are copied there --
 * pass host-mapped pinned memory and the host needs only a stream sync to read the step's losses. */
int clv_adamwn_step_range(const clv_cfg* cfg, float* params, const float* grads, float* state, double lr,
                          double beta_1, double beta_2, double epsilon, double grad_scale,
                          int32_t weightnorm, int32_t t_first, int32_t t_last, int32_t advance,
                          float* loss_mirror, void* stream);

/* Data-parallel form: gradient all-reduce FUSED into the optimizer over NVLink peer memory.
 * peer_grads = device array of n_peers pointers to every rank's [grads(P) | losses(8)] buffer in
 * symmetric (peer-mapped) memory, in rank order, own buffer included.  The caller places a
 * cross-GPU barrier before the call (all gradients written) and after it (all peers done reading).
 * Each rank sums the peers in rank order (bitwise identical everywhere), keeps the reduced gradient
 * in gsum[P] for the second pass and writes the reduced loss scalars to loss_out[8]. */
int clv_adamwn_step_p2p(const clv_cfg* cfg, float* params, const float* const* peer_grads,
                        int32_t n_peers, float* gsum, float* loss_out, float* state, double lr,
                        double beta_1, double beta_2, double epsilon, int32_t weightnorm, void* stream);

/* ---------------------------------------------------------------- fused training step ----- */
/* zero loss_acc[8] (zero_losses) and advance the device RNG call counter (bump). */
int clv_step_begin(float* loss_acc, uint64_t* rng_ctr, int32_t zero_losses, int32_t bump,
                   void* stream);
/* One train (or validation) step of model.fit's train_function (cl_vrnn/train.py:66-71):
 * forward, the four losses + accuracy, backward into grads[P].  loss_acc[8] =
 * {vae, w_kl, w_rec, z_kl, acc, -, -, -} already divided by B_global (sum over ranks gives the
 * Keras means).  roll / win_off describe the batch: window b starts at frame win_off[b]; with
 * use_x_prev the window is L+1 frames (history = frames 0..L-1, current = 1..L), otherwise L.
 * eps_w[B,C-1], eps_z[B*L,Z] are read (gen_noise=0) or written (gen_noise=1).
 * rng_ctr: device uint64 incremented once per call when gen_noise=1. */
int64_t clv_workspace_bytes(const clv_cfg* cfg);
int clv_train_step(const clv_cfg* cfg, const float* params, float* grads, float* loss_acc,
                   const uint8_t* roll, const int32_t* win_off, const int32_t* labels,
                   float* eps_w, float* eps_z, uint64_t* rng_ctr, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* clv_train_step followed by the optimizer update of the same step, as ONE schedule (the reference's
 * train_function = gradients + AdamWithWeightnorm updates in one session.run, cl_vrnn/train.py:66-71
 * + utils/weightnorm.py:75-143).  For the single-GPU case, where nothing sits between backward and
 * update: Adam-WN is launched per tensor range as soon as that range's gradients are final and no
 * later kernel of the step reads those parameters, so most of the update overlaps the encoder BPTT
 * and the weight-gradient kernels.  Result identical to clv_train_step + clv_adamwn_step.
 * Requires do_backward=1, accumulate=0. */
/* Peer-memory data parallelism with the hand-shake INSIDE the kernels (no collective launch, no host-side
 * barrier): every rank's [grads | losses] buffer and a small flag block live in peer-mapped (symmetric)
 * memory.  When a gradient bucket is final a one-warp kernel stores the step number into every peer's flag
 * block over NVLink (clv_p2p_signal); the Adam-WN kernel of that bucket polls its OWN flag block until all
 * peers have published, then reads the peers' gradients directly over NVLink, sums them in rank order
 * (bitwise identical on every rank) and applies the update -- all-reduce and optimizer are one kernel.  The
 * last kernel of a step tells the peers it has stopped reading (done flags); clv_p2p_wait_done at the start of
 * the next step keeps a rank from overwriting gradients a slower peer still reads.  Step numbers come from
 * the Adam state's `iterations`, so CUDA-graph replays need no host-side counters. */
typedef struct clv_p2p_args {
  const float* const* peer_grads;  /* device array [n_peers]: every rank's [grads(P) | losses(8)] buffer, rank order */
  int32_t* const* peer_flags;      /* device array [n_peers]: every rank's flag block, clv_p2p_flag_ints() int32,
                                      zero-initialised before the first step */
  int32_t n_peers, rank;
  float* gsum;                     /* local [P + 8]: the reduced [grads | losses] */
  float* loss_out;                 /* local [8]: reduced loss scalars (form 0: must be gsum + P) */
  int32_t form;                    /* 0: one-shot all-reduce kernel per bucket (clv_p2p_allreduce), then the ordinary
                                      Adam-WN update on gsum; 1: all-reduce fused into the Adam-WN kernels
                                      (clv_adamwn_step_range_p2p) */
  const float* mc_grads;           /* nullable: NVSwitch multicast address of the ranks' [grads | losses] buffers */
  float* mc_gsum;                  /* nullable: multicast address of the ranks' gsum buffers (gsum then lives in
                                      symmetric memory too).  Both set: clv_p2p_allreduce runs the two-shot form
                                      (multimem.ld_reduce of this rank's 1/N slice, multimem.st of the sum to all) */
} clv_p2p_args;
int clv_p2p_flag_ints(void);
int clv_p2p_signal(const clv_p2p_args* pp, const float* adam_state, const clv_cfg* cfg, int32_t slot, void* stream);
int clv_p2p_wait_done(const clv_p2p_args* pp, const float* adam_state, const clv_cfg* cfg, void* stream);
/* One-shot all-reduce of a gradient bucket over peer memory: gsum[first .. first+count) = sum over the ranks of
 * peer_grads[p][first .. first+count), summed in rank order (bit-identical on every rank).  The launch is
 * stream-ordered behind the producers of this rank's bucket; it publishes the bucket to the peers (flag slot
 * `slot`), waits for theirs inside the kernel (bounded: traps after ~2 minutes), and reads their buffers over NVLink.
 * last != 0 on the final bucket of a step: also raises this rank's "done reading" flag at every peer.
 * (the data-parallel exchange the reference does not have; replaces an NCCL all-reduce launch) */
int clv_p2p_allreduce(const clv_p2p_args* pp, const float* adam_state, const clv_cfg* cfg, int64_t first,
                      int64_t count, int32_t slot, int32_t last, void* stream);
/* Grid sizes, for callers that run the final launches of a step concurrently on two streams: `advance` of
 * clv_adamwn_step_range[_p2p] (and `last` of clv_p2p_allreduce) is 0 (not a final launch), 1 (the only final
 * launch) or the TOTAL number of blocks of the concurrent final launches -- the block that finishes last among
 * them advances `iterations` (raises the "done" flags). */
int clv_p2p_allreduce_blocks(int64_t count);
int clv_adamwn_range_blocks(const clv_cfg* cfg, int32_t weightnorm, int32_t t_first, int32_t t_last);
/* clv_adamwn_step_range on the sum of all peers' gradients; slot = the bucket's flag slot (0..3). */
int clv_adamwn_step_range_p2p(const clv_cfg* cfg, float* params, const clv_p2p_args* pp, float* state,
                              double lr, double beta_1, double beta_2, double epsilon, int32_t weightnorm,
                              int32_t t_first, int32_t t_last, int32_t slot, int32_t advance,
                              float* loss_mirror, void* stream);

/* Data-parallel hook: called by clv_train_step_opt on the HOST, while it enqueues the step, once per
 * gradient bucket at the point of the schedule where that bucket's gradients are final: sum-all-reduce
 * buf[0..count) over the ranks ON `stream` (stream-ordered; e.g. ncclAllReduce / torch.distributed.all_reduce
 * with `stream` current).  Currently ONE bucket, the whole [grads | 8 loss scalars] buffer after the last weight
 * gradient, followed by one Adam-WN update (a second NCCL bucket overlapped with the encoder BPTT measured
 * slower); callers must nevertheless honour (buf, count).  Return 0 on success. */
typedef int (*clv_exchange_fn)(void* user, float* buf, int64_t count, void* stream);
typedef struct clv_adam_args {
  float* state;                 /* clv_adamwn_state_floats() floats, initialised by clv_adamwn_init */
  double lr, beta_1, beta_2, epsilon, grad_scale;
  int32_t weightnorm;
  float* loss_mirror;           /* nullable: host-mapped copy of the step's 8 loss scalars (see above);
                                   requires loss_acc == grads + P (the [grads | losses] buffer) */
  clv_exchange_fn exchange;     /* nullable: gradient exchange between ranks (requires loss_acc == grads + P) */
  void* exchange_user;
  const clv_p2p_args* p2p;      /* nullable: peer-memory exchange fused into the Adam-WN kernels (takes
                                   precedence over `exchange`; grads must be this rank's peer-mapped buffer) */
} clv_adam_args;
int clv_train_step_opt(const clv_cfg* cfg, float* params, float* grads, float* loss_acc,
                       const uint8_t* roll, const int32_t* win_off, const int32_t* labels,
                       float* eps_w, float* eps_z, uint64_t* rng_ctr, void* workspace,
                       int64_t workspace_bytes, const clv_adam_args* opt, void* stream);
/* CL-VAE forward + the four losses (+ backward into grads) for one batch as ONE kernel: a CTA carries a tile
 * of 8 frames through all 8 Dense layers and back with every weight matrix resident in shared memory
 * (cl_vae/model.py:136-218 and its TF-autodiff backward); weight gradients are added to `grads` with red.add
 * (zero it first).  Wargs / W / Zargs are also written to the given workspace views.  Used by clv_train_step
 * for B <= 4096; returns CLV_E_UNSUPPORTED outside D, H, Hc <= 96, Z, C <= 16 (per-layer schedule then). */
int clv_vae_fused_step(const clv_cfg* cfg, const float* params, float* grads, float* loss_acc,
                       const uint8_t* roll, const int32_t* win_off, const int32_t* labels, float* eps_w,
                       float* eps_z, const uint64_t* rng_ctr, float* ws_Wargs, float* ws_W, float* ws_Zargs,
                       void* stream);
/* Named views into the workspace after a step (for parity tests): returns the float offset of
 * e.g. "W", "Zargs", "h_e", "h_d", "logits" or -1. */
int64_t clv_workspace_offset(const clv_cfg* cfg, const char* name);

/* ---------------------------------------------------------------- K5: samplers ------------ */
/* generate_sample (cl_vrnn/model.py:9-60) for S independent songs in one persistent kernel.
 * seed[S,T_seed,D] uint8 teacher-forces the first T_seed steps; w[S,C] is the key simplex (given
 * one-hot or inferred); enc_lstm = {kernel,recurrent_kernel,bias} of the z-encoder LSTM (quirk Q1:
 * pass params' encoder_h for the documented fix, or fresh weights for reference behaviour).
 * Noise: eps_z[S,T,Z], u[S,T,D] tapes (parity mode) or null => Philox(seed, song0+s, t).
 * out[S,T,D] uint8 (T = T_seed + nsteps; the host slices [T_seed:]); probs[S,T,D] optional. */
int clv_vrnn_sample(const clv_cfg* cfg, const float* params, const float* enc_kernel,
                    const float* enc_rkernel, const float* enc_bias, const uint8_t* seed_roll,
                    int32_t T_seed, int32_t nsteps, const float* w, const float* eps_z,
                    const float* u, uint64_t seed, int64_t song0, int32_t S, uint8_t* out,
                    float* probs, void* stream);
/* The same sampler with BIT-PACKED output: out_bits[S, T, ceil(D/8)] uint8, key d = bit (d % 8) of byte d / 8
 * (numpy.unpackbits(..., bitorder="little")) -- 11 bytes per 88-key frame instead of 88, the form in which
 * samples leave the GPU. */
int clv_vrnn_sample_bits(const clv_cfg* cfg, const float* params, const float* enc_kernel,
                         const float* enc_rkernel, const float* enc_bias, const uint8_t* seed_roll,
                         int32_t T_seed, int32_t nsteps, const float* w, const float* eps_z,
                         const float* u, uint64_t seed, int64_t song0, int32_t S, uint8_t* out_bits,
                         float* probs, void* stream);
/* generate_sample (cl_vae/model.py:9-42): x_seed[S,D]; optional use_z_prior. */
int clv_vae_sample(const clv_cfg* cfg, const float* params, const uint8_t* x_seed, int32_t nsteps,
                   const float* w, const float* eps_z, const float* u, uint64_t seed, int64_t song0,
                   int32_t S, int32_t use_z_prior, uint8_t* out, float* probs, void* stream);
/* mean over n_chunks consecutive rows: out[S,C] = mean_j in[s*n_chunks + j, C]
 * (key inference over seed chunks, cl_vrnn/model.py:34-41). */
int clv_chunk_mean(const float* in, float* out, int32_t S, int32_t n_chunks, int32_t C,
                   void* stream);

/* ---------------------------------------------------------------- diagnostics ------------ */
/* Measured fp32 FMA throughput of this device in TFLOP/s (bench.py's issue-bound roofline denominator;
 * MEASURED_PEAKS.json carries no fp32 figure).  mode 0: FMA-pipe peak (FFMA2, shared operands); mode 1: FFMA2
 * in the 5-distinct-register operand pattern of the register-resident LSTM mat-vecs (register-file bank
 * limited).  scratch: >= 2 * #SMs * 512 floats.  Unlike every other entry point it SYNCHRONISES. */
int clv_fp32_peak_probe(int32_t mode, float* scratch, int64_t scratch_floats, double* tflops_out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif
