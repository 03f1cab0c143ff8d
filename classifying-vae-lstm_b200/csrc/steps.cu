// Parameter layout, workspace carving and the fused training-step drivers: the sequence of kernel
// launches that replaces ONE tf.Session.run of Keras' train_function (cl_vrnn/train.py:66-71 ->
// graph of cl_vrnn/model.py:169-264; cl_vae/train.py:63-69 -> cl_vae/model.py:136-218).
// Host code only: no allocation, no synchronisation, everything stream-ordered (graph-capturable).
#include <string.h>
#include "common.cuh"
#include <cstdlib>

extern "C" int clv_lstm_fwd_fused(float*, int32_t, const float*, const float*, const float*, const float*,
                                  int32_t, const float*, const float*, int32_t, float*, float*, int32_t,
                                  int32_t, int32_t, void*);
extern "C" int clv_lstm_bwd_fused(float*, const float*, const float*, const float*, float*, const float*,
                                  int32_t, float*, int32_t, const float*, int32_t, float*, int32_t,
                                  int32_t, int32_t, void*);
extern "C" int clv_inproj_tc(const uint8_t*, const int32_t*, int32_t, int32_t, int32_t, const float*, int64_t,
                             int32_t, void*, float*, int64_t, int64_t, const float*, int64_t, int32_t, void*);
extern "C" int64_t clv_inproj_tc_scratch_bytes(void);
extern "C" int clv_lstm_wgrad_tc(const float*, const uint8_t*, const int32_t*, int32_t, int32_t, int32_t,
                                 const float*, const float*, int32_t, float*, float*, float*, int64_t,
                                 int32_t, void*);
extern "C" int64_t clv_lstm_fwd_tc_scratch_bytes(void);
extern "C" int clv_lstm_fwd_tc(float*, const float*, const float*, const float*, int32_t, float*, float*,
                               void*, int32_t, int32_t, int32_t, void*);
extern "C" int clv_xhead_fwd_bwd(const float*, const float*, const float*, const uint8_t*, const int32_t*,
                                 int32_t, int32_t, float*, float*, float*, int64_t, int32_t, int32_t,
                                 float, int32_t, void*);
extern "C" int64_t clv_xhead_tc_scratch_bytes(void);
extern "C" int clv_xhead_tc(const float*, const float*, const float*, const uint8_t*, const int32_t*, int32_t,
                            int32_t, float*, float*, float*, float*, float*, void*, int64_t, int32_t, int32_t, float,
                            void*);
extern "C" int clv_keyenc_fwd(const uint8_t*, const int32_t*, int32_t, int32_t, int32_t, const float*,
                              const float*, const float*, const float*, float*, const int32_t*, float*,
                              float*, float*, float*, int32_t, int32_t, float, float, int32_t, uint64_t,
                              const uint64_t*, void*);
extern "C" int clv_keyenc_bwd_full(const uint8_t*, const int32_t*, int32_t, int32_t, int32_t, const float*,
                                   const float*, const int32_t*, const float*, const float*, const float*,
                                   const float*, float*, float*, float*, float*, float*, float*, int32_t,
                                   int32_t, float, float, float, void*);
extern "C" int clv_keyenc_bwd(const float*, const float*, const int32_t*, const float*, const float*,
                              const float*, const float*, float*, float*, int32_t, int32_t, int32_t, float,
                              float, float, void*);
extern "C" int clv_vae_fused_step(const clv_cfg*, const float*, float*, float*, const uint8_t*, const int32_t*,
                                  const int32_t*, float*, float*, const uint64_t*, float*, float*, float*, void*);
extern "C" int clv_lstm_pair_bwd(float*, const float*, const float*, const float*, float*, const float*, const float*,
                                 float*, const float*, const float*, float, float*, float*, const float*, const float*,
                                 float*, const float*, const float*, const float*, float*, int32_t, int32_t, int32_t,
                                 int32_t, int32_t, void*);
extern "C" int clv_lstm_pair_fwd(float*, const float*, const float*, const float*, float*, float*, float*,
                                 int32_t, const float*, const float*, const float*, const float*, float*,
                                 float*, const float*, int32_t, const float*, const float*, const float*,
                                 const float*, float*, float*, float*, float*, float, int32_t, uint64_t,
                                 const uint64_t*, int32_t, int32_t, int32_t, int32_t, void*);

namespace {

// tensor indices in the flat buffer (Keras weighted-layer order)
enum { R_HW_K, R_HW_B, R_WA_K, R_WA_B, R_ENC_K, R_ENC_U, R_ENC_B, R_ZM_K, R_ZM_B, R_ZV_K, R_ZV_B,
       R_DEC_K, R_DEC_U, R_DEC_B, R_X_K, R_X_B };
enum { V_HW_K, V_HW_B, V_WM_K, V_WM_B, V_WV_K, V_WV_B, V_H_K, V_H_B, V_ZM_K, V_ZM_B, V_ZV_K, V_ZV_B,
       V_DH_K, V_DH_B, V_X_K, V_X_B };

int check_cfg(const clv_cfg* c) {
  if (!c) return CLV_E_INVALID;
  if (c->model != 0 && c->model != 1) return CLV_E_INVALID;
  if (c->B < 0 || c->B_global < c->B || c->L < 1 || c->D < 1 || c->H < 1 || c->Z < 1 || c->C < 2)
    return CLV_E_INVALID;
  if (c->C > 16 || c->Z > 16 || c->D > 128) return CLV_E_UNSUPPORTED;
  if (c->model == 0 && c->H != 88) return CLV_E_UNSUPPORTED;   // K3 register tiling is built for 88
  if (c->model == 1 && (c->L != 1 || c->H > 128 || c->Hc < 1)) return CLV_E_UNSUPPORTED;
  return CLV_OK;
}

struct WsItem { const char* name; int64_t off, n; };
struct Ws {
  WsItem it[32];
  int n = 0;
  int64_t total = 0;
  int64_t add(const char* name, int64_t count) {
    const int64_t off = total;
    it[n++] = {name, off, count};
    total += (count + 63) / 64 * 64;
    return off;
  }
  int64_t find(const char* name) const {
    for (int i = 0; i < n; ++i) if (!strcmp(it[i].name, name)) return it[i].off;
    return -1;
  }
};

Ws carve(const clv_cfg* c) {
  Ws w;
  const int64_t B = c->B, L = c->L, D = c->D, H = c->H, Z = c->Z, C = c->C, C1 = C - 1;
  if (c->model == 0) {
    const int64_t G = 4 * H, BL = B * L;
    w.add("hW", B * D); w.add("Wargs", B * 2 * C1); w.add("W", B * C);
    w.add("rb_e", B * G); w.add("gates_e", BL * G); w.add("h_e", BL * H); w.add("c_e", BL * H);
    w.add("Zargs", BL * 2 * Z); w.add("Zs", BL * Z);
    w.add("rb_d", B * G); w.add("gates_d", BL * G); w.add("h_d", BL * H); w.add("c_d", BL * H);
    w.add("logits", BL * D); w.add("dh", BL * H);
    w.add("dAsum_d", B * G); w.add("dAsum_e", B * G); w.add("dZ", BL * Z); w.add("dZa", BL * 2 * Z);
    w.add("dW_ext", B * C); w.add("dWargs", B * 2 * C1); w.add("dhW", B * D);
    w.add("wimg_e", clv_inproj_tc_scratch_bytes() / 4); w.add("wimg_d", clv_inproj_tc_scratch_bytes() / 4);
    w.add("uimg_e", clv_lstm_fwd_tc_scratch_bytes() / 4); w.add("uimg_d", clv_lstm_fwd_tc_scratch_bytes() / 4);
    w.add("ximg", clv_xhead_tc_scratch_bytes() / 4);
  } else {
    const int64_t Hc = c->Hc;
    w.add("h_w", B * Hc); w.add("Wargs", B * 2 * C1); w.add("W", B * C);
    w.add("h", B * H); w.add("Zargs", B * 2 * Z); w.add("Zs", B * Z);
    w.add("h_dec", B * H); w.add("logits", B * D);
    w.add("dpre", B * H); w.add("dh", B * H); w.add("dZ", B * Z);
    w.add("dW_ext", B * C); w.add("dWargs", B * 2 * C1); w.add("dh_w", B * Hc);
  }
  return w;
}

// ---- GEMM call helpers ---------------------------------------------------------------
clv_gemm_args gz() { clv_gemm_args a; memset(&a, 0, sizeof(a)); a.a_kmajor = 1; a.b_nmajor = 1; a.split_k = 1; return a; }

int pick_split(int M, int N, int K) {
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  int split = (2 * clv_num_sms() + tiles - 1) / tiles;
  const int maxs = (K + 63) / 64;
  if (split > maxs) split = maxs;
  return split < 1 ? 1 : split;
}

// C[M,N] = act(A_f32[M,K] @ B[K,N] + bias + rowadd)            (forward Dense)
int nn_f32(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M,
           int N, int K, const float* bias, int relu, int accumulate, cudaStream_t st) {
  clv_gemm_args a = gz();
  a.M = M; a.N = N; a.K = K; a.A = A; a.lda = lda; a.Bm = Bm; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.bias = bias; a.relu = relu; a.accumulate = accumulate;
  return clv_gemm(&a, st);
}
// C[M,N] = act(roll_u8 gather @ B + bias + rowadd[m/grp])       (forward Dense on piano-roll rows)
int nn_u8(const uint8_t* roll, const int32_t* off, int grp, int shift, int64_t lda, const float* Bm,
          int64_t ldb, float* C, int64_t ldc, int M, int N, int K, const float* bias,
          const float* rowadd, int64_t ldra, int ra_grp, int relu, int accumulate, cudaStream_t st) {
  clv_gemm_args a = gz();
  a.M = M; a.N = N; a.K = K; a.A = roll; a.lda = lda; a.a_u8 = 1; a.a_off = off; a.a_grp = grp;
  a.a_shift = shift; a.Bm = Bm; a.ldb = ldb; a.C = C; a.ldc = ldc; a.bias = bias; a.rowadd = rowadd;
  a.ldra = ldra; a.ra_grp = ra_grp; a.relu = relu; a.accumulate = accumulate;
  return clv_gemm(&a, st);
}
// dA[M,N] = (dC[M,K] @ W[N,K]^T) (* mask)                        (dgrad)
int nt_f32(const float* dC, int64_t ldd, const float* W, int64_t ldw, float* dA, int64_t lda, int M,
           int N, int K, const float* mask, int64_t ldmask, int accumulate, cudaStream_t st) {
  clv_gemm_args a = gz();
  a.M = M; a.N = N; a.K = K; a.A = dC; a.lda = ldd; a.Bm = W; a.ldb = ldw; a.b_nmajor = 0; a.C = dA;
  a.ldc = lda; a.relu_mask = mask; a.ldmask = ldmask; a.accumulate = accumulate;
  return clv_gemm(&a, st);
}
// dW[M,N] += A[K,M]^T @ dC[K,N]                                   (wgrad, fp32 A)
int tn_f32(const float* A, int64_t lda, const float* dC, int64_t ldd, float* dW, int64_t ldw, int M,
           int N, int K, int row_delta, int skip_grp, cudaStream_t st) {
  clv_gemm_args a = gz();
  a.M = M; a.N = N; a.K = K; a.A = A; a.lda = lda; a.a_kmajor = 0; a.a_row_delta = row_delta;
  a.a_skip_grp = skip_grp; a.Bm = dC; a.ldb = ldd; a.C = dW; a.ldc = ldw;
  a.split_k = pick_split(M, N, K);
  a.accumulate = 1;
  return clv_gemm(&a, st);
}
// dW[M,N] += roll_u8[K rows gathered, M]^T @ dC[K,N]              (wgrad, piano-roll A)
int tn_u8(const uint8_t* roll, const int32_t* off, int grp, int shift, int64_t lda, const float* dC,
          int64_t ldd, float* dW, int64_t ldw, int M, int N, int K, cudaStream_t st) {
  clv_gemm_args a = gz();
  a.M = M; a.N = N; a.K = K; a.A = roll; a.lda = lda; a.a_u8 = 1; a.a_kmajor = 0; a.a_off = off;
  a.a_grp = grp; a.a_shift = shift; a.Bm = dC; a.ldb = ldd; a.C = dW; a.ldc = ldw;
  a.split_k = pick_split(M, N, K);
  a.accumulate = 1;
  return clv_gemm(&a, st);
}

#define TRY(x) do { int rc__ = (x); if (rc__ != CLV_OK) return rc__; } while (0)
// launch whose stream predecessor is one of our kernels: allow programmatic dependent launch
// (the kernel's parameter-only prologue overlaps the predecessor; see common.cuh)
#define TRY_PDL(x) do { g_clv_pdl = pdl_enabled(); int rc__ = (x); g_clv_pdl = 0; if (rc__ != CLV_OK) return rc__; } while (0)
static int vae_fused_max_rows() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CLV_VAE_FUSED_MAX_ROWS"); v = e ? atoi(e) : 4096; }
  return v;
}
// rows from which the tensor-core X head replaces the SIMT one (one 128-row tile per SM; measured 0.032 vs 0.062
// ms at 32 768 rows, step 0.602 vs 0.635 ms at B = 1 024, L = 32); CLV_XHEAD_TC_MIN overrides
static int64_t xhead_tc_min_rows() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("CLV_XHEAD_TC_MIN"); v = e ? atoll(e) : 128LL * clv_num_sms(); }
  return v;
}
static int pair_disabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CLV_NO_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}
// profiling aid (profiles/skip_probe.sh): CLV_DIAG_SKIP=<bitmask> drops launches from the B=200 schedule so that
// differential timing shows which ones are on the critical path.  Results are WRONG with any bit set.
static int diag_skip() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CLV_DIAG_SKIP"); v = e ? atoi(e) : 0; }
  return v;
}
static int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CLV_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v;
}

// Auxiliary stream for the weight-gradient branch (off the critical path of the step).  Created by
// clv_runtime_init() OUTSIDE any stream capture; forked from / joined back into the caller's stream
// with events, so all work stays ordered on the caller's stream (and is captured with it).
constexpr int NAUX = 4;     // aux[0..2]: round-robin side work; aux[3]: optimizer launches
constexpr int NRR = 3;
struct SideStream {
  cudaStream_t aux[NAUX] = {};
  cudaEvent_t fork_ev[8] = {};
  cudaEvent_t join_ev[4][NAUX] = {};
  bool ready = false;
};
SideStream g_side[16];

static bool side_ready() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev < 16 && g_side[dev].ready;
}

// Independent side work (weight gradients, the decoder's roll projection) is spread round-robin
// over NAUX auxiliary streams so those launch-latency-bound kernels overlap each other as well as
// the critical path.
struct Fork {
  cudaStream_t main;
  SideStream* s;
  int n = 0, nj = 0, rr = 0;
  bool on;
  Fork(cudaStream_t st, bool want) : main(st), s(nullptr), on(false) {
    int dev = 0;
    if (want && cudaGetDevice(&dev) == cudaSuccess && dev < 16 && g_side[dev].ready) {
      s = &g_side[dev]; on = true;
    }
  }
  cudaStream_t next() {   // stream for the next independent side kernel
    if (!on) return main;
    rr = (rr + 1) % NRR;
    return s->aux[rr];
  }
  cudaStream_t opt_stream() { return on ? s->aux[NRR] : main; }
  int gather() {   // the optimizer stream waits for everything enqueued on the round-robin streams
    if (!on) return CLV_OK;
    for (int i = 0; i < NRR; ++i) {
      CLV_CUDA(cudaEventRecord(s->join_ev[nj][i], s->aux[i]));
      CLV_CUDA(cudaStreamWaitEvent(s->aux[NRR], s->join_ev[nj][i], 0));
    }
    nj = (nj + 1) % 4;
    return CLV_OK;
  }
  int fork() {   // side branches may now consume everything enqueued on main so far
    if (!on) return CLV_OK;
    CLV_CUDA(cudaEventRecord(s->fork_ev[n], main));
    for (int i = 0; i < NAUX; ++i) CLV_CUDA(cudaStreamWaitEvent(s->aux[i], s->fork_ev[n], 0));
    n = (n + 1) % 8;
    return CLV_OK;
  }
  int join() {   // main waits for everything enqueued on the side branches so far
    if (!on) return CLV_OK;
    for (int i = 0; i < NAUX; ++i) {
      CLV_CUDA(cudaEventRecord(s->join_ev[nj][i], s->aux[i]));
      CLV_CUDA(cudaStreamWaitEvent(main, s->join_ev[nj][i], 0));
    }
    nj = (nj + 1) % 4;
    return CLV_OK;
  }
};

// Adam-WN of the tensor range [t0, t1) on the peers' summed gradients (clv_p2p_args): either the one-shot
// all-reduce kernel over the range's elements followed by the ordinary update on gsum (form 0), or the update
// kernel that reads the peers itself (form 1).  The range [.., CLV_N_TENSORS) carries the 8 loss scalars.
static int adam_p2p(const clv_cfg* c, const float* P, const clv_adam_args* opt, int t0, int t1, int slot, int advance,
                    int last, bool mirror, cudaStream_t s_) {
  // advance / last: 0, 1 or the block totals of the step's concurrent final launches (clv_b200.h)
  const clv_p2p_args* pp = opt->p2p;
  float* lm = mirror ? opt->loss_mirror : nullptr;
  if (pp->form != 0)
    return clv_adamwn_step_range_p2p(c, const_cast<float*>(P), pp, opt->state, opt->lr, opt->beta_1, opt->beta_2,
                                     opt->epsilon, opt->weightnorm, t0, t1, slot, advance, lm, s_);
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  const int64_t Pn = clv_param_layout(c, po, pr, pc);
  const int64_t e0 = po[t0], e1 = t1 == CLV_N_TENSORS ? Pn + 8 : po[t1];
  if (pp->loss_out != pp->gsum + Pn) return CLV_E_INVALID;
  TRY(clv_p2p_allreduce(pp, opt->state, c, e0, e1 - e0, slot, last, s_));
  TRY_PDL(clv_adamwn_step_range(c, const_cast<float*>(P), pp->gsum, opt->state, opt->lr, opt->beta_1, opt->beta_2,
                                opt->epsilon, 1.0, opt->weightnorm, t0, t1, advance, lm, s_));
  return CLV_OK;
}

// opt != null: the optimizer is part of the schedule (single-GPU form, no exchange between backward
// and update).  Adam-WN runs per tensor range as soon as that range's gradients are complete and
// nothing later in the step reads those parameters: [decoder | X head] during the encoder BPTT,
// [key encoder] during the encoder weight gradients, [encoder LSTM | Z heads] last.
int vrnn_step(const clv_cfg* c, const float* P, float* Gr, float* loss, const uint8_t* roll,
              const int32_t* off, const int32_t* labels, float* eps_w, float* eps_z,
              uint64_t* ctr, float* ws, cudaStream_t st, const clv_adam_args* opt) {
  // Adam-WN of the tensor range [t0, t1).  Peer-memory data parallelism (opt->p2p): the range is a gradient
  // bucket -- the update kernel publishes it to the peers, waits for the peers' signals and sums their buffers
  // over NVLink (all-reduce fused into the optimizer, slot = bucket id)
  auto adam = [&](int t0, int t1, int advance, cudaStream_t s_, int slot, int last = -1, int mirror = -1) -> int {
    const bool mir = mirror < 0 ? advance != 0 : mirror != 0;
    if (opt->p2p) return adam_p2p(c, P, opt, t0, t1, slot, advance, last < 0 ? advance : last, mir, s_);
    return clv_adamwn_step_range(c, const_cast<float*>(P), Gr, opt->state, opt->lr, opt->beta_1, opt->beta_2,
                                 opt->epsilon, opt->grad_scale, opt->weightnorm, t0, t1, advance,
                                 mir ? opt->loss_mirror : nullptr, s_);
  };
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  clv_param_layout(c, po, pr, pc);
  const int B = c->B, L = c->L, D = c->D, H = c->H, Z = c->Z, C = c->C, C1 = C - 1, G = 4 * H;
  const int BL = B * L, xo = c->use_x_prev ? D : 0;
  const int sx = c->x_shift > 0 ? c->x_shift : (c->use_x_prev ? 1 : 0);
  const int sy = c->y_shift > 0 ? c->y_shift : sx;   // reconstruction target (--predict_next: next frame)
  const float sb = 1.0f / (float)c->B_global, sbl = 1.0f / ((float)c->B_global * (float)L);
  const Ws w = carve(c);
#define WSP(name) (ws + w.find(name))
  float *hW = WSP("hW"), *Wargs = WSP("Wargs"), *W = WSP("W"), *rb_e = WSP("rb_e"),
        *gates_e = WSP("gates_e"), *h_e = WSP("h_e"), *c_e = WSP("c_e"), *Zargs = WSP("Zargs"),
        *Zs = WSP("Zs"), *rb_d = WSP("rb_d"), *gates_d = WSP("gates_d"), *h_d = WSP("h_d"),
        *c_d = WSP("c_d"), *logits = WSP("logits"), *dh = WSP("dh"), *dAsum_d = WSP("dAsum_d"),
        *dAsum_e = WSP("dAsum_e"), *dZ = WSP("dZ"), *dZa = WSP("dZa"), *dW_ext = WSP("dW_ext"),
        *dWargs = WSP("dWargs"), *dhW = WSP("dhW"), *wimg_e = WSP("wimg_e"), *wimg_d = WSP("wimg_d"),
        *uimg_e = WSP("uimg_e"), *uimg_d = WSP("uimg_d"), *WSP_X = WSP("ximg");
#undef WSP
  // hoisted input projections on tensor cores (tcgen05) when asked for and the shape is the built one
  const bool tc = c->gemm_algo == 1 && G == 352 && D <= 96 && (D % 8) == 0;
  // tensor-core recurrence: wins once a CTA can be given 128 rows (large batches); below that the
  // register-resident FFMA kernel is latency-optimal
  // measured crossover on B200: 128-row CTAs are latency-bound (~30 us/step), so the chip must be
  // filled (>= 64 CTAs) before tensor cores beat the 2..8-row FFMA kernel
  const bool tcl = tc && c->use_x_prev && B >= (c->gemm_algo_tc_lstm_min > 0 ? c->gemm_algo_tc_lstm_min : 16384) && Z <= 16;
  const float *Khw = P + po[R_HW_K], *bhw = P + po[R_HW_B], *Kwa = P + po[R_WA_K],
              *bwa = P + po[R_WA_B], *Ke = P + po[R_ENC_K], *Ue = P + po[R_ENC_U],
              *be = P + po[R_ENC_B], *Kzm = P + po[R_ZM_K], *bzm = P + po[R_ZM_B],
              *Kzv = P + po[R_ZV_K], *bzv = P + po[R_ZV_B], *Kd = P + po[R_DEC_K],
              *Ud = P + po[R_DEC_U], *bd = P + po[R_DEC_B], *Kx = P + po[R_X_K], *bx = P + po[R_X_B];

  // encoder/decoder wavefront (lstm_pair.cu): both recurrences in one launch, the decoder a step or two
  // behind the encoder; needs the 2-latent head exchange and the FFMA recurrence
  // (one wave only: with more CTA pairs than SMs the 4-row CTAs lose to the 2..4-row kernels they replace --
  //  B = 1 024, L = 32: 203 us against 2 x 92 us, profiles/sweep_r2.jsonl)
  const bool pair = !tcl && Z <= 2 && H == 88 && C <= 16 && 2 * ((B + 3) / 4) <= clv_num_sms() && !pair_disabled();
  // zero the gradients; data parallel over peer memory: only once no peer still reads last step's (both off
  // the critical path when side streams exist: nothing writes a gradient before the first join)
  const bool side_zero = opt && opt->p2p && c->overlap_wgrad != 0 && side_ready() && c->do_backward && !c->accumulate;
  if (!side_zero) {
    if (opt && opt->p2p) TRY(clv_p2p_wait_done(opt->p2p, opt->state, c, st));
    if (c->do_backward && !c->accumulate)
      CLV_CUDA(cudaMemsetAsync(Gr, 0, sizeof(float) * (po[R_X_B] + pc[R_X_B]), st));
  }
  // wavefront hand-over: the consumer polls the producer's rows themselves; 0xFFFFFFFF = "not written yet"
  if (pair) CLV_CUDA(cudaMemsetAsync(h_e, 0xFF, sizeof(float) * (size_t)BL * H, st));
  const bool pairb = pair && c->do_backward && c->pair_bwd != 0;
  if (pairb) {
    CLV_CUDA(cudaMemsetAsync(dZa, 0xFF, sizeof(float) * (size_t)BL * 2 * Z, st));
    CLV_CUDA(cudaMemsetAsync(dW_ext, 0, sizeof(float) * (size_t)B * C, st));
  }
  TRY(clv_step_begin(loss, ctr, !c->accumulate, c->gen_noise, st));   // last: the key encoder chains on it

  Fork fk(st, c->overlap_wgrad != 0);
  if (side_zero) {
    TRY(fk.fork());
    TRY(clv_p2p_wait_done(opt->p2p, opt->state, c, fk.opt_stream()));
    CLV_CUDA(cudaMemsetAsync(Gr, 0, sizeof(float) * (po[R_X_B] + pc[R_X_B]), fk.opt_stream()));
  }
  const bool fused_ke = (int64_t)L * D <= 65535 && (D % 4) == 0;
  const float* Ke_w = Ke + (int64_t)D * G;            // rows of the kernels that multiply W
  const float* Kd_z = Kd + (int64_t)xo * G;           //                               ... Z
  const float* Kd_w = Kd + (int64_t)(xo + Z) * G;

  // ---- both roll projections depend on nothing but the batch -> side streams, concurrent with the
  //      key encoder (the encoder LSTM joins them)
  const bool enc_proj_side = !tcl && c->overlap_wgrad != 0;
  if (!tcl && (c->use_x_prev || enc_proj_side)) {
    TRY(fk.fork());
    if (enc_proj_side) {
      if (tc) TRY(clv_inproj_tc(roll, off, L, sx, D, Ke, G, G, wimg_e, gates_e, G, BL, nullptr, 0, 0, fk.next()));
      else TRY(nn_u8(roll, off, L, sx, D, Ke, G, gates_e, G, BL, G, D, nullptr, nullptr, 0, 0, 0, 0, fk.next()));
    }
    if (c->use_x_prev) {
      if (tc) TRY(clv_inproj_tc(roll, off, L, 0, D, Kd, G, G, wimg_d, gates_d, G, BL, nullptr, 0, 0, fk.next()));
      else TRY(nn_u8(roll, off, L, 0, D, Kd, G, gates_d, G, BL, G, D, nullptr, nullptr, 0, 0, 0, 0, fk.next()));
    }
  }
  // ---- key encoder: hW, Wargs, logistic-normal W + its losses (model.py:174-191,244-255,264)
  if (fused_ke) {
    // (measured: letting the key encoder start under step_begin costs 2 us -- its CTAs crowd out the
    //  two projection kernels on the side streams -- so this launch stays fully serialised)
    TRY(clv_keyenc_fwd(roll, off, sx, L, D, Khw, bhw, Kwa, bwa, eps_w, labels, hW, Wargs, W, loss, B, C,
                       c->w_log_var_prior, sb, c->gen_noise, c->seed, ctr, st));
  } else {
    clv_gemm_args a = gz();
    a.M = B; a.N = D; a.K = L * D; a.A = roll; a.lda = D; a.a_u8 = 1; a.a_off = off; a.a_grp = 1;
    a.a_shift = sx; a.Bm = Khw; a.ldb = D; a.C = hW; a.ldc = D;
    a.split_k = pick_split(B, D, L * D);
    if (a.split_k > 1) {
      CLV_CUDA(cudaMemsetAsync(hW, 0, sizeof(float) * (size_t)B * D, st));
      TRY(clv_gemm(&a, st));
      TRY(clv_bias_act(hW, D, B, D, bhw, 1, st));
    } else {
      a.bias = bhw; a.relu = 1;
      TRY(clv_gemm(&a, st));
    }
    TRY(nn_f32(hW, D, Kwa, 2 * C1, Wargs, 2 * C1, B, 2 * C1, D, bwa, 0, 0, st));
    TRY(clv_logitnormal_fwd(Wargs, 2 * C1, eps_w, labels, W, loss, B, C, c->w_log_var_prior, sb,
                            c->gen_noise, c->seed, ctr, st));
  }
  // ---- encoder LSTM (model.py:193-199): roll part hoisted as one GEMM; bias and the
  //      RepeatVector(W) columns are folded into the recurrent kernel's per-sequence constant
  if (tcl) {
    // bias + W term as a per-sequence addend of the tcgen05 projection, then the tcgen05 recurrence
    TRY(nn_f32(W, C, Ke_w, G, rb_e, G, B, G, C, be, 0, 0, st));
    TRY(nn_f32(W, C, Kd_w, G, rb_d, G, B, G, C, bd, 0, 0, st));
    TRY(clv_inproj_tc(roll, off, L, sx, D, Ke, G, G, wimg_e, gates_e, G, BL, rb_e, G, L, st));
    TRY(fk.fork());
    TRY(clv_inproj_tc(roll, off, L, 0, D, Kd, G, G, wimg_d, gates_d, G, BL, rb_d, G, L, fk.next()));
    TRY(clv_lstm_fwd_tc(gates_e, Ue, nullptr, nullptr, 0, h_e, c_e, uimg_e, B, L, H, st));
  } else {
    if (!enc_proj_side) {
      if (tc) TRY(clv_inproj_tc(roll, off, L, sx, D, Ke, G, G, wimg_e, gates_e, G, BL, nullptr, 0, 0, st));
      else TRY(nn_u8(roll, off, L, sx, D, Ke, G, gates_e, G, BL, G, D, nullptr, nullptr, 0, 0, 0, 0, st));
    } else {
      TRY(fk.join());   // both roll projections (issued before the key encoder) are done
    }
    if (pair) {
      TRY(clv_lstm_pair_fwd(gates_e, Ue, be, Ke_w, h_e, c_e, gates_d, c->use_x_prev, Ud, bd, Kd_w, Kd_z, h_d, c_d,
                            W, C, Kzm, bzm, Kzv, bzv, eps_z, Zargs, Zs, loss, sbl, c->gen_noise, c->seed, ctr,
                            B, L, H, Z, st));
    } else {
      TRY(clv_lstm_fwd_fused(gates_e, 1, Ue, be, W, Ke_w, C, nullptr, nullptr, 0, h_e, c_e, B, L, H, st));
    }
  }
  if (!pair) {
  // ---- Z heads + sample + kl (model.py:200-216,236-239)
  TRY_PDL(clv_gauss_heads_fwd(h_e, Kzm, bzm, Kzv, bzv, eps_z, Zargs, Zs, loss, BL, H, Z, sbl,
                              c->gen_noise, c->seed, ctr, st));
  // ---- decoder LSTM (model.py:218-228) on [Xp | Z | W]: Z enters as a rank-Z term per step
  if (c->use_x_prev && (tcl || !enc_proj_side)) TRY(fk.join());
  if (tcl) TRY(clv_lstm_fwd_tc(gates_d, Ud, Zs, Kd_z, Z, h_d, c_d, uimg_d, B, L, H, st));
  else if (c->use_x_prev && !enc_proj_side) TRY(clv_lstm_fwd_fused(gates_d, c->use_x_prev, Ud, bd, W, Kd_w, C, Zs, Kd_z, Z, h_d, c_d, B, L, H, st));
  else TRY_PDL(clv_lstm_fwd_fused(gates_d, c->use_x_prev, Ud, bd, W, Kd_w, C, Zs, Kd_z, Z, h_d, c_d, B, L, H, st));
  }
  // ---- X head + Bernoulli loss + dlogits + dgrad to h_d in one pass (model.py:229-234,241-242)
  const bool xtc = H == 88 && D == 88 && tc && c->do_backward && BL >= xhead_tc_min_rows();
  if (xtc) {
    // large batches: chained tcgen05 GEMMs per 128-row tile (xhead_tc.cu); the head's weight and bias
    // gradients come out of the same pass (third GEMM over the two tiles already in shared memory)
    TRY(clv_xhead_tc(h_d, Kx, bx, roll, off, L, sy, loss, logits, dh, Gr + po[R_X_K], Gr + po[R_X_B], WSP_X, BL, H,
                     D, sbl, st));
  } else if (H == 88 && D == 88) {
    TRY_PDL(clv_xhead_fwd_bwd(h_d, Kx, bx, roll, off, L, sy, loss, logits, dh, BL, H, D, sbl,
                              c->do_backward, st));
  } else {
    TRY(nn_f32(h_d, H, Kx, D, logits, D, BL, D, H, bx, 0, 0, st));
    TRY(clv_bernoulli_ce_fwd_bwd(logits, roll, off, L, sy, loss, BL, D, sbl, c->do_backward, st));
    if (c->do_backward) TRY(nt_f32(logits, D, Kx, D, dh, H, BL, H, D, nullptr, 0, 0, st));
  }
  if (!c->do_backward) return CLV_OK;

  // ---- backward.  Critical path on `st`; every weight gradient on the side branch.
  float *gKhw = Gr + po[R_HW_K], *gbhw = Gr + po[R_HW_B], *gKwa = Gr + po[R_WA_K],
        *gbwa = Gr + po[R_WA_B], *gKe = Gr + po[R_ENC_K], *gUe = Gr + po[R_ENC_U],
        *gbe = Gr + po[R_ENC_B], *gKzm = Gr + po[R_ZM_K], *gbzm = Gr + po[R_ZM_B],
        *gKzv = Gr + po[R_ZV_K], *gbzv = Gr + po[R_ZV_B], *gKd = Gr + po[R_DEC_K],
        *gUd = Gr + po[R_DEC_U], *gbd = Gr + po[R_DEC_B], *gKx = Gr + po[R_X_K],
        *gbx = Gr + po[R_X_B];
  TRY(fk.fork());
  if (!xtc && !(diag_skip() & 512)) {
    TRY(tn_f32(h_d, H, logits, D, gKx, D, H, D, BL, 0, 0, fk.next()));
    TRY(clv_colsum(logits, D, BL, D, gbx, 1, fk.next()));
  }
  // Z-head exchange fused into the two BPTT kernels (Z <= 2): the decoder BPTT also emits
  // dLoss/d(Z_mean|Z_log_var), the encoder BPTT turns it into dLoss/dh_e per cell, and the head
  // weight gradients move to a side stream -- no kernel between the two recurrences
  const bool fuse_heads = Z <= 2;
  const float klw = c->kl_weight * sbl;
  g_clv_pdl = (H == 88 && D == 88) ? pdl_enabled() : 0;
  {
    // both BPTTs as one wavefront launch (lstm_pair.cu), or the decoder BPTT alone
    const int rc__ = pairb
        ? clv_lstm_pair_bwd(gates_d, Ud, c_d, dh, dAsum_d, Kd_w, Kd_z, dZ, Zargs, eps_z, klw, dZa, gates_e, Ue, c_e,
                            dAsum_e, Ke_w, Kzm, Kzv, dW_ext, C, B, L, H, Z, st)
        : clv_lstm_bwd_heads(gates_d, Ud, c_d, dh, dAsum_d, Kd_w, C, dW_ext, 0, Kd_z, Z, dZ,
                             fuse_heads ? Zargs : nullptr, fuse_heads ? eps_z : nullptr, klw,
                             fuse_heads ? dZa : nullptr, nullptr, nullptr, nullptr, 0, B, L, H, st);
    g_clv_pdl = 0;
    if (rc__ != CLV_OK) return rc__;
  }
  TRY(fk.fork());
  const bool tcw = tc && H == 88 && Z <= 8;   // tcgen05 weight gradients
  const int dsk = diag_skip();
  if (dsk & 64) {
  } else if (tcw) {
    // (beside the encoder BPTT: 100 weight-gradient CTAs and 100 one-per-SM BPTT CTAs share 148 SMs, which costs
    //  the critical path 8.6 us -- profiles/skip_probe_r2.txt -- but capping this grid to the 48 free SMs, or
    //  folding the update that follows into the final launch, measured 1-3 us WORSE: the update chain behind
    //  this kernel has no slack either)
    TRY(clv_lstm_wgrad_tc(gates_d, roll, off, L, 0, D, h_d, Zs, Z, c->use_x_prev ? gKd : nullptr, gUd,
                          gKd + (int64_t)xo * G, BL, H, fk.next()));
  } else {
    if (c->use_x_prev) TRY(tn_u8(roll, off, L, 0, D, gates_d, G, gKd, G, D, G, BL, fk.next()));
    TRY(tn_f32(Zs, Z, gates_d, G, gKd + (int64_t)xo * G, G, Z, G, BL, 0, 0, fk.next()));
    TRY(tn_f32(h_d, H, gates_d, G, gUd, G, H, G, BL, -1, L, fk.next()));
  }
  if (!(dsk & 1024)) {
    TRY(tn_f32(W, C, dAsum_d, G, gKd + (int64_t)(xo + Z) * G, G, C, G, B, 0, 0, fk.next()));
    TRY(clv_colsum(dAsum_d, G, B, G, gbd, 1, fk.next()));
  }
  if (fuse_heads && (dsk & 2048)) {
  } else if (fuse_heads) {
    // head weight gradients only (dh = null), off the critical path
    TRY(clv_gauss_heads_bwd(h_e, Kzm, Kzv, eps_z, Zargs, dZ, nullptr, gKzm, gbzm, gKzv, gbzv, BL, H, Z, klw, 0,
                            fk.next()));
  } else {
    // K2b bwd overwrites dh: its only reader (decoder BPTT) is ordered before it on st
    TRY_PDL(clv_gauss_heads_bwd(h_e, Kzm, Kzv, eps_z, Zargs, dZ, dh, gKzm, gbzm, gKzv, gbzv, BL, H, Z, klw, 0, st));
  }
  // data parallel through the host callback (NCCL): ONE all-reduce of the whole buffer after the last weight
  // gradient, then one update (measured: a second all-reduce overlapped with the encoder BPTT costs more than
  // it hides -- 0.178 vs 0.163 ms/step at N=2); the peer-memory form keeps the three-bucket schedule of N=1
  const bool dp = opt && opt->exchange && !opt->p2p;
  if (opt) {   // decoder and X-head gradients are complete once the side branches drain; the Z heads
               // stay out of this range: the encoder BPTT below still reads their kernels
    TRY(fk.fork());
    TRY(fk.gather());
    if (!dp && !(dsk & 32)) TRY(adam(R_DEC_K, CLV_N_TENSORS, 0, fk.opt_stream(), 0));
  }
  if (pairb) {
    // (the encoder BPTT ran inside the wavefront launch above)
  } else if (fuse_heads)
    TRY_PDL(clv_lstm_bwd_heads(gates_e, Ue, c_e, nullptr, dAsum_e, Ke_w, C, dW_ext, 1, nullptr, 0, nullptr,
                               nullptr, nullptr, 0.f, nullptr, dZa, Kzm, Kzv, Z, B, L, H, st));
  else
    TRY_PDL(clv_lstm_bwd_fused(gates_e, Ue, c_e, dh, dAsum_e, Ke_w, C, dW_ext, 1, nullptr, 0, nullptr, B, L, H, st));
  TRY(fk.fork());
  if (dsk & 8) {
  } else if (tcw) {
    TRY(clv_lstm_wgrad_tc(gates_e, roll, off, L, sx, D, h_e, nullptr, 0, gKe, gUe, nullptr, BL, H, fk.next()));
  } else {
    TRY(tn_u8(roll, off, L, sx, D, gates_e, G, gKe, G, D, G, BL, fk.next()));
    TRY(tn_f32(h_e, H, gates_e, G, gUe, G, H, G, BL, -1, L, fk.next()));
  }
  if (!(dsk & 16)) {
    TRY(tn_f32(W, C, dAsum_e, G, gKe + (int64_t)D * G, G, C, G, B, 0, 0, fk.next()));
    TRY(clv_colsum(dAsum_e, G, B, G, gbe, 1, fk.next()));
  }
  // The step's tail.  One GPU / peer memory: the last two updates run CONCURRENTLY -- [encoder LSTM | Z heads]
  // on the optimizer stream as soon as the encoder weight gradients have drained, [key encoder] on the caller's
  // stream behind its backward kernel; whichever block finishes last advances `iterations` (group totals).
  const bool split_tail = opt && !dp && fused_ke;
  if (fused_ke) {
    // K2 backward + every key-encoder weight gradient in one kernel (sparse scatter for dK_hW)
    if (!(dsk & 4))
    TRY_PDL(clv_keyenc_bwd_full(roll, off, sx, L, D, Wargs, eps_w, labels, W, dW_ext, Kwa, hW, dWargs, dhW, gKhw,
                                gbhw, gKwa, gbwa, B, C, c->w_log_var_prior, c->class_weight * sb,
                                c->w_kl_weight * sb, st));
    if (split_tail && (dsk & 3)) {
      // (one of the two final launches dropped: the other advances alone)
      if (!(dsk & 1)) TRY_PDL(adam(R_HW_K, R_ENC_K, 1, st, 1, 1, 1));
      TRY(fk.gather());
      if (!(dsk & 2)) TRY(adam(R_ENC_K, R_DEC_K, 1, fk.opt_stream(), 2, 1, 1));
    } else if (split_tail) {
      const int nadv = clv_adamwn_range_blocks(c, opt->weightnorm, R_HW_K, R_ENC_K) +
                       clv_adamwn_range_blocks(c, opt->weightnorm, R_ENC_K, R_DEC_K);
      const int nlast = clv_p2p_allreduce_blocks(po[R_ENC_K] - po[R_HW_K]) +
                        clv_p2p_allreduce_blocks(po[R_DEC_K] - po[R_ENC_K]);
      TRY_PDL(adam(R_HW_K, R_ENC_K, nadv, st, 1, nlast, 0));
      // (the loss scalars were reduced with the first bucket on this stream: the mirror goes with this launch)
      TRY(fk.gather());
      TRY(adam(R_ENC_K, R_DEC_K, nadv, fk.opt_stream(), 2, nlast, 1));
    }
  } else {
    TRY(clv_keyenc_bwd(Wargs, eps_w, labels, W, dW_ext, Kwa, hW, dWargs, dhW, B, C, D,
                       c->w_log_var_prior, c->class_weight * sb, c->w_kl_weight * sb, st));
    TRY(fk.fork());
    TRY(tn_f32(hW, D, dWargs, 2 * C1, gKwa, 2 * C1, D, 2 * C1, B, 0, 0, fk.next()));
    TRY(clv_colsum(dWargs, 2 * C1, B, 2 * C1, gbwa, 1, fk.next()));
    TRY(tn_u8(roll, off, 1, sx, D, dhW, D, gKhw, D, L * D, D, B, fk.next()));
    TRY(clv_colsum(dhW, D, B, D, gbhw, 1, fk.next()));
  }
  TRY(fk.join());
  if (dp) {
    if (opt->exchange(opt->exchange_user, Gr, po[R_X_B] + pc[R_X_B] + 8, (void*)st) != 0) return CLV_E_CUDA;
    TRY(adam(R_HW_K, CLV_N_TENSORS, 1, st, 0));
  } else if (opt && !split_tail) {
    TRY(adam(R_HW_K, R_DEC_K, 1, st, 2));
  }
  return CLV_OK;
}

int vae_step(const clv_cfg* c, const float* P, float* Gr, float* loss, const uint8_t* roll,
             const int32_t* off, const int32_t* labels, float* eps_w, float* eps_z, uint64_t* ctr,
             float* ws, cudaStream_t st, const clv_adam_args* opt) {
  int64_t po[CLV_N_TENSORS]; int32_t pr[CLV_N_TENSORS], pc[CLV_N_TENSORS];
  clv_param_layout(c, po, pr, pc);
  const int B = c->B, D = c->D, H = c->H, Hc = c->Hc, Z = c->Z, C = c->C, C1 = C - 1;
  const int xo = c->use_x_prev ? D : 0;
  const int sx = c->x_shift > 0 ? c->x_shift : (c->use_x_prev ? 1 : 0);
  const int sy = c->y_shift > 0 ? c->y_shift : sx;
  const float sb = 1.0f / (float)c->B_global;
  const Ws w = carve(c);
#define WSP(name) (ws + w.find(name))
  float *h_w = WSP("h_w"), *Wargs = WSP("Wargs"), *W = WSP("W"), *h = WSP("h"),
        *Zargs = WSP("Zargs"), *Zs = WSP("Zs"), *h_dec = WSP("h_dec"), *logits = WSP("logits"),
        *dpre = WSP("dpre"), *dh = WSP("dh"), *dZ = WSP("dZ"), *dW_ext = WSP("dW_ext"),
        *dWargs = WSP("dWargs"), *dh_w = WSP("dh_w");
#undef WSP
  const float *Khw = P + po[V_HW_K], *bhw = P + po[V_HW_B], *Kwm = P + po[V_WM_K],
              *bwm = P + po[V_WM_B], *Kwv = P + po[V_WV_K], *bwv = P + po[V_WV_B],
              *Kh = P + po[V_H_K], *bh = P + po[V_H_B], *Kzm = P + po[V_ZM_K], *bzm = P + po[V_ZM_B],
              *Kzv = P + po[V_ZV_K], *bzv = P + po[V_ZV_B], *Kdh = P + po[V_DH_K],
              *bdh = P + po[V_DH_B], *Kx = P + po[V_X_K], *bx = P + po[V_X_B];

  if (opt && opt->p2p) TRY(clv_p2p_wait_done(opt->p2p, opt->state, c, st));
  TRY(clv_step_begin(loss, ctr, !c->accumulate, c->gen_noise, st));
  if (c->do_backward && !c->accumulate)
    CLV_CUDA(cudaMemsetAsync(Gr, 0, sizeof(float) * (po[V_X_B] + pc[V_X_B]), st));

  // ---- small / medium batches: the whole forward + backward as ONE kernel with the model resident in shared
  //      memory (vae_fused.cu); 3 launches per step instead of 37.  Large batches (where 8-frame tiles would
  //      pay 41 k red.adds per tile) and shapes outside the fused kernel use the per-layer schedule below.
  if (B <= vae_fused_max_rows()) {
    const int rc_ = clv_vae_fused_step(c, P, Gr, loss, roll, off, labels, eps_w, eps_z, ctr, Wargs, W, Zargs, st);
    if (rc_ == CLV_OK) {
      if (opt && c->do_backward) {
        if (opt->p2p) {
          TRY(adam_p2p(c, P, opt, 0, CLV_N_TENSORS, 0, 1, 1, true, st));
        } else {
          if (opt->exchange && opt->exchange(opt->exchange_user, Gr, po[V_X_B] + pc[V_X_B] + 8, (void*)st) != 0)
            return CLV_E_CUDA;
          TRY_PDL(clv_adamwn_step_range(c, const_cast<float*>(P), Gr, opt->state, opt->lr, opt->beta_1, opt->beta_2,
                                        opt->epsilon, opt->grad_scale, opt->weightnorm, 0, CLV_N_TENSORS, 1,
                                        opt->loss_mirror, st));
        }
      }
      return CLV_OK;
    }
    if (rc_ != CLV_E_UNSUPPORTED) return rc_;
  }
  // ---- forward (cl_vae/model.py:141-188)
  TRY(nn_u8(roll, off, 1, sx, D, Khw, Hc, h_w, Hc, B, Hc, D, bhw, nullptr, 0, 0, 1, 0, st));
  TRY(nn_f32(h_w, Hc, Kwm, C1, Wargs, 2 * C1, B, C1, Hc, bwm, 0, 0, st));
  TRY(nn_f32(h_w, Hc, Kwv, C1, Wargs + C1, 2 * C1, B, C1, Hc, bwv, 0, 0, st));
  TRY(clv_logitnormal_fwd(Wargs, 2 * C1, eps_w, labels, W, loss, B, C, c->w_log_var_prior, sb,
                          c->gen_noise, c->seed, ctr, st));
  TRY(nn_u8(roll, off, 1, sx, D, Kh, H, h, H, B, H, D, nullptr, nullptr, 0, 0, 0, 0, st));
  TRY(nn_f32(W, C, Kh + (int64_t)D * H, H, h, H, B, H, C, bh, 1, 1, st));
  TRY(clv_gauss_heads_fwd(h, Kzm, bzm, Kzv, bzv, eps_z, Zargs, Zs, loss, B, H, Z, sb, c->gen_noise,
                          c->seed, ctr, st));
  TRY(nn_f32(W, C, Kdh, H, h_dec, H, B, H, C, nullptr, 0, 0, st));
  if (c->use_x_prev)
    TRY(nn_u8(roll, off, 1, 0, D, Kdh + (int64_t)C * H, H, h_dec, H, B, H, D, nullptr, nullptr, 0, 0,
              0, 1, st));
  TRY(nn_f32(Zs, Z, Kdh + (int64_t)(C + xo) * H, H, h_dec, H, B, H, Z, bdh, 1, 1, st));
  TRY(nn_f32(h_dec, H, Kx, D, logits, D, B, D, H, bx, 0, 0, st));
  TRY(clv_bernoulli_ce_fwd_bwd(logits, roll, off, 1, sy, loss, B, D, sb, c->do_backward, st));
  if (!c->do_backward) return CLV_OK;

  // ---- backward
  float *gKhw = Gr + po[V_HW_K], *gbhw = Gr + po[V_HW_B], *gKwm = Gr + po[V_WM_K],
        *gbwm = Gr + po[V_WM_B], *gKwv = Gr + po[V_WV_K], *gbwv = Gr + po[V_WV_B],
        *gKh = Gr + po[V_H_K], *gbh = Gr + po[V_H_B], *gKzm = Gr + po[V_ZM_K],
        *gbzm = Gr + po[V_ZM_B], *gKzv = Gr + po[V_ZV_K], *gbzv = Gr + po[V_ZV_B],
        *gKdh = Gr + po[V_DH_K], *gbdh = Gr + po[V_DH_B], *gKx = Gr + po[V_X_K],
        *gbx = Gr + po[V_X_B];
  TRY(tn_f32(h_dec, H, logits, D, gKx, D, H, D, B, 0, 0, st));
  TRY(clv_colsum(logits, D, B, D, gbx, 1, st));
  TRY(nt_f32(logits, D, Kx, D, dpre, H, B, H, D, h_dec, H, 0, st));
  TRY(tn_f32(W, C, dpre, H, gKdh, H, C, H, B, 0, 0, st));
  if (c->use_x_prev) TRY(tn_u8(roll, off, 1, 0, D, dpre, H, gKdh + (int64_t)C * H, H, D, H, B, st));
  TRY(tn_f32(Zs, Z, dpre, H, gKdh + (int64_t)(C + xo) * H, H, Z, H, B, 0, 0, st));
  TRY(clv_colsum(dpre, H, B, H, gbdh, 1, st));
  TRY(nt_f32(dpre, H, Kdh, H, dW_ext, C, B, C, H, nullptr, 0, 0, st));
  TRY(nt_f32(dpre, H, Kdh + (int64_t)(C + xo) * H, H, dZ, Z, B, Z, H, nullptr, 0, 0, st));
  TRY(clv_gauss_heads_bwd(h, Kzm, Kzv, eps_z, Zargs, dZ, dh, gKzm, gbzm, gKzv, gbzv, B, H, Z,
                          c->kl_weight * sb, 1, st));
  TRY(tn_u8(roll, off, 1, sx, D, dh, H, gKh, H, D, H, B, st));
  TRY(tn_f32(W, C, dh, H, gKh + (int64_t)D * H, H, C, H, B, 0, 0, st));
  TRY(clv_colsum(dh, H, B, H, gbh, 1, st));
  TRY(nt_f32(dh, H, Kh + (int64_t)D * H, H, dW_ext, C, B, C, H, nullptr, 0, 1, st));
  TRY(clv_logitnormal_bwd(Wargs, 2 * C1, eps_w, labels, W, dW_ext, dWargs, B, C,
                          c->w_log_var_prior, c->class_weight * sb, c->w_kl_weight * sb, st));
  TRY(tn_f32(h_w, Hc, dWargs, 2 * C1, gKwm, C1, Hc, C1, B, 0, 0, st));
  TRY(tn_f32(h_w, Hc, dWargs + C1, 2 * C1, gKwv, C1, Hc, C1, B, 0, 0, st));
  TRY(clv_colsum(dWargs, 2 * C1, B, C1, gbwm, 1, st));
  TRY(clv_colsum(dWargs + C1, 2 * C1, B, C1, gbwv, 1, st));
  TRY(nt_f32(dWargs, 2 * C1, Kwm, C1, dh_w, Hc, B, Hc, C1, nullptr, 0, 0, st));
  TRY(nt_f32(dWargs + C1, 2 * C1, Kwv, C1, dh_w, Hc, B, Hc, C1, h_w, Hc, 1, st));
  TRY(tn_u8(roll, off, 1, sx, D, dh_w, Hc, gKhw, Hc, D, Hc, B, st));
  TRY(clv_colsum(dh_w, Hc, B, Hc, gbhw, 1, st));
  if (opt && opt->p2p) {
    TRY(adam_p2p(c, P, opt, 0, CLV_N_TENSORS, 0, 1, 1, true, st));
  } else if (opt) {
    if (opt->exchange && opt->exchange(opt->exchange_user, Gr, po[V_X_B] + pc[V_X_B] + 8, (void*)st) != 0)
      return CLV_E_CUDA;
    TRY(clv_adamwn_step_range(c, const_cast<float*>(P), Gr, opt->state, opt->lr, opt->beta_1, opt->beta_2,
                              opt->epsilon, opt->grad_scale, opt->weightnorm, 0, CLV_N_TENSORS, 1,
                              opt->loss_mirror, st));
  }
  return CLV_OK;
}

}  // namespace

unsigned long long g_clv_launches = 0;
thread_local int g_clv_pdl = 0;
extern "C" int clv_version(void) { return 100; }
extern "C" int64_t clv_launch_count(void) { return (int64_t)g_clv_launches; }

extern "C" int clv_runtime_init(void) {
  int dev = 0;
  CLV_CUDA(cudaGetDevice(&dev));
  if (dev >= 16) return CLV_E_UNSUPPORTED;
  SideStream& s = g_side[dev];
  if (s.ready) return CLV_OK;
  for (int i = 0; i < NAUX; ++i) CLV_CUDA(cudaStreamCreateWithFlags(&s.aux[i], cudaStreamNonBlocking));
  for (int i = 0; i < 8; ++i) CLV_CUDA(cudaEventCreateWithFlags(&s.fork_ev[i], cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < NAUX; ++k)
      CLV_CUDA(cudaEventCreateWithFlags(&s.join_ev[i][k], cudaEventDisableTiming));
  s.ready = true;
  return CLV_OK;
}

extern "C" const char* clv_error_string(int code) {
  switch (code) {
    case CLV_OK: return "ok";
    case CLV_E_INVALID: return "invalid argument";
    case CLV_E_UNSUPPORTED: return "unsupported shape (H must be 88 for the LSTM; C<=16, Z<=16, D<=128)";
    case CLV_E_CUDA: return "CUDA error (launch or runtime call failed; no CPU fallback exists)";
    case CLV_E_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

extern "C" int64_t clv_param_layout(const clv_cfg* c, int64_t* offs, int32_t* rows, int32_t* cols) {
  if (!c || !offs || !rows || !cols) return CLV_E_INVALID;
  const int D = c->D, H = c->H, Z = c->Z, C = c->C, L = c->L;
  int r[CLV_N_TENSORS], k[CLV_N_TENSORS];
  if (c->model == 0) {
    const int G = 4 * H, in_d = (c->use_x_prev ? D : 0) + Z + C;
    const int rr[] = {L * D, 0, D, 0, D + C, H, 0, H, 0, H, 0, in_d, H, 0, H, 0};
    const int cc[] = {D, D, 2 * (C - 1), 2 * (C - 1), G, G, G, Z, Z, Z, Z, G, G, G, D, D};
    memcpy(r, rr, sizeof(rr)); memcpy(k, cc, sizeof(cc));
  } else if (c->model == 1) {
    const int Hc = c->Hc, in_dec = C + (c->use_x_prev ? D : 0) + Z;
    const int rr[] = {D, 0, Hc, 0, Hc, 0, D + C, 0, H, 0, H, 0, in_dec, 0, H, 0};
    const int cc[] = {Hc, Hc, C - 1, C - 1, C - 1, C - 1, H, H, Z, Z, Z, Z, H, H, D, D};
    memcpy(r, rr, sizeof(rr)); memcpy(k, cc, sizeof(cc));
  } else {
    return CLV_E_INVALID;
  }
  int64_t o = 0;
  for (int i = 0; i < CLV_N_TENSORS; ++i) {
    offs[i] = o; rows[i] = r[i]; cols[i] = k[i];
    o += (r[i] > 0 ? (int64_t)r[i] : 1) * k[i];
  }
  return o;
}

extern "C" int64_t clv_workspace_bytes(const clv_cfg* cfg) {
  int rc = check_cfg(cfg);
  if (rc != CLV_OK) return rc;
  return carve(cfg).total * (int64_t)sizeof(float);
}

extern "C" int64_t clv_workspace_offset(const clv_cfg* cfg, const char* name) {
  if (check_cfg(cfg) != CLV_OK || !name) return -1;
  return carve(cfg).find(name);
}

static int train_step_impl(const clv_cfg* cfg, const float* params, float* grads, float* loss_acc,
                           const uint8_t* roll, const int32_t* win_off, const int32_t* labels,
                           float* eps_w, float* eps_z, uint64_t* rng_ctr, void* workspace,
                           int64_t workspace_bytes, const clv_adam_args* opt, void* stream) {
  int rc = check_cfg(cfg);
  if (rc != CLV_OK) return rc;
  if (!params || !loss_acc || !roll || !win_off || !labels || !eps_w || !eps_z || !workspace)
    return CLV_E_INVALID;
  if (cfg->do_backward && !grads) return CLV_E_INVALID;
  if (cfg->gen_noise && !rng_ctr) return CLV_E_INVALID;
  if (opt && (!opt->state || !cfg->do_backward || cfg->accumulate)) return CLV_E_INVALID;
  if (opt && (opt->loss_mirror || opt->exchange || opt->p2p)) {
    // both read the 8 loss scalars at grads[P..P+8): the caller must pass the [grads | losses] buffer
    int64_t po_[CLV_N_TENSORS]; int32_t pr_[CLV_N_TENSORS], pc_[CLV_N_TENSORS];
    const int64_t P_ = clv_param_layout(cfg, po_, pr_, pc_);
    if (P_ < 0 || loss_acc != grads + P_) return CLV_E_INVALID;
  }
  if (workspace_bytes < carve(cfg).total * (int64_t)sizeof(float)) return CLV_E_WORKSPACE;
  if (cfg->B == 0) return CLV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->model == 0)
    return vrnn_step(cfg, params, grads, loss_acc, roll, win_off, labels, eps_w, eps_z, rng_ctr,
                     (float*)workspace, st, opt);
  return vae_step(cfg, params, grads, loss_acc, roll, win_off, labels, eps_w, eps_z, rng_ctr,
                  (float*)workspace, st, opt);
}

extern "C" int clv_train_step(const clv_cfg* cfg, const float* params, float* grads,
                              float* loss_acc, const uint8_t* roll, const int32_t* win_off,
                              const int32_t* labels, float* eps_w, float* eps_z, uint64_t* rng_ctr,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  return train_step_impl(cfg, params, grads, loss_acc, roll, win_off, labels, eps_w, eps_z, rng_ctr,
                         workspace, workspace_bytes, nullptr, stream);
}

extern "C" int clv_train_step_opt(const clv_cfg* cfg, float* params, float* grads, float* loss_acc,
                                  const uint8_t* roll, const int32_t* win_off, const int32_t* labels,
                                  float* eps_w, float* eps_z, uint64_t* rng_ctr, void* workspace,
                                  int64_t workspace_bytes, const clv_adam_args* opt, void* stream) {
  if (!opt) return CLV_E_INVALID;
  return train_step_impl(cfg, params, grads, loss_acc, roll, win_off, labels, eps_w, eps_z, rng_ctr,
                         workspace, workspace_bytes, opt, stream);
}
