// K6: AdamWithWeightnorm.get_updates (utils/weightnorm.py:75-143) with
// get_weightnorm_params_and_grads (:146-166) and add_weightnorm_param_updates (:169-178), fused
// into ONE launch over the flat parameter / gradient / state buffers.
// Every >=2-D tensor is updated in its (V, g) weight-norm reparameterisation, norms taken over
// axis 0 (per output column, incl. each of the 4H LSTM gate columns); 1-D tensors get plain Keras
// Adam (p -= lr_t * m / (sqrt(v) + eps)) [K2-recall (7)].
// A block owns 8 adjacent columns of one matrix (32-byte sectors per row) and 32 row lanes.
#include "common.cuh"

namespace {

struct AdamPlan {
  int64_t off[CLV_N_TENSORS];
  int32_t rows[CLV_N_TENSORS], cols[CLV_N_TENSORS], coloff[CLV_N_TENSORS];
  int32_t first_block[CLV_N_TENSORS + 1];
  int64_t P;
  int32_t NC;
};

constexpr int CT = 8, RL = 128, NTH = CT * RL;

// P2P = true: the gradient all-reduce is FUSED into the optimizer.  Every rank's [grads | losses]
// buffer lives in symmetric (peer-mapped) memory; after a cross-GPU barrier each rank's kernel reads
// the N peer buffers over NVLink (plain ld.global on peer addresses), sums them in the fixed order
// p = 0..N-1 (bitwise identical on every rank, so the replicas stay in lock-step) and applies the
// update -- no separate collective launch, no extra pass over the gradient.
struct PeerSet {
  const float* const* peers;   // device array of n pointers (this rank's own buffer included)
  int n;
  float* gsum;                 // local scratch [P]: reduced gradient for the second pass
  float* loss_out;             // local [8]: reduced loss scalars
};

template <bool P2P>
__device__ __forceinline__ float load_grad(const float* __restrict__ G, const PeerSet& ps, int64_t e) {
  if (!P2P) return G[e];
  float g = 0.f;
  for (int p = 0; p < ps.n; ++p) g += ps.peers[p][e];
  return g;
}

template <bool P2P>
__global__ void __launch_bounds__(NTH) adamwn_kernel(const AdamPlan pl, float* __restrict__ W,
                                                     const float* __restrict__ G,
                                                     float* __restrict__ state, const double lr,
                                                     const double b1d, const double b2d,
                                                     const float eps, const float gscale,
                                                     const int weightnorm, const PeerSet ps,
                                                     const int block_base, const int advance) {
  pdl_wait();   // no-op unless launched as a programmatic dependent
  if (P2P && blockIdx.x == 0 && threadIdx.x < 8) {
    float v = 0.f;
    for (int p = 0; p < ps.n; ++p) v += ps.peers[p][pl.P + threadIdx.x];
    ps.loss_out[threadIdx.x] = v;
  }
  __shared__ float red[2][RL][CT];
  __shared__ float col[4][CT];
  float* m = state;
  float* v = state + pl.P;
  float* vsc = state + 2 * pl.P;
  float* mg = vsc + pl.NC;
  float* vg = mg + pl.NC;
  int* iter = reinterpret_cast<int*>(vg + pl.NC);
  unsigned* done = reinterpret_cast<unsigned*>(iter + 1);

  const int t = *iter + 1;
  const float lr_t = (float)(lr * sqrt(1.0 - pow(b2d, (double)t)) / (1.0 - pow(b1d, (double)t)));
  const float b1 = (float)b1d, b2 = (float)b2d;

  // block_base: first plan block of this launch (a launch may cover a sub-range of the tensors)
  const int gb = (int)blockIdx.x + block_base;
  int ti = 0;
#pragma unroll
  for (int i = 1; i < CLV_N_TENSORS; ++i)
    if (gb >= pl.first_block[i]) ti = i;
  const int lb = gb - pl.first_block[ti];
  const int64_t off = pl.off[ti];
  const int rows = pl.rows[ti], cols = pl.cols[ti];
  const int tid = threadIdx.x;

  if (rows == 0 || !weightnorm) {  // plain Adam on a flat chunk
    const int64_t n = (rows == 0) ? cols : (int64_t)rows * cols;
    const int64_t i = (int64_t)lb * NTH + tid;
    if (i < n) {
      const float g = load_grad<P2P>(G, ps, off + i) * gscale;
      const float mt = b1 * m[off + i] + (1.0f - b1) * g;
      const float vt = b2 * v[off + i] + (1.0f - b2) * g * g;
      m[off + i] = mt;
      v[off + i] = vt;
      W[off + i] -= lr_t * mt / (sqrtf(vt) + eps);
    }
  } else {
    const int cx = tid & (CT - 1), ry = tid >> 3;
    const int c = lb * CT + cx;
    const bool cv = c < cols;
    const int sc = pl.coloff[ti] + c;
    const float vs = cv ? vsc[sc] : 1.f;
    // pass 1: ||V||^2 and <G, V> per column
    float svv = 0.f, sgv = 0.f;
    if (cv)
#pragma unroll 4
      for (int r = ry; r < rows; r += RL) {
        const int64_t e = off + (int64_t)r * cols + c;
        const float graw = load_grad<P2P>(G, ps, e);
        if (P2P) ps.gsum[e] = graw;
        const float V = W[e] / vs, g = graw * gscale;
        svv = fmaf(V, V, svv);
        sgv = fmaf(g, V, sgv);
      }
    red[0][ry][cx] = svv;
    red[1][ry][cx] = sgv;
    __syncthreads();
    if (ry == 0) {
#pragma unroll
      for (int i = 1; i < RL; ++i) { svv += red[0][i][cx]; sgv += red[1][i][cx]; }
      const float V_norm = sqrtf(svv);
      const float grad_g = sgv / V_norm;
      const float g_param = vs * V_norm;
      float new_g = g_param;
      if (cv) {
        const float mgt = b1 * mg[sc] + (1.0f - b1) * grad_g;
        const float vgt = b2 * vg[sc] + (1.0f - b2) * grad_g * grad_g;
        mg[sc] = mgt;
        vg[sc] = vgt;
        new_g = g_param - lr_t * mgt / (sqrtf(vgt) + eps);
      }
      col[0][cx] = grad_g / V_norm;
      col[1][cx] = new_g;
    }
    __syncthreads();
    const float gg_over_norm = col[0][cx];
    // pass 2: Adam on V; new V parked in W.  Rows are processed in explicit batches of 4 (all loads,
    // then all stores) because W/m/v may alias as far as the compiler knows: without batching every
    // iteration would wait a full memory round trip behind the previous iteration's stores.
    float snn = 0.f;
    if (cv)
      for (int r0 = ry; r0 < rows; r0 += 4 * RL) {
        float Wv[4], Gv[4], Mv[4], Vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RL;
          if (r < rows) {
            const int64_t e = off + (int64_t)r * cols + c;
            Wv[u] = W[e]; Gv[u] = P2P ? ps.gsum[e] : G[e]; Mv[u] = m[e]; Vv[u] = v[e];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RL;
          if (r < rows) {
            const int64_t e = off + (int64_t)r * cols + c;
            const float V = Wv[u] / vs, g = Gv[u] * gscale;
            const float gV = vs * (g - gg_over_norm * V);
            const float mt = b1 * Mv[u] + (1.0f - b1) * gV;
            const float vt = b2 * Vv[u] + (1.0f - b2) * gV * gV;
            m[e] = mt;
            v[e] = vt;
            const float nV = V - lr_t * mt / (sqrtf(vt) + eps);
            W[e] = nV;
            snn = fmaf(nV, nV, snn);
          }
        }
      }
    red[0][ry][cx] = snn;
    __syncthreads();
    if (ry == 0) {
#pragma unroll
      for (int i = 1; i < RL; ++i) snn += red[0][i][cx];
      const float ns = col[1][cx] / sqrtf(snn);
      if (cv) vsc[sc] = ns;
      col[2][cx] = ns;
    }
    __syncthreads();
    // pass 3: W = V_scaler' * V'  (same batching)
    const float ns = col[2][cx];
    if (cv)
      for (int r0 = ry; r0 < rows; r0 += 4 * RL) {
        float Wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RL;
          if (r < rows) Wv[u] = W[off + (int64_t)r * cols + c];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * RL;
          if (r < rows) W[off + (int64_t)r * cols + c] = Wv[u] * ns;
        }
      }
  }
  // last block to finish advances `iterations` (only the final launch of a step is told to)
  if (!advance) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned d = atomicAdd(done, 1u);
    if (d == gridDim.x - 1) {
      *iter = t;
      *done = 0u;
    }
  }
}

__global__ void adamwn_init_kernel(float* state, int64_t P, int NC) {
  const int64_t n = 2 * P + 3 * (int64_t)NC + 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    state[i] = (i >= 2 * P && i < 2 * P + NC) ? 1.0f : 0.0f;
}

int make_plan(const clv_cfg* cfg, AdamPlan* pl, int weightnorm) {
  int64_t P = clv_param_layout(cfg, pl->off, pl->rows, pl->cols);
  if (P < 0) return (int)P;
  pl->P = P;
  int nc = 0, nb = 0;
  for (int i = 0; i < CLV_N_TENSORS; ++i) {
    pl->coloff[i] = nc;
    pl->first_block[i] = nb;
    if (pl->rows[i] > 0) nc += pl->cols[i];
    if (pl->rows[i] > 0 && weightnorm) nb += (pl->cols[i] + CT - 1) / CT;
    else {
      const int64_t n = pl->rows[i] > 0 ? (int64_t)pl->rows[i] * pl->cols[i] : pl->cols[i];
      nb += (int)((n + NTH - 1) / NTH);
    }
  }
  pl->first_block[CLV_N_TENSORS] = nb;
  pl->NC = nc;
  return CLV_OK;
}

}  // namespace

extern "C" int64_t clv_adamwn_state_floats(const clv_cfg* cfg) {
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, 1);
  if (rc != CLV_OK) return rc;
  return 2 * pl.P + 3 * (int64_t)pl.NC + 2;
}

extern "C" int clv_adamwn_init(const clv_cfg* cfg, float* state, void* stream) {
  if (!cfg || !state) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, 1);
  if (rc != CLV_OK) return rc;
  adamwn_init_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(state, pl.P, pl.NC);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_adamwn_step_range(const clv_cfg* cfg, float* params, const float* grads, float* state,
                                     double lr, double beta_1, double beta_2, double epsilon,
                                     double grad_scale, int32_t weightnorm, int32_t t_first,
                                     int32_t t_last, int32_t advance, void* stream) {
  if (!cfg || !params || !grads || !state) return CLV_E_INVALID;
  if (t_first < 0 || t_last > CLV_N_TENSORS || t_first >= t_last) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, weightnorm);
  if (rc != CLV_OK) return rc;
  PeerSet ps = {nullptr, 0, nullptr, nullptr};
  const int nb = pl.first_block[t_last] - pl.first_block[t_first];
  if (nb <= 0) return CLV_OK;
  CLV_CUDA(clv_launch(adamwn_kernel<false>, nb, NTH, 0, (cudaStream_t)stream, pl, params, grads, state, lr,
                      beta_1, beta_2, (float)epsilon, (float)grad_scale, (int)weightnorm, ps,
                      pl.first_block[t_first], (int)advance));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_adamwn_step(const clv_cfg* cfg, float* params, const float* grads, float* state,
                               double lr, double beta_1, double beta_2, double epsilon,
                               double grad_scale, int32_t weightnorm, void* stream) {
  return clv_adamwn_step_range(cfg, params, grads, state, lr, beta_1, beta_2, epsilon, grad_scale,
                               weightnorm, 0, CLV_N_TENSORS, 1, stream);
}

extern "C" int clv_adamwn_step_p2p(const clv_cfg* cfg, float* params, const float* const* peer_grads,
                                   int32_t n_peers, float* gsum, float* loss_out, float* state,
                                   double lr, double beta_1, double beta_2, double epsilon,
                                   int32_t weightnorm, void* stream) {
  if (!cfg || !params || !peer_grads || !gsum || !loss_out || !state) return CLV_E_INVALID;
  if (n_peers < 1 || n_peers > 16) return CLV_E_UNSUPPORTED;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, weightnorm);
  if (rc != CLV_OK) return rc;
  PeerSet ps = {peer_grads, n_peers, gsum, loss_out};
  adamwn_kernel<true><<<pl.first_block[CLV_N_TENSORS], NTH, 0, (cudaStream_t)stream>>>(
      pl, params, nullptr, state, lr, beta_1, beta_2, (float)epsilon, 1.0f, weightnorm, ps, 0, 1);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
