// K6: AdamWithWeightnorm.get_updates (utils/weightnorm.py:75-143) with
// get_weightnorm_params_and_grads (:146-166) and add_weightnorm_param_updates (:169-178), fused
// into ONE launch over the flat parameter / gradient / state buffers.
// Every >=2-D tensor is updated in its (V, g) weight-norm reparameterisation, norms taken over
// axis 0 (per output column, incl. each of the 4H LSTM gate columns); 1-D tensors get plain Keras
// Adam (p -= lr_t * m / (sqrt(v) + eps)) [K2-recall (7)].
// A block owns 8 (short tensors) or 4 (tall tensors) adjacent columns of one matrix; see below.
#include "common.cuh"
#include <cstdlib>

namespace {

struct AdamPlan {
  int64_t off[CLV_N_TENSORS];
  int32_t rows[CLV_N_TENSORS], cols[CLV_N_TENSORS], coloff[CLV_N_TENSORS];
  int32_t first_block[CLV_N_TENSORS + 1];
  int64_t P;
  int32_t NC;
};

constexpr int NTH = 1024;

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
  // in-kernel hand-shake (clv_p2p_args): flag blocks of every rank in peer-mapped memory
  int32_t* const* flags;       // device array of n pointers, or null (caller-side barriers)
  int rank, slot, reduce_losses;
};

// flag block of one rank: ready[slot][src rank] = step number whose gradients of bucket `slot` rank `src` has
// finished writing; done[src] = step number after which `src` no longer reads anyone's gradients
constexpr int P2P_MAXR = 16, P2P_SLOTS = 4;
__device__ __forceinline__ int32_t* flag_ready(int32_t* f, int slot, int src) { return f + slot * P2P_MAXR + src; }
// landed[slot][src] = step number for which rank `src` has delivered its slice of bucket `slot` (two-shot form)
__device__ __forceinline__ int32_t* flag_landed(int32_t* f, int slot, int src) { return f + (P2P_SLOTS + slot) * P2P_MAXR + src; }
__device__ __forceinline__ int32_t* flag_done(int32_t* f, int src) { return f + 2 * P2P_SLOTS * P2P_MAXR + src; }
// last row: local counters (never written by a peer)
__device__ __forceinline__ int32_t* flag_local(int32_t* f, int i) { return f + (2 * P2P_SLOTS + 1) * P2P_MAXR + i; }
__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, const int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool P2P>
__device__ __forceinline__ float load_grad(const float* __restrict__ G, const PeerSet& ps, int64_t e) {
  if (!P2P) return G[e];
  float g = 0.f;
  for (int p = 0; p < ps.n; ++p) g += ps.peers[p][e];
  return g;
}

// Block shapes: a block owns CT adjacent columns of one matrix and RL row lanes (CT * RL = NTH).
//   short tensors (rows <= 128; every LSTM / head kernel): CT = 8, RL = 128, one row per thread;
//   tall tensors (the 1408-row key-encoder kernel):        CT = 4, RL = 256, up to NRF rows per thread.
// In both cases a thread's rows of W, grad, m, v stay in REGISTERS across the three phases of the
// update (norms -> Adam on V -> rescale), so the tensor is read once and written once; taller
// matrices than NRF * 256 rows fall back to three passes over memory.
constexpr int NRF = 6;
__host__ __device__ __forceinline__ int adam_ct(int rows) { return rows > 128 ? 4 : 8; }

// sum over the row lanes of a block, per column: lanes of a warp that share a column first, then
// the 32 warp partials through shared memory.  Result valid in threads with ry == 0 (tid < ct).
template <int NV>
__device__ __forceinline__ void col_reduce(float (&val)[NV], float (*red)[32][8], const int ct, const int tid) {
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    for (int o = ct; o < 32; o <<= 1) val[i] += __shfl_xor_sync(0xffffffffu, val[i], o);
    if (lane < ct) red[i][wid][lane] = val[i];
  }
  __syncthreads();
  // warp c sums the 32 warp partials of column c with a shuffle tree (a serial loop over the 32 partials by
  // one thread per column was ~1 us per reduction, three times per launch on the step's tail)
  if (wid < ct) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = red[i][lane][wid];
      a = warp_sum(a);
      if (lane == 0) red[i][0][wid] = a;
    }
  }
  __syncthreads();
  if (tid < ct) {
#pragma unroll
    for (int i = 0; i < NV; ++i) val[i] = red[i][0][tid];
  }
}

template <bool P2P>
__global__ void __launch_bounds__(NTH) adamwn_kernel(const AdamPlan pl, float* __restrict__ W,
                                                     const float* __restrict__ G,
                                                     float* __restrict__ state, const double lr,
                                                     const double b1d, const double b2d,
                                                     const float eps, const float gscale,
                                                     const int weightnorm, const PeerSet ps,
                                                     const int block_base, const int advance,
                                                     float* __restrict__ loss_mirror) {
  // (the programmatic-dependent wait comes later: everything up to the first gradient read touches only
  //  state that no kernel of this step writes -- this range's W, m, v, V_scaler, the iteration counter)
  int tnow_p2p = 0;
  if (P2P) {
    pdl_wait();
    if (ps.flags) {
      // This launch is stream-ordered behind the kernels that produced bucket `slot` of this rank, so its first
      // block publishes the bucket: the step number goes into every peer's flag block over NVLink.  The wait
      // for the peers' signals comes later (p2p_wait_peers), after the local operands have been fetched, so
      // the one-way NVLink latency of the hand-shake overlaps useful loads.  A rank's signal never depends on
      // another rank's progress, so the waits cannot deadlock.
      tnow_p2p = *reinterpret_cast<const int*>(state + 2 * pl.P + 3 * (int64_t)pl.NC + 2) + 1;
      if (blockIdx.x == 0 && threadIdx.x < ps.n) {
        __threadfence_system();
        st_release_sys(flag_ready(ps.flags[threadIdx.x], ps.slot, ps.rank), tnow_p2p);
      }
    }
  }
  // every peer has published bucket `slot` of THIS step: poll LOCAL memory (the peers store into my block)
  auto p2p_wait_peers = [&]() {
    if (P2P && ps.flags) {
      if (threadIdx.x < ps.n) {
        // relaxed polls (an acquire per poll is a system-scope fence each time), one fence after the last
        const volatile int32_t* f = flag_ready(ps.flags[ps.rank], ps.slot, threadIdx.x);
        while (*f < tnow_p2p) { }
        __threadfence_system();
      }
      __syncthreads();
    }
  };
  const bool red_losses = P2P && ps.reduce_losses && blockIdx.x == 0;
  if (red_losses && threadIdx.x < 8) {     // (after the wait above: every block passes exactly one p2p_wait_peers)
    float lv = 0.f;
    for (int p = 0; p < ps.n; ++p) lv += ps.peers[p][pl.P + threadIdx.x];
    ps.loss_out[threadIdx.x] = lv;
  }
  // the advancing (last) launch of a step also mirrors the 8 loss scalars that follow the gradients
  // to `loss_mirror` (host-mapped pinned memory: the caller then needs a stream sync, not a D2H copy)
  __shared__ float red[2][32][8];
  __shared__ float col[4][8];
  __shared__ float lr_s;
  float* m = state;
  float* v = state + pl.P;
  float* vsc = state + 2 * pl.P;
  float* mg = vsc + pl.NC;
  float* vg = mg + pl.NC;
  int* fct = reinterpret_cast<int*>(vg + pl.NC);          // [t the cached factor is for | factor]
  int* iter = fct + 2;
  unsigned* done = reinterpret_cast<unsigned*>(iter + 1);
  const int tid = threadIdx.x;

  // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t): the double-precision factor is computed once per step by
  // the block that advances `iterations` (for the NEXT step) and only re-derived here if the cache
  // is for another t (first step) -- fp64 pow on every thread was a measurable part of this kernel
  const int t = *iter + 1;
  if (tid == 0) {
    float f;
    if (fct[0] == t) f = __int_as_float(fct[1]);
    else f = (float)(sqrt(1.0 - pow(b2d, (double)t)) / (1.0 - pow(b1d, (double)t)));
    lr_s = (float)lr * f;
  }
  __syncthreads();
  const float lr_t = lr_s;
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

  if (rows == 0 || !weightnorm) {  // plain Adam on a flat chunk
    const int64_t n = (rows == 0) ? cols : (int64_t)rows * cols;
    const int64_t i = (int64_t)lb * NTH + tid;
    float w0 = 0.f, m0 = 0.f, v0 = 0.f;
    if (i < n) { w0 = W[off + i]; m0 = m[off + i]; v0 = v[off + i]; }
    if (!P2P) pdl_wait();
    p2p_wait_peers();
    if (i < n) {
      const float g = load_grad<P2P>(G, ps, off + i) * gscale;
      const float mt = b1 * m0 + (1.0f - b1) * g;
      const float vt = b2 * v0 + (1.0f - b2) * g * g;
      m[off + i] = mt;
      v[off + i] = vt;
      W[off + i] = w0 - lr_t * mt / (sqrtf(vt) + eps);
    }
  } else {
    const int ct = adam_ct(rows), rl = NTH / ct;
    const int cx = tid & (ct - 1), ry = tid / ct;
    const int c = lb * ct + cx;
    const bool cv = c < cols;
    const int sc = pl.coloff[ti] + c;
    const float vs = cv ? vsc[sc] : 1.f;
    const bool in_regs = rows <= NRF * rl;
    float Wr[NRF], Gr[NRF], Mr[NRF], Vr[NRF];
    // phase 1: ||V||^2 and <G, V> per column
    float s2[2] = {0.f, 0.f};
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < NRF; ++u) {
        const int r = ry + u * rl;
        Wr[u] = Gr[u] = Mr[u] = Vr[u] = 0.f;
        if (cv && r < rows) {
          const int64_t e = off + (int64_t)r * cols + c;
          Wr[u] = W[e] / vs; Mr[u] = m[e]; Vr[u] = v[e];
        }
      }
      if (!P2P) pdl_wait();      // the gradients are the only operands the predecessor writes
      p2p_wait_peers();
#pragma unroll
      for (int u = 0; u < NRF; ++u) {
        const int r = ry + u * rl;
        if (cv && r < rows) {
          const int64_t e = off + (int64_t)r * cols + c;
          const float graw = load_grad<P2P>(G, ps, e);
          if (P2P) ps.gsum[e] = graw;
          Gr[u] = graw * gscale;
        }
      }
#pragma unroll
      for (int u = 0; u < NRF; ++u) { s2[0] = fmaf(Wr[u], Wr[u], s2[0]); s2[1] = fmaf(Gr[u], Wr[u], s2[1]); }
    } else {
      if (!P2P) pdl_wait();
      p2p_wait_peers();
    }
    if (!in_regs && cv) {
#pragma unroll 4
      for (int r = ry; r < rows; r += rl) {
        const int64_t e = off + (int64_t)r * cols + c;
        const float graw = load_grad<P2P>(G, ps, e);
        if (P2P) ps.gsum[e] = graw;
        const float V = W[e] / vs, g = graw * gscale;
        s2[0] = fmaf(V, V, s2[0]);
        s2[1] = fmaf(g, V, s2[1]);
      }
    }
    col_reduce(s2, red, ct, tid);
    if (ry == 0) {
      const float V_norm = sqrtf(s2[0]);
      const float grad_g = s2[1] / V_norm;
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
    // phase 2: Adam on V
    float s1[1] = {0.f};
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < NRF; ++u) {
        const int r = ry + u * rl;
        if (cv && r < rows) {
          const int64_t e = off + (int64_t)r * cols + c;
          const float gV = vs * (Gr[u] - gg_over_norm * Wr[u]);
          const float mt = b1 * Mr[u] + (1.0f - b1) * gV;
          const float vt = b2 * Vr[u] + (1.0f - b2) * gV * gV;
          m[e] = mt;
          v[e] = vt;
          Wr[u] = Wr[u] - lr_t * mt / (sqrtf(vt) + eps);
          s1[0] = fmaf(Wr[u], Wr[u], s1[0]);
        }
      }
    } else if (cv) {
      // new V parked in W.  Rows in explicit batches of 4 (all loads, then all stores) because
      // W/m/v may alias as far as the compiler knows: without batching every iteration would wait a
      // full memory round trip behind the previous iteration's stores.
      for (int r0 = ry; r0 < rows; r0 += 4 * rl) {
        float Wv[4], Gv[4], Mv[4], Vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rl;
          if (r < rows) {
            const int64_t e = off + (int64_t)r * cols + c;
            Wv[u] = W[e]; Gv[u] = P2P ? ps.gsum[e] : G[e]; Mv[u] = m[e]; Vv[u] = v[e];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rl;
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
            s1[0] = fmaf(nV, nV, s1[0]);
          }
        }
      }
    }
    __syncthreads();   // red[] is reused
    col_reduce(s1, red, ct, tid);
    if (ry == 0) {
      const float ns = col[1][cx] / sqrtf(s1[0]);
      if (cv) vsc[sc] = ns;
      col[2][cx] = ns;
    }
    __syncthreads();
    // phase 3: W = V_scaler' * V'
    const float ns = col[2][cx];
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < NRF; ++u) {
        const int r = ry + u * rl;
        if (cv && r < rows) W[off + (int64_t)r * cols + c] = Wr[u] * ns;
      }
    } else if (cv) {
      for (int r0 = ry; r0 < rows; r0 += 4 * rl) {
        float Wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rl;
          if (r < rows) Wv[u] = W[off + (int64_t)r * cols + c];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * rl;
          if (r < rows) W[off + (int64_t)r * cols + c] = Wv[u] * ns;
        }
      }
    }
  }
  if (red_losses && threadIdx.x < 8) {     // (after the wait above: every block passes exactly one p2p_wait_peers)
    float lv = 0.f;
    for (int p = 0; p < ps.n; ++p) lv += ps.peers[p][pl.P + threadIdx.x];
    ps.loss_out[threadIdx.x] = lv;
  }
  // the advancing (last) launch of a step also mirrors the 8 loss scalars that follow the gradients
  // to `loss_mirror` (host-mapped pinned memory: the caller then needs a stream sync, not a D2H copy)
  if (loss_mirror && advance && blockIdx.x == 0 && threadIdx.x < 8)
    loss_mirror[threadIdx.x] = P2P ? ps.loss_out[threadIdx.x] : G[pl.P + threadIdx.x];
  // last block to finish advances `iterations` (only the final launch of a step is told to) and
  // caches the bias-correction factor of the next step
  if (!advance) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned d = atomicAdd(done, 1u);
    if (d == (unsigned)advance - 1u) {       // advance = number of blocks of the step's final launch(es)
      const float f = (float)(sqrt(1.0 - pow(b2d, (double)(t + 1))) / (1.0 - pow(b1d, (double)(t + 1))));
      fct[1] = __float_as_int(f);
      fct[0] = t + 1;
      __threadfence();
      *iter = t;
      *done = 0u;
      if (P2P && ps.flags) {
        // this rank has finished reading every peer's gradients of step t (the other buckets' kernels were
        // joined before this launch): tell the peers their buffers may be overwritten
        __threadfence_system();
        for (int p = 0; p < ps.n; ++p) st_release_sys(flag_done(ps.flags[p], ps.rank), t);
      }
    }
  }
}

// ---- one-shot all-reduce over peer memory (clv_p2p_allreduce) ------------------------------------------------
// gsum[e] = sum_p peers[p][e] for e in [e0, e0 + cnt), summed in rank order (bitwise identical on every rank).
// The launch is stream-ordered behind the kernels that produced this rank's bucket: block 0 publishes the bucket
// (step number into every peer's flag block), every block waits for the peers' flags by polling LOCAL memory,
// then the grid streams the peers' buffers over NVLink with NP independent 16-byte loads in flight per thread.
// `last`: the final bucket of a step -- the block that finishes last tells the peers that this rank no longer
// reads their gradients ("done" flags, waited for by clv_p2p_wait_done before the next step zeroes them).
//
// No fences (each system-scope fence measured 1.7 us, profiles/nvl_probe_r2.txt; the first version of this
// kernel had four in series and took 11-16 us): the gradients were written by EARLIER kernels of the stream,
// so they are in this GPU's L2 -- the point of coherence for peer accesses -- before this kernel starts, and a
// plain system-scope store of the flag is enough (the start-of-collective barrier of NCCL's LL protocol and of
// vLLM's custom all-reduce rely on the same property); the consumer reads the peers' data with volatile loads
// (never from a stale L1 line), issued after it has seen the flag.  The "done" flag only orders READS that
// have already returned their values.  The polls are bounded (~2 minutes): a peer that never arrives
// traps the kernel instead of hanging the GPU for good.
constexpr int AR_NTH = 512;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(int32_t* p, const int v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int NP>
__global__ void __launch_bounds__(AR_NTH) p2p_allreduce_kernel(const float* const* __restrict__ peers,
                                                               int32_t* const* __restrict__ flags, const int n,
                                                               const int rank, const int slot, const int* iter,
                                                               float* __restrict__ gsum, const int64_t e0,
                                                               const int64_t cnt, const int last) {
  __shared__ const float* pp[NP];
  pdl_wait();
  const int t = *iter + 1;
  const int tid = threadIdx.x;
  // diagnostics (block 0): ns spent waiting for the peers / moving data, summed per slot in the local flag row
  const bool diag = blockIdx.x == 0 && tid == 0;
  unsigned long long tg0 = 0, tg1 = 0;
  if (diag) tg0 = globaltimer_ns();
  if (tid < NP) pp[tid] = tid < n ? peers[tid] : nullptr;
  if (blockIdx.x == 0 && tid < n) st_relaxed_sys(flag_ready(flags[tid], slot, rank), t);
  if (tid < n) {
    const volatile int32_t* f = flag_ready(flags[rank], slot, tid);
    const long long t0 = clock64();
    while (*f < t) {
      if (clock64() - t0 > 240000000000ll) __trap();     // ~2 minutes
    }
  }
  __syncthreads();
  if (diag) tg1 = globaltimer_ns();
  pdl_launch_dependents();      // the update that follows may start fetching its W / m / v
  // head up to the first 16-byte boundary, 16-byte body, tail
  const int64_t head = min(cnt, (int64_t)((4 - (e0 & 3)) & 3));
  const int64_t nv = (cnt - head) >> 2;
  const int64_t tail0 = head + (nv << 2);
  const int64_t gt = (int64_t)blockIdx.x * AR_NTH + tid, gs = (int64_t)gridDim.x * AR_NTH;
  for (int64_t i = gt; i < nv; i += gs) {
    const int64_t e = e0 + head + (i << 2);
    float4 v[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p)
      if (p < n) v[p] = ld_volatile_f4(pp[p] + e);
    float4 a = v[0];
#pragma unroll
    for (int p = 1; p < NP; ++p)
      if (p < n) { a.x += v[p].x; a.y += v[p].y; a.z += v[p].z; a.w += v[p].w; }
    *reinterpret_cast<float4*>(gsum + e) = a;
  }
  if (gt < head + (cnt - tail0)) {
    const int64_t e = e0 + (gt < head ? gt : tail0 + (gt - head));
    float a = *(const volatile float*)(pp[0] + e);
    for (int p = 1; p < n; ++p) a += *(const volatile float*)(pp[p] + e);
    gsum[e] = a;
  }
  if (diag && slot < 4) {
    int32_t* d = flag_local(flags[rank], 1 + 3 * slot);
    atomicAdd(d, (int)(tg1 - tg0));
    atomicAdd(d + 1, (int)(globaltimer_ns() - tg1));
    atomicAdd(d + 2, 1);
  }
  if (!last) return;
  __syncthreads();              // every load of this block has returned (its value went into a store)
  if (tid == 0) {
    int32_t* cntr = flag_local(flags[rank], 0);
    if (atomicAdd(cntr, 1) == last - 1) {     // last = number of blocks of the step's final bucket launch(es)
      *cntr = 0;
      for (int p = 0; p < n; ++p) st_relaxed_sys(flag_done(flags[p], rank), t);
    }
  }
}

// ---- two-shot all-reduce through the NVSwitch multicast mapping (NVLS); opt-in (clv_p2p_args.mc_*) ------------
// The one-shot kernel above makes every rank read all N-1 peers' buckets (N=8: copy 11-13 us per bucket).  Here
// rank r owns 1/N of the bucket: one multimem.ld_reduce per 16 bytes fetches that chunk from EVERY rank's
// gradient buffer and adds it in the switch, one multimem.st writes the sum into every rank's gsum -- a rank
// moves 1/N of a bucket in and out, whatever N is.  Every element has exactly one reducer, so the replicas stay
// bit-identical.  Two hand-shakes: "bucket ready" before the reads (as above, fence-free) and "slice landed"
// after the writes: each block fences its multicast stores at system scope, the block that finishes last tells
// every peer, and stays until every peer's slice has landed here -- so the kernel completes only when this
// rank's gsum is whole and the Adam-WN update behind it needs no further check.
// Measured: 7-9 us per bucket whatever N (one-shot: 4.5 us at N=2, 11-13 us at N=8); step time equal at N=8
// (0.178 vs 0.177 ms), worse at N=2 (0.176 vs 0.158) -- the two fences and the second hand-shake cost what the
// smaller transfer saves, so the one-shot form stays the default.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float multimem_ld_reduce_f1(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f1(float* mc, const float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

__global__ void __launch_bounds__(AR_NTH) p2p_allreduce_mc_kernel(const float* __restrict__ mc_grads,
                                                                  float* __restrict__ mc_gsum,
                                                                  int32_t* const* __restrict__ flags, const int n,
                                                                  const int rank, const int slot, const int* iter,
                                                                  const int64_t e0, const int64_t cnt, const int last) {
  pdl_wait();
  const int t = *iter + 1;
  const int tid = threadIdx.x;
  const bool diag = blockIdx.x == 0 && tid == 0;
  unsigned long long tg0 = 0, tg1 = 0;
  if (diag) tg0 = globaltimer_ns();
  if (blockIdx.x == 0 && tid < n) st_relaxed_sys(flag_ready(flags[tid], slot, rank), t);
  if (tid < n) {
    const volatile int32_t* f = flag_ready(flags[rank], slot, tid);
    const long long t0 = clock64();
    while (*f < t) {
      if (clock64() - t0 > 240000000000ll) __trap();
    }
  }
  __syncthreads();
  if (diag) tg1 = globaltimer_ns();
  pdl_launch_dependents();
  const int64_t head = min(cnt, (int64_t)((4 - (e0 & 3)) & 3));
  const int64_t nv = (cnt - head) >> 2;
  const int64_t tail0 = head + (nv << 2);
  const int64_t lo = nv * rank / n, hi = nv * (rank + 1) / n;      // this rank's 16-byte chunks
  const int64_t gt = (int64_t)blockIdx.x * AR_NTH + tid, gs = (int64_t)gridDim.x * AR_NTH;
  for (int64_t i = lo + gt; i < hi; i += gs) {
    const int64_t e = e0 + head + (i << 2);
    multimem_st_f4(mc_gsum + e, multimem_ld_reduce_f4(mc_grads + e));
  }
  if (rank == 0 && gt < head + (cnt - tail0)) {                      // unaligned ends: rank 0, one float each
    const int64_t e = e0 + (gt < head ? gt : tail0 + (gt - head));
    multimem_st_f1(mc_gsum + e, multimem_ld_reduce_f1(mc_grads + e));
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();                 // this block's multicast stores are performed everywhere
    int32_t* cl = flag_local(flags[rank], 13 + slot);
    const bool last_block = atomicAdd(cl, 1) == (int)gridDim.x - 1;
    if (diag && slot < 3) {
      int32_t* d = flag_local(flags[rank], 1 + 3 * slot);
      atomicAdd(d, (int)(tg1 - tg0));
      atomicAdd(d + 1, (int)(globaltimer_ns() - tg1));
      atomicAdd(d + 2, 1);
    }
    if (last_block) {
      *cl = 0;
      __threadfence_system();
      for (int p = 0; p < n; ++p) st_relaxed_sys(flag_landed(flags[p], slot, rank), t);
      const long long t0 = clock64();
      for (int p = 0; p < n; ++p) {
        const volatile int32_t* f = flag_landed(flags[rank], slot, p);
        while (*f < t) {
          if (clock64() - t0 > 240000000000ll) __trap();
        }
      }
      __threadfence_system();               // (the peers' slices are in this rank's memory before the kernel ends)
    }
    if (last) {
      // every read of the peers' gradients by this block has returned (its value went into a store)
      int32_t* cntr = flag_local(flags[rank], 0);
      if (atomicAdd(cntr, 1) == last - 1) {
        *cntr = 0;
        for (int p = 0; p < n; ++p) st_relaxed_sys(flag_done(flags[p], rank), t);
      }
    }
  }
}

// bucket `slot` of this rank's gradient buffer is final: publish the step number to every peer's flag block
__global__ void p2p_signal_kernel(int32_t* const* flags, const int n, const int rank, const int slot,
                                  const int* iter) {
  pdl_wait();
  const int t = *iter + 1;
  __threadfence_system();
  if ((int)threadIdx.x < n) st_release_sys(flag_ready(flags[threadIdx.x], slot, rank), t);
}
// start of a step: no peer is still reading this rank's gradient buffer of the previous step
__global__ void p2p_wait_done_kernel(int32_t* const* flags, const int n, const int rank, const int* iter) {
  const int tprev = *iter;
  if ((int)threadIdx.x < n) {
    const int32_t* f = flag_done(flags[rank], threadIdx.x);
    while (ld_acquire_sys(f) < tprev) __nanosleep(64);
  }
}

__global__ void adamwn_init_kernel(float* state, int64_t P, int NC) {
  const int64_t n = 2 * P + 3 * (int64_t)NC + 4;
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
    if (pl->rows[i] > 0 && weightnorm) nb += (pl->cols[i] + adam_ct(pl->rows[i]) - 1) / adam_ct(pl->rows[i]);
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
  return 2 * pl.P + 3 * (int64_t)pl.NC + 4;
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
                                     int32_t t_last, int32_t advance, float* loss_mirror, void* stream) {
  if (!cfg || !params || !grads || !state) return CLV_E_INVALID;
  if (t_first < 0 || t_last > CLV_N_TENSORS || t_first >= t_last) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, weightnorm);
  if (rc != CLV_OK) return rc;
  PeerSet ps = {nullptr, 0, nullptr, nullptr, nullptr, 0, 0, 0};
  const int nb = pl.first_block[t_last] - pl.first_block[t_first];
  if (nb <= 0) return CLV_OK;
  CLV_CUDA(clv_launch(adamwn_kernel<false>, nb, NTH, 0, (cudaStream_t)stream, pl, params, grads, state, lr,
                      beta_1, beta_2, (float)epsilon, (float)grad_scale, (int)weightnorm, ps,
                      pl.first_block[t_first], advance == 1 ? nb : (int)advance, loss_mirror));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_adamwn_step(const clv_cfg* cfg, float* params, const float* grads, float* state,
                               double lr, double beta_1, double beta_2, double epsilon,
                               double grad_scale, int32_t weightnorm, void* stream) {
  return clv_adamwn_step_range(cfg, params, grads, state, lr, beta_1, beta_2, epsilon, grad_scale,
                               weightnorm, 0, CLV_N_TENSORS, 1, nullptr, stream);
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
  PeerSet ps = {peer_grads, n_peers, gsum, loss_out, nullptr, 0, 0, 1};
  adamwn_kernel<true><<<pl.first_block[CLV_N_TENSORS], NTH, 0, (cudaStream_t)stream>>>(
      pl, params, nullptr, state, lr, beta_1, beta_2, (float)epsilon, 1.0f, weightnorm, ps, 0,
      pl.first_block[CLV_N_TENSORS], nullptr);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

// ---- peer-memory data parallelism with the hand-shake inside the kernels (clv_p2p_args) ------------------
extern "C" int clv_p2p_flag_ints(void) { return (2 * P2P_SLOTS + 2) * P2P_MAXR; }

extern "C" int clv_p2p_signal(const clv_p2p_args* pp, const float* state, const clv_cfg* cfg, int32_t slot,
                              void* stream) {
  if (!pp || !state || !cfg || slot < 0 || slot >= P2P_SLOTS || pp->n_peers < 1 || pp->n_peers > P2P_MAXR)
    return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, 1);
  if (rc != CLV_OK) return rc;
  const int* iter = reinterpret_cast<const int*>(state + 2 * pl.P + 3 * (int64_t)pl.NC + 2);
  CLV_CUDA(clv_launch(p2p_signal_kernel, 1, 32, 0, (cudaStream_t)stream, pp->peer_flags, (int)pp->n_peers,
                      (int)pp->rank, (int)slot, iter));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_p2p_wait_done(const clv_p2p_args* pp, const float* state, const clv_cfg* cfg, void* stream) {
  if (!pp || !state || !cfg || pp->n_peers < 1 || pp->n_peers > P2P_MAXR) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, 1);
  if (rc != CLV_OK) return rc;
  const int* iter = reinterpret_cast<const int*>(state + 2 * pl.P + 3 * (int64_t)pl.NC + 2);
  p2p_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp->peer_flags, (int)pp->n_peers, (int)pp->rank, iter);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_p2p_allreduce_blocks(int64_t count) {
  int nb = (int)((count / 4 + AR_NTH - 1) / AR_NTH);
  return nb < 1 ? 1 : (nb > clv_num_sms() ? clv_num_sms() : nb);
}

extern "C" int clv_adamwn_range_blocks(const clv_cfg* cfg, int32_t weightnorm, int32_t t_first, int32_t t_last) {
  if (!cfg || t_first < 0 || t_last > CLV_N_TENSORS || t_first >= t_last) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, weightnorm);
  if (rc != CLV_OK) return rc;
  return pl.first_block[t_last] - pl.first_block[t_first];
}

extern "C" int clv_p2p_allreduce(const clv_p2p_args* pp, const float* state, const clv_cfg* cfg, int64_t first,
                                 int64_t count, int32_t slot, int32_t last, void* stream) {
  if (!pp || !state || !cfg || !pp->peer_grads || !pp->peer_flags || !pp->gsum) return CLV_E_INVALID;
  if (slot < 0 || slot >= P2P_SLOTS || pp->n_peers < 1 || pp->n_peers > P2P_MAXR || pp->rank < 0 ||
      pp->rank >= pp->n_peers)
    return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, 1);
  if (rc != CLV_OK) return rc;
  if (first < 0 || count <= 0 || first + count > pl.P + 8) return CLV_E_INVALID;
  const int* iter = reinterpret_cast<const int*>(state + 2 * pl.P + 3 * (int64_t)pl.NC + 2);
  // enough threads for one 16-byte chunk each, at most one CTA per SM (the remote loads are latency-bound:
  // a wide grid with few chunks per thread finishes in about one NVLink round trip per chunk batch)
  const int nb = clv_p2p_allreduce_blocks(count);
  const int n = pp->n_peers;
  if (last == 1) last = nb;
  if (pp->mc_grads && pp->mc_gsum && slot < 3) {
    CLV_CUDA(clv_launch(p2p_allreduce_mc_kernel, nb, AR_NTH, 0, (cudaStream_t)stream, pp->mc_grads, pp->mc_gsum,
                        pp->peer_flags, n, (int)pp->rank, (int)slot, iter, first, count, (int)last));
    CLV_CHECK_LAUNCH();
    return CLV_OK;
  }
#define CLV_AR(NP)                                                                                             \
  CLV_CUDA(clv_launch(p2p_allreduce_kernel<NP>, nb, AR_NTH, 0, (cudaStream_t)stream, pp->peer_grads,            \
                      pp->peer_flags, n, (int)pp->rank, (int)slot, iter, pp->gsum, first, count, (int)last))
  if (n <= 2) CLV_AR(2);
  else if (n <= 4) CLV_AR(4);
  else if (n <= 8) CLV_AR(8);
  else CLV_AR(16);
#undef CLV_AR
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}

extern "C" int clv_adamwn_step_range_p2p(const clv_cfg* cfg, float* params, const clv_p2p_args* pp, float* state,
                                         double lr, double beta_1, double beta_2, double epsilon,
                                         int32_t weightnorm, int32_t t_first, int32_t t_last, int32_t slot,
                                         int32_t advance, float* loss_mirror, void* stream) {
  if (!cfg || !params || !pp || !state || !pp->peer_grads || !pp->peer_flags || !pp->gsum || !pp->loss_out)
    return CLV_E_INVALID;
  if (t_first < 0 || t_last > CLV_N_TENSORS || t_first >= t_last || slot < 0 || slot >= P2P_SLOTS)
    return CLV_E_INVALID;
  if (pp->n_peers < 1 || pp->n_peers > P2P_MAXR || pp->rank < 0 || pp->rank >= pp->n_peers) return CLV_E_INVALID;
  AdamPlan pl;
  int rc = make_plan(cfg, &pl, weightnorm);
  if (rc != CLV_OK) return rc;
  PeerSet ps = {pp->peer_grads, pp->n_peers, pp->gsum, pp->loss_out, pp->peer_flags, pp->rank, slot,
                t_last == CLV_N_TENSORS ? 1 : 0};

  const int nb = pl.first_block[t_last] - pl.first_block[t_first];
  if (nb <= 0) return CLV_OK;
  CLV_CUDA(clv_launch(adamwn_kernel<true>, nb, NTH, 0, (cudaStream_t)stream, pl, params, (const float*)nullptr,
                      state, lr, beta_1, beta_2, (float)epsilon, 1.0f, (int)weightnorm, ps,
                      pl.first_block[t_first], advance == 1 ? nb : (int)advance, loss_mirror));
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
