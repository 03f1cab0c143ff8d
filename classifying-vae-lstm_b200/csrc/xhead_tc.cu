// K4x (tensor-core form): the X head at large batch -- logits = h_dec @ Kx + bx, the 88-key Bernoulli loss,
// dLoss/dlogits and dh = dlogits @ Kx^T (cl_vrnn/model.py:229-234,241-242 and their TF-autodiff backward) -- as
// TWO chained tcgen05 GEMMs per 128-row tile with the loss math between them:
//   GEMM1  logits[128,88]  = h[128,88]       * Kx      accumulator 1 in TMEM
//   epilogue 1 (8 warps)   : TMEM -> registers, + bias, sigmoid / clip / Keras-BCE / dlogits, dlogits -> global (transposed through smem: 128-byte row
//                            segments) AND -> the A-operand tile in shared memory, as bf16 hi + mid
//   GEMM2  dh[128,88]      = dlogits[128,88] * Kx^T    accumulator 2 in TMEM
//   epilogue 2             : TMEM -> registers -> smem transpose -> global
//   GEMM3  gKx[88,88] (+ the bias gradient) += h^T * dlogits over the tile's 128 rows: both tiles are already in
//          shared memory and are read a second time MN-major; accumulator 3 stays in TMEM across all tiles of the
//          CTA and is added to the gradient buffer with red.add at the end (replaces a 0.68 ms SIMT GEMM + a 0.06
//          ms column sum at 524 k rows)
// fp32 operands are split into bf16 hi + mid (2 x 8 mantissa bits) and three products are accumulated in fp32
// (hi*hi + hi*mid + mid*hi): relative error ~2^-16 per operand pair, 2e-6 of max|dlogits| on the results
// (tested at 1e-4 like every other kernel).  Warp-specialised persistent CTAs: 4 producer warps load and
// split the h tile (its first half prefetched in registers while the previous tile owns the buffer), one MMA
// thread issues 2 x 18 + 24 MMAs per tile, 12 epilogue warps (the loss epilogue of tile i+1 runs during GEMM2 of tile i); both accumulators double-buffered in TMEM, so the MMA of tile i+1
// runs under the epilogues of tile i.  Replaces clv_xhead_fwd_bwd's SIMT kernels from 256 tiles up.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int TM = 128;              // rows per tile (UMMA M)
constexpr int XD = 88;               // D == H == 88
constexpr int NP = 96;               // N padded (UMMA N, multiple of 16)
constexpr int KP = 96;               // K padded
constexpr int LBO = 128;
constexpr int SBO = (KP / 8) * 128;  // 1536
constexpr int A_SPLIT = TM * KP * 2;           // 24 576
constexpr int A_STAGE = 2 * A_SPLIT;           // hi + mid
constexpr int B_SPLIT = NP * KP * 2;           // 18 432
constexpr int B_IMG = 2 * B_SPLIT;             // hi + mid of one matrix
constexpr int NEPI = 12;            // epilogue warps: 4 TMEM quadrants x 3 column parts of 32
constexpr int STAGE_FLOATS = 32 * 20;
constexpr int SMEM_BYTES = 2 * A_STAGE + 2 * B_IMG + NEPI * STAGE_FLOATS * 4 + 512 + 1024;
constexpr int THREADS = 17 * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float ex2_approx(const float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(const float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(const float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// fp32 pair -> packed bf16 hi pair and bf16 mid pair (x0 in the low half: the lower address)
__device__ __forceinline__ void split2(const float x0, const float x1, uint32_t& hi, uint32_t& mid) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 m = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  mid = *reinterpret_cast<const uint32_t*>(&m);
}

// ---- weight prep: Kx[k][d] fp32 -> two canonical K-major images, each bf16 hi + mid:
//   image 0 (GEMM1): N = d, K = k  -> element (n, kk) = Kx[kk][n]
//   image 1 (GEMM2): N = k, K = d  -> element (n, kk) = Kx[n][kk]
__global__ void xsplit_kernel(const float* __restrict__ Kx, __nv_bfloat16* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * NP * KP) return;
  const int which = idx / (NP * KP), r = idx - which * NP * KP;
  const int n = r / KP, kk = r - n * KP;
  float w = 0.f;
  if (n < XD && kk < XD) w = which == 0 ? __ldg(Kx + kk * XD + n) : __ldg(Kx + n * XD + kk);
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 mid = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t off = (size_t)(n >> 3) * (SBO / 2) + (size_t)(kk >> 3) * (LBO / 2) + (n & 7) * 8 + (kk & 7);
  __nv_bfloat16* base = img + (size_t)which * (B_IMG / 2);
  base[off] = hi;
  base[off + B_SPLIT / 2] = mid;
}

struct XArgs {
  const float* h; const float* bx; const __nv_bfloat16* img;
  const uint8_t* roll; const int32_t* x_off; int x_grp, x_shift;
  float* loss_acc; float* dlogits; float* dh; float* gKx; float* gbx;
  int64_t R; float scale; int tiles;
};

// MN-major view of a canonical K-major tile (GEMM3 reduces over the ROWS of both tiles): element (mn, k = row)
// sits at (row / 8) * SBO + (mn / 8) * LBO + (row % 8) * 16 + (mn % 8) * 2, i.e. the MN-major canonical form
// with the two strides swapped: "leading" (between k-groups) = SBO, "stride" (between mn-groups) = LBO.
__device__ __forceinline__ constexpr uint32_t umma_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

//   warps 0-3        producers : h tile fp32 -> bf16 hi/mid, canonical A tile
//   warps 4-7, 9-16  epilogue  : quadrant q = warp & 3 (TMEM lanes 32q..32q+31), column part cpart (32 columns)
//   warp  8          MMA       : weight images by bulk copy, then the MMAs of the three GEMMs
// TMEM columns: accumulator 1 (logits) at 0 / 96, accumulator 2 (dh) at 192 / 288 (both ping-pong), accumulator 3
// (the weight gradient, summed over all tiles of the CTA) at 384.
__global__ void __launch_bounds__(THREADS, 1) xhead_tc_kernel(const XArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a1_s = smem;                                  // h tile       (hi | mid)
  uint8_t* a2_s = smem + A_STAGE;                        // dlogits tile (hi | mid)
  uint8_t* b_s = smem + 2 * A_STAGE;                     // image 0 (hi | mid), image 1 (hi | mid)
  float* stage_all = reinterpret_cast<float*>(smem + 2 * A_STAGE + 2 * B_IMG);
  float* bias_s = stage_all + NEPI * STAGE_FLOATS;       // [96]
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum { B_FULL = 0, A1_FULL = 1, A1_EMPTY = 2, A2_FULL = 3, A2_EMPTY = 4, ACC1_FULL = 5, ACC1_EMPTY = 7,
         ACC2_FULL = 9, ACC2_EMPTY = 11, ACC3_FULL = 13 };
  const bool wg = a.gKx != nullptr;                       // GEMM3: weight and bias gradients

  if (tid == 0) {
    mbar_init(BAR(B_FULL), 1);
    mbar_init(BAR(A1_FULL), 4);
    mbar_init(BAR(A1_EMPTY), 1);
    mbar_init(BAR(A2_FULL), NEPI);
    mbar_init(BAR(A2_EMPTY), 1);
    mbar_init(BAR(ACC3_FULL), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(ACC1_FULL + s), 1);
      mbar_init(BAR(ACC1_EMPTY + s), NEPI);
      mbar_init(BAR(ACC2_FULL + s), 1);
      mbar_init(BAR(ACC2_EMPTY + s), NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < NP) bias_s[tid] = tid < XD ? __ldg(a.bx + tid) : 0.f;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int ntile = (a.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  pdl_wait();                    // (the weight image and h come from predecessors)
  pdl_launch_dependents();

  if (warp < 4) {
    // ================= producers: thread = row of the tile
    const int row = tid;
    for (int it = 0; it < ntile; ++it) {
      const int64_t m = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + row;
      // (the row in two halves of 44 floats: 22 float4 in flight at once cost registers the 12 epilogue warps
      //  need; the first half is in flight while the previous tile still owns the buffer)
      const bool mv = m < a.R;
      const float4* src = reinterpret_cast<const float4*>(a.h + (mv ? m : 0) * XD);
      float4 v[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) v[j] = mv ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(BAR(A1_EMPTY), (it & 1) ^ 1);     // GEMM1 and GEMM3 of the previous tile have read the buffer
      uint8_t* dst = a1_s + (row >> 3) * SBO + (row & 7) * 16;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 1) {
#pragma unroll
          for (int j = 0; j < 10; ++j) v[j] = mv ? __ldg(src + 12 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < (hf == 0 ? 6 : 5); ++j) {
          uint4 hi, mid;
          split2(v[2 * j].x, v[2 * j].y, hi.x, mid.x);
          split2(v[2 * j].z, v[2 * j].w, hi.y, mid.y);
          split2(v[2 * j + 1].x, v[2 * j + 1].y, hi.z, mid.z);
          split2(v[2 * j + 1].z, v[2 * j + 1].w, hi.w, mid.w);
          *reinterpret_cast<uint4*>(dst + (6 * hf + j) * LBO) = hi;
          *reinterpret_cast<uint4*>(dst + A_SPLIT + (6 * hf + j) * LBO) = mid;
        }
      }
      // k = 88..95: zero padding, except k = 88 := 1.0 (bf16 0x3F80) -- GEMM1 multiplies it with a zero weight
      // row, GEMM3 turns it into the column sums of dlogits (the bias gradient) as row 88 of its result
      *reinterpret_cast<uint4*>(dst + 11 * LBO) = make_uint4(0x3F80u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(dst + A_SPLIT + 11 * LBO) = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(A1_FULL));
    }
  } else if (warp != 8) {
    // ================= epilogue warps
    const int q = warp & 3, cpart = warp < 8 ? 0 : (warp < 13 ? 1 : 2);
    float* stage = stage_all + (q + 4 * cpart) * STAGE_FLOATS;
    const int cbase = 32 * cpart;                 // this warp's 32 columns: [cbase, cbase + 32)
    float lsum = 0.f;
    const float hi_p = 1.0f - CLV_EPS;
    const float lo_l = logf(CLV_EPS / (1.0f - CLV_EPS)), hi_l = logf(hi_p / (1.0f - hi_p));
    // transposed store of a 16-column chunk held row-per-lane: 4 lanes per 64-byte row segment
    auto store_chunk16 = [&](const float* vals, float* gbase, const int64_t m0, const int rows_valid, const int c0,
                             const int nvalid) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(stage + lane * 20 + 4 * i) =
            make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
      __syncwarp();
      const int rs = lane >> 2, cc = (lane & 3) * 4;
      if (cc < nvalid) {
#pragma unroll
        for (int rr = 0; rr < 32; rr += 8) {
          if (rr + rs < rows_valid)
            *reinterpret_cast<float4*>(gbase + (m0 + rr + rs) * XD + c0 + cc) =
                *reinterpret_cast<const float4*>(stage + (rr + rs) * 20 + cc);
        }
      }
      __syncwarp();
    };
    auto epi1 = [&](const int it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + q * 32;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.R - m0));
      const int64_t m = m0 + lane;
      const bool rv = m < a.R;
      // this row's target bits (32 bytes of the roll row, 8-byte aligned)
      uint2 xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xb[i] = make_uint2(0u, 0u);
      if (rv) {
        const uint32_t mu = (uint32_t)m, g = mu / (uint32_t)a.x_grp;
        const uint8_t* xr = a.roll + ((int64_t)__ldg(a.x_off + g) + a.x_shift + (mu - g * a.x_grp)) * XD + cbase;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (cbase + 8 * i < XD) xb[i] = __ldg(reinterpret_cast<const uint2*>(xr + 8 * i));
      }
      mbar_wait(BAR(ACC1_FULL + s), ph);
      mbar_wait(BAR(A2_EMPTY), (it & 1) ^ 1);     // GEMM2 and GEMM3 of the previous tile have read the dlogits tile
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t1 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 96) + (uint32_t)cbase;
      uint8_t* adst = a2_s + ((q * 32 + lane) >> 3) * SBO + ((q * 32 + lane) & 7) * 16;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {             // 2 chunks of 16 columns
        uint32_t r[16];
        tmem_ld16(t1 + 16 * ch, r);
        float dl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int d = cbase + 16 * ch + i;
          const uint2 w = xb[(16 * ch + i) >> 3];
          const uint32_t word = ((i & 7) < 4) ? w.x : w.y;
          const float x = (float)((word >> (8 * (i & 3))) & 0xffu);
          float v = 0.f;
          if (rv && d < XD) {
            // Keras-BCE of the clipped sigmoid, written on the LOGIT: clip(p, eps, 1-eps) <=> clip(z, lo, hi) with
            // lo/hi the logits of the bounds, so one exponential serves the sigmoid and the softplus term
            // (ex2/rcp/lg2.approx: abs. error < 3e-7 on p and on the loss term; the SIMT kernel's expf/logf/
            // log1pf chain cost ~100 instructions per element and made the epilogue warps the bottleneck)
            const float z = __uint_as_float(r[i]) + bias_s[d];
            const float l = fminf(fmaxf(z, lo_l), hi_l);
            const float e = ex2_approx(-fabsf(l) * 1.4426950408889634f);
            const float rc = rcp_approx(1.0f + e);
            const float pc = l >= 0.f ? rc : e * rc;
            lsum += fmaxf(l, 0.f) - l * x + lg2_approx(1.0f + e) * 0.6931471805599453f;
            v = (z >= lo_l && z <= hi_l) ? a.scale * (pc - x) : 0.f;
          }
          dl[i] = v;
        }
        // A-operand tile of GEMM2 (and B operand of GEMM3): k = d, 8 values per 16-byte store
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 hi, mid;
          split2(dl[8 * g8 + 0], dl[8 * g8 + 1], hi.x, mid.x);
          split2(dl[8 * g8 + 2], dl[8 * g8 + 3], hi.y, mid.y);
          split2(dl[8 * g8 + 4], dl[8 * g8 + 5], hi.z, mid.z);
          split2(dl[8 * g8 + 6], dl[8 * g8 + 7], hi.w, mid.w);
          const int j = (cbase + 16 * ch + 8 * g8) >> 3;
          *reinterpret_cast<uint4*>(adst + j * LBO) = hi;
          *reinterpret_cast<uint4*>(adst + A_SPLIT + j * LBO) = mid;
        }
        const int c0 = cbase + 16 * ch;
        if (c0 < XD) store_chunk16(dl, a.dlogits, m0, rows_valid, c0, min(16, XD - c0));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(BAR(ACC1_EMPTY + s)); mbar_arrive(BAR(A2_FULL)); }
    };
    auto epi2 = [&](const int it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TM + q * 32;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.R - m0));
      mbar_wait(BAR(ACC2_FULL + s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t2 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(192 + s * 96) + (uint32_t)cbase;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int c0 = cbase + 16 * ch;
        if (c0 >= XD) break;
        uint32_t r[16];
        tmem_ld16(t2 + 16 * ch, r);
        float vals[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) vals[i] = __uint_as_float(r[i]);
        store_chunk16(vals, a.dh, m0, rows_valid, c0, min(16, XD - c0));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(ACC2_EMPTY + s));
    };
    // software pipeline: the loss epilogue of tile i+1 runs while GEMM2 / GEMM3 of tile i are in flight
    if (ntile > 0) epi1(0);
    for (int it = 0; it < ntile; ++it) {
      if (it + 1 < ntile) epi1(it + 1);
      epi2(it);
    }
    lsum = warp_sum(lsum);
    if (lane == 0) atomicAdd(a.loss_acc, lsum * a.scale);
    // ---------- weight / bias gradient: accumulator 3, lane = row of gKx (k), lane 88 = the bias gradient
    if (wg && ntile > 0) {
      mbar_wait(BAR(ACC3_FULL), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int k = q * 32 + lane;
      const uint32_t t3 = tmem + ((uint32_t)(q * 32) << 16) + 384u + (uint32_t)cbase;
      if (q < 3) {                                  // (lanes 96..127 hold nothing)
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = cbase + 16 * ch;
          if (c0 >= XD) break;
          uint32_t r[16];
          tmem_ld16(t3 + 16 * ch, r);
          float* dstp = k < XD ? a.gKx + (int64_t)k * XD : (k == XD ? a.gbx : nullptr);
          if (dstp) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < XD) atomicAdd(dstp + c0 + i, __uint_as_float(r[i]));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  } else {
    // ================= MMA warp (one elected thread)
    if (lane == 0) {
      mbar_expect_tx(BAR(B_FULL), 2 * B_IMG);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.img);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        bulk_g2s(smem_u32(b_s) + i * B_SPLIT, src + (size_t)i * B_SPLIT, B_SPLIT, BAR(B_FULL));
      mbar_wait(BAR(B_FULL), 0);
      const uint32_t idesc = umma_idesc(TM, NP), idesc3 = umma_idesc_mn(TM, NP);
      const uint32_t b_addr = smem_u32(b_s), a1 = smem_u32(a1_s), a2 = smem_u32(a2_s);
      // three products: (A hi, B hi), (A hi, B mid), (A mid, B hi)
      auto gemm = [&](const uint32_t a_addr, const uint32_t bimg, const uint32_t tacc) {
        uint32_t acc = 0;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const uint32_t ao = a_addr + (pr == 2 ? A_SPLIT : 0);
          const uint32_t bo = bimg + (pr == 1 ? B_SPLIT : 0);
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk) {
            umma_bf16(tacc, umma_desc(ao + kk * 2 * LBO, LBO, SBO), umma_desc(bo + kk * 2 * LBO, LBO, SBO), idesc, acc);
            acc = 1;
          }
        }
      };
      uint32_t acc3 = 0;
      auto issue1 = [&](const int it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(BAR(A1_FULL), it & 1);
        mbar_wait(BAR(ACC1_EMPTY + s), ph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm(a1, b_addr, tmem + (uint32_t)(s * 96));
        umma_commit(BAR(ACC1_FULL + s));
      };
      auto issue23 = [&](const int it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(BAR(A2_FULL), it & 1);
        mbar_wait(BAR(ACC2_EMPTY + s), ph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm(a2, b_addr + B_IMG, tmem + (uint32_t)(192 + s * 96));
        umma_commit(BAR(ACC2_FULL + s));
        if (wg) {
          // GEMM3: gKx[k][d] += sum_rows h[row][k] * dlogits[row][d] -- both tiles read MN-major (strides swapped),
          // K = the 128 rows of the tile in 8 steps of 16; M = 128 covers k = 0..95 (+ 32 don't-care rows)
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t ao = a1 + (pr == 2 ? A_SPLIT : 0);
            const uint32_t bo = a2 + (pr == 1 ? A_SPLIT : 0);
#pragma unroll
            for (int kk = 0; kk < TM / 16; ++kk) {
              umma_bf16(tmem + 384u, umma_desc(ao + kk * 2 * SBO, SBO, LBO), umma_desc(bo + kk * 2 * SBO, SBO, LBO),
                        idesc3, acc3);
              acc3 = 1;
            }
          }
        }
        umma_commit(BAR(A1_EMPTY));
        umma_commit(BAR(A2_EMPTY));
      };
      if (ntile > 0) issue1(0);
      for (int it = 0; it < ntile; ++it) {
        issue23(it);
        if (it + 1 < ntile) issue1(it + 1);
      }
      if (wg && ntile > 0) umma_commit(BAR(ACC3_FULL));
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int64_t clv_xhead_tc_scratch_bytes(void) { return 2 * (int64_t)B_IMG; }

extern "C" int clv_xhead_tc(const float* h, const float* Kx, const float* bx, const uint8_t* roll,
                            const int32_t* x_off, int32_t x_grp, int32_t x_shift, float* loss_acc,
                            float* dlogits, float* dh, float* gKx, float* gbx, void* scratch, int64_t R, int32_t H,
                            int32_t D, float scale, void* stream) {
  if (!h || !Kx || !bx || !roll || !x_off || !loss_acc || !dlogits || !dh || !scratch || x_grp <= 0)
    return CLV_E_INVALID;
  if ((gKx == nullptr) != (gbx == nullptr)) return CLV_E_INVALID;
  if (H != XD || D != XD || R >= (1LL << 32) || (((uintptr_t)h | (uintptr_t)dlogits | (uintptr_t)dh |
                                                   (uintptr_t)scratch) & 15) || ((uintptr_t)roll & 7))
    return CLV_E_UNSUPPORTED;
  if (R <= 0) return CLV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(scratch);
  xsplit_kernel<<<(2 * NP * KP + 255) / 256, 256, 0, st>>>(Kx, img);
  CLV_CHECK_LAUNCH();
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(xhead_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[attr_set_dev] = true;
  }
  XArgs a;
  a.h = h; a.bx = bx; a.img = img; a.roll = roll; a.x_off = x_off; a.x_grp = x_grp; a.x_shift = x_shift;
  a.loss_acc = loss_acc; a.dlogits = dlogits; a.dh = dh; a.gKx = gKx; a.gbx = gbx; a.R = R; a.scale = scale;
  a.tiles = (int)((R + TM - 1) / TM);
  int gx = clv_num_sms();
  if (gx > a.tiles) gx = a.tiles;
  xhead_tc_kernel<<<gx, THREADS, SMEM_BYTES, st>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
