// K1 (tensor-core form, backward): the LSTM weight gradients that reduce over all B*L rows,
//     dK_x[D,4H]  = X^T      @ dA        (piano-roll columns of the input kernel)
//     dU  [H,4H]  = Hprev^T  @ dA        (recurrent kernel; Hprev[b,t] = h[b,t-1], 0 at t=0)
//     dK_z[Z,4H]  = Zs^T     @ dA        (decoder only)
// as ONE tcgen05 kernel: both operands are MN-major (the reduction index = row is the slow axis in
// HBM), accumulators live in TMEM ([128 x 176] x 2 per CTA), the row range is split over CTAs and
// partial results are added into the (pre-zeroed) gradient buffer with coalesced red.add.
// Precision (fp32 inputs, bf16 tensor cores, fp32 accumulate):
//     X is {0,1}: exact.  dA = hi + mid (16 mantissa bits).  h, Zs = hi + mid;  products kept:
//     hi*hi + hi*mid + mid*hi  -> relative error ~2^-16 per term, far inside the 1e-4 gradient bound.
// Replaces the TF-autodiff MatMul-gradients of the LSTM kernels (cl_vrnn/model.py:196-199,225-228).
#include <cuda_bf16.h>
#include "common.cuh"

#ifdef CLV_PROF
__device__ long long g_wprof[16];
extern "C" int clv_debug_wprof(long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_wprof, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_wprof, z, sizeof(z)); }
  return 0;
}
#define WPROF(i, cond) do { if ((cond) && blockIdx.x == 0 && blockIdx.y == 0) g_wprof[i] = clock64(); } while (0)
#else
#define WPROF(i, cond)
#endif

namespace {

constexpr int WM = 128, KS = 32;                    // UMMA M and rows per slab (= K of one slab)
// the 4H = 352 gate columns are split over two CTAs as 192 + 160 (both multiples of 32: whole
// epilogue chunks and whole groups of four 8-column chunks for the skewed converter below)
constexpr int WN0 = 192, WN1 = 160, WNMAX = 192, GFULL = 352;
constexpr int LBO = 128;                            // between the k-groups (8 rows) of a core column
constexpr int SBO = (KS / 8) * 128;                 // between mn-groups (8 columns)
constexpr int A_TILE = (WM / 8) * SBO;              // 8 192 B
constexpr int B_TILE = (WNMAX / 8) * SBO;           // 12 288 B
constexpr int STAGE = 3 * A_TILE + 2 * B_TILE;      // X | Hhi | Hmid | Dhi | Dmid = 49 152 B
constexpr int NSTAGE = 2;
// fp32 staging ring filled by the bulk-copy (TMA) engine with TWO copies per 32-row slab: the dA rows
// (full 352-column rows are contiguous: 45 056 B, of which this CTA converts its column range) and
// the 32 h_{t-1} rows (11 264 B).  History: lane-per-row global loads cost one L1 sector request per
// lane (3.8 k cycles per slab); one bulk copy per row segment cost ~65 cycles of issue EACH (4.1 k
// cycles per slab, measured with clock64); two large copies per slab are issue-free.
constexpr int RAW_D = KS * GFULL * 4;               // 45 056 B
constexpr int RAW_H = KS * 88 * 4;                  // 11 264 B
constexpr int RAW_STAGE = RAW_D + RAW_H;            // 56 320 B
constexpr int NRAW = 2;
constexpr int NPROD = 8;                            // converter warps
constexpr int WTHREADS = (NPROD + 2) * 32;          // + MMA warp + bulk-copy warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// MN-major, no swizzle, version 1: element (mn, k) at (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
// kind::f16, D=f32, A=B=bf16, A and B MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t umma_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 x8 -> bf16 hi (round-to-nearest) and mid (residual), packed as 2 x 16 bytes; packed
// two-at-a-time conversions (F2FP) instead of scalar F2F, which runs at a quarter of the rate
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp);
    const float r0 = v[2 * i] - __uint_as_float(hb << 16);
    const float r1 = v[2 * i + 1] - __uint_as_float(hb & 0xffff0000u);
    const __nv_bfloat162 mp = __floats2bfloat162_rn(r0, r1);
    h[i] = hb;
    m[i] = *reinterpret_cast<const uint32_t*>(&mp);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

struct WgArgs {
  const float* dA; const uint8_t* roll; const int32_t* off; const float* h; const float* Zs;
  float* gKx; float* gU; float* gKz;
  int64_t R; int L, shift, D, H, Z, G;
  int stages, stages_per_cta, nsplit;
};

__global__ void __launch_bounds__(WTHREADS, 1) lstm_wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 2 * NRAW + 1];   // full, empty, raw_full, raw_empty, done
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // column range of this CTA: 192|160 (2-way) or 96|96|96|64 (4-way, small M: the epilogue's red.add
  // rate is ~1.6 elements/clk per SM whatever the CTA count, so narrower tiles on more SMs finish sooner)
  const int ny = blockIdx.y;
  const int n0 = (a.nsplit == 2) ? ny * WN0 : ny * 96;
  const int wn = (a.nsplit == 2) ? (ny ? WN1 : WN0) : (ny == 3 ? 64 : 96);
  WPROF(0, tid == 0);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto RAW_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto RAW_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NRAW + s); };
  const uint32_t DONE = bar0 + 8u * (2 * NSTAGE + 2 * NRAW);
  uint8_t* raw = smem + NSTAGE * STAGE;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL(s), NPROD); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < NRAW; ++s) { mbar_init(RAW_FULL(s), 1); mbar_init(RAW_EMPTY(s), NPROD); }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the padding column groups of the three A tiles once (M rows beyond D / H+Z): they are never
  // written by the producers
  {
    const int gx0 = a.D / 8, gh0 = a.H / 8 + (a.Zs ? 1 : 0);
    for (int i = tid; i < NSTAGE * 3 * (WM / 8) * (SBO / 16); i += WTHREADS) {
      const int c16 = i % (SBO / 16), g = (i / (SBO / 16)) % (WM / 8);
      const int tile = (i / (SBO / 16) / (WM / 8)) % 3, s = i / (SBO / 16) / (WM / 8) / 3;
      if (g >= (tile == 0 ? gx0 : gh0))
        *reinterpret_cast<uint4*>(smem + s * STAGE + tile * A_TILE + g * SBO + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  WPROF(1, tid == 0);
  const int st0 = blockIdx.x * a.stages_per_cta;
  const int nst = max(0, min(a.stages_per_cta, a.stages - st0));
  const int ngx = a.D / 8, ngh = a.H / 8, ngz = a.Zs ? 1 : 0, ngb = wn / 8;   // 8-column chunks

  if (warp == NPROD + 1) {
    // ================= bulk-copy warp (one thread): two copies per slab
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int rs = it % NRAW, rph = (it / NRAW) & 1;
        const int64_t r0 = ((int64_t)(st0 + it)) * KS;
        const int nv = (int)min((int64_t)KS, a.R - r0);            // valid rows of the slab (>= 1)
        // h_{t-1} of row r is h row r-1: slab rows r0-1 .. r0+nv-2 land in slots 0 .. nv-1 (the very
        // first slab has no row -1: slots 1.. are filled, slot 0 is a t == 0 row and ignored)
        const int hskip = (r0 == 0) ? 1 : 0;
        const uint32_t dbytes = (uint32_t)nv * GFULL * 4, hbytes = (uint32_t)(nv - hskip) * a.H * 4;
        mbar_wait(RAW_EMPTY(rs), rph ^ 1);
        mbar_expect_tx(RAW_FULL(rs), dbytes + hbytes);
        uint8_t* rb = raw + rs * RAW_STAGE;
        bulk_g2s(smem_u32(rb), a.dA + r0 * a.G, dbytes, RAW_FULL(rs));
        if (hbytes)
          bulk_g2s(smem_u32(rb + RAW_D + hskip * a.H * 4), a.h + (r0 - 1 + hskip) * a.H, hbytes, RAW_FULL(rs));
        WPROF(9 + it, it < 2);
      }
    }
  } else if (warp < NPROD) {
    // ================= converters: lane = row of the slab.  The staged rows are dense (pitch = 0 mod
    // 128 B), so a plain lane-per-row read of one 8-column chunk would be an 8-way bank conflict;
    // instead the 8 lanes of a quarter-warp work on a GROUP of 4 chunks in 4 rotations: lane i takes
    // chunk ((i >> 1) + rotation) & 3 and reads its two 16-byte halves in the order given by i & 1,
    // which spreads both the reads and the MN-major 16-byte stores over all 32 banks.
    const int i8 = lane & 7;
    const int nsub_d = ngb, nsub_h = 4 * ((ngh + 3) / 4);    // sub-steps: 4 rotations per group of 4
    for (int it = 0; it < nst; ++it) {
      const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
      const int rs = it % NRAW, rph = (it / NRAW) & 1;
      const int64_t r = ((int64_t)(st0 + it)) * KS + lane;
      const bool rv = r < a.R;
      int64_t xrow = 0; bool tpos = false;
      if (rv) {
        const uint32_t ru = (uint32_t)r, b = ru / (uint32_t)a.L, t = ru - b * a.L;
        xrow = (int64_t)__ldg(a.off + b) + a.shift + t;
        tpos = t > 0;
      }
      // the small operands that still come through the LSU: piano-roll keys (88-byte rows are not a
      // bulk-copy size) and the latent columns; issued before the waits
      const int xt0 = warp, xt1 = warp + NPROD;     // roll chunks of this warp (ngx <= 16 = 2 NPROD)
      uint2 q0 = make_uint2(0u, 0u), q1 = make_uint2(0u, 0u);
      if (rv && xt0 < ngx) q0 = __ldg(reinterpret_cast<const uint2*>(a.roll + xrow * a.D + 8 * xt0));
      if (rv && xt1 < ngx) q1 = __ldg(reinterpret_cast<const uint2*>(a.roll + xrow * a.D + 8 * xt1));
      float zv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) zv[j] = 0.f;
      if (ngz && warp == NPROD - 1 && rv) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < a.Z) zv[j] = __ldg(a.Zs + r * a.Z + j);
      }
      mbar_wait(RAW_FULL(rs), rph);
      WPROF(11 + it, tid == 0 && it < 3);
      mbar_wait(EMPTY(s), ph ^ 1);
      uint8_t* sb = smem + s * STAGE + (lane >> 3) * LBO + (lane & 7) * 16;   // this row's slot
      const uint8_t* rd = raw + rs * RAW_STAGE + lane * (GFULL * 4) + n0 * 4;
      const uint8_t* rh = raw + rs * RAW_STAGE + RAW_D + lane * (88 * 4);
      // at most 36 sub-steps per slab = 5 per warp; fully unrolled so the five independent
      // load -> convert -> store chains overlap
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int u = warp + k * NPROD;
        if (u >= nsub_d + nsub_h) break;
        const bool isd = u < nsub_d;
        const int uu = isd ? u : u - nsub_d;
        const int c = 4 * (uu >> 2) + (((i8 >> 1) + uu) & 3);          // this lane's chunk
        const bool cvalid = c < (isd ? ngb : ngh);
        const bool live = cvalid && rv && (isd || tpos);
        const uint8_t* src = (isd ? rd : rh) + c * 32;
        const int first = i8 & 1;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (cvalid) {   // stale rows beyond R / t == 0 rows are read (conflict pattern) and discarded
          p0 = *reinterpret_cast<const float4*>(src + first * 16);
          p1 = *reinterpret_cast<const float4*>(src + (first ^ 1) * 16);
        }
        const float4 x = first ? p1 : p0, y = first ? p0 : p1;
        float v[8];
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        if (!live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        if (cvalid) {
          uint4 hi, mid;
          split8(v, hi, mid);
          uint8_t* dst = sb + (isd ? 3 * A_TILE : A_TILE) + c * SBO;
          *reinterpret_cast<uint4*>(dst) = hi;
          *reinterpret_cast<uint4*>(dst + (isd ? B_TILE : A_TILE)) = mid;
        }
      }
      if (ngz && warp == NPROD - 1) {     // the Z columns: one more chunk of the H tile
        uint4 hi, mid;
        split8(zv, hi, mid);
        *reinterpret_cast<uint4*>(sb + A_TILE + ngh * SBO) = hi;
        *reinterpret_cast<uint4*>(sb + 2 * A_TILE + ngh * SBO) = mid;
      }
      // piano-roll keys (exact in bf16)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int task = u ? xt1 : xt0;
        if (task < ngx) {
          const uint32_t x = u ? q1.x : q0.x, y = u ? q1.y : q0.y;
          const uint32_t w0 = (x & 1u) * 0x3F80u + ((x >> 8) & 1u) * 0x3F800000u;
          const uint32_t w1 = ((x >> 16) & 1u) * 0x3F80u + ((x >> 24) & 1u) * 0x3F800000u;
          const uint32_t w2 = (y & 1u) * 0x3F80u + ((y >> 8) & 1u) * 0x3F800000u;
          const uint32_t w3 = ((y >> 16) & 1u) * 0x3F80u + ((y >> 24) & 1u) * 0x3F800000u;
          *reinterpret_cast<uint4*>(sb + task * SBO) = make_uint4(w0, w1, w2, w3);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(FULL(s)); mbar_arrive(RAW_EMPTY(rs)); }
      WPROF(2 + it, tid == 0 && it < 4);
    }
  } else if (warp == NPROD && lane == 0) {
    // ================= MMA thread
    const uint32_t idesc = umma_idesc_mn(WM, wn);
    const bool has_x = a.gKx != nullptr;
    for (int it = 0; it < nst; ++it) {
      const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
      mbar_wait(FULL(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(smem + s * STAGE);
#pragma unroll
      for (int kk = 0; kk < KS / 16; ++kk) {
        const uint32_t ko = kk * 2 * LBO;
        const uint64_t xd = umma_desc(base + ko), hh = umma_desc(base + A_TILE + ko),
                       hm = umma_desc(base + 2 * A_TILE + ko), dh = umma_desc(base + 3 * A_TILE + ko),
                       dm = umma_desc(base + 3 * A_TILE + B_TILE + ko);
        const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
        if (has_x) {
          umma_bf16(tmem, xd, dh, idesc, acc);
          umma_bf16(tmem, xd, dm, idesc, 1u);
        }
        umma_bf16(tmem + 256, hh, dh, idesc, acc);
        umma_bf16(tmem + 256, hh, dm, idesc, 1u);
        umma_bf16(tmem + 256, hm, dh, idesc, 1u);
      }
      umma_commit(EMPTY(s));
    }
    umma_commit(DONE);
  }
  // ================= epilogue: everything joins; the 8 producer warps drain TMEM
  __syncwarp();
  mbar_wait(DONE, 0);
  WPROF(6, tid == 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < NPROD && nst > 0) {
    // warps w and w+4 share TMEM lane quadrant w%4 and split its column chunks; CTAs start at
    // different chunks / rows so that concurrent CTAs do not queue on the same L2 atomics
    // (red.add.v2 was measured slower than scalar red here: the L2 cost is per element)
    const int quad = warp & 3;
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 33);   // operand stages are free now
    const int NCH = wn / 32;
    for (int tile = (a.gKx ? 0 : 1); tile < 2; ++tile) {
      const uint32_t tacc = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tile * 256);
      const int mvalid = (tile == 0) ? a.D : a.H + a.Z;      // rows of this accumulator that exist
      if (quad * 32 >= mvalid) continue;
      const int nvalid = min(32, mvalid - quad * 32);
#pragma unroll 1
      for (int ci = (warp >> 2); ci < NCH; ci += 2) {
        const int c0 = 32 * ((ci + (int)blockIdx.x) % NCH);
        const int ncol = 32;
        uint32_t r[32];
        if (ncol == 32) {
          tmem_ld32(tacc + c0, r);
        } else {
          uint32_t r16[16];
          tmem_ld16(tacc + c0, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = __uint_as_float(r[i]);
        __syncwarp();
        const int rot = (5 * (int)blockIdx.x) % nvalid;
        if (lane < ncol) {
          int rr = rot;
          for (int i = 0; i < nvalid; ++i) {
            const int m = quad * 32 + rr;
            float* dst;
            if (tile == 0) dst = a.gKx + (int64_t)m * a.G;
            else if (m < a.H) dst = a.gU + (int64_t)m * a.G;
            else dst = a.gKz + (int64_t)(m - a.H) * a.G;
            atomicAdd(dst + n0 + c0 + lane, stage[rr * 33 + lane]);
            rr = (rr + 1 == nvalid) ? 0 : rr + 1;
          }
        }
        __syncwarp();
      }
    }
  }
  WPROF(7, tid == 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  WPROF(8, tid == 0);
  if (warp == NPROD) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int clv_lstm_wgrad_tc(const float* dA, const uint8_t* roll, const int32_t* win_off, int32_t L,
                                 int32_t shift, int32_t D, const float* h, const float* Zs, int32_t Z,
                                 float* gKx, float* gU, float* gKz, int64_t R, int32_t H, void* stream) {
  if (!dA || !h || !gU || L <= 0) return CLV_E_INVALID;
  if (gKx && (!roll || !win_off)) return CLV_E_INVALID;
  if (Zs && (!gKz || Z < 1)) return CLV_E_INVALID;
  if (H != 88 || (gKx && (D > WM || (D & 7))) || (Zs && Z > 8) || ((uintptr_t)dA & 15) || ((uintptr_t)h & 15) ||
      (gKx && ((uintptr_t)roll & 7)))
    return CLV_E_UNSUPPORTED;
  if (R <= 0) return CLV_OK;
  static bool attr_set[CLV_MAX_DEVICES] = {};   // per device: function attributes belong to a context
  const int attr_set_dev = clv_device_slot();
  const int smem = NSTAGE * STAGE + NRAW * RAW_STAGE + 1024;
  if (!attr_set[attr_set_dev]) {
    CLV_CUDA(cudaFuncSetAttribute(lstm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[attr_set_dev] = true;
  }
  WgArgs a;
  a.dA = dA; a.roll = roll; a.off = win_off; a.h = h; a.Zs = Zs; a.gKx = gKx; a.gU = gU; a.gKz = gKz;
  a.R = R; a.L = L; a.shift = shift; a.D = gKx ? D : 0; a.H = H; a.Z = Zs ? Z : 0; a.G = 4 * H;
  a.stages = (int)((R + KS - 1) / KS);
  // small problems: 4 column ranges x (>= 4 slabs per CTA); large ones: 2 column ranges x SMs/2 CTAs
  a.nsplit = (a.stages <= 16 * (clv_num_sms() / 4)) ? 4 : 2;
  int gx = a.stages / 4;
  const int cap = clv_num_sms() / a.nsplit;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  a.stages_per_cta = (a.stages + gx - 1) / gx;
  gx = (a.stages + a.stages_per_cta - 1) / a.stages_per_cta;
  lstm_wgrad_tc_kernel<<<dim3(gx, a.nsplit), WTHREADS, smem, (cudaStream_t)stream>>>(a);
  CLV_CHECK_LAUNCH();
  return CLV_OK;
}
