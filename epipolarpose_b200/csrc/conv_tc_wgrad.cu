// tcgen05 weight-gradient kernel of the tap-list implicit GEMM (include/epb.h):
//
//   dw[co][wt[t]][ci] += sum_m dout[pix_out(m)][co] * f(in[pix_in(m, t)][ci])
//
// GEMM view: D[co, ci] with the reduction over pixels m.  Both operands are
// "MN-major" for the tensor core (channels are contiguous in NHWC memory, the
// reduction index m strides over pixel rows), so no transposition is needed:
// producer warps gather 32 pixel rows per stage (dout rows and the tap-shifted
// input rows with the fused BatchNorm+ReLU of the producing layer), split them
// into TF32 hi (+ lo for 3xTF32) and store them in the SWIZZLE_128B_BASE32B
// MN-major layout (128-byte rows = 32 consecutive channels of one pixel; 4 pixel
// rows form a 512-byte atom; the only layout tcgen05 takes for 32-bit MN-major).  One thread issues tcgen05.mma kind::tf32 with
// M = 128 (co), N = ci tile, K = 8 pixels; FP32 accumulators live in TMEM for
// the CTA's whole pixel range (split-K over pixels across CTAs), then the
// epilogue adds the partial tile to dw with vector reductions
// (red.global.add.v4.f32).
//
// Work item = (co tile, tap, ci tile, pixel split); one work item per CTA.
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WM = 128;          // co rows per tile (UMMA M); rows >= Cout are zero
constexpr int KPIX = 32;         // pixels per pipeline stage (4 MMAs of K = 8)
constexpr int kProdWarps = 8;
constexpr int kThreadsW = 32 * (kProdWarps + 1);   // 8 producer/epilogue warps + 1 MMA warp

template <int BNW, int NS>
struct WCfg {
  static constexpr int PL = (NS == 3) ? 2 : 1;
  static constexpr int A_BYTES = WM * KPIX * 4 * PL;     // dout tile  (hi[,lo])
  static constexpr int B_BYTES = BNW * KPIX * 4 * PL;    // input tile (hi[,lo])
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int S_ = (200 * 1024) / STAGE;
  static constexpr int S = S_ > 6 ? 6 : S_;
  static constexpr int TMEM_COLS = BNW <= 32 ? 32 : (BNW <= 64 ? 64 : 128);
  static constexpr int SMEM = S * STAGE + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// Store one float4 (4 consecutive channels `ch4*4..` of pixel row r) into the MN-major
// tile.  For 32-bit MN-major operands tcgen05 accepts only the SWIZZLE_128B_BASE32B
// layout: 128-byte rows (32 consecutive channels of one pixel), pixel rows at a 128-byte
// pitch, atoms of 4 rows (512 B), and the 32-byte chunk index (address bits 5-6) XORed with
// the row index mod 4 (address bits 7-8).  chunk c = ch4 / 8 selects the 32-channel column
// block (LBO apart), ch4 % 8 the 16-byte slot inside the row.
__device__ __forceinline__ void st_mn(uint8_t* tile, int r, int ch4, float4 v) {
  const int slot = ch4 & 7;
  const uint32_t off = (uint32_t)(ch4 >> 3) * (KPIX * 128) + (uint32_t)r * 128u +
                       (uint32_t)((((slot >> 1) ^ (r & 3)) << 5) | ((slot & 1) << 4));
  *reinterpret_cast<float4*>(tile + off) = v;
}

template <int BNW, int NS>
__global__ void __launch_bounds__(kThreadsW, 1)
conv_wgrad_tc_kernel(const __grid_constant__ epb_conv_geom g, const float* __restrict__ in,
                     const float* __restrict__ dout, const float* __restrict__ in_scale,
                     const float* __restrict__ in_shift, float* __restrict__ dw, int co_tiles,
                     int ci_tiles, int rows_per_split) {
  using C = WCfg<BNW, NS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint8_t* ctrl = sm + C::S * C::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);       // full[S], empty[S], done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ctrl + 8 * 16);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (6 + s); };
  const uint32_t done_bar = bar0 + 8u * 12;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item decode
  int w = blockIdx.x;
  const int cit = w % ci_tiles; w /= ci_tiles;
  const int t = w % g.T; w /= g.T;
  const int cot = w % co_tiles;
  const int split = w / co_tiles;
  const int64_t M = (int64_t)g.N * g.Hp * g.Wp;
  const int64_t mbeg = (int64_t)split * rows_per_split;
  const int64_t mend = mbeg + rows_per_split < M ? mbeg + rows_per_split : M;
  const int nstages = (int)((mend - mbeg + KPIX - 1) / KPIX);
  const int co0 = cot * WM, ci0 = cit * BNW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::S; ++s) {
      tc::mbar_init(full_bar(s), kProdWarps);
      tc::mbar_init(empty_bar(s), 1);
    }
    tc::mbar_init(done_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == kProdWarps) tc::tmem_alloc<C::TMEM_COLS>(tc::smem_u32(tmem_ptr));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kProdWarps) {
    // ======================================== producers: 4 pixel rows per warp per stage
    const int dh = g.dh[t], dwv = g.dw[t];
    const int co = co0 + lane * 4;            // this lane's 4 dout channels
    const bool co_ok = co < g.Cout;
    const int ci = ci0 + lane * 4;            // this lane's 4 input channels
    const bool ci_ok = (lane * 4 < BNW) && ci < g.Cin;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_scale && ci_ok) {
      sc = *reinterpret_cast<const float4*>(in_scale + ci);
      sh = *reinterpret_cast<const float4*>(in_shift + ci);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int st = 0; st < nstages; ++st) {
      float4 a[4], b[4];
      bool bok[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t m = mbeg + (int64_t)st * KPIX + warp * 4 + q;
        a[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        b[q] = a[q];
        bok[q] = false;
        if (m < mend) {
          const int j = (int)(m % g.Wp);
          const int i = (int)((m / g.Wp) % g.Hp);
          const int n = (int)(m / ((int64_t)g.Wp * g.Hp));
          if (co_ok)
            a[q] = *reinterpret_cast<const float4*>(
                dout + (((int64_t)n * g.Ho + (i * g.os + g.ph)) * g.Wo + (j * g.os + g.pw)) * g.Cout + co);
          const int ih = i * g.is + dh, iw = j * g.is + dwv;
          if (ci_ok && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi) {
            b[q] = *reinterpret_cast<const float4*>(
                in + (((int64_t)n * g.Hi + ih) * g.Wi + iw) * g.Cin + ci);
            bok[q] = true;
          }
        }
      }
      tc::mbar_wait(empty_bar(stage), phase ^ 1);
      uint8_t* a_hi = sm + stage * C::STAGE;
      uint8_t* b_hi = a_hi + C::A_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = warp * 4 + q;
        {
          const float4 x = a[q];
          const float4 hi = make_float4(tc::to_tf32(x.x), tc::to_tf32(x.y), tc::to_tf32(x.z),
                                        tc::to_tf32(x.w));
          st_mn(a_hi, r, lane, hi);
          if (NS == 3)
            st_mn(a_hi + WM * KPIX * 4, r, lane,
                  make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
        }
        if (lane * 4 < BNW) {
          float4 x = b[q];
          if (in_scale && bok[q]) {
            x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y);
            x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
            if (g.in_relu) {
              x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f);
              x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
            }
          }
          const float4 hi = make_float4(tc::to_tf32(x.x), tc::to_tf32(x.y), tc::to_tf32(x.z),
                                        tc::to_tf32(x.w));
          st_mn(b_hi, r, lane, hi);
          if (NS == 3)
            st_mn(b_hi + BNW * KPIX * 4, r, lane,
                  make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(full_bar(stage));
      if (++stage == C::S) { stage = 0; phase ^= 1; }
    }
    // ======================================== epilogue (warps 0-3: one TMEM lane quarter each)
    if (warp < 4) {
      tc::mbar_wait(done_bar, 0);
      tc::tc_fence_after();
      const int row = warp * 32 + lane;
      const int corow = co0 + row;
      const int64_t wrow = (int64_t)g.Tw * g.Cin;
#pragma unroll 1
      for (int chunk = 0; chunk < BNW / 32; ++chunk) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + chunk * 32 + ((uint32_t)(warp * 32) << 16), r);
        tc::tmem_ld_wait();
        const int c0 = ci0 + chunk * 32;
        if (corow < g.Cout && c0 < g.Cin) {
          float* dst = dw + (int64_t)corow * wrow + (int64_t)g.wt[t] * g.Cin + c0;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c0 + c < g.Cin)
              red_add_v4(dst + c, __uint_as_float(r[c]), __uint_as_float(r[c + 1]),
                         __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        }
      }
    }
  } else {
    // ======================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(WM, BNW, 1, 1);     // both operands MN-major
      constexpr uint32_t LBO = KPIX * 128, SBO = 512;   // chunk stride, 4-row k-atom stride
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nstages; ++st) {
        tc::mbar_wait(full_bar(stage), phase);
        tc::tc_fence_after();
        const uint32_t a_hi = base + stage * C::STAGE;
        const uint32_t b_hi = a_hi + C::A_BYTES;
#pragma unroll
        for (int ks = 0; ks < KPIX / 8; ++ks) {
          const uint64_t ah = tc::desc_mnmajor_sw128(a_hi + ks * 1024, LBO, SBO);
          const uint64_t bh = tc::desc_mnmajor_sw128(b_hi + ks * 1024, LBO, SBO);
          if (NS == 3) {
            const uint64_t al = tc::desc_mnmajor_sw128(a_hi + WM * KPIX * 4 + ks * 1024, LBO, SBO);
            const uint64_t bl = tc::desc_mnmajor_sw128(b_hi + BNW * KPIX * 4 + ks * 1024, LBO, SBO);
            tc::mma_tf32(tmem_base, al, bh, idesc, (st | ks) != 0);
            tc::mma_tf32(tmem_base, ah, bl, idesc, 1);
            tc::mma_tf32(tmem_base, ah, bh, idesc, 1);
          } else {
            tc::mma_tf32(tmem_base, ah, bh, idesc, (st | ks) != 0);
          }
        }
        tc::mma_commit(empty_bar(stage));
        if (++stage == C::S) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit(done_bar);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == kProdWarps) {
    tc::tc_fence_after();
    tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BNW, int NS>
int launch_wgrad(const epb_conv_geom* g, const float* in, const float* dout, const float* in_scale,
                 const float* in_shift, float* dw, cudaStream_t st) {
  using C = WCfg<BNW, NS>;
  const int64_t M = (int64_t)g->N * g->Hp * g->Wp;
  const int co_tiles = (g->Cout + WM - 1) / WM;
  const int ci_tiles = (g->Cin + BNW - 1) / BNW;
  const int64_t tiles = (int64_t)co_tiles * ci_tiles * g->T;
  int64_t splits = (2 * kNumSMs + tiles - 1) / tiles;
  const int64_t max_splits = (M + 8 * KPIX - 1) / (8 * KPIX);     // >= 8 stages per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rows = (M + splits - 1) / splits;
  rows = (rows + KPIX - 1) / KPIX * KPIX;
  splits = (M + rows - 1) / rows;
  static bool attr_set = false;
  if (!attr_set) {
    EPB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BNW, NS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  const int64_t grid = tiles * splits;
  EPB_CHECK_ARG(grid < (1LL << 31));
  conv_wgrad_tc_kernel<BNW, NS><<<(unsigned)grid, kThreadsW, C::SMEM, st>>>(
      *g, in, dout, in_scale, in_shift, dw, co_tiles, ci_tiles, (int)rows);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

}  // namespace

bool epb_conv_wgrad_tc_supported(const epb_conv_geom* g) {
  return g->Cin % 32 == 0 && g->Cout % 4 == 0 && g->Cout >= 32;
}

int epb_conv_wgrad_tc(const epb_conv_geom* g, const float* in, const float* dout,
                      const float* in_scale, const float* in_shift, float* dw, cudaStream_t st) {
  const int ns = g->precision == 3 ? 3 : 1;
  const int bn = g->Cin >= 128 ? 128 : (g->Cin >= 64 ? 64 : 32);
#define EPB_WG_CASE(BN_, NS_) \
  if (bn == BN_ && ns == NS_) return launch_wgrad<BN_, NS_>(g, in, dout, in_scale, in_shift, dw, st);
  EPB_WG_CASE(32, 1) EPB_WG_CASE(64, 1) EPB_WG_CASE(128, 1)
  EPB_WG_CASE(32, 3) EPB_WG_CASE(64, 3) EPB_WG_CASE(128, 3)
#undef EPB_WG_CASE
  epb_set_error("no tcgen05 wgrad configuration for Cin=%d", g->Cin);
  return EPB_EINVAL;
}
