// tcgen05 weight-gradient kernel of the tap-list implicit GEMM (include/epb.h):
//
//   dw[co][wt[t]][ci] += sum_m dout[pix_out(m)][co] * f(in[pix_in(m, t)][ci])
//
// GEMM view: D[co, ci] with the reduction over pixels m.  Both operands are
// "MN-major" for the tensor core (channels are contiguous in NHWC memory, the
// reduction index m strides over pixel rows), so no transposition is needed:
// producer warps gather 32 pixel rows per tile (dout rows, and the tap-shifted
// input rows with the fused BatchNorm+ReLU of the producing layer), split them
// into TF32 hi (+ lo for 3xTF32) and store them in the SWIZZLE_128B_BASE32B
// MN-major layout (128-byte rows = 32 consecutive channels of one pixel; 4 pixel
// rows form a 512-byte atom; the only layout tcgen05 takes for 32-bit MN-major).
//
// One CTA = (co tile of 128, a group of NB "slots" = (tap, ci tile) pairs, a
// split of the pixel range).  The dout tile A(p) of pixel block p is produced
// ONCE and multiplied against the NB input tiles B(p, b) (two smem rings), each
// slot accumulating into its own TMEM columns (NB * BNW <= 512 fp32 columns), so
// dout is gathered once per NB taps instead of once per tap.  Four groups of four
// producer warps take tiles round-robin, so four tiles' global loads are in flight and
// every SM sub-partition has four warps to issue the convert/store work from.  One
// thread issues tcgen05.mma kind::tf32 (M = 128, N = BNW, K = 8 pixels); the
// epilogue adds the partial tiles to dw with red.global.add.v4.f32 (split-K).
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WM = 128;          // co rows per tile (UMMA M); rows >= Cout are zero
constexpr int KPIX = 32;         // pixels per tile (4 MMAs of K = 8)
constexpr int kProdWarps = 16;
constexpr int kGroups = 4;       // producer groups of 4 warps (one tile each, round-robin)
constexpr int kProd = 32 * kProdWarps;
constexpr int kThreadsW = kProd + 32;      // + 1 MMA warp
constexpr int kMaxSlots = 16;

template <int BNW, int NS>
struct WCfg {
  static constexpr int PL = (NS == 3) ? 2 : 1;
  static constexpr int A_TILE = WM * KPIX * 4 * PL;      // dout tile  (hi[,lo])
  static constexpr int B_TILE = BNW * KPIX * 4 * PL;     // input tile (hi[,lo])
  static constexpr int SA = 2;
  static constexpr int SB_ = (196 * 1024 - SA * A_TILE) / B_TILE;
  static constexpr int SB = SB_ > 8 ? 8 : SB_;
  static_assert(SB >= 2, "input ring too small");
  static constexpr int NB_MAX = (512 / BNW) > kMaxSlots ? kMaxSlots : (512 / BNW);
  static constexpr int SMEM = SA * A_TILE + SB * B_TILE + 1024 + 3072;
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

// Slots are awaited by mbarrier PARITY, which is ambiguous for a waiter two phases ahead of
// the barrier.  With tiles handed round-robin to four groups a group can get that far ahead,
// so every ring slot carries a use counter in smem: use k of a slot may start its parity
// wait only after use k-1 has PASSED its own wait (i.e. the release before last is known to
// have happened), and publishes k+1 once it has passed.
__device__ __forceinline__ void wait_seen(const volatile int* p, int need) {
  if (*p >= need) return;
  const long long t0 = clock64();
  while (*p < need)
    if (clock64() - t0 > 4000000000LL) __trap();
}

template <int BNW, int NS>
__global__ void __launch_bounds__(kThreadsW, 1)
conv_wgrad_tc_kernel(const __grid_constant__ epb_conv_geom g, const float* __restrict__ in,
                     const float* __restrict__ dout, const float* __restrict__ in_scale,
                     const float* __restrict__ in_shift, float* __restrict__ dw, int co_tiles,
                     int ci_tiles, int groups, int nb_per_group, int rows_per_split) {
  using C = WCfg<BNW, NS>;
  constexpr int QA = WM / 4;                 // float4 per dout row (32)
  constexpr int QB = BNW / 4;                // float4 per input row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint8_t* smB = sm + C::SA * C::A_TILE;
  uint8_t* ctrl = smB + C::SB * C::B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);   // fullA[2] emptyA[2] fullB[8] emptyB[8] done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ctrl + 8 * 24);
  int* slot_t = reinterpret_cast<int*>(ctrl + 256);          // [kMaxSlots] tap of each slot
  int* slot_cit = slot_t + kMaxSlots;                        // [kMaxSlots] ci tile of each slot
  int* rowtab = slot_cit + kMaxSlots;                        // [kGroups][4][KPIX]
  volatile int* seenA = rowtab + kGroups * 4 * KPIX;         // [SA] uses whose wait has passed
  volatile int* seenB = seenA + C::SA;                       // [SB]
  const uint32_t bar0 = tc::smem_u32(bars);
  auto fullA = [&](int s) { return bar0 + 8u * s; };
  auto emptyA = [&](int s) { return bar0 + 8u * (2 + s); };
  auto fullB = [&](int s) { return bar0 + 8u * (4 + s); };
  auto emptyB = [&](int s) { return bar0 + 8u * (12 + s); };
  const uint32_t done_bar = bar0 + 8u * 20;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item decode: blockIdx.x = ((split * co_tiles + cot) * groups + grp)
  int w = blockIdx.x;
  const int grp = w % groups; w /= groups;
  const int cot = w % co_tiles;
  const int split = w / co_tiles;
  const int total_slots = g.T * ci_tiles;
  const int slot0 = grp * nb_per_group;
  const int NB = min(nb_per_group, total_slots - slot0);
  const int M = g.N * g.Hp * g.Wp;
  const int mbeg = split * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  const int nblk = (mend - mbeg + KPIX - 1) / KPIX;
  const int co0 = cot * WM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::SA; ++s) { tc::mbar_init(fullA(s), kProdWarps / kGroups); tc::mbar_init(emptyA(s), 1); }
    for (int s = 0; s < C::SB; ++s) { tc::mbar_init(fullB(s), kProdWarps / kGroups); tc::mbar_init(emptyB(s), 1); }
    tc::mbar_init(done_bar, 1);
    tc::fence_barrier_init();
    for (int s = 0; s < C::SA; ++s) seenA[s] = 0;
    for (int s = 0; s < C::SB; ++s) seenB[s] = 0;
    for (int b = 0; b < NB; ++b) {
      slot_t[b] = (slot0 + b) / ci_tiles;
      slot_cit[b] = (slot0 + b) % ci_tiles;
    }
  }
  if (warp == kProdWarps) tc::tmem_alloc<512>(tc::smem_u32(tmem_ptr));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kProdWarps) {
    // ======================================================== producers (4 groups of 4 warps)
    // Tiles go round-robin to the groups (group = u % 4): each warp pays the per-tile fixed
    // costs (mbarrier wait, proxy fence, arrive) for every FOURTH tile, the loads of four
    // tiles are in flight at once, and each SM sub-partition always has other groups'
    // warps to issue from while one waits for its loads (slot hand-over: wait_seen).
    constexpr int GT = kProd / kGroups;            // threads per group (128)
    constexpr int RAg = GT / QA, PAg = KPIX / RAg; // dout rows per pass / passes (4, 8)
    constexpr int RBg = GT / QB, PBg = KPIX / RBg; // input rows per pass / passes
    constexpr int NREG = PAg > PBg ? PAg : PBg;
    const int grpi = warp >> 2;
    const int tg = threadIdx.x & (GT - 1);
    const int qa = tg % QA, ra0 = tg / QA;
    const int qb = tg % QB, rb0 = tg / QB;
    const int co = co0 + qa * 4;
    const bool co_ok = co < g.Cout;
    // dense: phase-grid pixel m IS the dout / input pixel (1x1 stride-1 layers and the
    // dout side of every stride-1 conv): no row table, no bounds checks
    const bool dense_out = g.os == 1 && g.Hp == g.Ho && g.Wp == g.Wo;
    const bool dense_in = g.is == 1 && g.Hp == g.Hi && g.Wp == g.Wi && g.T == 1 &&
                          g.dh[0] == 0 && g.dw[0] == 0;
    const bool need_tab = !(dense_out && dense_in);
    const float lb = g.in_relu ? 0.f : -INFINITY;
    int* rt = rowtab + grpi * 4 * KPIX;           // [dout pixel | n*Hi*Wi | i*is | j*is][KPIX]
    const int per_blk = NB + 1;
    const int total = nblk * per_blk;
    float4 buf[NREG];
    unsigned msk;
    int ie = grpi, iblk = 0, tab_blk = -1;        // tile cursor of this group
    while (ie >= per_blk) { ie -= per_blk; ++iblk; }
    auto split_store = [&](uint8_t* tile, int tile_plane_bytes, int r, int ch4, float4 x) {
      const float4 hi = make_float4(tc::to_tf32(x.x), tc::to_tf32(x.y), tc::to_tf32(x.z),
                                    tc::to_tf32(x.w));
      st_mn(tile, r, ch4, hi);
      if (NS == 3)
        st_mn(tile + tile_plane_bytes, r, ch4,
              make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
    };
    const int mine = (total - grpi + kGroups - 1) / kGroups;
    for (int k0 = 0; k0 < mine; ++k0) {
      // ---------------------------------------------------------------- loads of this tile
      if (tab_blk != iblk) {                      // the cursor entered a new pixel block
        if (need_tab) {
          asm volatile("bar.sync %0, 128;" ::"r"(2 + grpi) : "memory");
          if (tg < KPIX) {
            const int m = mbeg + iblk * KPIX + tg;
            if (m < mend) {
              const unsigned um = (unsigned)m;
              const unsigned j = um % (unsigned)g.Wp, qq = um / (unsigned)g.Wp;
              const unsigned i = qq % (unsigned)g.Hp, n = qq / (unsigned)g.Hp;
              rt[tg] = ((int)n * g.Ho + ((int)i * g.os + g.ph)) * g.Wo + ((int)j * g.os + g.pw);
              rt[KPIX + tg] = (int)n * g.Hi * g.Wi;
              rt[2 * KPIX + tg] = (int)i * g.is;
              rt[3 * KPIX + tg] = (int)j * g.is;
            } else {
              rt[tg] = -1;
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + grpi) : "memory");
        }
        tab_blk = iblk;
      }
      const int m0 = mbeg + iblk * KPIX;
      msk = 0;
      if (ie == 0) {
        // warm L2 eight pixel blocks ahead (rows of a block are contiguous when the phase
        // grid is dense): the register buffers alone keep too few bytes in flight for HBM
        const int pfm = m0 + 8 * KPIX;
        if (pfm < mend) {
          if (g.os == 1) {
            const int lines = (KPIX * WM * 4) / 128;
            for (int l = tg; l < lines; l += GT) {
              const int r = l >> 2, c = (l & 3) * 32;
              if (co0 + c < g.Cout)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dout + (int64_t)(pfm + r) * g.Cout + co0 + c));
            }
          }
          if (g.is == 1 && g.Hp == g.Hi && g.Wp == g.Wi) {
            const int lpr = g.Cin >> 5;
            for (int l = tg; l < KPIX * lpr; l += GT) {
              const int r = l / lpr, c = (l - r * lpr) * 32;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (int64_t)(pfm + r) * g.Cin + c));
            }
          }
        }
#pragma unroll
        for (int k = 0; k < PAg; ++k) {
          buf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int r = ra0 + k * RAg;
          const int dp = dense_out ? (m0 + r < mend ? m0 + r : -1) : rt[r];
          if (co_ok && dp >= 0)
            buf[k] = *reinterpret_cast<const float4*>(dout + (int64_t)dp * g.Cout + co);
        }
      } else {
        const int sl = ie - 1;
        const int ci = slot_cit[sl] * BNW + qb * 4;
        if (dense_in) {
#pragma unroll
          for (int k = 0; k < PBg; ++k) {
            buf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int mr = m0 + rb0 + k * RBg;
            if (ci < g.Cin && mr < mend) {
              buf[k] = *reinterpret_cast<const float4*>(in + (int64_t)mr * g.Cin + ci);
              msk |= 1u << k;
            }
          }
        } else {
          const int dh = g.dh[slot_t[sl]], dwv = g.dw[slot_t[sl]];
#pragma unroll
          for (int k = 0; k < PBg; ++k) {
            buf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int r = rb0 + k * RBg;
            const int ih = rt[2 * KPIX + r] + dh, iw = rt[3 * KPIX + r] + dwv;
            if (ci < g.Cin && rt[r] >= 0 && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi) {
              buf[k] = *reinterpret_cast<const float4*>(
                  in + ((int64_t)rt[KPIX + r] + (int64_t)ih * g.Wi + iw) * g.Cin + ci);
              msk |= 1u << k;
            }
          }
        }
      }
      // ---------------------------------------------------------------- convert + store
      if (ie == 0) {
        const int s = iblk % C::SA, use = iblk / C::SA;
        wait_seen(seenA + s, use);
        tc::mbar_wait(emptyA(s), (use & 1) ^ 1);
        if (lane == 0) seenA[s] = use + 1;
        uint8_t* tile = sm + s * C::A_TILE;
#pragma unroll
        for (int k = 0; k < PAg; ++k) split_store(tile, WM * KPIX * 4, ra0 + k * RAg, qa, buf[k]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(fullA(s));
      } else {
        const int qn = iblk * NB + (ie - 1);
        const int s = qn % C::SB;
        const int ci = slot_cit[ie - 1] * BNW + qb * 4;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_scale && ci < g.Cin) {
          sc = *reinterpret_cast<const float4*>(in_scale + ci);
          sh = *reinterpret_cast<const float4*>(in_shift + ci);
        }
        const int use = qn / C::SB;
        wait_seen(seenB + s, use);
        tc::mbar_wait(emptyB(s), (use & 1) ^ 1);
        if (lane == 0) seenB[s] = use + 1;
        uint8_t* tile = smB + s * C::B_TILE;
        if (in_scale) {
          // branch-free affine (+ReLU) on every row; rows that were not loaded (image border,
          // end of the pixel range) are post-activation zeros and are patched afterwards
#pragma unroll
          for (int k = 0; k < PBg; ++k) buf[k] = tc::bn_act4(buf[k], sc, sh, lb);
          if (msk != (1u << PBg) - 1u) {
#pragma unroll
            for (int k = 0; k < PBg; ++k)
              if (!((msk >> k) & 1u)) buf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < PBg; ++k) split_store(tile, BNW * KPIX * 4, rb0 + k * RBg, qb, buf[k]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(fullB(s));
      }
      ie += kGroups;
      while (ie >= per_blk) { ie -= per_blk; ++iblk; }
    }
    // ======================================================== epilogue (warps 0-3)
    if (warp < 4) {
      tc::mbar_wait(done_bar, 0);
      tc::tc_fence_after();
      const int corow = co0 + warp * 32 + lane;
      const int64_t wrow = (int64_t)g.Tw * g.Cin;
      for (int b = 0; b < NB; ++b) {
        const int t = slot_t[b], cit = slot_cit[b];
#pragma unroll 1
        for (int chunk = 0; chunk < BNW / 32; ++chunk) {
          uint32_t r[32];
          tc::tmem_ld32(tmem_base + b * BNW + chunk * 32 + ((uint32_t)(warp * 32) << 16), r);
          tc::tmem_ld_wait();
          const int c0 = cit * BNW + chunk * 32;
          if (nblk > 0 && corow < g.Cout && c0 < g.Cin) {
            float* dst = dw + (int64_t)corow * wrow + (int64_t)g.wt[t] * g.Cin + c0;
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              if (c0 + c < g.Cin)
                red_add_v4(dst + c, __uint_as_float(r[c]), __uint_as_float(r[c + 1]),
                           __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
          }
        }
      }
    }
  } else {
    // ======================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(WM, BNW, 1, 1);     // both operands MN-major
      constexpr uint32_t LBO = KPIX * 128, SBO = 512;   // chunk stride, 4-row k-atom stride
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int blk = 0; blk < nblk; ++blk) {
        tc::mbar_wait(fullA(sa), pha);
        const uint32_t a_hi = base + sa * C::A_TILE;
        for (int b = 0; b < NB; ++b) {
          tc::mbar_wait(fullB(sb), phb);
          tc::tc_fence_after();
          const uint32_t b_hi = base + C::SA * C::A_TILE + sb * C::B_TILE;
          const uint32_t d_tmem = tmem_base + b * BNW;
#pragma unroll
          for (int ks = 0; ks < KPIX / 8; ++ks) {
            const uint64_t ah = tc::desc_mnmajor_sw128(a_hi + ks * 1024, LBO, SBO);
            const uint64_t bh = tc::desc_mnmajor_sw128(b_hi + ks * 1024, LBO, SBO);
            if (NS == 3) {
              const uint64_t al = tc::desc_mnmajor_sw128(a_hi + WM * KPIX * 4 + ks * 1024, LBO, SBO);
              const uint64_t bl = tc::desc_mnmajor_sw128(b_hi + BNW * KPIX * 4 + ks * 1024, LBO, SBO);
              tc::mma_tf32(d_tmem, al, bh, idesc, (blk | ks) != 0);
              tc::mma_tf32(d_tmem, ah, bl, idesc, 1);
              tc::mma_tf32(d_tmem, ah, bh, idesc, 1);
            } else {
              tc::mma_tf32(d_tmem, ah, bh, idesc, (blk | ks) != 0);
            }
          }
          tc::mma_commit(emptyB(sb));
          if (++sb == C::SB) { sb = 0; phb ^= 1; }
        }
        tc::mma_commit(emptyA(sa));
        if (++sa == C::SA) { sa = 0; pha ^= 1; }
      }
      tc::mma_commit(done_bar);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == kProdWarps) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int BNW, int NS>
int launch_wgrad(const epb_conv_geom* g, const float* in, const float* dout, const float* in_scale,
                 const float* in_shift, float* dw, cudaStream_t st) {
  using C = WCfg<BNW, NS>;
  const int64_t M = (int64_t)g->N * g->Hp * g->Wp;
  EPB_CHECK_ARG(M < (1LL << 31));
  const int co_tiles = (g->Cout + WM - 1) / WM;
  const int ci_tiles = (g->Cin + BNW - 1) / BNW;
  const int total_slots = g->T * ci_tiles;
  const int groups = (total_slots + C::NB_MAX - 1) / C::NB_MAX;
  const int nb = (total_slots + groups - 1) / groups;          // balanced group size
  const int groups2 = (total_slots + nb - 1) / nb;
  const int64_t tiles = (int64_t)co_tiles * groups2;
  // Split the pixel range so that the grid fills whole waves of one CTA per SM (a
  // 2.16-wave grid idles most SMs in its last round).  Cost model per split count:
  // rounds x (pixel blocks per CTA + fixed prologue/epilogue cost worth ~6 blocks).
  const int64_t max_splits = (M + 8 * KPIX - 1) / (8 * KPIX);   // >= 8 pixel blocks per CTA
  int64_t splits = 1;
  int64_t best = -1;
  for (int64_t sp = 1; sp <= max_splits && sp * tiles <= 3 * kNumSMs + tiles; ++sp) {
    const int64_t rounds = (sp * tiles + kNumSMs - 1) / kNumSMs;
    const int64_t blocks = ((M + sp - 1) / sp + KPIX - 1) / KPIX;
    const int64_t cost = rounds * (blocks + 6);
    if (best < 0 || cost < best) { best = cost; splits = sp; }
  }
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
      *g, in, dout, in_scale, in_shift, dw, co_tiles, ci_tiles, groups2, nb, (int)rows);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

}  // namespace

bool epb_conv_wgrad_tc_supported(const epb_conv_geom* g) {
  // 32-bit element offsets into the input / pixel indices into dout
  return g->Cin % 32 == 0 && g->Cout % 4 == 0 && g->Cout >= 32 &&
         (int64_t)g->N * g->Hi * g->Wi * g->Cin < (1LL << 31) &&
         (int64_t)g->N * g->Ho * g->Wo < (1LL << 31);
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
