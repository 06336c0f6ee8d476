// tcgen05 tensor-core path of the tap-list implicit GEMM (include/epb.h).
//
//   out[m, co] = sum_k A[m, k] * W[co, k],  m = (n, i, j) phase-grid pixel,
//   k = (tap, ci);  A gathered from the NHWC input with the producing layer's
//   BatchNorm+ReLU applied on the fly.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-15 A producers: gather 128 pixel rows x 32 channels (128-byte rows),
//              fused BN+ReLU, split into TF32 hi (+ lo for the 3xTF32 mode) and
//              store into the SWIZZLE_128B K-major smem layout tcgen05 reads; the
//              k-blocks of all tiles of the CTA form ONE stream that the producer
//              groups take round-robin (warp-private row tables, no CTA barrier);
//   warp  16   B producer: TMA loads of the packed weight tile (hi / lo planes);
//   warp  17   MMA issuer: one thread issues tcgen05.mma kind::tf32 (M=128,
//              N=BN, K=8) with FP32 accumulators in TMEM (double buffered);
//   warps 18-21 epilogue: tcgen05.ld TMEM -> registers -> bias -> smem-staged
//              128-byte-line stores (row pointers from a per-tile smem table, rows
//              past M clamped: branch free) / accumulate, plus per-channel sum and
//              sum of squares for the following BatchNorm (warp-private smem
//              slices, one double atomic per column per tile).
// smem ring full/empty mbarriers, TMEM full/empty mbarriers; every wait is
// bounded (trap instead of hang).
//
// precision 1: TF32 single pass.  precision 3: 3xTF32 error-compensated
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, FP32 accumulate): fp32-grade results on
// the tensor pipe, which is what the 1e-3 end-to-end parity bar needs.
#include <cstdlib>
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;          // pixel rows per tile == UMMA M
constexpr int BKE = 32;          // tf32 elements per k-block (128 bytes)
constexpr int kProducerWarps = 16;      // G groups take k-blocks round-robin (Cfg::G)

constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (kProducerWarps + 2 + kEpiWarps);   // 704
constexpr int kSmemBudget = 200 * 1024;
constexpr int kRowTab = kProducerWarps * 32;   // one private table of <= 32 rows per producer warp

// PAIR: two CTAs of a cluster work on one 256-row M tile with tcgen05 cta_group::2; each
// holds its own 128 A rows and HALF of the B tile's N rows, which halves the weight bytes
// every SM pulls from L2 per MMA (the bound of the wide 3xTF32 tiles) and its smem stage.
template <int BN, int NS, bool PAIR = false>
struct Cfg {
  static constexpr int PL = (NS == 3) ? 2 : 1;                 // operand planes (hi, lo)
  static constexpr int BROWS = PAIR ? BN / 2 : BN;             // B rows held by this CTA
  static constexpr int A_BYTES = BM * 128 * PL;
  static constexpr int B_BYTES = BROWS * 128 * PL;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int S_ = kSmemBudget / STAGE;
  static constexpr int S = S_ > 8 ? 8 : S_;
  // Producer groups.  A group waits on a slot's empty barrier by PARITY, which is only
  // sound while it is less than two phases ahead of the barrier: G <= S.  With a deep
  // ring: 4 groups of 4 warps, one register buffer each; with 2-3 slots (3xTF32, wide
  // tiles): 2 groups of 8 warps with a register double buffer.  Either way the loads of
  // four k-blocks are in flight and every SM sub-partition has 4 producer warps.
  static constexpr int G = S >= 4 ? 4 : 2;
  static constexpr int D = 4 / G;                              // register buffers per group
  static constexpr int W = kProducerWarps / G;                 // warps per group
  static constexpr int NQ = 32 / W;                            // float4 per thread per k-block
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int EPI_PITCH = 36;                          // floats per staged row (144 B)
  static constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;       // one 32x32 block per epilogue warp
  static constexpr int STAT_BYTES = kEpiWarps * 2 * BN * 4;     // per-warp [sum | sum of squares][BN]
  static constexpr int PTAB_BYTES = kEpiWarps * 32 * 8;         // output row pointer of every TMEM lane
  static constexpr int SMEM = S * STAGE + 1024 /*align*/ + 1024 /*barriers*/ + kRowTab * 12 /*rowinfo*/ +
                              STAT_BYTES + EPI_BYTES + PTAB_BYTES;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct RowInfo {
  int off0[kRowTab];                  // ((n*Hi + i*is)*Wi + j*is)*Cin : element offset of the un-shifted pixel
  unsigned long long vmask[kRowTab];  // bit t set <=> tap t reads inside the image (0 for rows past M)
};

template <int BN, int NS, bool PAIR>
__device__ __forceinline__ void
fprop_body(const epb_conv_geom& g, const CUtensorMap* tmap_w, const float* __restrict__ in,
           const float* __restrict__ in_scale, const float* __restrict__ in_shift,
           const float* __restrict__ bias, float* __restrict__ out, double* __restrict__ stats,
           int npad, int m_tiles, int n_tiles, int tune) {
  using C = Cfg<BN, NS, PAIR>;
  // PAIR: `tile` enumerates (pair of M tiles, N tile); CTA `crank` of the cluster owns M tile
  // 2 * (tile / n_tiles) + crank (rows past M are zero rows / unwritten, as in the tail tile)
  const int crank = PAIR ? (int)tc::cluster_ctarank() : 0;
  const int tile0 = PAIR ? (int)tc::cluster_id_x() : (int)blockIdx.x;
  const int tstep = PAIR ? (int)tc::cluster_count_x() : (int)gridDim.x;
  auto m_tile_of = [&](int tile) { return PAIR ? 2 * (tile / n_tiles) + crank : tile / n_tiles; };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;               // SWIZZLE_128B needs 1024 B
  uint8_t* sm = smem_raw + (base - raw);
  uint8_t* ctrl = sm + C::S * C::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);          // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ctrl + 8 * (2 * 8 + 4));
  RowInfo* rows = reinterpret_cast<RowInfo*>(ctrl + 1024);
  float* sstat = reinterpret_cast<float*>(ctrl + 1024 + sizeof(RowInfo));   // [4 warps][2][BN]
  float* epi_stage = sstat + kEpiWarps * 2 * BN;                             // [4][32][EPI_PITCH]
  unsigned long long* eprow =                                                // [4][32] output row pointers
      reinterpret_cast<unsigned long long*>(epi_stage + kEpiWarps * 32 * C::EPI_PITCH);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (18 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t M = (int64_t)g.N * g.Hp * g.Wp;
  const int CB = g.Cin / BKE;              // channel blocks per tap
  const int KB = g.T * CB;                 // k-blocks per tile
  const int total_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::S; ++s) {
      tc::mbar_init(full_bar(s), (PAIR ? 2 : 1) * C::W + 1);   // PAIR: leader's barrier collects both CTAs
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar(a), 1);
      tc::mbar_init(tempty_bar(a), (PAIR ? 2 : 1) * kEpiWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == kProducerWarps + 1) {
    if (PAIR) tc::tmem_alloc_pair<C::TMEM_COLS>(tc::smem_u32(tmem_ptr));
    else tc::tmem_alloc<C::TMEM_COLS>(tc::smem_u32(tmem_ptr));
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync();             // the peer's barriers are initialised before any remote arrive
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kProducerWarps) {
    // =================================================== A producers (512 threads), flat stream
    // The k-blocks of ALL tiles of this CTA form one stream n = 0, 1, 2, ...; group g takes
    // n = g, g + G, ... (ring slot n % S).  A warp only ever touches its own RPW pixel rows,
    // so the row geometry lives in a WARP-PRIVATE table that the warp rebuilds when its issue
    // cursor enters a new tile: no CTA-wide barrier between tiles, and the loads of the next
    // tile's first k-blocks are in flight while the current tile is still being converted
    // (with two bar.sync of all 16 producer warps per tile, 1x1 layers with 2..8 k-blocks per
    // tile spent 20-30 % of their warp-stall samples at those barriers; measured gain 4-9 % on
    // the K <= 256 layers, profiles/r1_layer_sweep_flat_producer.md).
    //
    // L2 prefetch policy of the A stream (measured on the bench layers, tools/tune_sweep.sh):
    //  * dense 1x1 layers (phase-grid pixel == input pixel): ROLLING prefetch kPfDist k-blocks
    //    ahead of the loads, crossing into this CTA's next tiles -- bounded footprint;
    //  * otherwise the whole NEXT tile while the current one is processed, but only for rows
    //    of <= 1 KB: with wider rows 148 CTAs x 2 tiles overflow L2 and the prefetch costs
    //    more than it saves (Cin = 512: -9 %, Cin = 1024: -11 %).
    constexpr int kPfDist = 6;
    constexpr int G = C::G, D = C::D, NQ = C::NQ, RPW = BM / C::W;
    const bool roll = (tune & 1) && g.T == 1 && g.is == 1 && g.Hp == g.Hi && g.Wp == g.Wi;
    const bool pf_tile = ((tune & 1) && !roll && g.Cin <= 256) || (tune & 2);
    const int dmt = (n_tiles == 1 && tile0 + tstep < total_tiles)
                        ? m_tile_of(tile0 + tstep) - m_tile_of(tile0) : 0;   // M-tile stride of this CTA
    const bool once = (tune & 4) && g.T == 1 && n_tiles == 1;
    const float lb = g.in_relu ? 0.f : -INFINITY;
    uint64_t pol_first;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    const int c4 = lane & 7;                   // 16-byte chunk within the 128-byte row
    const int rsub = lane >> 3;                // 0..3
    const int grp = warp / C::W, wg = warp % C::W;
    int* w_off = rows->off0 + warp * 32;                   // this warp's rows wg*RPW .. +RPW-1
    unsigned long long* w_vm = rows->vmask + warp * 32;
    const int my_tiles = tile0 < total_tiles ? (total_tiles - tile0 + tstep - 1) / tstep : 0;
    const long long total_kb = (long long)my_tiles * KB;
    const int mine = (int)((total_kb - grp + G - 1) / G);  // k-blocks of this group (>= 0)
    int i_tile = tile0, i_kb = grp, it = 0, icb = 0, i_mt = 0;
    int p_kb = grp, pcb = 0;
    bool i_new = true;                         // the issue cursor stands on a tile whose rows are not tabled yet
    while (i_kb >= KB) { i_kb -= KB; i_tile += tstep; }
    it = i_kb / CB; icb = i_kb - it * CB;
    while (p_kb >= KB) p_kb -= KB;
    pcb = p_kb % CB;
    int st = grp;                              // G <= S: ring slot / phase of this group's next k-block
    uint32_t ph = 0;
    auto table_rows = [&]() {
      i_mt = m_tile_of(i_tile);
      __syncwarp();
      if (lane < RPW) {
        const int64_t m = (int64_t)i_mt * BM + wg * RPW + lane;
        unsigned long long vm = 0;
        int off = 0;
        if (m < M) {
          const unsigned um = (unsigned)m;
          const unsigned j = um % (unsigned)g.Wp, qq = um / (unsigned)g.Wp;
          const unsigned i = qq % (unsigned)g.Hp, n = qq / (unsigned)g.Hp;
          const int ih0 = (int)i * g.is, iw0 = (int)j * g.is;
          off = (((int)n * g.Hi + ih0) * g.Wi + iw0) * g.Cin;
          for (int t = 0; t < g.T; ++t) {
            const int ih = ih0 + g.dh[t], iw = iw0 + g.dw[t];
            if (ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi) vm |= 1ull << t;
          }
        }
        w_off[lane] = off;
        w_vm[lane] = vm;
        const int ntile = i_tile + tstep;
        if (pf_tile && grp == 0 && ntile < total_tiles && m_tile_of(ntile) != i_mt) {
          const int64_t mn = (int64_t)m_tile_of(ntile) * BM + wg * RPW + lane;
          if (mn < M) {
            const unsigned um = (unsigned)mn;
            const unsigned j = um % (unsigned)g.Wp, qq = um / (unsigned)g.Wp;
            const unsigned i = qq % (unsigned)g.Hp, n = qq / (unsigned)g.Hp;
            int ih = (int)i * g.is, iw = (int)j * g.is;
            ih = ih < g.Hi ? ih : g.Hi - 1;
            iw = iw < g.Wi ? iw : g.Wi - 1;
            const float* rowp = in + (((int64_t)n * g.Hi + ih) * g.Wi + iw) * g.Cin;
            for (int c = 0; c < g.Cin; c += 32)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + c));
          }
        }
      }
      __syncwarp();
      i_new = false;
    };
    float4 buf[D][NQ];
    unsigned okm[D];
    auto issue = [&](float4 (&dst)[NQ], unsigned& mask) {
      if (i_new) table_rows();
      const int delta = (g.dh[it] * g.Wi + g.dw[it]) * g.Cin + icb * BKE + c4 * 4;
      const int tap = it;
      if (roll) {
        int pb = icb + kPfDist, ta = 0;          // channel block / tiles ahead of this CTA
        while (pb >= CB) { pb -= CB; ++ta; }
        if ((ta == 0 || dmt > 0) && lane < RPW) {
          if ((int64_t)(i_mt + ta * dmt) * BM + wg * RPW + lane < M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(
                in + ((int64_t)w_off[lane] + (int64_t)ta * dmt * BM * g.Cin + pb * BKE)));
        }
      }
      mask = 0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int lr = q * 4 + rsub;
        dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((w_vm[lr] >> tap) & 1ull) {
          const float* ap = in + ((int64_t)w_off[lr] + delta);
          if (once) {
            asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(dst[q].x), "=f"(dst[q].y), "=f"(dst[q].z), "=f"(dst[q].w)
                         : "l"(ap), "l"(pol_first));
          } else {
            dst[q] = *reinterpret_cast<const float4*>(ap);
          }
          mask |= 1u << q;
        }
      }
      i_kb += G;
      icb += G;
      while (icb >= CB) { icb -= CB; ++it; }
      if (i_kb >= KB) {
        do { i_kb -= KB; i_tile += tstep; } while (i_kb >= KB);
        it = i_kb / CB; icb = i_kb - it * CB;
        i_new = true;
      }
    };
    auto process = [&](const float4 (&v)[NQ], unsigned mask) {
      const int ch = pcb * BKE + c4 * 4;
      p_kb += G;
      pcb += G;
      while (pcb >= CB) pcb -= CB;
      if (p_kb >= KB) {
        do { p_kb -= KB; } while (p_kb >= KB);
        pcb = p_kb % CB;
      }
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_scale) {
        sc = *reinterpret_cast<const float4*>(in_scale + ch);
        sh = *reinterpret_cast<const float4*>(in_shift + ch);
      }
      tc::mbar_wait(empty_bar(st), ph ^ 1);
      uint8_t* a_hi = sm + st * C::STAGE;
      float4 xv[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) xv[q] = v[q];
      if (in_scale) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) xv[q] = tc::bn_act4(xv[q], sc, sh, lb);
        if (mask != (1u << NQ) - 1u) {
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            if (!((mask >> q) & 1u)) xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int r = wg * RPW + q * 4 + rsub;
        const float4 x = xv[q];
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c4 ^ (r & 7)) << 4);
        float4 hi = make_float4(tc::to_tf32(x.x), tc::to_tf32(x.y), tc::to_tf32(x.z),
                                tc::to_tf32(x.w));
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        if (NS == 3) {
          float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
          *reinterpret_cast<float4*>(a_hi + BM * 128 + off) = lo;
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) {
          if (tune & 8) tc::mbar_arrive_cluster_cta(tc::mapa(full_bar(st), 0));
          else tc::mbar_arrive_cluster(tc::mapa(full_bar(st), 0));
        } else {
          tc::mbar_arrive(full_bar(st));
        }
      }
      st += G;
      ph ^= (uint32_t)((st / C::S) & 1);
      st %= C::S;
    };
#ifdef EPB_DBG_SKIP_A     // bottleneck probe (tools/build_variant.py): slots are handed over unfilled
    for (int k = 0; k < mine; ++k) {
      tc::mbar_wait(empty_bar(st), ph ^ 1);
      __syncwarp();
      if (lane == 0) {
        if (PAIR) tc::mbar_arrive_cluster(tc::mapa(full_bar(st), 0));
        else tc::mbar_arrive(full_bar(st));
      }
      st += G;
      ph ^= (uint32_t)((st / C::S) & 1);
      st %= C::S;
    }
#else
    if constexpr (D == 1) {
      for (int k = 0; k < mine; ++k) {
        issue(buf[0], okm[0]);
        process(buf[0], okm[0]);
      }
    } else {
      if (mine > 0) issue(buf[0], okm[0]);
      for (int k0 = 0; k0 < mine; k0 += 2) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int k = k0 + d;
          if (k < mine) {
            if (k + 1 < mine) issue(buf[d ^ 1], okm[d ^ 1]);
            process(buf[d], okm[d]);
          }
        }
      }
    }
#endif
  } else if (warp == kProducerWarps) {
    // =================================================== B producer (TMA)
    if (lane == 0) {
      tc::tma_prefetch_desc(tmap_w);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tstep) {
        const int nt = tile % n_tiles;
        int t = 0, cb = 0;
        for (int kb = 0; kb < KB; ++kb) {
          const int kx = g.wt[t] * g.Cin + cb * BKE;
          if (++cb == CB) { cb = 0; ++t; }
          tc::mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t b_dst = base + stage * C::STAGE + C::A_BYTES;
#ifdef EPB_DBG_SKIP_B
          if (!PAIR || crank == 0) tc::mbar_arrive(full_bar(stage));
          if (++stage == C::S) { stage = 0; phase ^= 1; }
          continue;
#endif
          if (PAIR) {
            // both halves complete on the LEADER's barrier; the leader posts the byte count
            if (crank == 0) tc::mbar_arrive_expect_tx(full_bar(stage), 2 * C::B_BYTES);
            const uint32_t lead_bar = tc::mapa(full_bar(stage), 0);
#pragma unroll
            for (int pl = 0; pl < C::PL; ++pl)
              tc::tma_load_2d_pair(b_dst + pl * C::BROWS * 128, tmap_w, lead_bar, kx,
                                   pl * npad + nt * BN + crank * C::BROWS);
          } else {
            tc::mbar_arrive_expect_tx(full_bar(stage), C::B_BYTES);
#pragma unroll
            for (int pl = 0; pl < C::PL; ++pl)
              tc::tma_load_2d(b_dst + pl * BN * 128, tmap_w, full_bar(stage), kx,
                              pl * npad + nt * BN);
          }
          if (++stage == C::S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kProducerWarps + 1) {
    // =================================================== MMA issuer
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(PAIR ? 2 * BM : BM, BN, 0, 0);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
#ifdef EPB_DBG_SKIP_MMA
        return;
#endif
        if (PAIR) tc::mma_tf32_pair(d, a, b, idesc, acc);
        else tc::mma_tf32(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if (PAIR) tc::mma_commit_pair(bar);
        else tc::mma_commit(bar);
      };
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tstep) {
        if (PAIR) tc::mbar_wait_cluster(tempty_bar(as), aphase ^ 1);
        else tc::mbar_wait(tempty_bar(as), aphase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < KB; ++kb) {
          if (PAIR) tc::mbar_wait_cluster(full_bar(stage), phase);
          else tc::mbar_wait(full_bar(stage), phase);
          tc::tc_fence_after();
          const uint32_t a_hi = base + stage * C::STAGE;
          const uint32_t b_hi = a_hi + C::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BKE / 8; ++kk) {
            const uint64_t ah = tc::desc_kmajor_sw128(a_hi + kk * 32);
            const uint64_t bh = tc::desc_kmajor_sw128(b_hi + kk * 32);
            if (NS == 3) {
              const uint64_t al = tc::desc_kmajor_sw128(a_hi + BM * 128 + kk * 32);
              const uint64_t bl = tc::desc_kmajor_sw128(b_hi + C::BROWS * 128 + kk * 32);
              mma(d_tmem, al, bh, (kb | kk) != 0);
              mma(d_tmem, ah, bl, 1);
              mma(d_tmem, ah, bh, 1);
            } else {
              mma(d_tmem, ah, bh, (kb | kk) != 0);
            }
          }
          commit(empty_bar(stage));                  // frees the smem slot when the MMAs retire
          if (++stage == C::S) { stage = 0; phase ^= 1; }
        }
        commit(tfull_bar(as));                        // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // =================================================== epilogue (4 warps)
    // One warp per TMEM lane quarter drains its 32 rows x BN columns in chunks of 32 columns.
    // The warp is alone on its scheduler slot, so the chunk is written to expose as few
    // dependent latencies as possible (short-K layers -- K <= 256 -- are bound by this loop):
    // row pointers come from a per-tile smem table (no shuffle + branch per row group), rows
    // past M are clamped onto the last valid row (a duplicate store of the same value) so the
    // store loop is branch free, and the per-column statistics go to a warp-private smem slice
    // with plain stores (no shared-memory CAS loops); the slices are summed at the tile end.
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int wq = warp - (kProducerWarps + 2);
    const int et = wq * 32 + lane;             // 0..127
    float* stg = epi_stage + wq * 32 * C::EPI_PITCH;
    float* sst = sstat + wq * 2 * BN;
    unsigned long long* ptab = eprow + wq * 32;
    const int c4 = lane & 7, rsub = lane >> 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tstep) {
      const int mt = m_tile_of(tile), nt = tile % n_tiles;
      const int64_t m = (int64_t)mt * BM + q * 32 + lane;
      const bool valid = m < M;
      float* orow = nullptr;
      if (valid) {
        const int j = (int)(m % g.Wp);
        const int i = (int)((m / g.Wp) % g.Hp);
        const int n = (int)(m / ((int64_t)g.Wp * g.Hp));
        orow = out + (((int64_t)n * g.Ho + (i * g.os + g.ph)) * g.Wo + (j * g.os + g.pw)) * g.Cout;
      }
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      const int nvalid = __popc(vmask);        // rows of a tile are valid in a prefix (m < M)
      ptab[lane] = reinterpret_cast<unsigned long long>(orow);
      __syncwarp();
      tc::mbar_wait(tfull_bar(as), aphase);
      tc::tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        const int col0 = nt * BN + chunk * 32;
        if (col0 >= g.Cout) break;             // N tail (Cout % 32 == 0)
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + as * BN + chunk * 32 + ((uint32_t)(q * 32) << 16), r);
        tc::tmem_ld_wait();
        // stage this warp's 32x32 block (row = TMEM lane) so that the global stores below
        // write whole 128-byte lines (4 rows per warp instruction) instead of 32 scattered
        // 16-byte pieces
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 x = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]),
                                 __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
          if (bias) {
            const float4 b = *reinterpret_cast<const float4*>(bias + col0 + c);
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          *reinterpret_cast<float4*>(stg + lane * C::EPI_PITCH + c) = x;
        }
        __syncwarp();
        if (stats) {
          // lane = column: sum over the staged valid rows (bank = 4*row + lane: conflict free);
          // four independent partial sums keep the add chains short
          float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const float x = ((vmask >> rr) & 1u) ? stg[rr * C::EPI_PITCH + lane] : 0.f;
            s1[rr & 3] += x;
            s2[rr & 3] = fmaf(x, x, s2[rr & 3]);
          }
          sst[chunk * 32 + lane] = (s1[0] + s1[1]) + (s1[2] + s1[3]);
          sst[BN + chunk * 32 + lane] = (s2[0] + s2[1]) + (s2[2] + s2[3]);
        }
        if (g.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            if (rr < nvalid) {
              float4 x = *reinterpret_cast<const float4*>(stg + rr * C::EPI_PITCH + c4 * 4);
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ptab[rr]) + col0) + c4;
              const float4 pv = *o;
              x.x += pv.x; x.y += pv.y; x.z += pv.z; x.w += pv.w;
              *o = x;
            }
          }
        } else if (nvalid > 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = min(i * 4 + rsub, nvalid - 1);
            const float4 x = *reinterpret_cast<const float4*>(stg + rr * C::EPI_PITCH + c4 * 4);
            *(reinterpret_cast<float4*>(reinterpret_cast<float*>(ptab[rr]) + col0) + c4) = x;
          }
        }
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) {
          if (tune & 8) tc::mbar_arrive_cluster_cta(tc::mapa(tempty_bar(as), 0));
          else tc::mbar_arrive_cluster(tc::mapa(tempty_bar(as), 0));
        } else {
          tc::mbar_arrive(tempty_bar(as));
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
      if (stats) {
        // every column of this N tile that lies below Cout was written by all four warps
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int c = et; c < 2 * BN; c += 128) {
          const int which = c / BN, col = nt * BN + (c % BN);
          if (col < g.Cout) {
            const float v = (sstat[c] + sstat[2 * BN + c]) + (sstat[4 * BN + c] + sstat[6 * BN + c]);
            atomicAdd(stats + (int64_t)which * g.Cout + col, (double)v);
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync();             // the peer may still read this CTA's smem / signal its barriers
  if (warp == kProducerWarps + 1) {
    tc::tc_fence_after();
    if (PAIR) tc::tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
    else tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int NS>
__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_tc_kernel(const __grid_constant__ epb_conv_geom g,
                     const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ in,
                     const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                     const float* __restrict__ bias, float* __restrict__ out,
                     double* __restrict__ stats, int npad, int m_tiles, int n_tiles, int tune) {
  fprop_body<BN, NS, false>(g, &tmap_w, in, in_scale, in_shift, bias, out, stats, npad, m_tiles,
                            n_tiles, tune);
}

// CTA-pair variant (cluster of 2, tcgen05 cta_group::2, M = 256 per pair)
template <int BN, int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_fprop_tc_pair_kernel(const __grid_constant__ epb_conv_geom g,
                          const __grid_constant__ CUtensorMap tmap_w,
                          const float* __restrict__ in, const float* __restrict__ in_scale,
                          const float* __restrict__ in_shift, const float* __restrict__ bias,
                          float* __restrict__ out, double* __restrict__ stats, int npad,
                          int m_tiles, int n_tiles, int tune) {
  fprop_body<BN, NS, true>(g, &tmap_w, in, in_scale, in_shift, bias, out, stats, npad, m_tiles,
                           n_tiles, tune);
}

// ---------------------------------------------------------------- weight prep
// packed [Cout][K] fp32 -> planes [PL][npad][K]: TF32 hi (RN) and fp32 residual lo;
// rows >= Cout are zero so N-tail tiles need no predication.
__global__ void prep_weight_tc(const float* __restrict__ w, float* __restrict__ dst, int Cout,
                               int64_t K, int npad, int planes) {
  const int64_t total = (int64_t)npad * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / K;
    const float x = row < Cout ? w[i] : 0.f;
    const float hi = tc::to_tf32(x);
    dst[i] = hi;
    if (planes == 2) dst[total + i] = x - hi;
  }
}

// EPB_TUNE: memory-system switches of the A stream (bit 0 prefetch policy, bit 1 force next-tile
// prefetch, bit 2 evict-first single-use loads, bit 3 CTA-scope arrive across the CTA pair)
int tune_flags() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EPB_TUNE");
    v = e ? atoi(e) : 13;      // prefetch policy on, evict-first single-use loads, CTA-scope pair arrive
  }
  return v;
}

template <int BN, int NS, bool PAIR>
int launch_fprop(const epb_conv_geom* g, const CUtensorMap& tmap, const float* in,
                 const float* in_scale, const float* in_shift, const float* bias, float* out,
                 double* stats, int npad, cudaStream_t st) {
  using C = Cfg<BN, NS, PAIR>;
  const int64_t M = (int64_t)g->N * g->Hp * g->Wp;
  const int m_tiles = (int)((M + BM - 1) / BM);
  const int n_tiles = npad / BN;
  auto kern = [] {
    if constexpr (PAIR) return conv_fprop_tc_pair_kernel<BN, NS>;
    else return conv_fprop_tc_kernel<BN, NS>;
  }();
  static bool attr_set = false;
  if (!attr_set) {
    EPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  int grid;
  if (PAIR) {
    const int64_t pairs = (int64_t)((m_tiles + 1) / 2) * n_tiles;       // one cluster per tile pair
    grid = 2 * (int)(pairs < kNumSMs / 2 ? pairs : kNumSMs / 2);
  } else {
    const int64_t tiles = (int64_t)m_tiles * n_tiles;
    grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  }
  kern<<<grid, kThreads, C::SMEM, st>>>(*g, tmap, in, in_scale, in_shift, bias, out, stats, npad,
                                        m_tiles, n_tiles, tune_flags());
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

// EPB_CTA_PAIR=0 keeps every tile on single-CTA MMAs
bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EPB_CTA_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

epb_encode_tiled_fn epb_get_encode_tiled() {
  static epb_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) ==
            cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<epb_encode_tiled_fn>(p);
  }
  return fn;
}

bool epb_conv_wgrad_tc_supported(const epb_conv_geom* g);

bool epb_conv_tc_supported(const epb_conv_geom* g, bool wgrad) {
  if (wgrad) return epb_conv_wgrad_tc_supported(g);
  // the producers address the input with 32-bit element offsets
  return g->Cin % 32 == 0 && g->Cout % 32 == 0 && g->Cout >= 32 &&
         (int64_t)g->N * g->Hi * g->Wi * g->Cin < (1LL << 31);
}

int epb_conv_fprop_tc(const epb_conv_geom* g, const float* in, const float* w,
                      const float* in_scale, const float* in_shift, const float* bias, float* out,
                      double* stats, cudaStream_t st) {
  const int ns = g->precision == 3 ? 3 : 1;
  const int planes = ns == 3 ? 2 : 1;
  const int bn = g->Cout >= 256 ? 256 : (g->Cout >= 128 ? 128 : 64);
  const int npad = (g->Cout + bn - 1) / bn * bn;
  const int64_t K = (int64_t)g->Tw * g->Cin;
  // hi / lo operand planes of this call: scratch of this (device, stream); one slot per
  // stream, so concurrent calls on different streams do not share it
  float* wws = nullptr;
  int rc = epb_workspace(EPB_WS_WPLANES, (size_t)planes * npad * K * sizeof(float), st, (void**)&wws);
  if (rc) return rc;
  {
    int64_t blocks = ((int64_t)npad * K + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    prep_weight_tc<<<(int)blocks, 256, 0, st>>>(w, wws, g->Cout, K, npad, planes);
    EPB_LAUNCH_CHECK();
  }
  epb_encode_tiled_fn enc = epb_get_encode_tiled();
  if (!enc) {
    epb_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return EPB_ECUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)((int64_t)planes * npad)};
  const cuuint64_t strides[1] = {(cuuint64_t)(K * sizeof(float))};
  // wide tiles over at least one full pair of M tiles run on CTA pairs (each CTA loads half
  // of the tile's weight rows)
  const bool pair = bn == 256 && pair_enabled() &&
                    (int64_t)g->N * g->Hp * g->Wp > BM;
  const cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)(pair ? bn / 2 : bn)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, wws, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    epb_set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return EPB_ECUDA;
  }
#define EPB_TC_CASE(BN_, NS_, PAIR_)                   \
  if (bn == BN_ && ns == NS_ && pair == PAIR_)         \
    return launch_fprop<BN_, NS_, PAIR_>(g, tmap, in, in_scale, in_shift, bias, out, stats, npad, st);
  EPB_TC_CASE(64, 1, false) EPB_TC_CASE(128, 1, false) EPB_TC_CASE(256, 1, false)
  EPB_TC_CASE(64, 3, false) EPB_TC_CASE(128, 3, false) EPB_TC_CASE(256, 3, false)
  EPB_TC_CASE(256, 1, true) EPB_TC_CASE(256, 3, true)
#undef EPB_TC_CASE
  epb_set_error("no tcgen05 tile configuration for Cout=%d", g->Cout);
  return EPB_EINVAL;
}

