// fp32 CUDA-core implementation of the tap-list implicit GEMM (include/epb.h,
// epb_conv_geom).  This is the exact-fp32 path (precision == 0): it backs the
// layers whose shapes the tcgen05 path does not take (Cin % 32 != 0: the
// 7x7 stem) and serves as the on-device cross-check for the tensor-core path.
// Reference call sites: the cuDNN fprop/dgrad/wgrad behind nn.Conv2d /
// nn.ConvTranspose2d in lib/models/pose3d_resnet.py.
#include "common.cuh"
#include "conv_common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, THREADS = 256;

__global__ void __launch_bounds__(THREADS)
conv_fprop_simt(const __grid_constant__ epb_conv_geom g, const float* __restrict__ in,
                const float* __restrict__ w, const float* __restrict__ in_scale,
                const float* __restrict__ in_shift, const float* __restrict__ bias,
                float* __restrict__ out, double* __restrict__ stats) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.Hp * g.Wp;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = g.T * g.Cin;

  // A loader: thread owns rows lr and lr+64, k-quad lq
  const int lq = tid & 3, lr = tid >> 2;
  int rn[2], ri[2], rj[2];
  bool rv[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t m = m0 + lr + h * 64;
    rv[h] = m < M;
    const int64_t mm = rv[h] ? m : 0;
    rj[h] = (int)(mm % g.Wp);
    ri[h] = (int)((mm / g.Wp) % g.Hp);
    rn[h] = (int)(mm / ((int64_t)g.Wp * g.Hp));
  }
  // B loader: thread owns cout bo, k-quad bq
  const int bq = tid & 3, bo = tid >> 2;
  const int64_t wrow = (int64_t)g.Tw * g.Cin;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- gather A (with the fused BN+ReLU of the producing layer)
    {
      const int k = k0 + lq * 4;
      float4 v[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      if (k < K) {
        const int t = k / g.Cin, c = k - t * g.Cin;
        const int dh = g.dh[t], dw = g.dw[t];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ih = ri[h] * g.is + dh, iw = rj[h] * g.is + dw;
          if (rv[h] && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi) {
            float4 x = *reinterpret_cast<const float4*>(
                in + (((int64_t)rn[h] * g.Hi + ih) * g.Wi + iw) * g.Cin + c);
            if (in_scale) {
              const float4 s = *reinterpret_cast<const float4*>(in_scale + c);
              const float4 b = *reinterpret_cast<const float4*>(in_shift + c);
              x.x = fmaf(x.x, s.x, b.x); x.y = fmaf(x.y, s.y, b.y);
              x.z = fmaf(x.z, s.z, b.z); x.w = fmaf(x.w, s.w, b.w);
              if (g.in_relu) {
                x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f);
                x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
              }
            }
            v[h] = x;
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        As[lq * 4 + 0][lr + h * 64] = v[h].x;
        As[lq * 4 + 1][lr + h * 64] = v[h].y;
        As[lq * 4 + 2][lr + h * 64] = v[h].z;
        As[lq * 4 + 3][lr + h * 64] = v[h].w;
      }
    }
    // ---- load B
    {
      const int k = k0 + bq * 4;
      float4 x = {0, 0, 0, 0};
      const int co = n0 + bo;
      if (k < K && co < g.Cout) {
        const int t = k / g.Cin, c = k - t * g.Cin;
        x = *reinterpret_cast<const float4*>(w + (int64_t)co * wrow + (int64_t)g.wt[t] * g.Cin + c);
      }
      Bs[bq * 4 + 0][bo] = x.x; Bs[bq * 4 + 1][bo] = x.y;
      Bs[bq * 4 + 2][bo] = x.z; Bs[bq * 4 + 3][bo] = x.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias, accumulate, store, per-channel statistics
  const int co = n0 + tx * 4;
  float4 bv = {0, 0, 0, 0};
  if (bias && co < g.Cout) bv = *reinterpret_cast<const float4*>(bias + co);
  float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t m = m0 + ty * 8 + r;
    if (m >= M || co >= g.Cout) continue;
    const int j = (int)(m % g.Wp);
    const int i = (int)((m / g.Wp) % g.Hp);
    const int n = (int)(m / ((int64_t)g.Wp * g.Hp));
    float* o = out + (((int64_t)n * g.Ho + (i * g.os + g.ph)) * g.Wo + (j * g.os + g.pw)) * g.Cout + co;
    float4 v = make_float4(acc[r][0] + bv.x, acc[r][1] + bv.y, acc[r][2] + bv.z, acc[r][3] + bv.w);
    if (g.accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(o);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    *reinterpret_cast<float4*>(o) = v;
    cs[0] += v.x; cs[1] += v.y; cs[2] += v.z; cs[3] += v.w;
    cq[0] += v.x * v.x; cq[1] += v.y * v.y; cq[2] += v.z * v.z; cq[3] += v.w * v.w;
  }
  if (stats) {
    // reuse As as [2][16][64] scratch
    float* red = &As[0][0];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[(0 * 16 + ty) * 64 + tx * 4 + c] = cs[c];
      red[(1 * 16 + ty) * 64 + tx * 4 + c] = cq[c];
    }
    __syncthreads();
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      if (n0 + c < g.Cout) {
        double s = 0;
        for (int y = 0; y < 16; ++y) s += red[(which * 16 + y) * 64 + c];
        atomicAdd(stats + (int64_t)which * g.Cout + n0 + c, s);
      }
    }
  }
}

// dw[co][wt[t]][ci] += sum_m dout[m][co] * f(in[m shifted by tap t][ci])
constexpr int WM = 64, WN = 64, WK = 16;

__global__ void __launch_bounds__(THREADS)
conv_wgrad_simt(const __grid_constant__ epb_conv_geom g, const float* __restrict__ in,
                const float* __restrict__ dout, const float* __restrict__ in_scale,
                const float* __restrict__ in_shift, float* __restrict__ dw, int rows_per_split) {
  __shared__ __align__(16) float As[WK][WM + 4];  // dout  [m][co]
  __shared__ __align__(16) float Bs[WK][WN + 4];  // input [m][ci]
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.Hp * g.Wp;
  const int ci0 = blockIdx.x * WN;
  const int co_tiles = (g.Cout + WM - 1) / WM;
  const int t = blockIdx.y / co_tiles;
  const int co0 = (blockIdx.y % co_tiles) * WM;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t mend = min(M, mbeg + rows_per_split);
  const int dh = g.dh[t], dwv = g.dw[t];

  const int lrow = tid >> 4, lq = tid & 15;  // 16 rows x 16 float4 (64 floats)
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  float4 sc = {1, 1, 1, 1}, sh = {0, 0, 0, 0};
  const int cil = ci0 + lq * 4;
  if (in_scale && cil < g.Cin) {
    sc = *reinterpret_cast<const float4*>(in_scale + cil);
    sh = *reinterpret_cast<const float4*>(in_shift + cil);
  }

  for (int64_t mb = mbeg; mb < mend; mb += WK) {
    const int64_t m = mb + lrow;
    float4 a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
    if (m < mend) {
      const int j = (int)(m % g.Wp);
      const int i = (int)((m / g.Wp) % g.Hp);
      const int n = (int)(m / ((int64_t)g.Wp * g.Hp));
      const int col = co0 + lq * 4;
      if (col < g.Cout)
        a = *reinterpret_cast<const float4*>(
            dout + (((int64_t)n * g.Ho + (i * g.os + g.ph)) * g.Wo + (j * g.os + g.pw)) * g.Cout + col);
      const int ih = i * g.is + dh, iw = j * g.is + dwv;
      if (cil < g.Cin && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi) {
        b = *reinterpret_cast<const float4*>(in + (((int64_t)n * g.Hi + ih) * g.Wi + iw) * g.Cin + cil);
        if (in_scale) {
          b.x = fmaf(b.x, sc.x, sh.x); b.y = fmaf(b.y, sc.y, sh.y);
          b.z = fmaf(b.z, sc.z, sh.z); b.w = fmaf(b.w, sc.w, sh.w);
          if (g.in_relu) {
            b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
          }
        }
      }
    }
    *reinterpret_cast<float4*>(&As[lrow][lq * 4]) = a;
    *reinterpret_cast<float4*>(&Bs[lrow][lq * 4]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(aa[r], bb[c], acc[r][c]);
    }
    __syncthreads();
  }
  const int64_t wrow = (int64_t)g.Tw * g.Cin;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int co = co0 + ty * 4 + r;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ci = ci0 + tx * 4 + c;
      if (ci < g.Cin) atomicAdd(dw + (int64_t)co * wrow + (int64_t)g.wt[t] * g.Cin + ci, acc[r][c]);
    }
  }
}

}  // namespace

int epb_conv_geom_check(const epb_conv_geom* g) {
  EPB_CHECK_ARG(g != nullptr);
  EPB_CHECK_ARG(g->N > 0 && g->Hi > 0 && g->Wi > 0 && g->Cin > 0);
  EPB_CHECK_ARG(g->Ho > 0 && g->Wo > 0 && g->Cout > 0 && g->Hp > 0 && g->Wp > 0);
  EPB_CHECK_ARG(g->os >= 1 && g->is >= 1 && g->ph >= 0 && g->ph < g->os && g->pw >= 0 && g->pw < g->os);
  EPB_CHECK_ARG((g->Hp - 1) * g->os + g->ph < g->Ho && (g->Wp - 1) * g->os + g->pw < g->Wo);
  EPB_CHECK_ARG(g->T >= 1 && g->T <= EPB_MAX_TAPS && g->Tw >= 1);
  for (int t = 0; t < g->T; ++t) EPB_CHECK_ARG(g->wt[t] >= 0 && g->wt[t] < g->Tw);
  EPB_CHECK_ARG(g->Cin % 4 == 0 && g->Cout % 4 == 0);
  return EPB_OK;
}

int epb_conv_fprop_simt(const epb_conv_geom* g, const float* in, const float* w,
                        const float* in_scale, const float* in_shift, const float* bias,
                        float* out, double* stats, cudaStream_t st) {
  const int64_t M = (int64_t)g->N * g->Hp * g->Wp;
  dim3 grid((unsigned)((M + BM - 1) / BM), (g->Cout + BN - 1) / BN);
  conv_fprop_simt<<<grid, THREADS, 0, st>>>(*g, in, w, in_scale, in_shift, bias, out, stats);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

int epb_conv_wgrad_simt(const epb_conv_geom* g, const float* in, const float* dout,
                        const float* in_scale, const float* in_shift, float* dw, cudaStream_t st) {
  const int64_t M = (int64_t)g->N * g->Hp * g->Wp;
  const int tiles = ((g->Cin + WN - 1) / WN) * ((g->Cout + WM - 1) / WM) * g->T;
  int64_t splits = (4 * kNumSMs + tiles - 1) / tiles;
  const int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows = (M + splits - 1) / splits;
  rows = (rows + WK - 1) / WK * WK;
  splits = (M + rows - 1) / rows;
  dim3 grid((g->Cin + WN - 1) / WN, ((g->Cout + WM - 1) / WM) * g->T, (unsigned)splits);
  conv_wgrad_simt<<<grid, THREADS, 0, st>>>(*g, in, dout, in_scale, in_shift, dw, (int)rows);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}
