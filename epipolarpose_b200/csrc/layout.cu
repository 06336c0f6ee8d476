// Module-boundary layout conversion (NCHW <-> NHWC), weight packing between
// the reference's state_dict layouts (Conv2d [O][I][kh][kw], ConvTranspose2d
// [I][O][kh][kw]; lib/models/pose3d_resnet.py:99,116-122,171-178) and the
// packed GEMM operand [X][T][Ypad], and the fused optimiser steps
// (torch.optim.Adam / SGD call sites lib/utils/utils.py:45-61).
#include "common.cuh"

namespace {

// [N][C][HW] -> [N][HW][Cpad] (zero padded) through a 32x32 smem tile
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                    int HW, int Cpad) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((int64_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < Cpad) dst[((int64_t)n * HW + p) * Cpad + c] = tile[threadIdx.x][i];
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                    int HW, int Cpad) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < Cpad) ? src[((int64_t)n * HW + p) * Cpad + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((int64_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// src [A][B][T] (T = kh*kw), packed [X][T][Ypad]: swap=0 -> X=A,Y=B ; swap=1 -> X=B,Y=A
__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A,
                                   int B, int T, int swap, int Ypad, int unpack) {
  const int X = swap ? B : A, Y = swap ? A : B;
  const int64_t total = (int64_t)X * T * Ypad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i % Ypad);
    const int t = (int)((i / Ypad) % T);
    const int x = (int)(i / ((int64_t)Ypad * T));
    const int a = swap ? y : x, b = swap ? x : y;
    if (!unpack) {
      dst[i] = (y < Y) ? src[((int64_t)a * B + b) * T + t] : 0.f;
    } else if (y < Y) {
      // inverse: `src` is the packed tensor, `dst` the state_dict layout
      dst[((int64_t)a * B + b) * T + t] = src[i];
    }
  }
}

// Many pack / unpack jobs in one launch: a block finds its job by binary search over the
// jobs' first blocks and converts 1024 packed elements of it.
__global__ void __launch_bounds__(256)
pack_weight_batch_kernel(const epb_pack_job* __restrict__ jobs, int njobs) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const epb_pack_job j = jobs[lo];
  const int X = j.swap ? j.B : j.A, Y = j.swap ? j.A : j.B;
  const int64_t total = (int64_t)X * j.T * j.Ypad;
  const int64_t i0 = ((int64_t)blockIdx.x - j.first_block) * 1024;
  const int64_t i1 = i0 + 1024 < total ? i0 + 1024 : total;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    const int y = (int)(i % j.Ypad);
    const int t = (int)((i / j.Ypad) % j.T);
    const int x = (int)(i / ((int64_t)j.Ypad * j.T));
    const int a = j.swap ? y : x, b = j.swap ? x : y;
    const int64_t pk = (int64_t)x * j.x_pitch + (int64_t)t * j.Ypad + y;
    if (!j.unpack) {
      j.dst[pk] = (y < Y) ? j.src[((int64_t)a * j.B + b) * j.T + t] : 0.f;
    } else if (y < Y) {
      j.dst[((int64_t)a * j.B + b) * j.T + t] = j.src[pk];
    }
  }
}

// col[m][kk], kk = (r*kw + s)*C + c (zero for kk >= kh*kw*C and for padding pixels):
// the 7x7 stem (C = 3) becomes a K-major GEMM operand for the tensor-core path.
__global__ void im2col_kernel(const float* __restrict__ in, float4* __restrict__ col, int N, int Hi,
                              int Wi, int pitch, int C, int kh, int kw, int stride, int pad, int Ho,
                              int Wo, int Kpad) {
  const int K4 = Kpad >> 2, K = kh * kw * C;
  const int64_t total = (int64_t)N * Ho * Wo * K4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k4 = (int)(i % K4);
    int64_t m = i / K4;
    const int ow = (int)(m % Wo); m /= Wo;
    const int oh = (int)(m % Ho);
    const int n = (int)(m / Ho);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = k4 * 4 + e;
      v[e] = 0.f;
      if (kk < K) {
        const int t = kk / C, c = kk - t * C;
        const int r = t / kw, sx = t - r * kw;
        const int ih = oh * stride - pad + r, iw = ow * stride - pad + sx;
        if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
          v[e] = __ldg(in + ((int64_t)(n * Hi + ih) * Wi + iw) * pitch + c);
      }
    }
    col[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                            float4* __restrict__ m, float4* __restrict__ v, int64_t n4, float lr,
                            float b1, float b2, float eps, float wd, float bc1, float rsqrt_bc2,
                            float gscale) {
  const float step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* Mv = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = G[k] * gscale;
      if (wd != 0.f) gr += wd * P[k];
      Mv[k] = b1 * Mv[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[k]) * rsqrt_bc2 + eps;
      P[k] -= step * (Mv[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

__global__ void adam_kernel1(float* __restrict__ p, const float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float rsqrt_bc2,
                             float gscale) {
  const float step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale;
    if (wd != 0.f) gr += wd * p[i];
    const float mm = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm; v[i] = vv;
    p[i] -= step * (mm / (sqrtf(vv) * rsqrt_bc2 + eps));
  }
}

// Device-side hyper-parameters (CUDA-graph friendly): hyper = [lr, beta1, beta2, eps, wd,
// grad_scale], *step_dev = 1-based step count (incremented by the caller before the launch).
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                float* __restrict__ m, float* __restrict__ v, int64_t n,
                                const float* __restrict__ hyper, const int* __restrict__ step_dev) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4],
              gscale = hyper[5];
  const double st = (double)*step_dev;
  const float bc1 = (float)(1.0 - pow((double)b1, st));
  const float rsqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, st)));
  const float step = lr / bc1;
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float* P = &pp.x; float* G = &gg.x; float* Mv = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = G[k] * gscale;
      if (wd != 0.f) gr += wd * P[k];
      Mv[k] = b1 * Mv[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      P[k] -= step * (Mv[k] / (sqrtf(V[k]) * rsqrt_bc2 + eps));
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
}

// hyper = [lr, momentum, wd, nesterov, grad_scale]; first step when *step_dev == 1
__global__ void sgd_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                               float* __restrict__ buf, int64_t n, const float* __restrict__ hyper,
                               const int* __restrict__ step_dev) {
  const float lr = hyper[0], mom = hyper[1], wd = hyper[2], gscale = hyper[4];
  const bool nesterov = hyper[3] != 0.f, first = (*step_dev == 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale;
    if (wd != 0.f) gr += wd * p[i];
    if (mom != 0.f) {
      const float b = first ? gr : mom * buf[i] + gr;
      buf[i] = b;
      gr = nesterov ? gr + mom * b : b;
    }
    p[i] -= lr * gr;
  }
}

__global__ void sgd_kernel1(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ buf, int64_t n, float lr, float mom, float wd,
                            int nesterov, int first, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * gscale;
    if (wd != 0.f) gr += wd * p[i];
    if (mom != 0.f) {
      const float b = first ? gr : mom * buf[i] + gr;
      buf[i] = b;
      gr = nesterov ? gr + mom * b : b;
    }
    p[i] -= lr * gr;
  }
}

__global__ void sgd_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                           float4* __restrict__ buf, int64_t n4, float lr, float mom, float wd,
                           int nesterov, int first, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i];
    float4 bb = buf ? buf[i] : make_float4(0, 0, 0, 0);
    float* P = &pp.x; float* G = &gg.x; float* Bf = &bb.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = G[k] * gscale;
      if (wd != 0.f) gr += wd * P[k];
      if (mom != 0.f) {
        Bf[k] = first ? gr : mom * Bf[k] + gr;
        gr = nesterov ? gr + mom * Bf[k] : Bf[k];
      }
      P[k] -= lr * gr;
    }
    p[i] = pp;
    if (buf) buf[i] = bb;
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int epb_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int Cpad,
                                epb_stream_t stream) {
  EPB_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C);
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (Cpad + 31) / 32, N);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, C, HW, Cpad);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int Cpad,
                                epb_stream_t stream) {
  EPB_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C);
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (Cpad + 31) / 32, N);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, C, HW, Cpad);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_im2col(const float* in, float* col, int N, int Hi, int Wi, int pitch, int C, int kh,
                          int kw, int stride, int pad, int Ho, int Wo, int Kpad,
                          epb_stream_t stream) {
  EPB_CHECK_ARG(in && col && N > 0 && Hi > 0 && Wi > 0 && C > 0 && pitch >= C);
  EPB_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && Ho > 0 && Wo > 0);
  EPB_CHECK_ARG(Kpad % 4 == 0 && Kpad >= kh * kw * C);
  const int64_t total = (int64_t)N * Ho * Wo * (Kpad / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  im2col_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(
      in, reinterpret_cast<float4*>(col), N, Hi, Wi, pitch, C, kh, kw, stride, pad, Ho, Wo, Kpad);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_pack_weight(const float* src, float* dst, int A, int B, int kh, int kw, int swap,
                               int Ypad, int unpack, epb_stream_t stream) {
  EPB_CHECK_ARG(src && dst && A > 0 && B > 0 && kh > 0 && kw > 0);
  EPB_CHECK_ARG(Ypad >= (swap ? A : B));
  const int64_t total = (int64_t)(swap ? B : A) * kh * kw * Ypad;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_weight_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, dst, A, B, kh * kw, swap,
                                                                 Ypad, unpack);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_pack_weight_batch(const epb_pack_job* jobs, int njobs, long long total_blocks,
                                     epb_stream_t stream) {
  EPB_CHECK_ARG(jobs && njobs > 0 && total_blocks > 0 && total_blocks < (1LL << 31));
  pack_weight_batch_kernel<<<(unsigned)total_blocks, 256, 0, as_stream(stream)>>>(jobs, njobs);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, float grad_scale, epb_stream_t stream) {
  EPB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  const bool vec = (n % 4 == 0) && (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg |
                                     (uintptr_t)exp_avg_sq) % 16 == 0);
  if (!vec) {
    blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    adam_kernel1<<<(int)blocks, 256, 0, as_stream(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
        (float)(1.0 / sqrt(bc2)), grad_scale);
    EPB_LAUNCH_CHECK();
    return EPB_OK;
  }
  adam_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
      reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), n / 4, lr, beta1,
      beta2, eps, weight_decay, (float)bc1, (float)(1.0 / sqrt(bc2)), grad_scale);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n,
                            float lr, float momentum, float weight_decay, int nesterov,
                            int first_step, float grad_scale, epb_stream_t stream) {
  EPB_CHECK_ARG(param && grad && n > 0);
  EPB_CHECK_ARG(momentum == 0.f || momentum_buf);
  const bool vec = (n % 4 == 0) &&
                   (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) % 16 == 0);
  if (!vec) {
    int64_t b1 = (n + 255) / 256;
    if (b1 > kNumSMs * 8) b1 = kNumSMs * 8;
    sgd_kernel1<<<(int)b1, 256, 0, as_stream(stream)>>>(param, grad, momentum_buf, n, lr, momentum,
                                                        weight_decay, nesterov, first_step,
                                                        grad_scale);
    EPB_LAUNCH_CHECK();
    return EPB_OK;
  }
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  sgd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
      reinterpret_cast<float4*>(momentum_buf), n / 4, lr, momentum, weight_decay, nesterov,
      first_step, grad_scale);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                 int64_t n, const float* hyper, const int* step_dev,
                                 epb_stream_t stream) {
  EPB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper && step_dev && n > 0 && n % 4 == 0);
  EPB_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0);
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  adam_dev_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                              hyper, step_dev);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_sgd_step_dev(float* param, const float* grad, float* momentum_buf, int64_t n,
                                const float* hyper, const int* step_dev, epb_stream_t stream) {
  EPB_CHECK_ARG(param && grad && momentum_buf && hyper && step_dev && n > 0);
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  sgd_dev_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(param, grad, momentum_buf, n, hyper,
                                                             step_dev);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

// ------------------------------------------------------------------ element-wise helpers of the
// refiner MLP (refiner/model.py): residual sums and dropout.  HBM bound, 4 B per element per stream.
namespace {
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            const float* __restrict__ c, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = c ? (a[i] + b[i]) + c[i] : a[i] + b[i];
}
__global__ void mask_scale_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                  float scale, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mask[i] ? x[i] * scale : 0.f;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int epb_add3(const float* a, const float* b, const float* c,
                                                              float* out, int64_t n, epb_stream_t stream) {
  EPB_CHECK_ARG(a && b && out && n >= 0);
  if (n == 0) return EPB_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  add3_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(a, b, c, out, n);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_mask_scale(const float* x, const uint8_t* mask, float scale,
                                                                    float* out, int64_t n, epb_stream_t stream) {
  EPB_CHECK_ARG(x && mask && out && n >= 0);
  if (n == 0) return EPB_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  mask_scale_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, mask, scale, out, n);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}


// torch.nn.utils.clip_grad_norm_(parameters, max_norm) (reference refiner/main.py:57) over a list
// of gradient tensors: epb_sumsq accumulates sum(x^2) of one tensor into a float64 scalar;
// epb_clip_scale multiplies one tensor by min(1, max_norm / (sqrt(total) + 1e-6)).
namespace {
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ total) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double v = (double)x[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) acc += sm[w];
    atomicAdd(total, acc);
  }
}
__global__ void __launch_bounds__(256)
clip_scale_kernel(float* __restrict__ x, int64_t n, const double* __restrict__ total, double max_norm) {
  const double coef = max_norm / (sqrt(*total) + 1e-6);
  if (coef >= 1.0) return;                      // clamp(coef, max=1): nothing to do
  const float c = (float)coef;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    x[i] *= c;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int epb_sumsq(const float* x, int64_t n, double* total,
                                                            epb_stream_t stream) {
  EPB_CHECK_ARG(x && total && n >= 0);
  if (n == 0) return EPB_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, total);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_clip_scale(float* x, int64_t n, const double* total,
                                                                 double max_norm, epb_stream_t stream) {
  EPB_CHECK_ARG(x && total && n >= 0 && max_norm > 0);
  if (n == 0) return EPB_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  clip_scale_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, total, max_norm);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}
