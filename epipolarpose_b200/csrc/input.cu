// Input pipeline on the device: the crop / colour / normalisation half of the reference
// lib/utils/img_utils.py:246-298 (get_single_patch_sample) once the frame is decoded:
//   generate_patch_image_cv (:114-127): gen_trans_from_patch_cv (:72-105, float32 point
//   triplets, cv2.getAffineTransform = 6x6 LU in float64) + cv2.warpAffine(INTER_LINEAR,
//   constant border 0) on uint8 BGR -> BGR->RGB (:268) -> per-channel colour scale, clip,
//   (x - mean) / std (:277-281) -> float32 [3][ph][pw];
// and the joint half (:283-296): joints through the same affine, depth scaling and
// generate_joint_location_label (lib/core/integral_loss.py:170-177).
//
// Byte / integer work, bit-exact against OpenCV: warpAffine's fixed-point source coordinates
// (10 fractional bits rounded per term, reduced to 5) and integer bilinear weights of sum 2^15
// are reproduced exactly (imgproc/imgwarp.cpp).  HBM bound: 12 B written per output pixel plus
// the touched source bytes; one thread per output pixel, x fastest (coalesced plane writes).
// Compiled with --fmad=false: the float64 affine arithmetic rounds as on the CPU.
#include "common.cuh"

namespace {

// cv::solve(A, B, DECOMP_LU) on the 6x6 system of cv2.getAffineTransform (hal::LU64f: partial
// pivoting by largest magnitude, row operations with alpha = A[j][i] * (-1 / A[i][i]),
// back substitution); src / dst are the three float32 point pairs.
__host__ __device__ inline bool affine_lu6(const float (*src)[2], const float (*dst)[2], double* M) {
  double A[6][6], b[6];
  for (int i = 0; i < 3; ++i) {
    const int r0 = 2 * i, r1 = 2 * i + 1;
    for (int k = 0; k < 6; ++k) { A[r0][k] = 0.0; A[r1][k] = 0.0; }
    A[r0][0] = A[r1][3] = (double)src[i][0];
    A[r0][1] = A[r1][4] = (double)src[i][1];
    A[r0][2] = A[r1][5] = 1.0;
    b[r0] = (double)dst[i][0];
    b[r1] = (double)dst[i][1];
  }
  const double eps = 2.220446049250313e-16 * 100;
  for (int i = 0; i < 6; ++i) {
    int k = i;
    for (int j = i + 1; j < 6; ++j)
      if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
    if (fabs(A[k][i]) < eps) return false;
    if (k != i) {
      for (int j = i; j < 6; ++j) { const double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
      const double t = b[i]; b[i] = b[k]; b[k] = t;
    }
    const double d = -1.0 / A[i][i];
    for (int j = i + 1; j < 6; ++j) {
      const double alpha = A[j][i] * d;
      for (int kk = i + 1; kk < 6; ++kk) A[j][kk] += alpha * A[i][kk];
      b[j] += alpha * b[i];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    for (int kk = i + 1; kk < 6; ++kk) s -= A[i][kk] * b[kk];
    b[i] = s / A[i][i];
  }
  for (int k = 0; k < 6; ++k) M[k] = b[k];
  return true;
}

// img_utils.py:72-105 (inv = False): image -> patch transform, row-major 2x3 in M.
// box = (c_x, c_y, bb_width, bb_height, scale, rot)
__host__ __device__ inline bool patch_affine_fwd(const double* box, double patch_w, double patch_h,
                                                 double* M) {
  const double c_x = box[0], c_y = box[1], scale = box[4], rot = box[5];
  const double src_w = box[2] * scale, src_h = box[3] * scale;
  const double rot_rad = 3.141592653589793 * rot / 180;
  const double sn = sin(rot_rad), cs = cos(rot_rad);
  // rotate_2d(np.array([0, src_h*0.5], f32), rot_rad) -> f32 (:63-69)
  const double dy_ = (double)(float)(src_h * 0.5), rx_ = (double)(float)(src_w * 0.5);
  const float down_x = (float)(0.0 * cs - dy_ * sn), down_y = (float)(0.0 * sn + dy_ * cs);
  const float right_x = (float)(rx_ * cs - 0.0 * sn), right_y = (float)(rx_ * sn + 0.0 * cs);
  float s[3][2], d[3][2];
  s[0][0] = (float)c_x;                       s[0][1] = (float)c_y;
  s[1][0] = (float)(c_x + (double)down_x);    s[1][1] = (float)(c_y + (double)down_y);
  s[2][0] = (float)(c_x + (double)right_x);   s[2][1] = (float)(c_y + (double)right_y);
  const float dcx = (float)(patch_w * 0.5), dcy = (float)(patch_h * 0.5);
  d[0][0] = dcx;        d[0][1] = dcy;
  d[1][0] = dcx + 0.f;  d[1][1] = dcy + (float)(patch_h * 0.5);
  d[2][0] = dcx + (float)(patch_w * 0.5);  d[2][1] = dcy + 0.f;
  return affine_lu6(s, d, M);
}

// cv::warpAffine without WARP_INVERSE_MAP: invert the 2x3 map in float64 (imgwarp.cpp)
__host__ __device__ inline void invert_affine(const double* M, double* iM) {
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1.0 / D : 0.0;
  const double A11 = M[4] * D, A22 = M[0] * D;
  iM[0] = A11; iM[1] = M[1] * (-D);
  iM[3] = M[3] * (-D); iM[4] = A22;
  const double b1 = -iM[0] * M[2] - iM[1] * M[5];
  const double b2 = -iM[3] * M[2] - iM[4] * M[5];
  iM[2] = b1; iM[5] = b2;
}

// saturate_cast<int>(double) of OpenCV = cvRound: round half to even, saturating
__host__ __device__ inline long long cv_round(double v) {
  const double r = rint(v);
  if (r >= 2147483647.0) return 2147483647LL;
  if (r <= -2147483648.0) return -2147483648LL;
  return (long long)r;
}

// one output pixel of the uint8 BGR patch (warpAffine INTER_LINEAR, constant border 0).
// flip: the source is the horizontally mirrored frame (img[:, ::-1, :], img_utils.py:119).
__host__ __device__ inline void warp_pixel_u8(const uint8_t* img, int H, int W, int64_t pitch, int flip,
                                              const double* iM, int x, int y, int (&bgr)[3]) {
  const long long adelta = cv_round(iM[0] * x * 1024.0), bdelta = cv_round(iM[3] * x * 1024.0);
  const long long X0 = cv_round((iM[1] * y + iM[2]) * 1024.0) + 16;
  const long long Y0 = cv_round((iM[4] * y + iM[5]) * 1024.0) + 16;
  const long long X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const long long sx = X >> 5, sy = Y >> 5;
  const int ax = (int)(X & 31), ay = (int)(Y & 31);
  const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32,
            w11 = ax * ay * 32;
  int acc[3] = {0, 0, 0};
  auto tap = [&](long long yy, long long xx, int wgt) {
    if (wgt == 0 || yy < 0 || yy >= H || xx < 0 || xx >= W) return;
    const long long xs = flip ? (W - 1 - xx) : xx;
    const uint8_t* p = img + yy * pitch + xs * 3;
    acc[0] += wgt * p[0];
    acc[1] += wgt * p[1];
    acc[2] += wgt * p[2];
  };
  tap(sy, sx, w00);
  tap(sy, sx + 1, w01);
  tap(sy + 1, sx, w10);
  tap(sy + 1, sx + 1, w11);
  for (int c = 0; c < 3; ++c) {
    const int v = (acc[c] + (1 << 14)) >> 15;
    bgr[c] = v < 0 ? 0 : (v > 255 ? 255 : v);
  }
}

// img_utils.py:268-281 for output channel c (RGB order) of one pixel
__host__ __device__ inline float finish_pixel(int v_u8, float color_scale, bool norm, double mean,
                                              double stdv) {
  float f = (float)v_u8 * color_scale;              // float32 array * python float
  f = f < 0.f ? 0.f : (f > 255.f ? 255.f : f);      // np.clip(., 0, 255)
  if (!norm) return f;
  return (float)(((double)f - mean) / stdv);        // float64 scalars promote; stored as float32
}

struct PatchArgs {
  double mean[3], stdv[3];
  int norm;
};

// lib/utils/augmentation.py:81-114 paste_over for ONE patch pixel (x, y) and one occluder: the
// RGBA occluder (w x h) centred at (cx, cy) (already np.round'ed), alpha-blended in float32
//   alpha * src + (1 - alpha) * dst   with alpha = a / 255 (float32),
// and stored back into the uint8 image (C truncation), exactly as numpy evaluates :112-113.
__host__ __device__ inline void paste_pixel(const uint8_t* occ, int w, int h, int cx, int cy, int x,
                                            int y, int (&rgb)[3]) {
  const int sx = x - (cx - w / 2), sy = y - (cy - h / 2);     // raw_start_dst = center - wh // 2
  if (sx < 0 || sx >= w || sy < 0 || sy >= h) return;
  const uint8_t* p = occ + ((int64_t)sy * w + sx) * 4;
  const float alpha = (float)p[3] / 255.f;
  const float om = 1.f - alpha;
  for (int c = 0; c < 3; ++c) {
    const float v = alpha * (float)p[c] + om * (float)rgb[c];
    rgb[c] = (int)(uint8_t)v;
  }
}

constexpr int kMaxOccluders = 7;       // count = np.random.randint(1, 8)  (:67)

// per sample, once: the forward patch affine (returned as `trans`) and its inverse (what warpAffine
// walks); the pixel kernel's blocks only read the six inverse coefficients
__global__ void patch_affine_kernel(const int32_t* __restrict__ img_hwp, const double* __restrict__ box,
                                    const int32_t* __restrict__ flip, int B, int patch_w, int patch_h,
                                    double* __restrict__ trans, double* __restrict__ inv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int W = img_hwp[b * 3 + 1];
  double bx[6];
  for (int k = 0; k < 6; ++k) bx[k] = box[b * 6 + k];
  if (flip && flip[b]) bx[0] = (double)W - bx[0] - 1.0;          // c_x = img_width - c_x - 1 (:120)
  double M[6] = {0, 0, 0, 0, 0, 0};
  patch_affine_fwd(bx, (double)patch_w, (double)patch_h, M);
  double iM[6];
  invert_affine(M, iM);
  for (int k = 0; k < 6; ++k) inv[b * 6 + k] = iM[k];
  if (trans)
    for (int k = 0; k < 6; ++k) trans[b * 6 + k] = M[k];
}

__global__ void __launch_bounds__(256)
patch_sample_kernel(const uint8_t* __restrict__ img_base, const int64_t* __restrict__ img_off,
                    const int32_t* __restrict__ img_hwp, const double* __restrict__ inv,
                    const int32_t* __restrict__ flip, const float* __restrict__ color, PatchArgs pa,
                    int patch_w, int patch_h, const uint8_t* __restrict__ occ_base,
                    const int64_t* __restrict__ occ_desc, const int32_t* __restrict__ occ_count,
                    float* __restrict__ out) {
  const int b = blockIdx.z;
  const int H = img_hwp[b * 3 + 0], W = img_hwp[b * 3 + 1];
  const int64_t pitch = img_hwp[b * 3 + 2];
  const int fl = flip ? flip[b] : 0;
  double siM[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) siM[k] = inv[b * 6 + k];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= patch_w || y >= patch_h) return;
  int bgr[3];
  warp_pixel_u8(img_base + img_off[b], H, W, pitch, fl, siM, x, y, bgr);
  int rgb[3] = {bgr[2], bgr[1], bgr[0]};              // image[:, :, ::-1] (:268)
  if (occ_count) {                                     // occlude_with_objects (:269-270), in order
    const int cnt = occ_count[b];
    for (int k = 0; k < cnt && k < kMaxOccluders; ++k) {
      const int64_t* d = occ_desc + ((int64_t)b * kMaxOccluders + k) * 5;
      paste_pixel(occ_base + d[0], (int)d[1], (int)d[2], (int)d[3], (int)d[4], x, y, rgb);
    }
  }
  const int64_t plane = (int64_t)patch_w * patch_h;
  float* o = out + (int64_t)b * 3 * plane + (int64_t)y * patch_w + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float cs = color ? color[b * 3 + c] : 1.f;
    o[c * plane] = finish_pixel(rgb[c], cs, pa.norm != 0, pa.mean[c], pa.stdv[c]);
  }
}

// joints half: one thread per (sample, joint)
__host__ __device__ inline void patch_joint(const double* jt, const double* M, double patch_w,
                                            double patch_h, double depth_den, double* label) {
  const double x = (M[0] * jt[0] + M[1] * jt[1]) + M[2] * 1.0;     // np.dot(trans, [x, y, 1])
  const double y = (M[3] * jt[0] + M[4] * jt[1]) + M[5] * 1.0;
  const double z = jt[2] / depth_den * patch_w;                    // :291-293
  label[0] = x / patch_w - 0.5;                                    // integral_loss.py:171-173
  label[1] = y / patch_h - 0.5;
  label[2] = z / patch_w;
}

__global__ void patch_joints_kernel(const double* __restrict__ joints, const double* __restrict__ box,
                                    const double* __restrict__ trans, int B, int J, double patch_w,
                                    double patch_h, double rect_3d_w, int depth_in_image,
                                    double* __restrict__ label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * J) return;
  const int b = i / J;
  const double scale = box[b * 6 + 4];
  const double den = depth_in_image ? box[b * 6 + 2] * scale : rect_3d_w * scale;
  patch_joint(joints + (int64_t)i * 3, trans + b * 6, patch_w, patch_h, den, label + (int64_t)i * 3);
}


// ---------------------------------------------------------------------------------------
// lib/core/inference.py:43-68 get_final_preds on the device: hard argmax per (n, j) map
// (first index on ties, masked to (0,0) where max <= 0, :12-40), +-0.25 px toward the higher
// neighbour (:49-61), and transform_preds (lib/utils/transforms.py:39-44): the heat-map ->
// image affine of get_affine_transform(center, scale, 0, (W, H), inv=1) (:47-79: float32
// point triplets, cv2.getAffineTransform = the 6x6 LU above), applied in float64 and stored
// as float32 like the reference's `preds[i] = ...` assignment.  One warp per map.
__device__ __forceinline__ bool fp_better(float v, int i, float bv, int bi) {
  return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256)
final_preds_kernel(const float* __restrict__ hm, int NJ, int J, int H, int W,
                   const double* __restrict__ center, const double* __restrict__ scale,
                   int post_process, float* __restrict__ preds, float* __restrict__ maxvals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= NJ) return;
  const int HW = H * W;
  const float* p = hm + (int64_t)warp * HW;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < HW; i += 32) {
    const float v = p[i];
    if (fp_better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (fp_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane != 0) return;
  if (bi == 0x7fffffff) bi = 0;
  const float mask = (bv > 0.f) ? 1.f : 0.f;
  float cx = (float)(bi % W) * mask, cy = floorf((float)bi / (float)W) * mask;
  if (post_process) {
    const int px = (int)floor((double)cx + 0.5), py = (int)floor((double)cy + 0.5);
    if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
      const float dx = p[py * W + px + 1] - p[py * W + px - 1];
      const float dy = p[(py + 1) * W + px] - p[(py - 1) * W + px];
      cx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
      cy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
    }
  }
  // get_affine_transform(center, scale, rot = 0, [W, H], inv = 1)
  const int n = warp / J;
  const double c0 = center[2 * n], c1 = center[2 * n + 1];
  const double src_w = scale[2 * n] * 200.0;
  const double dir0 = 0.0 * 1.0 - (src_w * -0.5) * 0.0, dir1 = 0.0 * 0.0 + (src_w * -0.5) * 1.0;   // get_dir, rot 0
  float src[3][2], dst[3][2];
  src[0][0] = (float)(c0 + src_w * 0.0);   src[0][1] = (float)(c1 + scale[2 * n + 1] * 200.0 * 0.0);
  src[1][0] = (float)(c0 + dir0 + src_w * 0.0);
  src[1][1] = (float)(c1 + dir1 + scale[2 * n + 1] * 200.0 * 0.0);
  const float dst_dir1 = (float)((double)W * -0.5);
  dst[0][0] = (float)((double)W * 0.5);    dst[0][1] = (float)((double)H * 0.5);
  dst[1][0] = (float)((double)W * 0.5 + 0.0);
  dst[1][1] = (float)((double)H * 0.5 + (double)dst_dir1);
  // get_3rd_point(a, b) = b + (-(a - b)[1], (a - b)[0]) in float32
  {
    const float d0 = src[0][0] - src[1][0], d1 = src[0][1] - src[1][1];
    src[2][0] = src[1][0] + (-d1);  src[2][1] = src[1][1] + d0;
    const float e0 = dst[0][0] - dst[1][0], e1 = dst[0][1] - dst[1][1];
    dst[2][0] = dst[1][0] + (-e1);  dst[2][1] = dst[1][1] + e0;
  }
  double M[6];
  if (!affine_lu6(dst, src, M)) {
    for (int k = 0; k < 6; ++k) M[k] = 0.0;
  }
  // affine_transform: np.dot(t, [x, y, 1.]) in float64, stored to the float32 result
  const double x = (double)cx, y = (double)cy;
  preds[warp * 2 + 0] = (float)(M[0] * x + M[1] * y + M[2] * 1.0);
  preds[warp * 2 + 1] = (float)(M[3] * x + M[4] * y + M[5] * 1.0);
  if (maxvals) maxvals[warp] = bv;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int epb_patch_sample_occ(
    const uint8_t* img_base, const int64_t* img_off, const int32_t* img_hwp, const double* box,
    const int32_t* flip, const float* color, const double* mean_std_host, int B, int patch_w,
    int patch_h, const uint8_t* occ_base, const int64_t* occ_desc, const int32_t* occ_count,
    float* out, double* trans, epb_stream_t stream) {
  EPB_CHECK_ARG(img_base && img_off && img_hwp && box && out);
  EPB_CHECK_ARG(B >= 0 && patch_w > 0 && patch_h > 0 && B <= 65535);
  EPB_CHECK_ARG((occ_base == nullptr) == (occ_desc == nullptr) && (occ_desc == nullptr) == (occ_count == nullptr));
  if (B == 0) return EPB_OK;
  PatchArgs pa;
  pa.norm = mean_std_host ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    pa.mean[c] = mean_std_host ? mean_std_host[c] : 0.0;
    pa.stdv[c] = mean_std_host ? mean_std_host[3 + c] : 1.0;
    EPB_CHECK_ARG(pa.stdv[c] != 0.0);
  }
  const dim3 block(64, 4);
  const dim3 grid((patch_w + 63) / 64, (patch_h + 3) / 4, B);
  cudaStream_t st = as_stream(stream);
  void* inv = nullptr;
  int rc = epb_workspace(EPB_WS_PATCHINV, (size_t)65536 * 6 * sizeof(double), st, &inv);
  if (rc) return rc;
  patch_affine_kernel<<<(B + 127) / 128, 128, 0, st>>>(img_hwp, box, flip, B, patch_w, patch_h, trans,
                                                      static_cast<double*>(inv));
  EPB_LAUNCH_CHECK();
  patch_sample_kernel<<<grid, block, 0, st>>>(img_base, img_off, img_hwp, static_cast<const double*>(inv),
                                              flip, color, pa, patch_w, patch_h, occ_base, occ_desc,
                                              occ_count, out);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_patch_sample(
    const uint8_t* img_base, const int64_t* img_off, const int32_t* img_hwp, const double* box,
    const int32_t* flip, const float* color, const double* mean_std_host, int B, int patch_w,
    int patch_h, float* out, double* trans, epb_stream_t stream) {
  return epb_patch_sample_occ(img_base, img_off, img_hwp, box, flip, color, mean_std_host, B, patch_w,
                              patch_h, nullptr, nullptr, nullptr, out, trans, stream);
}

extern "C" __attribute__((visibility("default"))) int epb_patch_joints(
    const double* joints, const double* box, const double* trans, int B, int J, double patch_w,
    double patch_h, double rect_3d_w, int depth_in_image, double* label, epb_stream_t stream) {
  EPB_CHECK_ARG(joints && box && trans && label);
  EPB_CHECK_ARG(B >= 0 && J >= 0 && patch_w > 0 && patch_h > 0 && rect_3d_w != 0);
  if (B * J == 0) return EPB_OK;
  const int n = B * J;
  patch_joints_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(joints, box, trans, B, J, patch_w,
                                                                     patch_h, rect_3d_w, depth_in_image,
                                                                     label);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_final_preds(
    const float* hm, int N, int J, int H, int W, const double* center, const double* scale,
    int post_process, float* preds, float* maxvals, epb_stream_t stream) {
  EPB_CHECK_ARG(hm && center && scale && preds);
  EPB_CHECK_ARG(N >= 0 && J > 0 && H > 0 && W > 0 && (int64_t)H * W < (1LL << 31));
  if (N == 0) return EPB_OK;
  const int NJ = N * J;
  const int threads = 256;
  const int blocks = (NJ * 32 + threads - 1) / threads;
  final_preds_kernel<<<blocks, threads, 0, as_stream(stream)>>>(hm, NJ, J, H, W, center, scale,
                                                                post_process, preds, maxvals);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}
