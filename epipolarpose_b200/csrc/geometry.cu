// Epipolar geometry in float64: patch->image affine, two-view triangulation
// (homogeneous DLT / linear LS / iterative LS) and projection to labels.
// Reference arithmetic: lib/utils/triangulation.py, lib/utils/img_utils.py:63-111,
// 141-155,193-243, lib/utils/prep_h36m.py:170-204, lib/core/integral_loss.py:170-205.
// The SVDs the reference obtains from OpenCV (cv2.triangulatePoints,
// cv2.solve(DECOMP_SVD)) are one-sided Jacobi SVDs on the un-squared matrix,
// as in OpenCV's JacobiSVD (so the condition number is not squared).
//
// One thread per (pair, joint): the problem is a few hundred bytes per joint
// and latency-bound at real sizes (64 pairs x 17 joints); no tensor cores.
// Compiled with --fmad=false so that a*b+c rounds twice as on the CPU.
#include "common.cuh"
#include <float.h>

namespace {

// One-sided (Hestenes) Jacobi: rotate columns of A (R x C) until mutually
// orthogonal; V (C x C) accumulates the right rotations.  On exit
// A = U*diag(sigma) (columns), sigma_j = ||A[:,j]||.
template <int R, int C>
__host__ __device__ void jacobi_onesided(double (&A)[R][C], double (&V)[C][C]) {
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool changed = false;
#pragma unroll
    for (int p = 0; p < C - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < C; ++q) {
        double a = 0, b = 0, g = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          a += A[i][p] * A[i][p];
          b += A[i][q] * A[i][q];
          g += A[i][p] * A[i][q];
        }
        if (fabs(g) <= DBL_EPSILON * sqrt(a * b) || g == 0.0) continue;
        changed = true;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const double x = A[i][p], y = A[i][q];
          A[i][p] = c * x - s * y;
          A[i][q] = s * x + c * y;
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const double x = V[i][p], y = V[i][q];
          V[i][p] = c * x - s * y;
          V[i][q] = s * x + c * y;
        }
      }
    }
    if (!changed) break;
  }
}

// least-squares solve of A(4x3) x = b via SVD, as cv2.solve(.., DECOMP_SVD)
__device__ void solve_ls_4x3(const double (&A0)[4][3], const double (&b)[4], double (&x)[3]) {
  double A[4][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = A0[i][j];
  jacobi_onesided<4, 3>(A, V);
  double s2[3], y[3], ssum = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double n2 = 0, d = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { n2 += A[i][j] * A[i][j]; d += A[i][j] * b[i]; }
    s2[j] = n2;
    y[j] = d;
    ssum += sqrt(n2);
  }
  const double thr = DBL_EPSILON * 2.0 * ssum;  // cv::SVD::backSubst threshold
#pragma unroll
  for (int j = 0; j < 3; ++j) y[j] = (sqrt(s2[j]) > thr) ? y[j] / s2[j] : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = V[i][0] * y[0] + V[i][1] * y[1] + V[i][2] * y[2];
}

// triangulation.py:139-150 / :80-92: rows C*P[:3,:3], b = -(C*P[:3,3]),
// C = [[-1,0,u],[0,-1,v]]
__device__ void build_Ab(const double* u1, const double* P1, const double* u2, const double* P2,
                         double (&A)[4][3], double (&b)[4]) {
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const double* u = v ? u2 : u1;
    const double* P = v ? P2 : P1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      A[2 * v + 0][k] = -P[0 * 4 + k] + u[0] * P[2 * 4 + k];
      A[2 * v + 1][k] = -P[1 * 4 + k] + u[1] * P[2 * 4 + k];
    }
    b[2 * v + 0] = -(-P[0 * 4 + 3] + u[0] * P[2 * 4 + 3]);
    b[2 * v + 1] = -(-P[1 * 4 + 3] + u[1] * P[2 * 4 + 3]);
  }
}

// --------------------------------------------------------------- polynomial (optimal) correction
// lib/utils/triangulation.py:184-220: F from the two projection matrices, cv2.correctMatches
// (Hartley-Sturm, Hartley & Zisserman Alg. 12.1) and the homogeneous DLT on the corrected
// matches.  OpenCV finds the six roots of the stationarity polynomial with cvSolvePoly; here:
// Laguerre iterations with deflation and a polishing pass on the un-deflated polynomial
// (complex float64), which reaches the same roots to rounding.
struct cplx { double re, im; };
__host__ __device__ inline cplx cmk(double r, double i) { cplx c; c.re = r; c.im = i; return c; }
__host__ __device__ inline cplx cadd(cplx a, cplx b) { return cmk(a.re + b.re, a.im + b.im); }
__host__ __device__ inline cplx csub(cplx a, cplx b) { return cmk(a.re - b.re, a.im - b.im); }
__host__ __device__ inline cplx cmul(cplx a, cplx b) {
  return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__host__ __device__ inline cplx cscale(cplx a, double k) { return cmk(a.re * k, a.im * k); }
__host__ __device__ inline double cabs2(cplx a) { return hypot(a.re, a.im); }
__host__ __device__ inline cplx cdiv(cplx a, cplx b) {       // Smith's algorithm
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, den = b.re + r * b.im;
    return cmk((a.re + r * a.im) / den, (a.im - r * a.re) / den);
  }
  const double r = b.re / b.im, den = b.im + r * b.re;
  return cmk((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}
__host__ __device__ inline cplx csqrt2(cplx z) {
  if (z.re == 0.0 && z.im == 0.0) return cmk(0.0, 0.0);
  const double x = fabs(z.re), y = fabs(z.im);
  double w;
  if (x >= y) { const double r = y / x; w = sqrt(x) * sqrt(0.5 * (1.0 + sqrt(1.0 + r * r))); }
  else { const double r = x / y; w = sqrt(y) * sqrt(0.5 * (r + sqrt(1.0 + r * r))); }
  if (z.re >= 0.0) return cmk(w, z.im / (2.0 * w));
  return cmk(y / (2.0 * w), z.im >= 0.0 ? w : -w);
}

// one root of sum_{k<=m} a[k] x^k by Laguerre's method, starting from (and returned in) x
__host__ __device__ inline void laguerre(const cplx* a, int m, cplx& x) {
  const double frac[9] = {0.0, 0.5, 0.25, 0.75, 0.13, 0.38, 0.62, 0.88, 1.0};
  for (int iter = 1; iter <= 80; ++iter) {
    cplx b = a[m], d = cmk(0, 0), f = cmk(0, 0);
    double err = cabs2(b);
    const double abx = cabs2(x);
    for (int j = m - 1; j >= 0; --j) {
      f = cadd(cmul(x, f), d);
      d = cadd(cmul(x, d), b);
      b = cadd(cmul(x, b), a[j]);
      err = cabs2(b) + abx * err;
    }
    if (cabs2(b) <= err * 2.0e-16) return;                 // on the root to rounding
    const cplx g = cdiv(d, b), g2 = cmul(g, g);
    const cplx h = csub(g2, cscale(cdiv(f, b), 2.0));
    const cplx sq = csqrt2(cscale(csub(cscale(h, (double)m), g2), (double)(m - 1)));
    cplx gp = cadd(g, sq);
    const cplx gm = csub(g, sq);
    const double abp = cabs2(gp), abm = cabs2(gm);
    if (abp < abm) gp = gm;
    cplx dx;
    if (fmax(abp, abm) > 0.0) dx = cdiv(cmk((double)m, 0.0), gp);
    else dx = cscale(cmk(cos((double)iter), sin((double)iter)), 1.0 + abx);
    const cplx x1 = csub(x, dx);
    if (x.re == x1.re && x.im == x1.im) return;
    if (iter % 10) x = x1;
    else x = csub(x, cscale(dx, frac[(iter / 10) % 9]));
  }
}

// real parts of the roots of the real polynomial sum_{k<=6} k[k] t^k (leading zeros stripped)
__host__ __device__ inline int poly6_root_reals(const double* k, double* re) {
  int m = 6;
  double big = 0.0;
  for (int i = 0; i <= 6; ++i) big = fmax(big, fabs(k[i]));
  while (m > 0 && fabs(k[m]) <= big * 1e-300) --m;
  if (m == 0) return 0;
  cplx a[7], ad[7];
  for (int i = 0; i <= m; ++i) { a[i] = cmk(k[i], 0.0); ad[i] = a[i]; }
  cplx roots[6];
  for (int j = m; j >= 1; --j) {
    cplx x = cmk(0.0, 0.0);
    laguerre(ad, j, x);
    if (fabs(x.im) <= 4.0e-16 * fabs(x.re)) x.im = 0.0;
    roots[j - 1] = x;
    cplx b = ad[j];
    for (int jj = j - 1; jj >= 0; --jj) {                  // forward deflation
      const cplx c = ad[jj];
      ad[jj] = b;
      b = cadd(cmul(x, b), c);
    }
  }
  for (int j = 0; j < m; ++j) {
    laguerre(a, m, roots[j]);                              // polish on the full polynomial
    re[j] = roots[j].re;
  }
  return m;
}

// [t]x R of the canonical pair P2_full * inv(P1_full)  (triangulation.py:198-204)
__host__ __device__ inline void fundamental_from_P(const double* P1, const double* P2, double* F) {
  // inv([M p; 0 1]) = [M^-1  -M^-1 p; 0 1]
  const double m00 = P1[0], m01 = P1[1], m02 = P1[2], m10 = P1[4], m11 = P1[5], m12 = P1[6],
               m20 = P1[8], m21 = P1[9], m22 = P1[10];
  const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
  const double det = m00 * c00 + m01 * c01 + m02 * c02;
  double Mi[3][3];
  Mi[0][0] = c00 / det; Mi[0][1] = (m02 * m21 - m01 * m22) / det; Mi[0][2] = (m01 * m12 - m02 * m11) / det;
  Mi[1][0] = c01 / det; Mi[1][1] = (m00 * m22 - m02 * m20) / det; Mi[1][2] = (m02 * m10 - m00 * m12) / det;
  Mi[2][0] = c02 / det; Mi[2][1] = (m01 * m20 - m00 * m21) / det; Mi[2][2] = (m00 * m11 - m01 * m10) / det;
  double R[3][3], t[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      R[r][c] = P2[r * 4 + 0] * Mi[0][c] + P2[r * 4 + 1] * Mi[1][c] + P2[r * 4 + 2] * Mi[2][c];
    t[r] = P2[r * 4 + 3] - (R[r][0] * P1[3] + R[r][1] * P1[7] + R[r][2] * P1[11]);
    // a translation that is pure cancellation noise (identical camera centres) is exactly zero:
    // the reference's F is then the zero matrix and its correction all-NaN (:213-217)
    const double mag = fabs(P2[r * 4 + 3]) + fabs(R[r][0] * P1[3]) + fabs(R[r][1] * P1[7]) +
                       fabs(R[r][2] * P1[11]);
    if (fabs(t[r]) <= 64.0 * DBL_EPSILON * mag) t[r] = 0.0;
  }
  for (int c = 0; c < 3; ++c) {                            // F[:,c] = t x R[:,c]
    F[0 * 3 + c] = t[1] * R[2][c] - t[2] * R[1][c];
    F[1 * 3 + c] = t[2] * R[0][c] - t[0] * R[2][c];
    F[2 * 3 + c] = t[0] * R[1][c] - t[1] * R[0][c];
  }
}

// cv2.correctMatches for one match (u1, u2 updated in place)
__host__ __device__ inline void correct_match(const double* F, double* u1, double* u2) {
  const double x1 = u1[0], y1 = u1[1], x2 = u2[0], y2 = u2[1];
  // TFT = T2i^T F T1i with T = [1 0 x; 0 1 y; 0 0 1]
  double G[3][3];
  for (int r = 0; r < 3; ++r) {
    G[r][0] = F[r * 3 + 0];
    G[r][1] = F[r * 3 + 1];
    G[r][2] = F[r * 3 + 0] * x1 + F[r * 3 + 1] * y1 + F[r * 3 + 2];
  }
  double T[3][3];
  for (int c = 0; c < 3; ++c) {
    T[0][c] = G[0][c];
    T[1][c] = G[1][c];
    T[2][c] = x2 * G[0][c] + y2 * G[1][c] + G[2][c];
  }
  // epipoles: right null vector = cross product of the two most independent rows, left null
  // vector likewise from the columns (OpenCV takes them from the SVD of the rank-2 matrix)
  auto null_of = [&](bool rows, double (&e)[3]) {
    double best = -1.0;
    for (int i = 0; i < 3; ++i) {
      const int j = (i + 1) % 3;
      double a[3], b[3], c[3];
      for (int k = 0; k < 3; ++k) { a[k] = rows ? T[i][k] : T[k][i]; b[k] = rows ? T[j][k] : T[k][j]; }
      c[0] = a[1] * b[2] - a[2] * b[1];
      c[1] = a[2] * b[0] - a[0] * b[2];
      c[2] = a[0] * b[1] - a[1] * b[0];
      const double n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      if (n > best) { best = n; e[0] = c[0]; e[1] = c[1]; e[2] = c[2]; }
    }
    const double s = sqrt(e[0] * e[0] + e[1] * e[1]);
    e[0] /= s; e[1] /= s; e[2] /= s;
  };
  double e1[3], e2[3];
  null_of(true, e1);
  null_of(false, e2);
  // RF = R2 TFT R1^T, R = [ex ey 0; -ey ex 0; 0 0 1]
  double H[3][3];
  for (int r = 0; r < 3; ++r) {                            // H = TFT R1^T
    H[r][0] = T[r][0] * e1[0] + T[r][1] * e1[1];
    H[r][1] = -T[r][0] * e1[1] + T[r][1] * e1[0];
    H[r][2] = T[r][2];
  }
  const double a = -e2[1] * H[0][1] + e2[0] * H[1][1], b = -e2[1] * H[0][2] + e2[0] * H[1][2];
  const double c = H[2][1], d = H[2][2];
  const double f1 = e1[2], f2 = e2[2];
  // g(t) = t((at+b)^2 + f2^2(ct+d)^2)^2 - (ad-bc)(1+f1^2 t^2)^2 (at+b)(ct+d), ascending powers
  const double q0 = b * b + f2 * f2 * d * d, q1 = 2.0 * (a * b + f2 * f2 * c * d),
               q2 = a * a + f2 * f2 * c * c;
  const double qq[5] = {q0 * q0, 2.0 * q0 * q1, 2.0 * q0 * q2 + q1 * q1, 2.0 * q1 * q2, q2 * q2};
  const double w = a * d - b * c, ff = f1 * f1;
  const double r0 = b * d, r1 = a * d + b * c, r2 = a * c;              // (at+b)(ct+d)
  const double p4[5] = {1.0, 0.0, 2.0 * ff, 0.0, ff * ff};               // (1+f1^2 t^2)^2
  double k[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 5; ++i) k[i + 1] += qq[i];
  for (int i = 0; i < 5; ++i) {
    k[i] -= w * p4[i] * r0;
    k[i + 1] -= w * p4[i] * r1;
    k[i + 2] -= w * p4[i] * r2;
  }
  double re[6];
  const int nr = poly6_root_reals(k, re);
  double smin = DBL_MAX, tmin = 0.0;
  for (int i = 0; i < nr; ++i) {
    const double t = re[i];
    const double n1 = c * t + d, n0 = a * t + b;
    const double sv = t * t / (1.0 + ff * t * t) + n1 * n1 / (n0 * n0 + f2 * f2 * n1 * n1);
    if (sv < smin) { smin = sv; tmin = t; }
  }
  const double sinf = 1.0 / ff + c * c / (a * a + f2 * f2 * c * c);
  double l1[3], l2[3];
  if (sinf < smin) {                                       // minimum at t = infinity
    l1[0] = f1; l1[1] = 0.0; l1[2] = -1.0;
    l2[0] = -f2 * c; l2[1] = a; l2[2] = c;
  } else {
    l1[0] = tmin * f1; l1[1] = 1.0; l1[2] = -tmin;
    l2[0] = -f2 * (c * tmin + d); l2[1] = a * tmin + b; l2[2] = c * tmin + d;
  }
  // closest points to the origin on the two lines, rotated and translated back
  const double h1[3] = {-l1[0] * l1[2], -l1[1] * l1[2], l1[0] * l1[0] + l1[1] * l1[1]};
  const double h2[3] = {-l2[0] * l2[2], -l2[1] * l2[2], l2[0] * l2[0] + l2[1] * l2[1]};
  const double g1[3] = {e1[0] * h1[0] - e1[1] * h1[1], e1[1] * h1[0] + e1[0] * h1[1], h1[2]};   // R1^T h1
  const double g2[3] = {e2[0] * h2[0] - e2[1] * h2[1], e2[1] * h2[0] + e2[0] * h2[1], h2[2]};
  u1[0] = (g1[0] + x1 * g1[2]) / g1[2];
  u1[1] = (g1[1] + y1 * g1[2]) / g1[2];
  u2[0] = (g2[0] + x2 * g2[2]) / g2[2];
  u2[1] = (g2[1] + y2 * g2[2]) / g2[2];
}

// cv2.findFundamentalMat(u1, u2, FM_8POINT) for one pair (OpenCV calib3d fundam.cpp run8Point):
// points rounded to float32 (findFundamentalMat converts its inputs to CV_32F), isotropic
// normalisation (centroid, mean distance sqrt 2), the 9x9 normal matrix A^T A, its eigenvector of
// the smallest eigenvalue, rank-2 projection through the SVD of the 3x3, de-normalisation,
// F[2][2] = 1.  Returns false where OpenCV returns no matrix (degenerate point sets).
__host__ __device__ inline bool fundamental_8point(const double* u1, const double* u2, int stride_u,
                                                   int J, double* F) {
  double c1[2] = {0, 0}, c2[2] = {0, 0};
  for (int i = 0; i < J; ++i) {
    c1[0] += (double)(float)u1[(int64_t)i * stride_u];
    c1[1] += (double)(float)u1[(int64_t)i * stride_u + 1];
    c2[0] += (double)(float)u2[(int64_t)i * stride_u];
    c2[1] += (double)(float)u2[(int64_t)i * stride_u + 1];
  }
  const double t = 1.0 / J;
  c1[0] *= t; c1[1] *= t; c2[0] *= t; c2[1] *= t;
  double s1 = 0, s2 = 0;
  for (int i = 0; i < J; ++i) {
    const double x1 = (double)(float)u1[(int64_t)i * stride_u] - c1[0];
    const double y1 = (double)(float)u1[(int64_t)i * stride_u + 1] - c1[1];
    const double x2 = (double)(float)u2[(int64_t)i * stride_u] - c2[0];
    const double y2 = (double)(float)u2[(int64_t)i * stride_u + 1] - c2[1];
    s1 += sqrt(x1 * x1 + y1 * y1);
    s2 += sqrt(x2 * x2 + y2 * y2);
  }
  s1 *= t; s2 *= t;
  if (s1 < 1.1920929e-07 || s2 < 1.1920929e-07) return false;     // FLT_EPSILON
  s1 = sqrt(2.0) / s1;
  s2 = sqrt(2.0) / s2;
  double G[9][9], V[9][9];
  for (int a = 0; a < 9; ++a)
    for (int b = 0; b < 9; ++b) G[a][b] = 0.0;
  for (int i = 0; i < J; ++i) {
    const double x1 = ((double)(float)u1[(int64_t)i * stride_u] - c1[0]) * s1;
    const double y1 = ((double)(float)u1[(int64_t)i * stride_u + 1] - c1[1]) * s1;
    const double x2 = ((double)(float)u2[(int64_t)i * stride_u] - c2[0]) * s2;
    const double y2 = ((double)(float)u2[(int64_t)i * stride_u + 1] - c2[1]) * s2;
    const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    for (int a = 0; a < 9; ++a)
      for (int b = 0; b < 9; ++b) G[a][b] += r[a] * r[b];
  }
  // eigenvectors of the symmetric PSD normal matrix = its right singular vectors
  jacobi_onesided<9, 9>(G, V);
  int best = 0, nz = 0;
  double bn = DBL_MAX;
  for (int j = 0; j < 9; ++j) {
    double n2 = 0;
    for (int i = 0; i < 9; ++i) n2 += G[i][j] * G[i][j];
    if (sqrt(n2) >= DBL_EPSILON) ++nz;
    if (n2 < bn) { bn = n2; best = j; }
  }
  if (nz < 8) return false;                         // OpenCV: fewer than 8 non-zero eigenvalues
  double F0[3][3], W[3][3];
  for (int i = 0; i < 9; ++i) {
    double v = V[i][0];
    for (int j = 1; j < 9; ++j)
      if (best == j) v = V[i][j];
    F0[i / 3][i % 3] = v;
  }
  // rank 2: drop the smallest singular value
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = F0[i][j];
  jacobi_onesided<3, 3>(A, W);                      // A = U diag(sigma) (columns), F0 = A W^T
  int sm = 0;
  double sn = DBL_MAX;
  for (int j = 0; j < 3; ++j) {
    const double n2 = A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j];
    if (n2 < sn) { sn = n2; sm = j; }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int k = 0; k < 3; ++k)
        if (k != sm) v += A[i][k] * W[j][k];
      F0[i][j] = v;
    }
  // F = T2^T F0 T1, T = [s 0 -s*cx; 0 s -s*cy; 0 0 1]
  const double T1[3][3] = {{s1, 0, -s1 * c1[0]}, {0, s1, -s1 * c1[1]}, {0, 0, 1}};
  const double T2[3][3] = {{s2, 0, -s2 * c2[0]}, {0, s2, -s2 * c2[1]}, {0, 0, 1}};
  double M[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int k = 0; k < 3; ++k) v += F0[i][k] * T1[k][j];
      M[i][j] = v;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int k = 0; k < 3; ++k) v += T2[k][i] * M[k][j];
      F[i * 3 + j] = v;
    }
  if (fabs(F[8]) > 1.1920929e-07) {
    const double inv = 1.0 / F[8];
    for (int k = 0; k < 9; ++k) F[k] *= inv;
  }
  return true;
}

// Fundamental matrix per pair for the polynomial method (triangulation.py:198-217): from the
// projection matrices; when the optimal correction with it is NaN for EVERY joint of the pair
// (identical / degenerate cameras) -- or always, for mode 4 -- the 8-point estimate from the
// matches themselves, as the reference falls back to.  One thread per pair.
__global__ void pair_fundamental_kernel(const double* __restrict__ u1, const double* __restrict__ u2,
                                        int stride_u, const double* __restrict__ P1,
                                        const double* __restrict__ P2, int NP, int J, int method,
                                        double* __restrict__ Fout) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= NP) return;
  double F[9];
  fundamental_from_P(P1 + pair * 12, P2 + pair * 12, F);
  bool fallback = method == 4;
  if (!fallback) {
    bool all1 = true, all2 = true;
    for (int j = 0; j < J && (all1 || all2); ++j) {
      const int64_t o = ((int64_t)pair * J + j) * stride_u;
      double a1[2] = {u1[o], u1[o + 1]}, a2[2] = {u2[o], u2[o + 1]};
      correct_match(F, a1, a2);
      if (!(isnan(a1[0]) && isnan(a1[1]))) all1 = false;
      if (!(isnan(a2[0]) && isnan(a2[1]))) all2 = false;
    }
    fallback = all1 || all2;
  }
  if (fallback) {
    double F8[9];
    if (fundamental_8point(u1 + (int64_t)pair * J * stride_u, u2 + (int64_t)pair * J * stride_u,
                           stride_u, J, F8))
      for (int k = 0; k < 9; ++k) F[k] = F8[k];
  }
  for (int k = 0; k < 9; ++k) Fout[pair * 9 + k] = F[k];
}

__global__ void triangulate_kernel(const double* __restrict__ u1, const double* __restrict__ u2,
                                   int stride_u, const double* __restrict__ P1,
                                   const double* __restrict__ P2, int NP, int J, int method,
                                   double tol, const double* __restrict__ Fpair,
                                   double* __restrict__ X, int32_t* __restrict__ status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NP * J) return;
  const int pair = idx / J;
  double p1[12], p2[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) { p1[k] = P1[pair * 12 + k]; p2[k] = P2[pair * 12 + k]; }
  double a1[2] = {u1[(int64_t)idx * stride_u], u1[(int64_t)idx * stride_u + 1]};
  double a2[2] = {u2[(int64_t)idx * stride_u], u2[(int64_t)idx * stride_u + 1]};
  double x[3];
  int st;
  if (method >= 3) {   // triangulation.py:184-220: optimal correction, then the homogeneous DLT
    double F[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) F[k] = Fpair[pair * 9 + k];
    correct_match(F, a1, a2);
  }
  if (method == 0 || method >= 3) {
    // triangulation.py:22 cv2.triangulatePoints: A rows x*P[2]-P[0], y*P[2]-P[1]
    double A[4][4], V[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[0][k] = a1[0] * p1[8 + k] - p1[0 + k];
      A[1][k] = a1[1] * p1[8 + k] - p1[4 + k];
      A[2][k] = a2[0] * p2[8 + k] - p2[0 + k];
      A[3][k] = a2[1] * p2[8 + k] - p2[4 + k];
    }
    jacobi_onesided<4, 4>(A, V);
    int best = 0;
    double bn = DBL_MAX;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double n2 = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) n2 += A[i][j] * A[i][j];
      if (n2 < bn) { bn = n2; best = j; }
    }
    double h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = V[i][0];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (best == j) h[i] = V[i][j];
    }
    x[0] = h[0] / h[3]; x[1] = h[1] / h[3]; x[2] = h[2] / h[3];  // :24
    const double mx = fmax(fabs(x[0]), fmax(fabs(x[1]), fabs(x[2])));
    st = (mx <= 1.e16) ? 1 : 0;  // :25 (NaN compares false)
  } else {
    double A[4][3], b[4];
    build_Ab(a1, p1, a2, p2, A, b);
    if (method == 1) {
      solve_ls_4x3(A, b, x);
      st = 1;
    } else {
      double d1 = 1.0, d2 = 1.0, d1n = 1.0, d2n = 1.0;
      for (int it = 0; it < 10; ++it) {   // :152
        solve_ls_4x3(A, b, x);
        d1n = ((p1[8] * x[0] + p1[9] * x[1]) + p1[10] * x[2]) + p1[11];   // :158
        d2n = ((p2[8] * x[0] + p2[9] * x[1]) + p2[10] * x[2]) + p2[11];
        if (fabs(d1n - d1) <= tol && fabs(d2n - d2) <= tol) break;  // :161-163
        const double r1 = 1.0 / d1n, r2 = 1.0 / d2n;
        // :165-169 CUMULATIVE re-weighting of the already weighted rows
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          A[0][k] *= r1; A[1][k] *= r1; A[2][k] *= r2; A[3][k] *= r2;
        }
        b[0] *= r1; b[1] *= r1; b[2] *= r2; b[3] *= r2;
        d1 = d1n; d2 = d2n;
      }
      st = (d1n > 0 && d2n > 0) ? 1 : 0;  // :175-176 (i < 10 always holds)
      if (d1n <= 0) st -= 1;
      if (d2n <= 0) st -= 2;
    }
  }
  X[(int64_t)idx * 3 + 0] = x[0];
  X[(int64_t)idx * 3 + 1] = x[1];
  X[(int64_t)idx * 3 + 2] = x[2];
  if (status) status[idx] = st;
}

// --------------------------------------------------------------- N-view homogeneous DLT
// the V-view generalisation of triangulation.py:8-27 (not in the reference,
// which only pairs two views): A (2V x 4) rows u*P[2]-P[0], v*P[2]-P[1] per view, X = right
// singular vector of the smallest singular value, de-homogenised.  V <= 4 (one tuple).
template <int V>
__host__ __device__ inline int dlt_nview(const double* u /*[V][2]*/, const double* P /*[V][12]*/,
                                         double* x /*[3]*/) {
  double A[2 * V][4], Vm[4][4];
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < 4; ++k) {
      A[2 * v + 0][k] = u[v * 2 + 0] * P[v * 12 + 8 + k] - P[v * 12 + 0 + k];
      A[2 * v + 1][k] = u[v * 2 + 1] * P[v * 12 + 8 + k] - P[v * 12 + 4 + k];
    }
  jacobi_onesided<2 * V, 4>(A, Vm);
  int best = 0;
  double bn = DBL_MAX;
  for (int j = 0; j < 4; ++j) {
    double n2 = 0;
    for (int i = 0; i < 2 * V; ++i) n2 += A[i][j] * A[i][j];
    if (n2 < bn) { bn = n2; best = j; }
  }
  double h[4];
  for (int i = 0; i < 4; ++i) h[i] = Vm[i][best];
  x[0] = h[0] / h[3]; x[1] = h[1] / h[3]; x[2] = h[2] / h[3];
  const double mx = fmax(fabs(x[0]), fmax(fabs(x[1]), fabs(x[2])));
  return (mx <= 1.e16) ? 1 : 0;
}

// u [NT][V][J][stride_u], P [NT][V][12] -> X [NT][J][3], status [NT][J]; one thread per (tuple, joint)
__global__ void triangulate_nview_kernel(const double* __restrict__ u, int stride_u,
                                         const double* __restrict__ P, int NT, int V, int J,
                                         double* __restrict__ X, int32_t* __restrict__ status) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NT * J) return;
  const int t = idx / J, j = idx - t * J;
  double uu[8], pp[48], x[3];
  for (int v = 0; v < V; ++v) {
    const double* up = u + (((int64_t)t * V + v) * J + j) * stride_u;
    uu[v * 2 + 0] = up[0];
    uu[v * 2 + 1] = up[1];
    for (int k = 0; k < 12; ++k) pp[v * 12 + k] = P[((int64_t)t * V + v) * 12 + k];
  }
  int st;
  if (V == 2) st = dlt_nview<2>(uu, pp, x);
  else if (V == 3) st = dlt_nview<3>(uu, pp, x);
  else st = dlt_nview<4>(uu, pp, x);
  X[(int64_t)idx * 3 + 0] = x[0];
  X[(int64_t)idx * 3 + 1] = x[1];
  X[(int64_t)idx * 3 + 2] = x[2];
  if (status) status[idx] = st;
}

// img_utils.py:72-105 with its float32 roundings.  Returns the 2x3 transform
// mapping src->dst (inv=0: image->patch) or dst->src (inv=1: patch->image).
__device__ void patch_affine(const double* box, double patch_w, double patch_h, int inv,
                             double (&M)[2][3]) {
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
  const float (*from)[2] = inv ? d : s;
  const float (*to)[2] = inv ? s : d;
  // cv2.getAffineTransform: solve [x y 1] m = to, float64
  const double e1x = (double)from[1][0] - from[0][0], e1y = (double)from[1][1] - from[0][1];
  const double e2x = (double)from[2][0] - from[0][0], e2y = (double)from[2][1] - from[0][1];
  const double det = e1x * e2y - e1y * e2x;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double f1 = (double)to[1][r] - to[0][r], f2 = (double)to[2][r] - to[0][r];
    const double a = (f1 * e2y - f2 * e1y) / det;
    const double b = (e1x * f2 - e2x * f1) / det;
    M[r][0] = a;
    M[r][1] = b;
    M[r][2] = (double)to[0][r] - a * from[0][0] - b * from[0][1];
  }
}

__global__ void patch_to_image_kernel(const float* __restrict__ coords,
                                      const double* __restrict__ box, int B, int J,
                                      double patch_w, double patch_h, double rect3d_w,
                                      double* __restrict__ kps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * J) return;
  const int b = idx / J;
  // integral_loss.py:196-201
  const double px = ((double)coords[idx * 3 + 0] + 0.5) * patch_w;
  const double py = ((double)coords[idx * 3 + 1] + 0.5) * patch_h;
  const double pz = (double)coords[idx * 3 + 2] * patch_w;
  double M[2][3];
  patch_affine(box + b * 6, patch_w, patch_h, 1, M);
  // img_utils.py:108-111 np.dot(trans, [x, y, 1])
  kps[(int64_t)idx * 4 + 0] = (M[0][0] * px + M[0][1] * py) + M[0][2];
  kps[(int64_t)idx * 4 + 1] = (M[1][0] * px + M[1][1] * py) + M[1][2];
  kps[(int64_t)idx * 4 + 2] = pz / patch_w * rect3d_w;  // img_utils.py:154
  kps[(int64_t)idx * 4 + 3] = 1.0;
}

__global__ void project_labels_kernel(const double* __restrict__ X, const double* __restrict__ cam,
                                      const double* __restrict__ box, int B, int J,
                                      double patch_w, double patch_h, double rect3d_w,
                                      float* __restrict__ label, float* __restrict__ weight) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * J) return;
  const int b = idx / J;
  const double* c = cam + b * 16;   // R(9) T(3) f(2) c(2)
  const double* x = X + (int64_t)idx * 3;
  const double* x0 = X + (int64_t)b * J * 3;  // root joint 0 (prep_h36m.py:181)
  // prep_h36m.py:186 np.dot(rot, keypoints - trans)
  const double dx = x[0] - c[9], dy = x[1] - c[10], dz = x[2] - c[11];
  const double cx = (c[0] * dx + c[1] * dy) + c[2] * dz;
  const double cy = (c[3] * dx + c[4] * dy) + c[5] * dz;
  const double cz = (c[6] * dx + c[7] * dy) + c[8] * dz;
  const double rx = x0[0] - c[9], ry = x0[1] - c[10], rz = x0[2] - c[11];
  const double pelvis_z = (c[6] * rx + c[7] * ry) + c[8] * rz;
  // CamProj :170-175
  double u = cx / cz * c[12] + c[14];
  double v = cy / cz * c[13] + c[15];
  double z = cz - pelvis_z;  // :199
  double M[2][3];
  patch_affine(box + b * 6, patch_w, patch_h, 0, M);
  const double pu = (M[0][0] * u + M[0][1] * v) + M[0][2];   // img_utils.py:235
  const double pv = (M[1][0] * u + M[1][1] * v) + M[1][2];
  z = z / (rect3d_w * box[b * 6 + 4]) * patch_w;               // :236
  // integral_loss.py:171-173
  label[idx * 3 + 0] = (float)(pu / patch_w - 0.5);
  label[idx * 3 + 1] = (float)(pv / patch_h - 0.5);
  label[idx * 3 + 2] = (float)(z / patch_w);
  weight[idx * 3 + 0] = 1.f;
  weight[idx * 3 + 1] = 1.f;
  weight[idx * 3 + 2] = 1.f;
}

// --------------------------------------------------------------- evaluation (H36M protocol)
// lib/dataset/h36m.py:168-378 per sample: back-projection of image-space joints
// (lib/utils/prep_h36m.py:85-89), similarity (Procrustes) alignment with optimal scale
// (:108-168, numpy SVD there; here the one-sided Jacobi on the 3x3 covariance -- V*U^T does
// not depend on the ordering / paired signs of the singular triplets), root alignment and
// the nine protocol means.  One thread per sample, float64, J <= 32.
//   metrics[s] = { e, e_align, e_norm, e14, e14_align, e14_norm, ex, ey, ez }   (means over joints)
// (also compiled for the host: tests/harness/host_geometry.cu runs this exact code on the CPU)
__host__ __device__ void h36m_eval_sample(const double* p, const double* q, const double* cam,
                                          int J, int root, unsigned j14mask, double pck_thr,
                                          double* metrics, double* per_joint, int32_t* pck,
                                          double* poses) {
  const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  const double zr = cam[4];
  // back projection (h36m.py:228-240): X = gt (targets), Y = prediction (inputs)
  auto bp = [&](const double* a, int j, double (&o)[3]) {
    const double d = a[j * 3 + 2] + zr;
    o[0] = (a[j * 3 + 0] - cx) / fx * d;
    o[1] = (a[j * 3 + 1] - cy) / fy * d;
    o[2] = d;
  };
  double muX[3] = {0, 0, 0}, muY[3] = {0, 0, 0};
  for (int j = 0; j < J; ++j) {
    double x[3], y[3];
    bp(q, j, x);
    bp(p, j, y);
    for (int k = 0; k < 3; ++k) { muX[k] += x[k]; muY[k] += y[k]; }
  }
  for (int k = 0; k < 3; ++k) { muX[k] /= J; muY[k] /= J; }
  double ssX = 0, ssY = 0;
  for (int j = 0; j < J; ++j) {
    double x[3], y[3];
    bp(q, j, x);
    bp(p, j, y);
    for (int k = 0; k < 3; ++k) {
      const double a = x[k] - muX[k], b = y[k] - muY[k];
      ssX += a * a;
      ssY += b * b;
    }
  }
  const double normX = sqrt(ssX), normY = sqrt(ssY);
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};     // X0^T Y0 (unit Frobenius norm each)
  for (int j = 0; j < J; ++j) {
    double x[3], y[3];
    bp(q, j, x);
    bp(p, j, y);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        A[a][b] += ((x[a] - muX[a]) / normX) * ((y[b] - muY[b]) / normY);
  }
  double V[3][3];
  jacobi_onesided<3, 3>(A, V);                            // A = U diag(sg) (columns)
  double sg[3], U[3][3];
  for (int c = 0; c < 3; ++c) {
    sg[c] = sqrt(A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c]);
    for (int r = 0; r < 3; ++r) U[r][c] = sg[c] > 0 ? A[r][c] / sg[c] : 0.0;
  }
  int jmin = 0;
  for (int c = 1; c < 3; ++c) if (sg[c] < sg[jmin]) jmin = c;
  if (sg[jmin] == 0.0) {
    // rank-deficient covariance: complete U with the cross product of the other two columns
    const int a = (jmin + 1) % 3, b = (jmin + 2) % 3;
    U[0][jmin] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
    U[1][jmin] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
    U[2][jmin] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
  }
  double T[3][3];
  auto make_T = [&]() {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        T[r][c] = V[r][0] * U[c][0] + V[r][1] * U[c][1] + V[r][2] * U[c][2];   // V U^T
  };
  make_T();
  const double det = T[0][0] * (T[1][1] * T[2][2] - T[1][2] * T[2][1]) -
                     T[0][1] * (T[1][0] * T[2][2] - T[1][2] * T[2][0]) +
                     T[0][2] * (T[1][0] * T[2][1] - T[1][1] * T[2][0]);
  const double sgn = det > 0 ? 1.0 : (det < 0 ? -1.0 : 0.0);   // np.sign
  for (int r = 0; r < 3; ++r) V[r][jmin] *= sgn;               // V[:,-1] *= sign(detT)
  sg[jmin] *= sgn;
  make_T();
  const double trace = sg[0] + sg[1] + sg[2];
  const double bsc = trace * normX / normY;                    // optimal scale
  double cvec[3];
  for (int c = 0; c < 3; ++c)
    cvec[c] = muX[c] - bsc * (muY[0] * T[0][c] + muY[1] * T[1][c] + muY[2] * T[2][c]);
  // root joint of each variant (h36m.py:247-251)
  double xr[3], yr[3], yar[3], ynr[3];
  bp(q, root, xr);
  bp(p, root, yr);
  for (int c = 0; c < 3; ++c) {
    yar[c] = bsc * (yr[0] * T[0][c] + yr[1] * T[1][c] + yr[2] * T[2][c]) + cvec[c];
    ynr[c] = bsc * yr[c];
  }
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int n14 = 0;
  for (int j = 0; j < J; ++j) {
    double x[3], y[3], d[3], da[3], dn[3];
    bp(q, j, x);
    bp(p, j, y);
    for (int c = 0; c < 3; ++c) {
      const double ya = bsc * (y[0] * T[0][c] + y[1] * T[1][c] + y[2] * T[2][c]) + cvec[c];
      const double yn = bsc * y[c];
      const double g0 = x[c] - xr[c];
      d[c] = g0 - (y[c] - yr[c]);
      da[c] = g0 - (ya - yar[c]);
      dn[c] = g0 - (yn - ynr[c]);
      if (poses) {
        double* o = poses + j * 9;
        o[c] = y[c] - yr[c];            // pred
        o[3 + c] = ya - yar[c];         // align_pred
        o[6 + c] = g0;                  // gt
      }
    }
    const double e = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double ea = sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]);
    const double en = sqrt(dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2]);
    m[0] += e; m[1] += ea; m[2] += en;
    if ((j14mask >> j) & 1u) { m[3] += e; m[4] += ea; m[5] += en; ++n14; }
    m[6] += fabs(d[0]); m[7] += fabs(d[1]); m[8] += fabs(d[2]);
    if (per_joint) per_joint[j] = e;
    if (pck) pck[j] = e >= pck_thr ? 0 : 1;
  }
  for (int k = 0; k < 9; ++k) {
    const int div = (k >= 3 && k < 6) ? (n14 > 0 ? n14 : 1) : J;
    metrics[k] = m[k] / div;
  }
}

__global__ void h36m_eval_kernel(const double* __restrict__ pred, const double* __restrict__ gt,
                                 const double* __restrict__ cam, int S, int J, int root,
                                 unsigned j14mask, double pck_thr, double* __restrict__ metrics,
                                 double* __restrict__ per_joint, int32_t* __restrict__ pck,
                                 double* __restrict__ poses) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  h36m_eval_sample(pred + (int64_t)s * J * 3, gt + (int64_t)s * J * 3, cam + (int64_t)s * 5, J, root,
                   j14mask, pck_thr, metrics + (int64_t)s * 9,
                   per_joint ? per_joint + (int64_t)s * J : nullptr,
                   pck ? pck + (int64_t)s * J : nullptr,
                   poses ? poses + (int64_t)s * J * 9 : nullptr);
}

// --------------------------------------------------------------- argmax
// inference.py:24-39: one warp per (n,j) map; (value, index) reduction with
// smallest-index tie-break == numpy argmax first-occurrence.  NaN: numpy
// treats the first NaN as the maximum; reproduced by ordering NaN above all.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  const bool vn = (v != v), bn = (bv != bv);
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

__global__ void argmax2d_kernel(const float* __restrict__ hm, int NJ, int HW, int W,
                                int32_t* __restrict__ idx_out, float* __restrict__ maxval,
                                float* __restrict__ preds) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= NJ) return;
  const float* p = hm + (int64_t)warp * HW;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < HW; i += 32) {
    const float v = p[i];
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    if (bi == 0x7fffffff) bi = 0;  // all -inf: numpy returns index 0
    if (idx_out) idx_out[warp] = bi;
    if (maxval) maxval[warp] = bv;
    if (preds) {
      const float mask = (bv > 0.f) ? 1.f : 0.f;   // :35-39
      preds[warp * 2 + 0] = (float)(bi % W) * mask;
      preds[warp * 2 + 1] = floorf((float)bi / (float)W) * mask;
    }
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int epb_triangulate(const double* u1, const double* u2, int stride_u, const double* P1,
                               const double* P2, int NP, int J, int method, double tol, double* X,
                               int32_t* status, epb_stream_t stream) {
  EPB_CHECK_ARG(u1 && u2 && P1 && P2 && X);
  EPB_CHECK_ARG(NP >= 0 && J >= 0 && stride_u >= 2);
  EPB_CHECK_ARG(method >= 0 && method <= 4);
  if (NP * J == 0) return EPB_OK;
  const int n = NP * J;
  cudaStream_t st = as_stream(stream);
  double* Fpair = nullptr;
  if (method >= 3) {
    // one fundamental matrix per pair (from the projection matrices, or the 8-point fallback)
    int rc = epb_workspace(EPB_WS_FPAIR, (size_t)NP * 9 * sizeof(double), st, (void**)&Fpair);
    if (rc) return rc;
    pair_fundamental_kernel<<<(NP + 63) / 64, 64, 0, st>>>(u1, u2, stride_u, P1, P2, NP, J, method,
                                                           Fpair);
    EPB_LAUNCH_CHECK();
  }
  triangulate_kernel<<<(n + 63) / 64, 64, 0, st>>>(u1, u2, stride_u, P1, P2, NP, J, method, tol,
                                                   Fpair, X, status);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_patch_to_image(const float* coords, const double* box, int B, int J,
                                  double patch_w, double patch_h, double rect3d_w, double* kps,
                                  epb_stream_t stream) {
  EPB_CHECK_ARG(coords && box && kps);
  EPB_CHECK_ARG(B >= 0 && J >= 0);
  if (B * J == 0) return EPB_OK;
  const int n = B * J;
  patch_to_image_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(coords, box, B, J, patch_w,
                                                                       patch_h, rect3d_w, kps);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_project_labels(const double* X, const double* cam, const double* box, int B,
                                  int J, double patch_w, double patch_h, double rect3d_w,
                                  float* label, float* weight, epb_stream_t stream) {
  EPB_CHECK_ARG(X && cam && box && label && weight);
  EPB_CHECK_ARG(B >= 0 && J >= 0);
  if (B * J == 0) return EPB_OK;
  const int n = B * J;
  project_labels_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(
      X, cam, box, B, J, patch_w, patch_h, rect3d_w, label, weight);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_argmax2d(const float* hm, int NJ, int H, int W, int32_t* idx, float* maxval,
                            float* preds, epb_stream_t stream) {
  EPB_CHECK_ARG(hm && NJ >= 0 && H > 0 && W > 0);
  if (NJ == 0) return EPB_OK;
  const int threads = 256;
  const int blocks = (NJ * 32 + threads - 1) / threads;
  argmax2d_kernel<<<blocks, threads, 0, as_stream(stream)>>>(hm, NJ, H * W, W, idx, maxval, preds);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_h36m_eval(
    const double* pred, const double* gt, const double* cam, int S, int J, int root,
    uint32_t j14mask, double pck_thr, double* metrics, double* per_joint, int32_t* pck,
    double* poses, epb_stream_t stream) {
  EPB_CHECK_ARG(pred && gt && cam && metrics);
  EPB_CHECK_ARG(S >= 0 && J > 0 && J <= 32 && root >= 0 && root < J);
  if (S == 0) return EPB_OK;
  h36m_eval_kernel<<<(S + 63) / 64, 64, 0, as_stream(stream)>>>(pred, gt, cam, S, J, root, j14mask,
                                                               pck_thr, metrics, per_joint, pck,
                                                               poses);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}

extern "C" __attribute__((visibility("default"))) int epb_triangulate_nview(
    const double* u, int stride_u, const double* P, int NT, int V, int J, double* X, int32_t* status,
    epb_stream_t stream) {
  EPB_CHECK_ARG(u && P && X);
  EPB_CHECK_ARG(NT >= 0 && J >= 0 && stride_u >= 2 && V >= 2 && V <= 4);
  if (NT * J == 0) return EPB_OK;
  const int n = NT * J;
  triangulate_nview_kernel<<<(n + 63) / 64, 64, 0, as_stream(stream)>>>(u, stride_u, P, NT, V, J, X, status);
  EPB_LAUNCH_CHECK();
  return EPB_OK;
}
