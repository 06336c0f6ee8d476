// TEST INFRASTRUCTURE: runs the __host__ __device__ per-sample bodies of csrc/geometry.cu on the
// CPU (same source, same --fmad=false arithmetic) so that the CUDA algorithm is checked against
// the oracle in the CPU test suite too.  Binary protocol on stdin/stdout (little-endian doubles):
//   "eval" S J root mask  then pred[S*J*3] gt[S*J*3] cam[S*5]  ->  metrics[S*9] per_joint[S*J] poses[S*J*9]
//   "correct" N  then P1[12] P2[12] u1[N*2] u2[N*2]  ->  F[9] u1'[N*2] u2'[N*2]
//   "f8" N  then u1[N*2] u2[N*2]  ->  ok[1] F[9]   (cv2.findFundamentalMat FM_8POINT)
//   "nview" V J  then u[V*J*2] P[V*12]  ->  X[J*3]
//   "patch" H W pw ph flip J  then box[6] color[3] mean_std[6] depth_den[1] joints[J*3] img[H*W*3 as doubles]
//           ->  trans[6] patch[3*ph*pw] (float32 values widened) label[J*3]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
void epb_set_error(const char*, ...) {}
int epb_workspace(int, size_t, struct CUstream_st*, void**) { return -1; }   // entry points are not run here
#include "../../epipolarpose_b200/csrc/geometry.cu"
#include "../../epipolarpose_b200/csrc/input.cu"

static void rd(void* p, size_t n) { if (fread(p, 1, n, stdin) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  if (!strcmp(argv[1], "eval")) {
    const int S = atoi(argv[2]), J = atoi(argv[3]), root = atoi(argv[4]);
    const unsigned mask = (unsigned)strtoul(argv[5], nullptr, 10);
    std::vector<double> pred(S * J * 3), gt(S * J * 3), cam(S * 5), met(S * 9), pj(S * J), poses(S * J * 9);
    std::vector<int32_t> pck(S * J);
    rd(pred.data(), pred.size() * 8); rd(gt.data(), gt.size() * 8); rd(cam.data(), cam.size() * 8);
    for (int s = 0; s < S; ++s)
      h36m_eval_sample(&pred[s * J * 3], &gt[s * J * 3], &cam[s * 5], J, root, mask, 150.0, &met[s * 9],
                       &pj[s * J], &pck[s * J], &poses[s * J * 9]);
    fwrite(met.data(), 8, met.size(), stdout);
    fwrite(pj.data(), 8, pj.size(), stdout);
    fwrite(poses.data(), 8, poses.size(), stdout);
    return 0;
  }
  if (!strcmp(argv[1], "nview")) {
    const int V = atoi(argv[2]), J = atoi(argv[3]);
    std::vector<double> u(V * J * 2), P(V * 12), X(J * 3);
    rd(u.data(), u.size() * 8); rd(P.data(), P.size() * 8);
    for (int j = 0; j < J; ++j) {
      double uu[8];
      for (int v = 0; v < V; ++v) { uu[v * 2] = u[(v * J + j) * 2]; uu[v * 2 + 1] = u[(v * J + j) * 2 + 1]; }
      if (V == 2) dlt_nview<2>(uu, P.data(), &X[j * 3]);
      else if (V == 3) dlt_nview<3>(uu, P.data(), &X[j * 3]);
      else dlt_nview<4>(uu, P.data(), &X[j * 3]);
    }
    fwrite(X.data(), 8, X.size(), stdout);
    return 0;
  }
  if (!strcmp(argv[1], "patch")) {
    const int H = atoi(argv[2]), W = atoi(argv[3]), pw = atoi(argv[4]), ph = atoi(argv[5]);
    const int flip = atoi(argv[6]), J = atoi(argv[7]);
    std::vector<double> box(6), color(3), ms(6), den(1), joints(J * 3), imgd((size_t)H * W * 3);
    rd(box.data(), 48); rd(color.data(), 24); rd(ms.data(), 48); rd(den.data(), 8);
    rd(joints.data(), joints.size() * 8); rd(imgd.data(), imgd.size() * 8);
    std::vector<uint8_t> img(imgd.size());
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)imgd[i];
    double bx[6];
    for (int k = 0; k < 6; ++k) bx[k] = box[k];
    if (flip) bx[0] = (double)W - bx[0] - 1.0;
    double M[6], iM[6];
    if (!patch_affine_fwd(bx, (double)pw, (double)ph, M)) return 3;
    invert_affine(M, iM);
    std::vector<double> patch((size_t)3 * ph * pw), label(J * 3);
    for (int y = 0; y < ph; ++y)
      for (int x = 0; x < pw; ++x) {
        int bgr[3];
        warp_pixel_u8(img.data(), H, W, (int64_t)W * 3, flip, iM, x, y, bgr);
        for (int c = 0; c < 3; ++c)
          patch[((size_t)c * ph + y) * pw + x] =
              (double)finish_pixel(bgr[2 - c], (float)color[c], true, ms[c], ms[3 + c]);
      }
    for (int j = 0; j < J; ++j) patch_joint(&joints[j * 3], M, (double)pw, (double)ph, den[0], &label[j * 3]);
    fwrite(M, 8, 6, stdout);
    fwrite(patch.data(), 8, patch.size(), stdout);
    fwrite(label.data(), 8, label.size(), stdout);
    return 0;
  }
  if (!strcmp(argv[1], "f8")) {
    const int N = atoi(argv[2]);
    std::vector<double> u1(N * 2), u2(N * 2), out(10, 0.0);
    rd(u1.data(), N * 16); rd(u2.data(), N * 16);
    out[0] = fundamental_8point(u1.data(), u2.data(), 2, N, &out[1]) ? 1.0 : 0.0;
    fwrite(out.data(), 8, 10, stdout);
    return 0;
  }
  if (!strcmp(argv[1], "correct")) {
    const int N = atoi(argv[2]);
    std::vector<double> P1(12), P2(12), u1(N * 2), u2(N * 2), F(9);
    rd(P1.data(), 96); rd(P2.data(), 96); rd(u1.data(), N * 16); rd(u2.data(), N * 16);
    fundamental_from_P(P1.data(), P2.data(), F.data());
    for (int i = 0; i < N; ++i) correct_match(F.data(), &u1[i * 2], &u2[i * 2]);
    fwrite(F.data(), 8, 9, stdout);
    fwrite(u1.data(), 8, u1.size(), stdout);
    fwrite(u2.data(), 8, u2.size(), stdout);
    return 0;
  }
  return 1;
}
