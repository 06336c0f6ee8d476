// TEST INFRASTRUCTURE: runs the __host__ __device__ per-sample bodies of csrc/geometry.cu on the
// CPU (same source, same --fmad=false arithmetic) so that the CUDA algorithm is checked against
// the oracle in the CPU test suite too.  Binary protocol on stdin/stdout (little-endian doubles):
//   "eval" S J root mask  then pred[S*J*3] gt[S*J*3] cam[S*5]  ->  metrics[S*9] per_joint[S*J] poses[S*J*9]
//   "correct" N  then P1[12] P2[12] u1[N*2] u2[N*2]  ->  F[9] u1'[N*2] u2'[N*2]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
void epb_set_error(const char*, ...) {}
#include "../../epipolarpose_b200/csrc/geometry.cu"

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
