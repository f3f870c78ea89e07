// shim_replay.cpp -- test driver for include/rbis_batch.hpp: replays an ARRIVAL-ordered update list
// (written by tests/test_cpp_shim.py) through MavStateEst::batch::MavStateEstimator exactly the way a
// reference handler would feed MavStateEst::MavStateEstimator::addUpdate (MSE/lcm_front_end.hpp:139-181),
// then dumps the head state of every filter.  The reference has no tests for this path; the expected
// values come from the CPU oracle in the Python test.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/rbis_batch.hpp"

using namespace MavStateEst::batch;

static void rd(FILE* f, void* p, size_t n) {
  if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(3); }
}
static std::vector<double> rdv(FILE* f, size_t n) {
  std::vector<double> v(n);
  rd(f, v.data(), n * sizeof(double));
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: shim_replay <in> <out>\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int64_t hdr[6];
  rd(f, hdr, sizeof(hdr));
  const int64_t N = hdr[0], utime0 = hdr[1], span = hdr[2], slots = hdr[3], period = hdr[4], n_events = hdr[5];
  try {
    auto vec = rdv(f, 21 * N), quat = rdv(f, 4 * N), cov = rdv(f, 441 * N);
    MavStateEstimator est(N, new RBISResetUpdate(vec, quat, cov, RBISUpdateInterface::reset, utime0), span, (int32_t)slots, period);
    int64_t dropped = 0;
    for (int64_t e = 0; e < n_events; e++) {
      int64_t kind, utime;
      rd(f, &kind, 8); rd(f, &utime, 8);
      RBISUpdateInterface* u = nullptr;
      if (kind == 0) {
        double dq[5];
        rd(f, dq, sizeof(dq));
        auto gyro = rdv(f, 3 * N), acc = rdv(f, 3 * N);
        u = new RBISIMUProcessStep(gyro, acc, dq[1], dq[2], dq[3], dq[4], dq[0], utime);
      } else {
        int64_t m;
        rd(f, &m, 8);
        std::vector<int64_t> idx64(m);
        rd(f, idx64.data(), 8 * m);
        std::vector<int32_t> idx(idx64.begin(), idx64.end());
        auto R = rdv(f, m * m), z = rdv(f, m * N);
        if (kind == 1) u = new RBISIndexedMeasurement(idx, z, R, RBISUpdateInterface::legodo, utime);
        else {
          auto q = rdv(f, 4 * N);
          u = new RBISIndexedPlusOrientationMeasurement(idx, z, R, q, RBISUpdateInterface::pose_meas, utime);
        }
      }
      if (!est.addUpdate(u, true)) dropped++;
    }
    fclose(f);
    std::vector<double> ovec(21 * N), oquat(4 * N), ocov(441 * N), oll(N);
    est.filters().getState(ovec.data(), oquat.data(), ocov.data(), oll.data(), nullptr);
    RBIS s0; RBIM c0;
    est.getHeadState(0, s0, c0);
    for (int i = 0; i < 21; i++)
      if (s0.vec[i] != ovec[i * N]) { fprintf(stderr, "getHeadState disagrees with getState\n"); return 4; }
    FILE* o = fopen(argv[2], "wb");
    fwrite(ovec.data(), 8, ovec.size(), o); fwrite(oquat.data(), 8, oquat.size(), o);
    fwrite(ocov.data(), 8, ocov.size(), o); fwrite(oll.data(), 8, oll.size(), o);
    int64_t tail[2] = {est.launches(), dropped};
    fwrite(tail, 8, 2, o);
    fclose(o);
    printf("shim_replay: N=%lld events=%lld launches=%lld dropped=%lld\n", (long long)N, (long long)n_events, (long long)tail[0], (long long)dropped);
  } catch (const Error& e) {
    fprintf(stderr, "rbis error %d: %s\n", e.code, e.what());
    return 5;
  }
  return 0;
}
