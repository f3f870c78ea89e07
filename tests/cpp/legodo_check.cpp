// legodo_check.cpp -- CPU-only check of MavStateEst::batch::LegOdoCommon (measurement formation; no GPU work):
// prints index sets, R diagonals and z for the three supported modes and the certain / uncertain / no-position cases.
#include <cstdio>
#include "../../include/rbis_batch.hpp"
using namespace MavStateEst::batch;

static void dump(const char* tag, RBISUpdateInterface* u, int64_t N) {
  auto* m = static_cast<RBISIndexedMeasurement*>(u);
  const size_t k = m->index.size();
  printf("%s utime=%lld sensor=%d m=%zu idx=", tag, (long long)m->utime, (int)m->sensor_id, k);
  for (auto i : m->index) printf("%d,", i);
  printf(" Rdiag=");
  for (size_t a = 0; a < k; a++) printf("%.17g,", m->measurement_cov[a + k * a]);
  double off = 0;
  for (size_t a = 0; a < k; a++) for (size_t b = 0; b < k; b++) if (a != b) off += m->measurement_cov[a + k * b];
  printf(" offdiag=%g z=", off);
  for (size_t a = 0; a < k * (size_t)N; a++) printf("%.17g,", m->measurement[a]);
  printf("\n");
  delete u;
}

int main() {
  const int64_t N = 2;
  std::vector<double> pos = {1, 2, 3, 4, 5, 6}, dxyz = {0.002, 0.004, -0.001, 0.003, 0.0005, 0.0015};
  std::vector<double> dq = {0.99999, 0.99998, 0.002, -0.001, 0.001, 0.003, 0.004, 0.002};  // [4][N], w x y z rows
  LegOdoCommon lin(LegOdoCommon::MODE_LIN_RATE, 0.01, 0.1, 0.2, 0.5, 0.9);
  dump("lin_certain", lin.createMeasurement(pos, dxyz, dq, N, 1002000, 1000000, 1, 0.0f), N);
  dump("lin_uncertain", lin.createMeasurement(pos, dxyz, dq, N, 1002000, 1000000, 1, 0.7f), N);
  LegOdoCommon pv(LegOdoCommon::MODE_POSITION_AND_LIN_RATE, 0.01, 0.1, 0.2, 0.5, 0.9);
  dump("posvel", pv.createMeasurement(pos, dxyz, dq, N, 1002000, 1000000, 1, 0.2f), N);
  dump("posvel_nopos", pv.createMeasurement(pos, dxyz, dq, N, 1002000, 1000000, 0, 0.2f), N);
  LegOdoCommon lr(LegOdoCommon::MODE_LIN_AND_ROT_RATE, 0.01, 0.1, 0.2, 0.5, 0.9);
  dump("linrot", lr.createMeasurement(pos, dxyz, dq, N, 1004000, 1000000, 1, 0.9f), N);
  return 0;
}
