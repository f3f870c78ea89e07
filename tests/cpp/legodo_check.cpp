// legodo_check.cpp -- CPU-only check of MavStateEst::batch::LegOdoCommon (measurement formation; no GPU work):
// prints index sets, R diagonals and z for the three supported modes and the certain / uncertain / no-position cases.
#include <cstdio>
#include <cstdlib>
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

// With arguments: one case, one filter --
//   legodo_check MODE r_xyz r_vxyz r_vang r_vxyz_uncertain r_vang_uncertain utime prev_utime pos_status delta_status  px py pz  dx dy dz  qw qx qy qz
// (MODE = 0 lin_rate, 2 lin_rot_rate, 3 pos_and_lin_rate) -- for the comparison with the reference's compiled LegOdoCommon.
static int one_case(char** a) {
  LegOdoCommon c((LegOdoCommon::LegOdoCommonMode)atoi(a[1]), atof(a[2]), atof(a[3]), atof(a[4]), atof(a[5]), atof(a[6]));
  std::vector<double> pos = {atof(a[11]), atof(a[12]), atof(a[13])}, dxyz = {atof(a[14]), atof(a[15]), atof(a[16])};
  std::vector<double> dq = {atof(a[17]), atof(a[18]), atof(a[19]), atof(a[20])};
  dump("case", c.createMeasurement(pos, dxyz, dq, 1, atoll(a[7]), atoll(a[8]), atoi(a[9]), (float)atof(a[10])), 1);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 21) return one_case(argv);
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
