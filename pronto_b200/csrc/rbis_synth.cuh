// rbis_synth.cuh -- on-device synthesis of the per-filter sensor streams of a Monte-Carlo ensemble.
//
// A noise sweep feeds every filter the SAME noise-free log plus its own noise realisation: 60.6 B per filter step of input
// for 12 KB of actual information per 200-step chunk.  Copying those streams from the host pins the end-to-end rate to PCIe
// (53 GB/s = 0.88 G filter-steps/s); here the host sends the noise-free rows and a seed, and the rows every filter reads are
// produced in HBM by these kernels, in the layout rbis_batch_run_fused consumes ([rows][6][N], [rows][m][N], [rows][4][N]).
//
// Generator (pronto_b200/synth.py:normal is the numpy statement of mode 0):
//   key = seed ^ filter * K1 ^ step * K2 ^ channel * K3;  a = splitmix64(key), b = splitmix64(a)
//   mode 0 (exact):  u1 = ((a >> 11) + 1) / (2^53 + 1), u2 = (b >> 11) / 2^53, n = sqrt(-2 ln u1) cos(2 pi u2) in double
//                    -- the integer part is bit-exact against numpy, the three libm calls agree to their last ulp or two;
//   mode 1 (fast):   the same counters and hash, ONE hash per pair of channels (radius from its top 24 bits, angle from the next
//                    24; even channel r cos, odd channel r cos(. - pi/2)), ln / sqrt / cos in single precision with the SFU
//                    approximations (n carries ~1e-6 relative error, irrelevant for a noise sample).
// `filter` is the GLOBAL filter index (first_filter + n), so a sharded ensemble draws the same noise for any GPU count.
#pragma once
#include "rbis_kernels.cuh"

namespace rbisk {

// The generator itself (splitmix64, syn_normal, syn_quat) lives in rbis_kernels.cuh: the SYN instantiations of the fused kernels
// draw the same rows inside the kernel, without materialising them (rbis_batch_run_fused_synth, mode 1).

struct SynStreamDev {
  int m, has_orient, channel, channel_rot;
  const double* mean;       // device [rows][m]
  const double* mean_quat;  // device [rows][4]
  const long long* step;    // device [rows]
  double sigma[MAX_MEAS];
  double sigma_rot[3];
  double* z;                // device out [rows][m][N]
  double* quat;             // device out [rows][4][N]
  long long rows;
};

// imu [rows][6][N]: row r, channel c of filter n = mean[r][c] + sigma_c(n) * normal(first_filter + n, step[r], c)
// sigma < 0: per filter sqrt(q / dt) from the handle's process noise, so that the sample noise matches Qd = q dt (rbis.cpp:116)
template <int MODE>
__global__ void __launch_bounds__(256) synth_imu_kernel(double* __restrict__ out, const double* __restrict__ mean, const long long* __restrict__ step,
                                                        long long rows, long long N, unsigned long long seed, long long first_filter,
                                                        double sigma_gyro, double sigma_accel, const double* __restrict__ q_gyro,
                                                        const double* __restrict__ q_accel, double dt) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = blockIdx.y;
  if (n >= N || r >= rows) return;
  const double sg = sigma_gyro >= 0 ? sigma_gyro : sqrt(q_gyro[n] / dt), sa = sigma_accel >= 0 ? sigma_accel : sqrt(q_accel[n] / dt);
  const unsigned long long f = (unsigned long long)(first_filter + n), s = (unsigned long long)step[r];
  double nn[6];
  syn_normals_k<MODE, 6>(seed ^ (f * SYN_K1) ^ (s * SYN_K2), 0u, nn);
#pragma unroll
  for (int c = 0; c < 6; c++) out[(r * 6 + c) * N + n] = fma(c < 3 ? sg : sa, nn[c], mean[r * 6 + c]);
}

// z [rows][m][N] = mean + sigma * normal(channel + a); quat [rows][4][N] = mean_quat (x) Exp(sigma_rot * normal(channel_rot + 0..2))
template <int MODE>
__global__ void __launch_bounds__(256) synth_stream_kernel(SynStreamDev st, long long N, unsigned long long seed, long long first_filter) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = blockIdx.y;
  if (n >= N || r >= st.rows) return;
  const unsigned long long f = (unsigned long long)(first_filter + n), s = (unsigned long long)st.step[r];
  for (int a = 0; a < st.m; a++) {
    const double nn = syn_normal<MODE>(seed, f, s, (unsigned)(st.channel + a));
    st.z[(r * st.m + a) * N + n] = fma(st.sigma[a], nn, st.mean[r * st.m + a]);
  }
  if (st.has_orient) {
    const unsigned long long kfs = seed ^ (f * SYN_K1) ^ (s * SYN_K2);
    const Q4 q = syn_quat<MODE>(st.mean_quat + r * 4, st.sigma_rot, kfs, (unsigned)st.channel_rot);
    st.quat[(r * 4 + 0) * N + n] = q.w;
    st.quat[(r * 4 + 1) * N + n] = q.x;
    st.quat[(r * 4 + 2) * N + n] = q.y;
    st.quat[(r * 4 + 3) * N + n] = q.z;
  }
}

}  // namespace rbisk
