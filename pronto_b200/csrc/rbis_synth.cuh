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
//   mode 1 (fast):   the same counters and hash; u1, u2 from the top 24 bits, ln / sqrt / cos in single precision with the
//                    SFU approximations (n carries ~1e-6 relative error, irrelevant for a noise sample) -- 4x cheaper.
// `filter` is the GLOBAL filter index (first_filter + n), so a sharded ensemble draws the same noise for any GPU count.
#pragma once
#include "rbis_kernels.cuh"

namespace rbisk {

constexpr unsigned long long SYN_K1 = 0x9E3779B97F4A7C15ull, SYN_K2 = 0xC2B2AE3D27D4EB4Full, SYN_K3 = 0x165667B19E3779F9ull;

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <int MODE>
__device__ __forceinline__ double syn_normal(unsigned long long seed, unsigned long long filter, unsigned long long step, unsigned channel) {
  const unsigned long long key = seed ^ (filter * SYN_K1) ^ (step * SYN_K2) ^ ((unsigned long long)channel * SYN_K3);
  const unsigned long long a = splitmix64(key), b = splitmix64(a);
  if constexpr (MODE == 0) {
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2);
  } else {
    const float u1 = ((float)(unsigned)(a >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float)(unsigned)(b >> 40) * (1.0f / 16777216.0f);
    return (double)(sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853071795865f * u2));
  }
}

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
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const double nn = syn_normal<MODE>(seed, f, s, (unsigned)c);
    out[(r * 6 + c) * N + n] = fma(c < 3 ? sg : sa, nn, mean[r * 6 + c]);
  }
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
    const V3 chi{st.sigma_rot[0] * syn_normal<MODE>(seed, f, s, (unsigned)st.channel_rot), st.sigma_rot[1] * syn_normal<MODE>(seed, f, s, (unsigned)(st.channel_rot + 1)),
                 st.sigma_rot[2] * syn_normal<MODE>(seed, f, s, (unsigned)(st.channel_rot + 2))};
    const double nrm = sqrt(chi.x * chi.x + chi.y * chi.y + chi.z * chi.z);
    double sn = 0.0, cs = 1.0;
    if (nrm > 0) { sincos(0.5 * nrm, &sn, &cs); sn /= nrm; }
    const Q4 dq{cs, sn * chi.x, sn * chi.y, sn * chi.z};
    const double* mq = st.mean_quat + r * 4;
    const Q4 t{mq[0], mq[1], mq[2], mq[3]};
    // plain products and sums in the order of pronto_b200/synth.py (no contraction: the library is built with -fmad=false)
    st.quat[(r * 4 + 0) * N + n] = t.w * dq.w - t.x * dq.x - t.y * dq.y - t.z * dq.z;
    st.quat[(r * 4 + 1) * N + n] = t.w * dq.x + t.x * dq.w + t.y * dq.z - t.z * dq.y;
    st.quat[(r * 4 + 2) * N + n] = t.w * dq.y + t.y * dq.w + t.z * dq.x - t.x * dq.z;
    st.quat[(r * 4 + 3) * N + n] = t.w * dq.z + t.z * dq.w + t.x * dq.y - t.y * dq.x;
  }
}

}  // namespace rbisk
