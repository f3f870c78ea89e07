// dev/kbench.cu -- developer micro-benchmark (not part of the product): times variants of the fused
// kernel and a few FP64-pipe probes on one B200.  Each variant is rbis_kernels.cuh compiled into its
// own namespace with different -D knobs (see dev/build_kbench.sh).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Variant {
  const char* name;
  int tpb;
  int smem;
  void (*geom)(long long N, int sms, int* grid, int* tpb, int* smem);
  void (*launch)(void* kparams_blob, int grid, int tpb, int smem, cudaStream_t st);
  size_t kparams_size;
  void (*fill)(void* blob, long long N, double* vec, double* quat, double* P, double* ll, double* q4, const double* imu,
               const void* ops, long long n_ops, const double* z0, const double* z1, const double* q1, const double* R0,
               const double* R1);
  void (*prep)(int smem);
};
std::vector<Variant>& registry() { static std::vector<Variant> r; return r; }

// ---- FP64 probes ----
__global__ void __launch_bounds__(256) dfma_probe(double* out, int iters, double a, int active_lanes) {
  if ((threadIdx.x & 31) >= active_lanes) return;
  double x[8];
#pragma unroll
  for (int c = 0; c < 8; c++) x[c] = 1e-3 * threadIdx.x + c;
  const double b = a * 0.5 - 0.5;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int c = 0; c < 8; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) s += x[c];
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-issue latency of DFMA: one chain per thread, one warp per SM
__global__ void dfma_latency(double* out, int iters, double a, long long* cycles) {
  double x = threadIdx.x;
  const double b = a * 0.5 - 0.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 64; u++) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main(int argc, char** argv) {
  long long N = argc > 1 ? atoll(argv[1]) : 65536;
  int T = argc > 2 ? atoi(argv[2]) : 200;
  const char* only = argc > 3 ? argv[3] : nullptr;
  if (only && !strcmp(only, "-")) only = nullptr;
  const bool decoupled = argc > 4 ? atoi(argv[4]) != 0 : true;  // zero omega / a couplings in the initial covariance
  const int schedule = argc > 5 ? atoi(argv[5]) : 3;            // 3: configs[2] (IMU + leg odometry + pose), 1: IMU only (configs[1])
  const bool probes = argc > 6 ? atoi(argv[6]) != 0 : false;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, N=%lld T=%d\n", prop.name, prop.multiProcessorCount, N, T);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  // ---- probes ----
  if (probes) {
    double* d; CK(cudaMalloc(&d, 592 * 256 * 8 * 8));
    for (int lanes : {32, 16, 8}) {
      for (int wps : {4, 8, 16, 32}) {  // warps per SM (one block per SM)
        int threads = wps * 32;
        if (threads > 1024) continue;
        int iters = 2000;
        dfma_probe<<<148, threads>>>(d, 10, 1.0000001, lanes);
        CK(cudaEventRecord(e0));
        dfma_probe<<<148, threads>>>(d, iters, 1.0000001, lanes);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 148.0 * wps * lanes * 8 * 16 * 2.0 * iters;
        double instr_per_clk_sm = (148.0 * wps * 8 * 16 * iters) / (ms * 1e-3 * 1.965e9) / 148.0;
        printf("probe dfma lanes=%2d warps/SM=%2d : %.2f TFLOP/s, %.3f warp-DFMA/clk/SM (at 1.965 GHz)\n", lanes, wps,
               flops / (ms * 1e-3) / 1e12, instr_per_clk_sm);
      }
    }
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    dfma_latency<<<1, 32>>>(d, 100, 1.0000001, cyc);
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("probe dfma dependent latency: %.2f cycles\n", (double)h / (100 * 64));
    cudaFree(d); cudaFree(cyc);
  }
  // ---- workload: config-3 schedule, T steps ----
  struct Op { int kind, stream; long long row; double dt; };
  std::vector<Op> ops;
  int li = 0, pi = 0;
  for (int k = 0; k < T; k++) {
    ops.push_back({0, 0, k, 1e-3});
    if (schedule == 3 && k % 2 == 0) ops.push_back({1, 0, li++, 0});
    if (schedule == 3 && k % 100 == 0) ops.push_back({1, 1, pi++, 0});
  }
  srand(1);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2 - 1; };
  std::vector<double> vec(21 * N), quat(4 * N), P(231 * N), q4(4 * N), imu((size_t)T * 6 * N), z0((size_t)(li + 1) * 3 * N),
      z1((size_t)(pi + 1) * 6 * N), q1((size_t)(pi + 1) * 4 * N);
  for (long long n = 0; n < N; n++) {
    for (int i = 0; i < 21; i++) vec[i * N + n] = 0.1 * rnd();
    vec[6 * N + n] = vec[7 * N + n] = vec[8 * N + n] = 0;
    double q[4] = {1 + 0.1 * rnd(), 0.1 * rnd(), 0.1 * rnd(), 0.1 * rnd()};
    double nn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) quat[i * N + n] = q[i] / nn;
    for (int j = 0; j < 21; j++)
      for (int i = 0; i <= j; i++) {
        auto passive = [](int k) { return k < 3 || (k >= 12 && k < 15); };
        const bool blk = (j < 3) || (i >= 12 && j < 15);
        double v = (i == j) ? 0.01 * (1 + 0.1 * rnd()) : 1e-4 * rnd();
        if (decoupled && (passive(i) || passive(j)) && !blk) v = 0.0;
        P[(size_t)(j * (j + 1) / 2 + i) * N + n] = v;
      }
    q4[n] = 7.6e-5; q4[N + n] = 1e-2; q4[2 * N + n] = 3e-10; q4[3 * N + n] = 1e-6;
    for (int i = 0; i < pi; i++) for (int c = 0; c < 4; c++) q1[((size_t)i * 4 + c) * N + n] = quat[c * N + n];
  }
  for (auto& v : imu) v = 0.05 * rnd();
  for (long long k = 0; k < T; k++) for (long long n = 0; n < N; n++) imu[(k * 6 + 5) * N + n] += 9.8;
  for (auto& v : z0) v = 0.1 * rnd();
  for (auto& v : z1) v = 0.1 * rnd();
  double R0[9] = {0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.01};
  double R1[36] = {0};
  for (int i = 0; i < 6; i++) R1[i * 7] = i < 3 ? 0.0025 : 0.0012;
  double *d_vec, *d_quat, *d_P, *d_ll, *d_q4, *d_imu, *d_z0, *d_z1, *d_q1, *d_R0, *d_R1;
  void* d_ops;
  CK(cudaMalloc(&d_vec, vec.size() * 8)); CK(cudaMalloc(&d_quat, quat.size() * 8)); CK(cudaMalloc(&d_P, P.size() * 8));
  CK(cudaMalloc(&d_ll, N * 8)); CK(cudaMalloc(&d_q4, q4.size() * 8)); CK(cudaMalloc(&d_imu, imu.size() * 8));
  CK(cudaMalloc(&d_z0, z0.size() * 8)); CK(cudaMalloc(&d_z1, z1.size() * 8)); CK(cudaMalloc(&d_q1, q1.size() * 8));
  CK(cudaMalloc(&d_R0, 81 * 8)); CK(cudaMalloc(&d_R1, 81 * 8)); CK(cudaMalloc(&d_ops, ops.size() * sizeof(Op)));
  CK(cudaMemcpy(d_q4, q4.data(), q4.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_imu, imu.data(), imu.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_z0, z0.data(), z0.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_z1, z1.data(), z1.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_q1, q1.data(), q1.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_R0, R0, 72, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_R1, R1, 36 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ops, ops.data(), ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
  std::vector<double> ref_vec, ref_P, out_vec(vec.size()), out_P(P.size());
  for (auto& v : registry()) {
    if (only && !strstr(v.name, only)) continue;
    std::vector<char> blob(v.kparams_size, 0);
    int grid, tpb, smem;
    v.geom(N, prop.multiProcessorCount, &grid, &tpb, &smem);
    v.prep(smem);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaMemcpy(d_vec, vec.data(), vec.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_quat, quat.data(), quat.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
      CK(cudaMemset(d_ll, 0, N * 8));
      v.fill(blob.data(), N, d_vec, d_quat, d_P, d_ll, d_q4, d_imu, d_ops, (long long)ops.size(), d_z0, d_z1, d_q1, d_R0, d_R1);
      CK(cudaEventRecord(e0));
      v.launch(blob.data(), grid, tpb, smem, 0);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    CK(cudaMemcpy(out_vec.data(), d_vec, vec.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_P.data(), d_P, P.size() * 8, cudaMemcpyDeviceToHost));
    double ev = 0, eP = 0;
    if (ref_vec.empty()) { ref_vec = out_vec; ref_P = out_P; }
    for (size_t i = 0; i < out_vec.size(); i++) ev = fmax(ev, fabs(out_vec[i] - ref_vec[i]) / fmax(1.0, fabs(ref_vec[i])));
    for (size_t i = 0; i < out_P.size(); i++) eP = fmax(eP, fabs(out_P[i] - ref_P[i]) / 0.01);
    bool finite = true;
    for (size_t i = 0; i < out_vec.size(); i++) if (!std::isfinite(out_vec[i])) finite = false;
    printf("%-28s grid %4d x %3d thr, %6d B smem: %8.3f ms  %7.3f G filter-steps/s  (vs first: dvec %.2e dP %.2e finite=%d)\n", v.name, grid, tpb, smem, best,
           (double)N * T / (best * 1e-3) / 1e9, ev, eP, (int)finite);
  }
  return 0;
}
