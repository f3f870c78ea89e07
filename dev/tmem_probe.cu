// dev/tmem_probe.cu -- can TMEM serve as a per-thread scratchpad next to shared memory?  Measures
// tcgen05.ld / tcgen05.st (32x32b, one double per thread) latency and throughput, alone and
// interleaved with LDS.64, with 8 warps per SM (two warps share each 32-lane TMEM slice).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tm_st2(uint32_t taddr, double v) {
  const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, uint32_t& lo, uint32_t& hi) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: correctness; 5: LDTM latency; 10+: throughput mixes, per iteration nL LDS.64 + nT LDTM.x2 + nS STTM.x2 + nF DFMA
template <int nL, int nT, int nS, int nF>
__device__ __forceinline__ double mix_loop(int iters, const double* ps, uint32_t my) {
  double a[8];
  int x = 0;
  const uint32_t ps_addr = (uint32_t)__cvta_generic_to_shared(ps);
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = k + threadIdx.x;
  for (int i = 0; i < iters; i++) {
    uint32_t lo[nT > 0 ? nT : 1], hi[nT > 0 ? nT : 1];
    double v[nL > 0 ? nL : 1];
#pragma unroll
    for (int k = 0; k < nT; k++) tm_ld2(my + 2 * ((k * 5) & 127), lo[k], hi[k]);
#pragma unroll
    for (int k = 0; k < nL; k++) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v[k]) : "r"(ps_addr + ((k * 3) % 100) * 2048));
#pragma unroll
    for (int k = 0; k < nF; k++) a[k & 7] = fma(a[k & 7], 1.0000001, 0.5);
#pragma unroll
    for (int k = 0; k < nS; k++) tm_st2(my + 2 * ((k * 5 + 64) & 127), a[k & 7]);
#pragma unroll
    for (int k = 0; k < nL; k++) x ^= __double2loint(v[k]);
    if (nT > 0) tm_wait_ld();
#pragma unroll
    for (int k = 0; k < nT; k++) x ^= (int)(lo[k] + hi[k]);
    if (nS > 0) tm_wait_st();
    asm volatile("" ::: "memory");
  }
  double s = x;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  return s;
}

__global__ void __launch_bounds__(256, 1) tmem_probe(int mode, int iters, double* out, long long* cyc, int* errors) {
  extern __shared__ double smem[];
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s;
  const uint32_t my = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
  double acc = 0;
  long long t0 = 0, t1 = 0;
  double* ps = smem + threadIdx.x;
  for (int k = 0; k < 100; k++) ps[k * 256] = k;
  for (int k = 0; k < 128; k++) tm_st2(my + 2 * k, 1000.0 * threadIdx.x + k + 0.5 * blockIdx.x);
  tm_wait_st();
  __syncthreads();
  if (mode == 0) {
    int bad = 0;
    for (int k = 127; k >= 0; k--) {
      uint32_t lo, hi;
      tm_ld2(my + 2 * k, lo, hi);
      tm_wait_ld();
      const double v = __hiloint2double((int)hi, (int)lo);
      if (v != 1000.0 * threadIdx.x + k + 0.5 * blockIdx.x) bad++;
    }
    if (bad) atomicAdd(errors, bad);
  } else if (mode == 1) {
    // scoreboard check: consume tcgen05.ld results WITHOUT tcgen05.wait::ld, values change every round
    int bad = 0;
    for (int it = 0; it < iters; it++) {
      for (int k = 0; k < 128; k++) tm_st2(my + 2 * k, 1000.0 * threadIdx.x + k + 0.25 * it);
      tm_wait_st();
      double sum = 0, expect = 0;
#pragma unroll
      for (int k = 0; k < 128; k += 8) {
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; j++) tm_ld2(my + 2 * (k + j), lo[j], hi[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) sum = fma(__hiloint2double((int)hi[j], (int)lo[j]), 1.0 + j, sum);
#pragma unroll
        for (int j = 0; j < 8; j++) expect = fma(1000.0 * threadIdx.x + (k + j) + 0.25 * it, 1.0 + j, expect);
      }
      if (sum != expect) bad++;
    }
    if (bad) atomicAdd(errors, bad);
  } else if (mode == 5) {
    if (warp == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          uint32_t lo, hi;
          tm_ld2(my + 2 * ((k * 7 + i) & 127), lo, hi);
          tm_wait_ld();
          acc += __hiloint2double((int)hi, (int)lo);
        }
      }
      t1 = clock64();
    }
  } else {
    t0 = clock64();
    switch (mode) {
      case 10: acc = mix_loop<32, 0, 0, 0>(iters, ps, my); break;
      case 11: acc = mix_loop<0, 32, 0, 0>(iters, ps, my); break;
      case 12: acc = mix_loop<0, 0, 32, 0>(iters, ps, my); break;
      case 13: acc = mix_loop<32, 16, 0, 0>(iters, ps, my); break;
      case 14: acc = mix_loop<32, 32, 0, 0>(iters, ps, my); break;
      case 15: acc = mix_loop<0, 0, 0, 64>(iters, ps, my); break;
      case 16: acc = mix_loop<16, 0, 0, 64>(iters, ps, my); break;
      case 17: acc = mix_loop<0, 16, 0, 64>(iters, ps, my); break;
      case 18: acc = mix_loop<16, 8, 8, 64>(iters, ps, my); break;
      case 19: acc = mix_loop<32, 0, 0, 64>(iters, ps, my); break;
      case 20: acc = mix_loop<24, 12, 8, 64>(iters, ps, my); break;
    }
    t1 = clock64();
  }
  out[(size_t)blockIdx.x * 256 + threadIdx.x] = acc;
  if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
  double* out; long long* cyc; int* err;
  CK(cudaMalloc(&out, 148 * 256 * 8)); CK(cudaMalloc(&cyc, 148 * 8 * 8)); CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  const int smem = 100 * 256 * 8;
  CK(cudaFuncSetAttribute(tmem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct M { int mode; const char* name; } modes[] = {{0, "correctness"}, {1, "no-wait"}, {5, "LDTM latency (ld+wait, 1 warp)"},
    {10, "32 LDS"}, {11, "32 LDTM"}, {12, "32 STTM"}, {13, "32 LDS + 16 LDTM"}, {14, "32 LDS + 32 LDTM"}, {15, "64 DFMA"},
    {16, "64 DFMA + 16 LDS"}, {17, "64 DFMA + 16 LDTM"}, {18, "64 DFMA + 16 LDS + 8 LDTM + 8 STTM"}, {19, "64 DFMA + 32 LDS"},
    {20, "64 DFMA + 24 LDS + 12 LDTM + 8 STTM"}};
  for (auto& m : modes) {
    const int iters = m.mode == 0 ? 1 : (m.mode == 1 ? 300 : 4000);
    CK(cudaMemset(cyc, 0, 148 * 8 * 8));
    CK(cudaEventRecord(e0));
    tmem_probe<<<148, 256, smem>>>(m.mode, iters, out, cyc, err);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[8]; CK(cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost));
    int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
    if (m.mode == 0) { printf("tmem correctness: errors=%d\n", herr); continue; }
    if (m.mode == 1) { printf("tmem loads consumed without wait::ld (scoreboard only), 300 rounds x 128 slots x 8 warps x 148 SMs: errors=%d\n", herr); continue; }
    if (m.mode == 5) { printf("tmem ld+wait dependent latency: %.1f cycles\n", (double)h[0] / iters / 16); continue; }
    printf("mix %-40s: %7.1f cycles per iteration of 8 warps/SM (all 8 warps do the mix)\n", m.name, (double)h[0] / iters);
  }
  return 0;
}
