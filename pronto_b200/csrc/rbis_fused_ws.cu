// Warp-specialised kernels (rbis_ws.cuh): one state warp + eight covariance warps per 32 filters, decoupled ensembles.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define rbisk rbisk_ws
#define RBIS_FUSED_ONLY 1
#include "rbis_kernels.cuh"
#include "rbis_group.cuh"
#include "rbis_ws.cuh"
#undef rbisk
#include "rbis_fused_tu.h"

namespace {
using namespace rbisk_ws;
cudaError_t tu_prepare() {
  cudaError_t e = cudaFuncSetAttribute(ws::rbis_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws::team_smem_bytes(ws::MAX_TEAMS));
  if (e != cudaSuccess) return e;
  if (getenv("RBIS_DEBUG_ATTR")) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, ws::rbis_ws_kernel<false>) == cudaSuccess)
      fprintf(stderr, "rbis_ws_kernel<false>: regs %d, maxThreadsPerBlock %d, static smem %zu, max dynamic smem %d, local %zu\n", a.numRegs,
              a.maxThreadsPerBlock, a.sharedSizeBytes, a.maxDynamicSharedSizeBytes, a.localSizeBytes);
  }
  return cudaFuncSetAttribute(ws::rbis_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ws::team_smem_bytes(ws::MAX_TEAMS));
}
// decoupled programs whose measurement chunks are all aligned index triples only (the host checks)
cudaError_t tu_launch(int blocks_variant, int decoupled, int syn, unsigned grid, int threads, int smem, cudaStream_t st, const void* kparams) {
  KParams kp;
  std::memcpy(&kp, kparams, sizeof(kp));
  if (blocks_variant || !decoupled) return cudaErrorInvalidValue;
  if (syn) ws::rbis_ws_kernel<true><<<grid, threads, smem, st>>>(kp);
  else ws::rbis_ws_kernel<false><<<grid, threads, smem, st>>>(kp);
  return cudaGetLastError();
}
}  // namespace
// filters_per_warp: filters of a team; smem_doubles_per_filter: doubles of a TEAM; max_warps: teams per CTA at most
extern "C" const rbis_fused_tu_t rbis_fused_tu_ws = {
    /*lanes_per_filter=*/32, /*decoupled=*/1, /*threads=*/rbisk_ws::ws::TEAM_THREADS, /*smem=*/rbisk_ws::ws::TEAM_DOUBLES * 8,
    /*max_warps=*/rbisk_ws::ws::MAX_TEAMS, /*filters_per_warp=*/rbisk_ws::ws::TEAM_FILTERS, /*smem_doubles_per_filter=*/rbisk_ws::ws::TEAM_DOUBLES,
    sizeof(rbisk_ws::KParams), tu_prepare, tu_launch};
