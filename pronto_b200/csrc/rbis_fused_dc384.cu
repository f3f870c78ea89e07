// Decoupled lane-per-filter kernels, 384 filters per CTA (3 warps per scheduler, 168 registers): ensembles that fill the GPU.
#define RBIS_TU_NAME dc384
#define RBIS_TU_NS rbisk_dc
#define RBIS_TPB 384
#define RBIS_PLACEMENT 2
#define RBIS_LATE_LOADS 1
#define RBIS_PARK_STATE 2              // the filter state waits in spare tensor memory during a measurement sweep
#define RBIS_PARK_COV_KEEP 0x3ffffffu  // ... and stays in registers during the covariance step
#include "rbis_fused_tu.inc"
