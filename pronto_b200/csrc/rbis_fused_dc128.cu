// Decoupled lane-per-filter kernels, 128 filters per CTA (1 warp per scheduler), all 120 active covariance slots in shared
// memory (no tensor memory): ensembles of 10-25 thousand filters.
#define RBIS_TU_NAME dc128
#define RBIS_TU_NS rbisk_dc128
#define RBIS_TPB 128
#define RBIS_PLACEMENT 3
#include "rbis_fused_tu.inc"
