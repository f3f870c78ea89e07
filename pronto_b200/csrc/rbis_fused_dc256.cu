// Decoupled lane-per-filter kernels, 256 filters per CTA (2 warps per scheduler, 255 registers, nothing parked): ensembles of
// 25-49 thousand filters, which leave 40 % of the SMs idle at 384 per CTA.
#define RBIS_TU_NAME dc256
#define RBIS_TU_NS rbisk_dc256
#define RBIS_TPB 256
#define RBIS_PLACEMENT 2
#include "rbis_fused_tu.inc"
