// stand-in for libbot2's BotParam: MavStateEstimator reads exactly one key, state_estimator.utime_history_span
// (MSE/mav_state_est.cpp:15); the value comes from a process-wide variable set by oracle/ref_capi.cpp.
#pragma once
#include <stdint.h>
typedef struct _BotParam BotParam;
#ifdef __cplusplus
extern "C" {
#endif
extern int64_t rbis_ref_shim_history_span;
static inline int bot_param_get_int_or_fail(BotParam*, const char*) { return (int)rbis_ref_shim_history_span; }
/* LegOdoCommon's constructor reads state_estimator.legodo.{mode, r_xyz, r_vxyz, r_vang, r_vxyz_uncertain, r_vang_uncertain}
 * (rbis_legodo_common.cpp:9-33): values come from process-wide variables set by oracle/ref_capi.cpp. */
extern const char* rbis_ref_shim_legodo_mode;
extern double rbis_ref_shim_legodo_r[5]; /* r_xyz, r_vxyz, r_vang, r_vxyz_uncertain, r_vang_uncertain */
char* rbis_ref_shim_param_str(const char* key);
double rbis_ref_shim_param_double(const char* key);
static inline char* bot_param_get_str_or_fail(BotParam*, const char* key) { return rbis_ref_shim_param_str(key); }
static inline double bot_param_get_double_or_fail(BotParam*, const char* key) { return rbis_ref_shim_param_double(key); }
#ifdef __cplusplus
}
#endif
