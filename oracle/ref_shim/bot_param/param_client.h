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
#ifdef __cplusplus
}
#endif
