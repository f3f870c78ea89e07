/* rbis_batch.h -- C ABI of the B200-native batched RBIS EKF (drop-in boundary).
 *
 * One handle = one ensemble of N independent 21-state RBIS filters resident on one B200.  Every
 * entry point replaces, for a whole ensemble at once, one piece of the reference's per-filter
 * interface under /root/reference/state-estimator/src/mav_state_est/ (cited per function as
 * MSE/<file>:<lines>).  Plain pointers and sizes only; no C++/torch types; never throws.
 *
 * Conventions
 *   - All floating point is IEEE double.  Indices are int32, time is int64 microseconds.
 *   - Ensemble arrays are structure-of-arrays with the FILTER INDEX FASTEST:
 *       vec  [21][N]   state vector, index layout of RBIS (MSE/rbis.hpp:22-24 + eigen_utils):
 *                      0-2 angular velocity, 3-5 body velocity, 6-8 chi, 9-11 position,
 *                      12-14 acceleration, 15-17 gyro bias, 18-20 accel bias
 *       quat [4][N]    orientation (w,x,y,z)
 *       cov  [441][N]  RBIM, element (r,c) at row r + 21*c (Eigen column-major, MSE/rbis.hpp:30,123)
 *       imu  [rows][6][N]  gyro xyz then accelerometer xyz per IMU sample row
 *       z    [rows][m][N], meas quat [rows][4][N]     (N -> cols under rbis_batch_set_column_map)
 *   - `mem` says where the caller's ensemble arrays live: host memory (copied by the library on the
 *     handle's stream; use pinned memory for truly asynchronous copies) or device memory (used in
 *     place; must stay valid until the call's work has completed).
 *   - Lifetime of host inputs: a call that takes host arrays with mem == RBIS_MEM_HOST returns once the copies are ENQUEUED
 *     (pageable memory is staged by the CUDA runtime before the call returns; PINNED memory is read later, by the copy
 *     engine).  Pinned input arrays of rbis_batch_run_fused must therefore stay valid and unchanged until a later
 *     rbis_batch_synchronize(), or a rbis_batch_wait() on a ticket recorded after the call, has returned.
 *   - Calls are stream-ordered on the handle and return once enqueued unless stated otherwise;
 *     rbis_batch_synchronize() waits.  A handle is not thread-safe (the reference is single
 *     threaded too: MSE/lcm_front_end.cpp:223-229).
 *   - Return value: 0 on success, negative rbis_status_t on error; rbis_last_error() gives the text.
 *   - The covariance is kept symmetric-packed on the device (upper triangle of `cov` is read on
 *     input; both triangles are written on output).  The reference never symmetrises, but its
 *     covariance is symmetric to rounding by construction.
 *   - There is no CPU fallback: every compute entry point runs CUDA kernels built for sm_100a and
 *     fails with RBIS_ERR_CUDA when no such device is present.
 */
#ifndef RBIS_BATCH_H_
#define RBIS_BATCH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBIS_NUM_STATES 21
#define RBIS_COV_ELEMS 441
#define RBIS_MAX_MEAS 9      /* largest index set any reference handler emits (SE/gpf/laser_gpf_lib.cpp:108-110) */
#define RBIS_MAX_STREAMS 8
#define RBIS_NUM_STATS 96    /* doubles per statistics chunk, see rbis_batch_stats */

typedef struct rbis_batch rbis_batch_t;

typedef enum {
  RBIS_OK = 0,
  RBIS_ERR_INVALID = -1,  /* bad argument */
  RBIS_ERR_CUDA = -2,     /* CUDA runtime failure, or no sm_100 device */
  RBIS_ERR_ALLOC = -3,
  RBIS_ERR_STATE = -4     /* call sequence error (e.g. restore of an empty snapshot slot) */
} rbis_status_t;

typedef enum {
  RBIS_MEM_HOST = 0,
  RBIS_MEM_DEVICE = 1,
  /* rbis_batch_run_fused only, OR-ed with one of the above: the sensor rows (imu, z, quat) are FLOAT arrays of the same shapes.  They
   * are widened to double exactly on the device, so the results equal those of the same values passed as doubles; a host log that is
   * float-valued (most sensor data is) then crosses PCIe in half the bytes.  Per-filter R arrays stay double. */
  RBIS_MEM_F32_ROWS = 4
} rbis_mem_t;

/* Fused-program op kinds. */
typedef enum {
  RBIS_OP_IMU = 0,       /* RBISIMUProcessStep::updateFilter, MSE/rbis_update_interface.cpp:30-52 */
  RBIS_OP_MEAS = 1,      /* RBISIndexed[PlusOrientation]Measurement::updateFilter, :54-107 */
  RBIS_OP_SNAPSHOT = 2,  /* save (state, cov, loglik) of every filter into ring slot `row` */
  RBIS_OP_RESTORE = 3    /* load them back: the history rewind of MSE/mav_state_est.cpp:35-70 */
} rbis_op_kind_t;

typedef enum {
  RBIS_R_SHARED_FULL = 0,      /* one m x m column-major matrix for all filters (HOST pointer always) */
  RBIS_R_PER_FILTER_DIAG = 1   /* diagonal, [m][N], located as `mem` says */
} rbis_r_mode_t;

typedef struct {
  double g_val;            /* eigen_utils g_val, g_vec = (0,0,-g_val); default 9.8 */
  double chi_tol;          /* eigen_utils chiToQuat tolerance; default 1e-6 */
  int32_t ctor_folds_chi;  /* RigidBodyState(VectorXd) ctor folds chi into its quaternion; default 1 */
  int32_t renormalize_quat;/* 0 = reference behaviour (never renormalises, MSE/rbis.cpp:219-227);
                              1 = renormalise the quaternion after every applied delta */
  int32_t snapshot_slots;  /* ring slots for RBIS_OP_SNAPSHOT / RESTORE; 0 = none */
  int32_t device;          /* CUDA device ordinal */
  int32_t launch_groups;   /* fused launches are split into this many CTA ranges on separate streams so that
                              consecutive rbis_batch_run_fused calls overlap and the last, partially filled wave of
                              one launch does not idle SMs; 0 = automatic (6 when the CTAs do not fill whole
                              waves, else 1), 1 = off, max 8 */
  int32_t dense_only;      /* 0 (default) = automatic: fused programs run the "decoupled" kernel variant (only the
                              15x15 block of v, chi, p, b_g, b_a on chip, 384 filters per SM) whenever every filter's
                              covariance couplings to the omega / a rows are exactly zero -- true for every filter
                              started from the reference's diagonal initial covariance (MSE/rbis_initializer.cpp:85-91)
                              until a measurement indexes omega or a -- and no stream of the program indexes omega or a;
                              results are bit-identical to the dense variant.
                              1 = always run the dense variant (whole covariance on chip, 256 filters per SM). */
  int32_t mapping;         /* lanes that cooperate on one filter in the fused kernels.  1 = the lane-per-filter kernels (a
                              B200 needs ~57,000 filters to fill them); 2, 4, 8, 16 = the warp-group kernels for small or
                              split ensembles (BASELINE configs[1], a 65,536-filter ensemble sharded over 8 GPUs): the
                              covariance products of MSE/rbis.cpp:113-118,134-140 are split over the lanes of a group, the
                              covariance is carried as a full matrix in shared memory.
                              0 (default) = automatic by ensemble size.  Results of the two mappings agree
                              BIT FOR BIT (every covariance element is computed by the same expression in both;
                              tests/test_gpu_group.py), so the choice is invisible in the results. */
  int32_t lane_filters_per_cta; /* filters per CTA of the decoupled lane-per-filter kernels: 384 (three warps per scheduler, for
                              ensembles that fill the GPU), 256 or 128 (more SMs busy for 10-50 thousand filters); 0 = automatic.
                              Same bits for every value. */
  int32_t piece_ops;       /* with launch groups, a fused program of >= 2 * piece_ops ops is cut into consecutive pieces of about
                              this many ops over the same inputs, each piece one set of launches, so that the partially filled last
                              wave of CTAs of a piece overlaps the next piece (rbis_batch.cu).  0 = automatic (256).  Smaller
                              pieces (64-100) pay when the caller reads a result after every call, which keeps consecutive CALLS
                              from overlapping; results do not depend on it. */
  int32_t synth_materialize; /* rbis_batch_run_fused_synth: 0 (default) = decoupled programs with the fast generator (rbis_synth_t::mode 1)
                              draw their input rows INSIDE the fused kernel -- no per-filter input exists in HBM; 1 = always generate
                              the rows into device buffers first and run the ordinary fused kernels over them.  Same bits either way
                              (tests/test_gpu_synth.py). */
} rbis_batch_config_t;

/* One measurement stream = the constant part of an RBISIndexedMeasurement /
 * RBISIndexedPlusOrientationMeasurement constructor (MSE/rbis_update_interface.hpp:90-96,111-117)
 * plus where its per-row data live. */
typedef struct {
  int32_t m;                      /* 1..RBIS_MAX_MEAS */
  int32_t has_orientation;        /* 0 indexedMeasurement, 1 indexedPlusOrientationMeasurement */
  int32_t r_mode;                 /* rbis_r_mode_t */
  int32_t sensor_id;              /* RBISUpdateInterface::sensor_enum value, carried only */
  int32_t idx[RBIS_MAX_MEAS];     /* state indices, 0..20 */
  int32_t reserved;
  const double* z;                /* [rows][m][N]; entries at chi indices are ignored when has_orientation */
  const double* quat;             /* [rows][4][N] or NULL */
  const double* R;                /* see rbis_r_mode_t */
  int64_t rows;
} rbis_stream_t;

typedef struct {
  int32_t kind;    /* rbis_op_kind_t */
  int32_t stream;  /* RBIS_OP_MEAS: stream number */
  int64_t row;     /* IMU / MEAS: row in the stream; SNAPSHOT / RESTORE: ring slot */
  int64_t utime;   /* becomes the ensemble's utime (MSE/mav_state_est.cpp:60) */
  double dt;       /* RBIS_OP_IMU only */
} rbis_op_t;

const char* rbis_last_error(void);
void rbis_default_config(rbis_batch_config_t* cfg);

/* ---- lifetime: replaces `new MavStateEstimator(new RBISResetUpdate(...))`, MSE/mav_state_est.cpp:12-22.
 * The handle owns all device memory; the caller owns every buffer it passes in. */
int rbis_batch_create(rbis_batch_t** out, int64_t n_filters, const rbis_batch_config_t* cfg);
int rbis_batch_destroy(rbis_batch_t* h);
int rbis_batch_synchronize(rbis_batch_t* h);
int64_t rbis_batch_num_filters(const rbis_batch_t* h);
/* Raw CUDA stream (cudaStream_t) of the handle, for event timing by the caller.  With launch groups the fused
 * kernels run on internal streams that this stream joins at the next non-fused call: to time fused work, call
 * rbis_batch_record() (it joins) and then record the timing event on this stream. */
void* rbis_batch_stream(rbis_batch_t* h);
/* Kernels launched by this handle so far (bench.py's gpu_launches). */
int64_t rbis_batch_launch_count(const rbis_batch_t* h);
/* Kernel variant the last rbis_batch_run_fused / single-op call launched: 0 dense, 2 decoupled (see
 * rbis_batch_config_t::dense_only), +1 when the program has measurement chunks other than uncorrelated aligned index
 * triples (the instantiations that also contain the one-row and the correlated-block updates), +4 when the kernel drew its
 * input rows itself (rbis_batch_run_fused_synth, SYN instantiations), + 16 * lanes per filter for the warp-group kernels
 * (rbis_batch_config_t::mapping > 1); -1 before the first launch. */
int rbis_batch_last_kernel_variant(const rbis_batch_t* h);

/* ---- RBISResetUpdate::updateFilter (MSE/rbis_update_interface.cpp:23-28): posterior := given,
 * loglikelihood := 0 (or `loglik` [N] when non-NULL).  cov may be NULL to keep the current one. */
int rbis_batch_set_state(rbis_batch_t* h, const double* vec, const double* quat, const double* cov,
                         const double* loglik, int64_t utime, int mem);
/* MavStateEstimator::getHeadState + getMeasurementsLogLikelihood (MSE/mav_state_est.cpp:82-96).
 * Any output pointer may be NULL.  Synchronises when mem == RBIS_MEM_HOST. */
int rbis_batch_get_state(rbis_batch_t* h, double* vec, double* quat, double* cov, double* loglik,
                         int64_t* utime, int mem);
/* Single-filter (array-of-structures) access for the C++ shim: vec[21], quat[4], cov[441]. Synchronous. */
int rbis_batch_set_filter(rbis_batch_t* h, int64_t n, const double* vec, const double* quat, const double* cov,
                          double loglik);
int rbis_batch_get_filter(rbis_batch_t* h, int64_t n, double* vec, double* quat, double* cov, double* loglik);

/* ---- process noise = the q_* constructor arguments of RBISIMUProcessStep
 * (MSE/rbis_update_interface.hpp:70-75); variances, as InsHandler squares them
 * (MSE/sensor_handlers.cpp:18-25).  Either one value for all filters or one per filter. */
int rbis_batch_set_process_noise(rbis_batch_t* h, double q_gyro, double q_accel, double q_gyro_bias,
                                 double q_accel_bias);
int rbis_batch_set_process_noise_per_filter(rbis_batch_t* h, const double* q_gyro, const double* q_accel,
                                            const double* q_gyro_bias, const double* q_accel_bias, int mem);

/* ---- one update for every filter (each is a one-op fused program) ----
 * RBISIMUProcessStep::updateFilter: insUpdateState (MSE/rbis.cpp:37-75) + insUpdateCovariance
 * linearised at the PRIOR state (MSE/rbis.cpp:77-122, rbis_update_interface.cpp:38-39).
 * gyro/accel: [3][N]. */
int rbis_batch_ins_step(rbis_batch_t* h, const double* gyro, const double* accel, double dt, int64_t utime,
                        int mem);
/* indexedMeasurement + rbisApplyDelta (MSE/rbis.cpp:160-178,219-227).  z [m][N]; R per r_mode. */
int rbis_batch_indexed_update(rbis_batch_t* h, int m, const int32_t* idx, const double* z, const double* R,
                              int r_mode, int64_t utime, int mem);
/* indexedPlusOrientationMeasurement + rbisApplyDelta (MSE/rbis.cpp:189-227).  quat [4][N]. */
int rbis_batch_indexed_orient_update(rbis_batch_t* h, int m, const int32_t* idx, const double* z,
                                     const double* quat, const double* R, int r_mode, int64_t utime, int mem);

/* ---- column maps: filters that share input data (parameter sweeps in the style of the reference's noise
 * identification, state-estimator/matlab/ins_noise_opt_script_mex.m:38-61, where every parameter point replays
 * the SAME log; or one noise realisation evaluated under many parameter sets).  While a map is set, the
 * corresponding array handed to rbis_batch_run_fused has `cols` columns instead of N and filter n reads column
 * map[n]:  imu [rows][6][cols]  (which = -1),  z [rows][m][cols] and quat [rows][4][cols] of stream `which` >= 0.
 * map: HOST int32[N] with entries in [0, cols), copied by the call; NULL restores the identity (cols ignored).
 * Per-filter R ([m][N]) and everything else stay per filter.  The one-op entry points above ignore the maps.
 * Synchronises the handle. */
int rbis_batch_set_column_map(rbis_batch_t* h, int which, const int32_t* map, int64_t cols);

/* ---- the fused hot path: run `n_ops` ops (host array, executed in order) for every filter in ONE
 * kernel launch, state and covariance resident on chip throughout.  This is the batch form of the
 * roll-forward loop MSE/mav_state_est.cpp:50-70.  imu may be NULL when no IMU op is present. */
int rbis_batch_run_fused(rbis_batch_t* h, int64_t n_ops, const rbis_op_t* ops, const double* imu,
                         int64_t imu_rows, int n_streams, const rbis_stream_t* streams, int mem);

/* ---- on-device synthesis of per-filter sensor streams for Monte-Carlo noise sweeps: every filter reads the
 * same noise-free log plus its own noise realisation, so the host passes the noise-free rows and a seed and the per-filter
 * rows are produced in HBM, never crossing PCIe.  Counter-based generator, one draw per (seed, GLOBAL filter index, step
 * counter, channel): key = seed ^ filter*K1 ^ step*K2 ^ channel*K3, a = splitmix64(key), b = splitmix64(a), Box-Muller.
 * mode 0 (exact): u1, u2 from 53 bits, sqrt(-2 ln u1) cos(2 pi u2) in double -- pronto_b200/synth.py:normal is its numpy
 * statement (bit-exact integers, libm calls within an ulp or two); mode 1 (fast): same counters and hash, 24-bit uniforms,
 * single-precision SFU ln / sqrt / cos.  Channels follow SURVEY.md 8d: 0-2 gyro, 3-5 accelerometer, then per stream. */
typedef struct {
  int32_t m;                      /* columns of z */
  int32_t has_orientation;        /* also draw a measured orientation: mean_quat (x) Exp(sigma_rot * n) */
  int32_t channel;                /* column a draws channel + a */
  int32_t channel_rot;            /* the rotation perturbation draws channel_rot + 0..2 */
  const double* mean;             /* HOST [rows][m] noise-free rows */
  const double* mean_quat;        /* HOST [rows][4] (w,x,y,z) or NULL */
  const int64_t* step;            /* HOST [rows] counter value of each row (the global time-step index) */
  double sigma[RBIS_MAX_MEAS];    /* standard deviation per column */
  double sigma_rot[3];
  int64_t rows;
} rbis_synth_stream_t;
typedef struct {
  uint64_t seed;
  int64_t first_filter;           /* global index of the handle's filter 0: shards of one ensemble draw the same noise for any GPU count */
  int32_t mode;                   /* 0 exact (double precision, one hash pair per sample: pronto_b200/synth.py:normal), 1 fast (one hash per
                                     PAIR of channels, single-precision Box-Muller; the mode the fused kernels can draw in-kernel) */
  int32_t n_streams;
  const double* imu_mean;         /* HOST [imu_rows][6]: noise-free gyro xyz, accelerometer xyz readings */
  const int64_t* imu_step;        /* HOST [imu_rows] */
  int64_t imu_rows;
  double sigma_gyro, sigma_accel; /* >= 0: standard deviations for all filters; < 0: per filter sqrt(q_gyro / dt), sqrt(q_accel / dt) from
                                     the handle's process noise, i.e. sample noise consistent with Qd = q dt (MSE/rbis.cpp:116) */
  double dt;
  const rbis_synth_stream_t* streams;
} rbis_synth_t;
/* Materialise the streams into DEVICE arrays imu_out [imu_rows][6][N], z_out[s] [rows][m][N], quat_out[s] [rows][4][N]
 * (enqueued on the handle's stream; the host arrays of `syn` are consumed before the call returns).  This is also how a
 * test feeds the CPU oracle exactly what the device drew. */
int rbis_batch_synthesize(rbis_batch_t* h, const rbis_synth_t* syn, double* imu_out, double* const* z_out, double* const* quat_out);
/* rbis_batch_run_fused with synthesised inputs: streams[s].z / .quat / .rows are ignored (taken from syn->streams[s]; a per-filter R
 * is a DEVICE array on this path).  Mode 1 on a decoupled ensemble (rbis_batch_config_t::dense_only): the fused kernel draws every row
 * it consumes itself, from the noise-free row and the filter's counters -- nothing is materialised, the launch reads no per-filter input.
 * Otherwise (mode 0, dense kernels, rbis_batch_config_t::synth_materialize) the rows are generated into library-owned device buffers
 * (double buffered across calls) and consumed by the ordinary fused launch; both ways give the same bits as rbis_batch_synthesize +
 * rbis_batch_run_fused.  Host -> device traffic per call: the op list and the noise-free rows. */
int rbis_batch_run_fused_synth(rbis_batch_t* h, int64_t n_ops, const rbis_op_t* ops, int n_streams, const rbis_stream_t* streams,
                               const rbis_synth_t* syn);

/* ---- delayed measurements: the batch form of MavStateEstimator::addUpdate's out-of-order insert +
 * replay (MSE/mav_state_est.cpp:28-80) over updateHistory's time-ordered multimap
 * (MSE/update_history.cpp:16-54), for an ensemble whose filters share one arrival schedule.
 * The planner keeps the time-ordered list of updates (equal utime keeps arrival order, as the hinted
 * multimap insert does) and turns arrivals into an op program for rbis_batch_run_fused: in-order
 * arrivals append ops; an arrival older than the head becomes RESTORE(nearest earlier snapshot) +
 * the replay of every update from there to the head.  Where the reference keeps a full posterior in
 * every history node, the device keeps a ring of `snapshot_slots` ensemble snapshots taken every
 * `snapshot_period_us` (at utimes u with (u - snapshot_phase_us) % period == 0, after the last
 * update stamped u), plus one at the start.  Updates older than the oldest retained history entry are
 * discarded as in update_history.cpp:28-39; so are updates that no retained snapshot precedes.
 * History is truncated to `history_span_us` behind the newest update after every roll-forward
 * (mav_state_est.cpp:74-77).  The rewind window is therefore ~ snapshot_slots x snapshot_period_us: choose them so that it
 * covers history_span_us when every update the reference would replay must be replayed (rbis_planner_create leaves a
 * warning in rbis_last_error() when it does not; the C++ shim's default does).  Host-side only; no device work.
 *
 * ONE SCHEDULE PER HANDLE.  A planner serves one rbis_batch_t: all of its filters see the same arrivals in the same order --
 * the shape of every workload in BASELINE.json (a noise sweep or a parameter sweep replays ONE log; delayed pose fixes are
 * late for every realisation alike).  Filters whose measurements arrive on DIFFERENT schedules belong in different handles,
 * one planner each: handles are independent (own streams, own snapshot ring), their launches run concurrently on the device,
 * and the warp-group kernels keep small sub-ensembles efficient; tests/test_gpu_parity.py
 * (test_two_arrival_schedules_as_two_handles) runs two schedules side by side against the oracle's per-filter histories. */
typedef struct rbis_planner rbis_planner_t;
int rbis_planner_create(rbis_planner_t** out, int64_t utime0, int32_t snapshot_slots, int64_t snapshot_period_us,
                        int64_t snapshot_phase_us, int64_t history_span_us);
int rbis_planner_destroy(rbis_planner_t* p);
/* addUpdate(update, roll_forward): update->kind is RBIS_OP_IMU or RBIS_OP_MEAS.  Returns 0 = accepted,
 * 1 = discarded (too old), negative = error.  With roll_forward == 0 the update only enters the
 * history; the next roll-forward processes everything outstanding in one replay. */
int rbis_planner_add_update(rbis_planner_t* p, const rbis_op_t* update, int roll_forward);
/* Ops planned so far and not yet taken. */
int64_t rbis_planner_pending(const rbis_planner_t* p);
/* Copy out (and clear) the pending program; fails with RBIS_ERR_INVALID if cap is too small. */
int rbis_planner_take(rbis_planner_t* p, rbis_op_t* out, int64_t cap, int64_t* n_out);
/* Counters: [0] updates accepted, [1] discarded, [2] rewinds, [3] updates replayed (re-applied),
 * [4] snapshots taken, [5] history entries retained. */
int rbis_planner_counters(const rbis_planner_t* p, int64_t out[6]);

/* ---- EKF smoother ("next" row 1 of SURVEY.md 8f): ekfSmoothingStep (MSE/rbis.cpp:234-266) driven as
 * MavStateEstimator::EKFSmoothBackwardsPass (MSE/mav_state_est.cpp:98-189).  Where the reference keeps a posterior in
 * every history node, the ensemble keeps the posteriors it wants smoothed in the snapshot ring: the forward program
 * stores the posterior of update u into ring slot slot[u] with RBIS_OP_SNAPSHOT.
 *
 * rbis_smooth_plan: host-side mirror of the reference's backwards traversal over a history of n updates in history
 * order (is_ins[u] != 0 for IMU process steps; every other update counts as a measurement, as the reset at the front
 * of the reference's history does).  It reproduces the traversal statement by statement, including the extra
 * decrement after trailing measurements (mav_state_est.cpp:130).  Outputs: the two slots the recursion starts from,
 * the smoothing steps in execution (backward-in-time) order (capacity n), and alias[u] = the slot that holds update
 * u's posterior after smoothing (the reference copies the smoothed posterior into the measurement updates that follow
 * an IMU step, :163-169; here they alias the IMU step's slot).  Returns the number of steps, or a negative status. */
typedef struct {
  int32_t cur_slot;       /* posterior being smoothed: the last measurement of the time step, or the IMU step's own */
  int32_t cur_pred_slot;  /* the IMU step's posterior of that time step (becomes the next step's prediction) */
  int32_t out_slot;       /* where the smoothed state and covariance are written (the IMU step's slot) */
  int32_t reserved;
} rbis_smooth_step_t;
int64_t rbis_smooth_plan(int64_t n, const uint8_t* is_ins, const int32_t* slot, int32_t* next_pred_slot,
                         int32_t* next_slot, rbis_smooth_step_t* steps, int32_t* alias);
/* Runs the steps on the device, one warp per filter (rbis_smooth.cuh); dt as passed to EKFSmoothBackwardsPass.
 * Slots must hold snapshots; `steps` is a HOST array.  Stream-ordered like every other call. */
int rbis_batch_smooth_backward(rbis_batch_t* h, int32_t next_pred_slot, int32_t next_slot, int64_t n_steps,
                               const rbis_smooth_step_t* steps, double dt);
/* Read a ring slot back: vec [21][N], quat [4][N], cov [441][N], loglik [N]; any may be NULL. */
int rbis_batch_get_snapshot(rbis_batch_t* h, int32_t slot, double* vec, double* quat, double* cov, double* loglik, int mem);

/* ---- IMU conditioning ("next" row 4 of SURVEY.md 8f): the accelerometer notch cascade of the Atlas INS path,
 * InsHandler::doFilter (MSE/sensor_handlers.cpp:155-162) over IIRNotch (estimate_tools/src/estimate_tools/
 * iir_notch.cpp:3-60), for every column of an IMU chunk at once.
 * configure: n_stages filters at notch_freq * 2^i, i = 0.., sample rate fs (the reference: 3 stages, fs = 1000,
 * MSE/sensor_handlers.cpp:29-41), for chunks of `cols` columns (N, or the column count of the IMU column map); resets
 * the carried samples to zero as the IIRNotch constructor does.
 * filter: filters rows 3..5 (accelerometer x,y,z) of imu [rows][6][cols] IN PLACE, continuing from the state the
 * previous call left, so a log can be conditioned chunk by chunk before rbis_batch_run_fused consumes it. */
int rbis_batch_notch_configure(rbis_batch_t* h, double notch_freq, double fs, int n_stages, int64_t cols);
int rbis_batch_notch_filter(rbis_batch_t* h, double* imu, int64_t rows, int mem);

/* ---- ensemble statistics against a truth state (error definition of
 * SE/noise_id/noise_id.cpp:37-38; NEES over velocity+chi+position as roll_forward.cpp:54-57).
 * truth_vec [21] / truth_quat [4] (host) shared by all filters, or per-filter [21][N] / [4][N] (`mem`)
 * when per_filter != 0.  Filters are reduced in chunks of `chunk` filters (power of two, <= 1024)
 * in a fixed tree order; out_chunks (host) receives [n_chunks][RBIS_NUM_STATS]:
 *   [0..20] sum e_i, [21..41] sum e_i^2, [42] sum NEES, [43] sum NEES^2, [44] sum loglik,
 *   [45] count non-finite filters, [46] count filters, [47] count NEES within the 95% chi2(9) bounds,
 *   rest reserved (0).
 * out_per_filter (optional, `mem`): [23][N] = e[21], NEES, loglik per filter.  Synchronous. */
int rbis_batch_stats(rbis_batch_t* h, const double* truth_vec, const double* truth_quat, int per_filter,
                     int chunk, double* out_chunks, int64_t* n_chunks, double* out_per_filter, int mem);
/* Same reduction, enqueued on the handle's stream without waiting: truth is shared ([21], [4], host),
 * out_chunks must be PINNED host memory and is valid once a ticket recorded after this call has been
 * waited for.  Lets a caller read per-chunk results every step without stalling the input copies of
 * the next step. */
int rbis_batch_stats_enqueue(rbis_batch_t* h, const double* truth_vec, const double* truth_quat, int chunk,
                             double* out_chunks, int64_t* n_chunks);
/* The same statistics over SNAPSHOT SLOT `slot` (written by an RBIS_OP_SNAPSHOT of an earlier fused program) instead of the live
 * ensemble, on a side stream: the pass is ordered after the launch that wrote the slot and nothing waits for it -- the fused
 * launches that follow keep overlapping the ones before, which a read of the LIVE state (rbis_batch_stats_enqueue) prevents.
 * A later program that snapshots into the same slot waits for this pass.  out_chunks must be PINNED host memory; it is valid
 * once rbis_batch_wait(*ticket) returns.  This is how a Monte-Carlo driver reads the ensemble error / NEES every step of a long
 * replay (append RBIS_OP_SNAPSHOT to the step's last program, alternate two slots) without draining the device. */
int rbis_batch_stats_snapshot_enqueue(rbis_batch_t* h, int32_t slot, const double* truth_vec, const double* truth_quat, int chunk,
                                      double* out_chunks, int64_t* n_chunks, int32_t* ticket);
/* ---- statistics of a SHARDED ensemble: one process per GPU, each handle holds the contiguous filter range
 * [first_chunk * chunk, first_chunk * chunk + N) of an ensemble of total_chunks chunks.  The shard's chunk partials are written
 * into its rows of a zero-initialised DEVICE table [total_chunks][RBIS_NUM_STATS], the table is summed over the ranks with
 * ncclAllReduce on the handle's stream (exact: every row has one non-zero contributor), and reduced on the device in
 * ascending chunk order -- no host round trip before the collective.  The totals (and the table) are therefore bit-identical
 * on every rank and for every GPU count, and equal to rbis_batch_stats + rbis_stats_reduce_chunks on one GPU.
 *   nccl_comm: the caller's ncclComm_t (as void*), one rank per handle, created on the handle's device; NULL = single
 *              process (no collective; first_chunk = 0, total_chunks = the shard's own count).  The library resolves
 *              ncclAllReduce from the NCCL already loaded in the calling process and has no link-time NCCL dependency.
 *   truth_vec [21] / truth_quat [4]: host, shared by all filters.  out_totals: host [RBIS_NUM_STATS].
 *   out_table: host [total_chunks][RBIS_NUM_STATS] or NULL.  Synchronous (collective: every rank of the communicator calls it). */
int rbis_batch_stats_allreduce(rbis_batch_t* h, void* nccl_comm, const double* truth_vec, const double* truth_quat, int chunk,
                               int64_t first_chunk, int64_t total_chunks, double* out_totals, double* out_table);

/* ---- windowed noise-identification likelihood: the per-window term of sampleProcessForward + negLogLikelihood
 * (state-estimator/src/noise_id/noise_id.cpp:36-40,44-65) for every filter at once.  With the
 * filters' heads = states rolled forward over one window from the truth, truth_vec [21][N] / truth_quat [4][N] = the
 * truth at the window ends, and base_cov [441][base_cols] = the covariance the same window accumulates under ZERO
 * process noise (column base_map[n] for filter n; base_map HOST int32[N], NULL = column n):
 *   e = head (-) truth (subtractState + quatToChi),  C = cov - base_cov[:, base_map[n]],
 *   out[n] = -loglike_normalized(e_A, 0, C_AA) = log det C_AA + e_A^T C_AA^-1 e_A     (A = active_idx, host int32[n_active])
 * Summing out over the windows of one parameter point gives negLogLikelihood.  truth, base_cov, out live as `mem`
 * says.  Synchronous. */
int rbis_batch_window_neg_loglik(rbis_batch_t* h, const double* truth_vec, const double* truth_quat, const double* base_cov,
                                 const int32_t* base_map, int64_t base_cols, int n_active, const int32_t* active_idx,
                                 double* out, int mem);

/* ---- wire structs and IMU decode -- host side.  Struct-level layouts of the reference's LCM types;
 * the LCM byte encoding (big-endian marshalling + fingerprint) is not produced here (no LCM, no lcm-gen, no log to check
 * an encoder against).
 * filter_state_t (pronto-lcmtypes/lcmtypes/pronto_filter_state_t.lcm) as rbisCreateFilterStateMessage writes it
 * (MSE/rbis.cpp:268-285: quat (w,x,y,z), state = vec, cov = Map<RBIM> i.e. COLUMN-major although the .lcm comment says row
 * major) and RBIS(const pronto_filter_state_t*) reads it (MSE/rbis.hpp:58-67). */
typedef struct {
  int64_t utime;
  double quat[4];
  int32_t num_states;        /* 21 */
  int32_t reserved0;
  double state[RBIS_NUM_STATES];
  int32_t num_cov_elements;  /* 441 */
  int32_t reserved1;
  double cov[RBIS_COV_ELEMS];
} rbis_filter_state_t;
/* One message per filter first..first+count-1 (synchronous; reads the whole ensemble: a logging path). */
int rbis_batch_get_filter_states(rbis_batch_t* h, int64_t first, int64_t count, rbis_filter_state_t* out);
/* Filters first..first+count-1 := the messages (state, quaternion, covariance; log-likelihood 0), as RBIS(msg) +
 * Map<const RBIM>(msg->cov) of the reference's log loader (state-estimator/src/noise_id/noise_id.cpp:83-86). */
int rbis_batch_set_filter_states(rbis_batch_t* h, int64_t first, int64_t count, const rbis_filter_state_t* msgs);

/* indexed_measurement_t (pronto-lcmtypes/lcmtypes/pronto_indexed_measurement_t.lcm), bounded to RBIS_MAX_MEAS rows. */
typedef struct {
  int64_t utime, state_utime;
  int32_t measured_dim;
  int32_t measured_cov_dim;                      /* measured_dim^2 */
  double z_effective[RBIS_MAX_MEAS];
  int32_t z_indices[RBIS_MAX_MEAS];
  int32_t reserved;
  double R_effective[RBIS_MAX_MEAS * RBIS_MAX_MEAS];
} rbis_indexed_measurement_t;
/* IndexedMeasurementHandler::processMessage (MSE/sensor_handlers.cpp:576-582): the constant part of a stream (index set,
 * R = Map<MatrixXd>(R_effective, m, m), column-major, copied to R_out[m*m]; sensor_id = indexed_sensor).  z rows are the
 * messages' z_effective; out->z / rows stay for the caller to fill. */
int rbis_stream_from_indexed_measurement(const rbis_indexed_measurement_t* msg, rbis_stream_t* out, double* R_out);

/* KVH raw IMU batches of the Atlas INS path.  rbis_kvh_packet_t = bot_core::kvh_raw_imu_t; a batch message carries
 * num_packets of them, NEWEST FIRST, most of which were already seen in earlier batches.
 * rbis_kvh_decode_batch = IMUStream::convertFromLCMBatch (estimate_tools/src/estimate_tools/imu_stream.cpp:62-97): the
 * packets newer than the last one seen come back in out_new, oldest first, with utime_delta = time since the previous new
 * packet; the others in out_old (optional) with the reference's obfuscated delta; a packet counter that runs backwards
 * resets the stream.  Both output arrays need room for num_packets entries. */
typedef struct {
  int64_t utime;
  int64_t packet_count;
  double delta_rotation[3];
  double linear_acceleration[3];
} rbis_kvh_packet_t;
typedef struct {
  int64_t utime_raw, utime_batch, utime, utime_delta, packet_count;
  double delta_rotation[3];
  double linear_acceleration[3];
} rbis_imu_packet_t;
typedef struct rbis_kvh_stream rbis_kvh_stream_t;
int rbis_kvh_stream_create(rbis_kvh_stream_t** out);
int rbis_kvh_stream_destroy(rbis_kvh_stream_t* s);
int rbis_kvh_decode_batch(rbis_kvh_stream_t* s, int64_t batch_utime, int32_t num_packets, const rbis_kvh_packet_t* raw,
                          rbis_imu_packet_t* out_new, int32_t* n_new, rbis_imu_packet_t* out_old, int32_t* n_old);
/* InsHandler::processMessageAtlas after the decode (MSE/sensor_handlers.cpp:186-251): the newest new packet of a batch
 * (after the notch cascade, rbis_batch_notch_filter, when the Atlas filter is on) -> gyro = R (delta_rotation / raw_dt),
 * accel = T linear_acceleration (rotation and translation, as the reference applies bot_trans_apply_vec), and the
 * integration dt = default_dt for the first message, else the batch utime difference.  *prev_utime: 0 before the first. */
int rbis_kvh_imu_step(const rbis_imu_packet_t* newest, int64_t batch_utime, const double ins_to_body_quat[4],
                      const double ins_to_body_trans[3], double default_dt, int64_t* prev_utime, double gyro[3], double accel[3],
                      double* dt);

/* Stream-ordered completion tickets: record marks "everything enqueued so far", wait blocks the host
 * until that point has completed.  Up to 8 tickets may be outstanding. */
int rbis_batch_record(rbis_batch_t* h, int32_t* ticket);
int rbis_batch_wait(rbis_batch_t* h, int32_t ticket);
/* Fixed-order final reduction of chunk partials (ascending chunk index): out[RBIS_NUM_STATS]. */
int rbis_stats_reduce_chunks(const double* chunks, int64_t n_chunks, double* out);

/* ---- FP64 roofline denominators measured on the handle's device: register-resident independent
 * DFMA chains, and mma.sync.m8n8k4.f64 (DMMA) for comparison.  TFLOP/s.  Synchronous. */
int rbis_measure_fp64_peak(int device, int iters, double* dfma_tflops, double* dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* RBIS_BATCH_H_ */
