// rbis_smooth.cuh -- EKF (Rauch-Tung-Striebel) smoother backward pass for an ensemble, sm_100a FP64.
//
// Reference semantics restated (paths under /root/reference/state-estimator/src/mav_state_est/):
//   smooth_step()       = ekfSmoothingStep                         rbis.cpp:234-266
//   rbis_smooth_kernel  = the recursion of EKFSmoothBackwardsPass  mav_state_est.cpp:131-187 over a step list that the
//                         host derives from the history (rbis_smooth_plan, rbis_planner.cpp), reading and writing
//                         posteriors in the snapshot ring ([slot][257][N], written by RBIS_OP_SNAPSHOT).
//
// Mapping: ONE WARP PER FILTER, lane j owns column j of every 21x21 matrix (21 of 32 lanes carry data; a dense
// 21x21 LDL^T solve with 21 right-hand sides per step does not fit the lane-per-filter layout of the forward kernel:
// it needs ~1,900 doubles of working set per filter).  Per warp, in shared memory: the factor of the predicted
// covariance S, D = next_cov - next_cov_pred, G = S^-1 Ad cur_cov (= L^T), the carried next_cov / next_cov_pred and
// the four states.  All matrix products are column-local:
//     M[:,j]  = Ad cur_cov[:,j]               (Ad's three non-identity block rows, as in the forward kernel)
//     G[:,j]  = S^-1 M[:,j]                   (every lane substitutes against the shared factor; broadcast reads)
//     w       = D G[:,j]
//     cov'[:,j] = cur_cov[:,j] + G^T w        (= cur_cov + L D L^T, rbis.cpp:256)
//     dx[j]   = G[:,j] . resid                (= (L resid)[j], rbis.cpp:263)
// A CTA holds four warps = four ADJACENT filters, so the 32-byte sectors of the filter-fastest ring layout are shared
// through L1/L2.  Everything is explicit fma (the library is compiled with -fmad=false).
#pragma once
#include "rbis_kernels.cuh"

namespace rbisk {

struct SmoothStep {
  int cur_slot, cur_pred_slot, out_slot, reserved;
};

constexpr int SM_LD = NS;                      // leading dimension of the shared 21x21 matrices
constexpr int SM_MAT = NS * NS;                // 441
constexpr int SM_WARP_DOUBLES = 5 * SM_MAT + 4 * 26 + 2 * NS + 1;  // Sf, Dm, G, Pp, Pn + 4 states + innovation + 1/d = 2,352
constexpr int SMOOTH_WARPS = 4;
constexpr int SMOOTH_SMEM_BYTES = SMOOTH_WARPS * SM_WARP_DOUBLES * 8;  // 75,264 B per CTA, three CTAs per SM

// state record in shared memory: 21 vec, 4 quat (w,x,y,z), 1 loglik
__device__ __forceinline__ void load_state_rec(const double* __restrict__ slot_base, long long N, int lane, double* rec) {
  if (lane < 26) rec[lane] = slot_base[(long long)lane * N];
}

__global__ void __launch_bounds__(SMOOTH_WARPS * 32, 3)
rbis_smooth_kernel(double* __restrict__ snap, long long N, int next_pred_slot, int next_slot, const SmoothStep* __restrict__ steps,
                   long long n_steps, double dt, double g_val, double chi_tol, int ctor_folds_chi) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * SMOOTH_WARPS + warp;
  if (n >= N) return;  // whole warp; only warp-level synchronisation below
  const bool on = lane < NS;
  const int j = on ? lane : NS - 1;  // idle lanes shadow column 20 and never store
  double* Sf = smem + (size_t)warp * SM_WARP_DOUBLES;
  double* Dm = Sf + SM_MAT;
  double* G = Dm + SM_MAT;
  double* Pp = G + SM_MAT;   // next_cov_pred, full symmetric
  double* Pn = Pp + SM_MAT;  // next_cov, full
  double* st_np = Pn + SM_MAT;  // next_state_pred
  double* st_n = st_np + 26;    // next_state
  double* st_c = st_n + 26;     // cur_state
  double* st_cp = st_c + 26;    // cur_state_pred
  double* innov = st_cp + 26;
  double* rdiag = innov + NS;  // 1 / d_k of the factorisation

  auto slot_ptr = [&](int s) { return snap + (long long)s * SNAP_ROWS * N + n; };
  // column j of the symmetric packed covariance of a ring slot
  auto load_col = [&](int s, double (&col)[NS]) {
    const double* b = slot_ptr(s) + 26 * N;
#pragma unroll
    for (int i = 0; i < NS; i++) col[i] = b[(long long)slot(i, j) * N];
  };

  // ---- start of the recursion: next_pred and next from their slots (mav_state_est.cpp:126-129) ----
  {
    double c[NS];
    load_col(next_pred_slot, c);
#pragma unroll
    for (int i = 0; i < NS; i++) Pp[i + SM_LD * j] = c[i];
    load_col(next_slot, c);
#pragma unroll
    for (int i = 0; i < NS; i++) Pn[i + SM_LD * j] = c[i];
    load_state_rec(slot_ptr(next_pred_slot), N, lane, st_np);
    load_state_rec(slot_ptr(next_slot), N, lane, st_n);
  }
  __syncwarp();

  for (long long si = 0; si < n_steps; si++) {
    const SmoothStep stp = steps[si];
    // ---- S = corrected next_cov_pred (rbis.cpp:243-250), D = next_cov - next_cov_pred ----
    bool fix_g = false, fix_a = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      fix_g = fix_g || (Pp[(15 + k) * (SM_LD + 1)] < .00000000001);
      fix_a = fix_a || (Pp[(18 + k) * (SM_LD + 1)] < .00000000001);
    }
    double ccur[NS], ccp[NS];
    load_col(stp.cur_slot, ccur);                       // issued early, consumed after the factorisation
    const bool same = stp.cur_slot == stp.cur_pred_slot;
    if (!same) load_col(stp.cur_pred_slot, ccp);
    load_state_rec(slot_ptr(stp.cur_slot), N, lane, st_c);
    load_state_rec(slot_ptr(stp.cur_pred_slot), N, lane, st_cp);
    double sc[NS];  // column j of S (rows >= j are the ones the factorisation uses)
#pragma unroll
    for (int i = 0; i < NS; i++) {
      const double pp = Pp[i + SM_LD * j];
      Dm[i + SM_LD * j] = Pn[i + SM_LD * j] - pp;
      double s = pp;
      if (fix_g && i >= 15 && i < 18 && j >= 15 && j < 18) s = (i == j) ? 1.0 : 0.0;
      if (fix_a && i >= 18 && j >= 18) s = (i == j) ? 1.0 : 0.0;
      sc[i] = s;
    }
    __syncwarp();
    // next_cov_pred of the NEXT step = this step's cur_cov_pred (mav_state_est.cpp:175)
#pragma unroll
    for (int i = 0; i < NS; i++) Pp[i + SM_LD * j] = same ? ccur[i] : ccp[i];
    // ---- LDL^T of S, no pivoting: lane j keeps column j in registers; at step k lane k publishes its scaled column
    // (unit L below the diagonal of Sf, 1/d_k in rdiag) and the lanes right of it take their rank-1 update ----
    static_for<NS>([&](auto kc) {
      constexpr int k = kc;
      if (lane == k) {
        const double rd = 1.0 / sc[k];
        rdiag[k] = rd;
        Sf[k * (SM_LD + 1)] = sc[k];
#pragma unroll
        for (int i = k + 1; i < NS; i++) Sf[i + SM_LD * k] = sc[i] * rd;
      }
      __syncwarp();
      if constexpr (k + 1 < NS) {
        if (j > k) {
          const double f = Sf[j + SM_LD * k] * Sf[k * (SM_LD + 1)];  // l_jk d_k
#pragma unroll
          for (int i = k + 1; i < NS; i++)
            if (i >= j) sc[i] = fma(-Sf[i + SM_LD * k], f, sc[i]);
        }
      }
    });
    // ---- M[:,j] = Ad cur_cov[:,j], Ad = I + dt Ac at cur_state (rbis.cpp:236-239, 12-35) ----
    double m[NS];
    {
      const V3 w{st_c[0], st_c[1], st_c[2]}, v{st_c[3], st_c[4], st_c[5]};
      const Q4 q{st_c[21], st_c[22], st_c[23], st_c[24]};
      const V3 gb = qrot(qinv(q), V3{0.0, 0.0, -g_val});
      const V3 cv{ccur[3], ccur[4], ccur[5]}, cc{ccur[6], ccur[7], ccur[8]}, cg{ccur[15], ccur[16], ccur[17]},
          ca{ccur[18], ccur[19], ccur[20]};
#pragma unroll
      for (int i = 0; i < NS; i++) m[i] = ccur[i];
      // rows v:  -w x cv + gb x cc - v x cg - ca
      V3 t = sub_cross(V3{-ca.x, -ca.y, -ca.z}, w, cv);
      t = add_cross(t, gb, cc);
      t = sub_cross(t, v, cg);
      m[3] = fma(dt, t.x, ccur[3]); m[4] = fma(dt, t.y, ccur[4]); m[5] = fma(dt, t.z, ccur[5]);
      // rows chi:  -w x cc - cg
      V3 u = sub_cross(V3{-cg.x, -cg.y, -cg.z}, w, cc);
      m[6] = fma(dt, u.x, ccur[6]); m[7] = fma(dt, u.y, ccur[7]); m[8] = fma(dt, u.z, ccur[8]);
      // rows p:  R (cv - v x cc)
      const V3 r = qrot(q, sub_cross(cv, v, cc));
      m[9] = fma(dt, r.x, ccur[9]); m[10] = fma(dt, r.y, ccur[10]); m[11] = fma(dt, r.z, ccur[11]);
    }
    // ---- G[:,j] = S^-1 M[:,j]: forward, diagonal, backward substitution against the shared factor ----
    // column-oriented, so that the 20 updates of a stage are independent; the warp barriers are there for ptxas, which
    // otherwise hoists the loads of the WHOLE factor (210 doubles) to the top and spills them
    static_for<NS - 1>([&](auto kc) {
      constexpr int k = kc;
#pragma unroll
      for (int i = k + 1; i < NS; i++) m[i] = fma(-Sf[i + SM_LD * k], m[k], m[i]);
      __syncwarp();
    });
#pragma unroll
    for (int i = 0; i < NS; i++) m[i] *= rdiag[i];
    static_for<NS - 1>([&](auto kc) {
      constexpr int k = NS - 1 - kc;
#pragma unroll
      for (int i = 0; i < k; i++) m[i] = fma(-Sf[k + SM_LD * i], m[k], m[i]);
      __syncwarp();
    });
    if (on) {
#pragma unroll
      for (int i = 0; i < NS; i++) G[i + SM_LD * j] = m[i];
    }
    // ---- smoothing residual (rbis.cpp:258-261) and innovation (L resid)[j] = G[:,j] . resid ----
    {
      const Q4 qn{st_n[21], st_n[22], st_n[23], st_n[24]}, qp{st_np[21], st_np[22], st_np[23], st_np[24]};
      const V3 dchi = subtract_quats(qmul(qinv(qp), qn), Q4{1.0, 0.0, 0.0, 0.0});  // subtractState then quatToChi
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NS; k++) {
        double rk = st_n[k] - st_np[k];
        if (k == 6) rk = dchi.x;
        if (k == 7) rk = dchi.y;
        if (k == 8) rk = dchi.z;
        acc = fma(m[k], rk, acc);
      }
      if (on) innov[j] = acc;
    }
    __syncwarp();
    // ---- w = D G[:,j];  cov'[:,j] = cur_cov[:,j] + G^T w ----
    double wv[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NS; k++) acc = fma(Dm[i + SM_LD * k], m[k], acc);
      wv[i] = acc;
      if (i % 3 == 2) __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NS; k++) acc = fma(G[k + SM_LD * i], wv[k], acc);
      ccur[i] += acc;
      if (i % 3 == 2) __syncwarp();
    }
    // ---- cur_state.addState(RBIS(L resid))  (rbis.cpp:263-265); every lane carries the same state ----
    FilterState s;
#pragma unroll
    for (int k = 0; k < NS; k++) s.x[k] = st_c[k];
    s.qw = st_c[21]; s.qx = st_c[22]; s.qy = st_c[23]; s.qz = st_c[24];
    s.ll = st_c[25];
    {
#pragma unroll
      for (int k = 0; k < NS; k++)
        if (k < 6 || k > 8) s.x[k] += innov[k];
      const V3 dchi{innov[6], innov[7], innov[8]};
      const double nn = sqrt(sumsq3(dchi.x, dchi.y, dchi.z));
      Q4 dq{1, 0, 0, 0};
      const bool folded = ctor_folds_chi && (nn > chi_tol);
      if (folded) dq = qexp(dchi, nn);
      add_state_tail(s, dchi, folded, dq, chi_tol, 0);
    }
    __syncwarp();  // everyone has read st_c, st_n, st_np, innov, Dm, G
    // ---- roll: next := smoothed cur, next_pred := cur_pred ----
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NS; k++) st_n[k] = s.x[k];
      st_n[21] = s.qw; st_n[22] = s.qx; st_n[23] = s.qy; st_n[24] = s.qz;
      st_n[25] = s.ll;
    }
    if (lane < 26) st_np[lane] = st_cp[lane];
    if (on) {
#pragma unroll
      for (int i = 0; i < NS; i++) Pn[i + SM_LD * j] = ccur[i];
    }
    __syncwarp();
    // ---- store the smoothed posterior into its ring slot (mav_state_est.cpp:160-161) ----
    {
      double* o = slot_ptr(stp.out_slot);
      if (lane < 25) o[(long long)lane * N] = st_n[lane];  // loglik row (25) of the slot stays
      if (on) {
        double* oc = o + 26 * N;
#pragma unroll
        for (int i = 0; i < NS; i++)
          if (i <= j) oc[(long long)slot(i, j) * N] = ccur[i];
      }
    }
    __syncwarp();
  }
}

}  // namespace rbisk
