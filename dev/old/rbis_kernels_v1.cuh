// rbis_kernels.cuh -- sm_100a FP64 device code of the batched RBIS EKF hot path.
//
// Mapping: ONE LANE PER FILTER, 128 filters per CTA, one CTA per SM.  A filter's
// 21x21 covariance is kept symmetric-packed (231 doubles): 225 of them live in shared memory as
// Ps[slot][lane] (conflict-free, 225 KB per CTA), the 6 of the angular-velocity block plus the
// 21+4 state doubles and the log-likelihood live in registers.  State and covariance stay on chip
// for the whole fused program; per-op inputs (IMU rows, measurement rows) are coalesced
// structure-of-arrays loads issued at the top of each op and consumed at its end, so their HBM
// latency hides behind the covariance work of the same op.  No cross-lane communication, no
// barriers: filters are independent.
//
// Reference semantics restated (paths under /root/reference/state-estimator/src/mav_state_est/):
//   cov_propagate()  = insUpdateCovariance + getIMUProcessLinearizationContinuous  rbis.cpp:12-35,77-122
//   state_propagate()= insUpdateState                                               rbis.cpp:37-75
//   meas_chunk<M>()  = matrixMeasurementGetKandCovDelta + indexed[PlusOrientation]Measurement
//                      + the covariance half of rbisApplyDelta                      rbis.cpp:124-227
//   meas_finish()    = the state half of rbisApplyDelta (addState)                  rbis.cpp:219-227
#pragma once
#include <cstdint>
#include <utility>

// Tuning knobs (dev/kbench.cu compiles several settings side by side).
#ifndef RBIS_FENCE
#define RBIS_FENCE 1   // 1: compiler memory fences between column groups (bounds load hoisting -> no spills)
#endif
#if RBIS_FENCE
#define RBIS_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define RBIS_SCHED_FENCE() do {} while (0)
#endif

namespace rbisk {

constexpr int NS = 21;       // rbis_num_states
constexpr int NP = 231;      // packed upper triangle
constexpr int NPW = 6;       // slots 0..5 (angular-velocity block) are register resident
constexpr int TPB = 128;     // filters (= threads) per CTA
constexpr int SMEM_BYTES = (NP - NPW) * TPB * 8;
constexpr int MAX_MEAS = 9;
constexpr int MAX_STREAMS = 8;
constexpr int MAX_CHUNKS = 9;

__host__ __device__ constexpr int slot(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }

struct StreamDesc {
  int m, has_orient, r_mode, n_chunks;
  int idx[MAX_MEAS];
  int chunk_start[MAX_CHUNKS];
  int chunk_len[MAX_CHUNKS];
  const double* z;     // [rows][m][N]
  const double* quat;  // [rows][4][N]
  const double* R;     // r_mode 0: device copy of m*m column-major; 1: [m][N]
};

struct Op {
  int kind, stream;
  long long row;
  double dt;
};

struct KParams {
  long long N;
  double* vec;     // [21][N]
  double* quat;    // [4][N]
  double* P;       // [231][N]
  double* loglik;  // [N]
  const double* q_gyro;  // [N] each
  const double* q_accel;
  const double* q_gyro_bias;
  const double* q_accel_bias;
  const double* imu;  // [rows][6][N]
  double* snap;       // [slots][257][N]
  const Op* ops;
  long long n_ops;
  double g_val, chi_tol;
  int ctor_folds_chi, renorm, n_snap;
  StreamDesc streams[MAX_STREAMS];
};

// ------------------------------------------------------------------------------------------------
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Per-lane view of the covariance: compile-time (i,j) accessors resolve to an immediate shared
// memory offset or to one of the six register-resident slots.
struct Cov {
  double* Ps;       // shared base + lane; slot s (>= NPW) at Ps[(s - NPW) * TPB]
  double pw[NPW];
  template <int I, int J>
  __device__ __forceinline__ double get() const {
    constexpr int s = slot(I, J);
    if constexpr (s < NPW) return pw[s];
    else return Ps[(s - NPW) * TPB];
  }
  template <int I, int J>
  __device__ __forceinline__ void set(double v) {
    constexpr int s = slot(I, J);
    if constexpr (s < NPW) pw[s] = v;
    else Ps[(s - NPW) * TPB] = v;
  }
  template <int S>
  __device__ __forceinline__ double gets() const {
    if constexpr (S < NPW) return pw[S];
    else return Ps[(S - NPW) * TPB];
  }
  template <int S>
  __device__ __forceinline__ void sets(double v) {
    if constexpr (S < NPW) pw[S] = v;
    else Ps[(S - NPW) * TPB] = v;
  }
  // runtime slot (slow paths only)
  __device__ __forceinline__ double getr(int s) const {
    if (s >= NPW) return Ps[(s - NPW) * TPB];
    double v = pw[0];
#pragma unroll
    for (int k = 1; k < NPW; k++) v = (s == k) ? pw[k] : v;
    return v;
  }
  __device__ __forceinline__ void setr(int s, double v) {
    if (s >= NPW) { Ps[(s - NPW) * TPB] = v; return; }
#pragma unroll
    for (int k = 0; k < NPW; k++) pw[k] = (s == k) ? v : pw[k];
  }
  // column C (compile time), runtime row r: element (r, C)
  template <int C>
  __device__ __forceinline__ double getrc(int r) const {
    const int s = (r <= C) ? (C * (C + 1) / 2 + r) : (r * (r + 1) / 2 + C);
    if constexpr (C * (C + 1) / 2 < NPW) return getr(s);  // only columns 0..2 can hit a register slot
    else return Ps[(s - NPW) * TPB];
  }
  template <int R0, int C>
  __device__ __forceinline__ V3 col3() const {
    return {get<R0, C>(), get<R0 + 1, C>(), get<R0 + 2, C>()};
  }
  template <int R0, int C>
  __device__ __forceinline__ void setcol3(const V3& v) {
    set<R0, C>(v.x); set<R0 + 1, C>(v.y); set<R0 + 2, C>(v.z);
  }
};

struct FilterState {
  double x[NS];
  double qw, qx, qy, qz;
  double ll;
};

// ---- quaternion helpers (Eigen semantics) ----
struct Q4 {
  double w, x, y, z;
};
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ Q4 qinv(const Q4& q) {
  const double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (n2 > 0) {
    const double r = 1.0 / n2;  // Eigen divides each coefficient; 1 ulp apart at most
    return {q.w * r, -q.x * r, -q.y * r, -q.z * r};
  }
  return {0, 0, 0, 0};
}
__device__ __forceinline__ V3 qrot(const Q4& q, const V3& v) {
  V3 u{q.x, q.y, q.z};
  V3 uv = cross(u, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  V3 c = cross(u, uv);
  return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
// quaternion of AngleAxis(|chi|, chi/|chi|)
__device__ __forceinline__ Q4 qexp(const V3& chi, double n) {
  double s, c;
  sincos(0.5 * n, &s, &c);
  return {c, s * (chi.x / n), s * (chi.y / n), s * (chi.z / n)};
}
// subtractQuats(q1, q2) = axis*angle of q2^-1 * q1 (Eigen >= 3.3 AngleAxis, bot_mod2pi)
__device__ __forceinline__ V3 subtract_quats(const Q4& q1, const Q4& q2) {
  const Q4 r = qmul(qinv(q2), q1);
  double n = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
  if (n == 0.0) return {0, 0, 0};
  double angle = 2.0 * atan2(n, fabs(r.w));
  if (r.w < 0) n = -n;
  const double PI = 3.14159265358979323846;
  if (angle >= PI) angle -= 2 * PI;  // bot_mod2pi maps exactly pi to -pi, identity below
  return {r.x / n * angle, r.y / n * angle, r.z / n * angle};
}

// ------------------------------------------------------------------------------------------------
// Covariance propagation.  Ad = I + dt*Ac has non-identity block rows v, chi, p only, and because
// the p-columns of Ac are zero and chi's row has no v-column, Ad factors EXACTLY as
//     Ad = E_chi * E_v * E_p,   E_X = I + (block row X of dt*Ac),
// so  Ad P Ad^T = E_chi (E_v (E_p P E_p^T) E_v^T) E_chi^T : three in-place symmetric congruences,
// each touching one block row/column.  For E = I + N (N supported on block row I):
//     z_c   = P[I,c] + N[I,:] P[:,c]            for every column c
//     P'[I,c] = z_c (c outside I),   P'[I,I] = Z_I + sum_K Z_K N[I,K]^T
// Products with the structural zeros/ones of Ad are skipped; everything else is the same
// arithmetic as the dense product up to summation order.
// ------------------------------------------------------------------------------------------------
struct Lin {  // linearisation point quantities, pre-scaled by dt where the reference scales Ac by dt
  V3 v;        // body velocity (unscaled)
  V3 wd;       // omega * dt
  V3 vd;       // v * dt
  V3 gd;       // (q^-1 g_vec) * dt
  double Rd[9];  // R * dt, row-major Rd[3*r + c]
  double dt;
};

// (Z * skew(u))[:, j] for Z given as three columns
__device__ __forceinline__ void mul_skew(const V3 Z[3], const V3& u, V3 out[3]) {
  out[0] = {u.z * Z[1].x - u.y * Z[2].x, u.z * Z[1].y - u.y * Z[2].y, u.z * Z[1].z - u.y * Z[2].z};
  out[1] = {u.x * Z[2].x - u.z * Z[0].x, u.x * Z[2].y - u.z * Z[0].y, u.x * Z[2].z - u.z * Z[0].z};
  out[2] = {u.y * Z[0].x - u.x * Z[1].x, u.y * Z[0].y - u.x * Z[1].y, u.y * Z[0].z - u.x * Z[1].z};
}

template <int C>
__device__ __forceinline__ V3 ep_z(const Cov& P, const Lin& L) {
  // rows p:  z = P[p,c] + R dt (P[v,c] - v x P[chi,c])
  const V3 pv = P.col3<3, C>(), pc = P.col3<6, C>(), pp = P.col3<9, C>();
  const V3 k = cross(L.v, pc);
  const V3 t{pv.x - k.x, pv.y - k.y, pv.z - k.z};
  return {pp.x + L.Rd[0] * t.x + L.Rd[1] * t.y + L.Rd[2] * t.z, pp.y + L.Rd[3] * t.x + L.Rd[4] * t.y + L.Rd[5] * t.z,
          pp.z + L.Rd[6] * t.x + L.Rd[7] * t.y + L.Rd[8] * t.z};
}
template <int C>
__device__ __forceinline__ V3 ev_z(const Cov& P, const Lin& L) {
  // rows v:  z = P[v,c] - wd x P[v,c] + gd x P[chi,c] - vd x P[bg,c] - dt P[ba,c]
  const V3 pv = P.col3<3, C>(), pc = P.col3<6, C>(), pg = P.col3<15, C>(), pa = P.col3<18, C>();
  const V3 a = cross(L.wd, pv), b = cross(L.gd, pc), c = cross(L.vd, pg);
  return {pv.x - a.x + b.x - c.x - L.dt * pa.x, pv.y - a.y + b.y - c.y - L.dt * pa.y,
          pv.z - a.z + b.z - c.z - L.dt * pa.z};
}
template <int C>
__device__ __forceinline__ V3 ec_z(const Cov& P, const Lin& L) {
  // rows chi:  z = P[chi,c] - wd x P[chi,c] - dt P[bg,c]
  const V3 pc = P.col3<6, C>(), pg = P.col3<15, C>();
  const V3 a = cross(L.wd, pc);
  return {pc.x - a.x - L.dt * pg.x, pc.y - a.y - L.dt * pg.y, pc.z - a.z - L.dt * pg.z};
}

__device__ __forceinline__ void cov_propagate(Cov& P, const Lin& L, double q_gyro, double q_accel,
                                              double q_gyro_bias, double q_accel_bias) {
  const double dt = L.dt;
  // ---------------- E_p : block row p (9..11), sources v (3..5), chi (6..8) ----------------
  {
    V3 Zv[3], Zc[3], Zp[3];
    static_for<3>([&](auto k) { Zv[k] = ep_z<3 + k>(P, L); }); RBIS_SCHED_FENCE();
    static_for<3>([&](auto k) { Zc[k] = ep_z<6 + k>(P, L); });
    static_for<3>([&](auto k) { Zp[k] = ep_z<9 + k>(P, L); });
    // T = Zv + Zc * skew(v);  P'[p,p] = Zp + T (R dt)^T
    V3 S[3];
    mul_skew(Zc, L.v, S);
    V3 T[3];
#pragma unroll
    for (int k = 0; k < 3; k++) T[k] = {Zv[k].x + S[k].x, Zv[k].y + S[k].y, Zv[k].z + S[k].z};
    // (T Rd^T)[i][j] = sum_k T[k].i * Rd[3j+k]
    const double t0[3] = {T[0].x, T[1].x, T[2].x}, t1[3] = {T[0].y, T[1].y, T[2].y}, t2[3] = {T[0].z, T[1].z, T[2].z};
    auto dotR = [&](const double t[3], int j) { return t[0] * L.Rd[3 * j] + t[1] * L.Rd[3 * j + 1] + t[2] * L.Rd[3 * j + 2]; };
    P.set<9, 9>(Zp[0].x + dotR(t0, 0));
    P.set<9, 10>(Zp[1].x + dotR(t0, 1));
    P.set<9, 11>(Zp[2].x + dotR(t0, 2));
    P.set<10, 10>(Zp[1].y + dotR(t1, 1));
    P.set<10, 11>(Zp[2].y + dotR(t1, 2));
    P.set<11, 11>(Zp[2].z + dotR(t2, 2));
    static_for<3>([&](auto k) { P.setcol3<9, 3 + k>(Zv[k]); });
    static_for<3>([&](auto k) { P.setcol3<9, 6 + k>(Zc[k]); });
    // remaining columns: omega (0..2), a (12..14), bg (15..17), ba (18..20)
    static_for<3>([&](auto k) { P.setcol3<9, 0 + k>(ep_z<0 + k>(P, L)); RBIS_SCHED_FENCE(); });
    static_for<9>([&](auto k) { P.setcol3<9, 12 + k>(ep_z<12 + k>(P, L)); RBIS_SCHED_FENCE(); });
  }
  RBIS_SCHED_FENCE();
  // ---------------- E_v : block row v (3..5), sources v, chi, bg (15..17), ba (18..20) ----------------
  {
    V3 Zv[3], Zc[3], Zg[3], Za[3];
    static_for<3>([&](auto k) { Zv[k] = ev_z<3 + k>(P, L); });
    static_for<3>([&](auto k) { Zc[k] = ev_z<6 + k>(P, L); });
    static_for<3>([&](auto k) { Zg[k] = ev_z<15 + k>(P, L); });
    static_for<3>([&](auto k) { Za[k] = ev_z<18 + k>(P, L); });
    // P'[v,v] = Zv + Zv skew(wd) - Zc skew(gd) + Zg skew(vd) - dt Za    (+ Qd[v,v], rbis.cpp:116)
    V3 A[3], B[3], C[3];
    mul_skew(Zv, L.wd, A);
    mul_skew(Zc, L.gd, B);
    mul_skew(Zg, L.vd, C);
    V3 N[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
      N[k] = {Zv[k].x + A[k].x - B[k].x + C[k].x - dt * Za[k].x, Zv[k].y + A[k].y - B[k].y + C[k].y - dt * Za[k].y,
              Zv[k].z + A[k].z - B[k].z + C[k].z - dt * Za[k].z};
    // Qd[v,v] = dt (q_gyro (|v|^2 I - v v^T) + q_accel I)
    const double vv = L.v.x * L.v.x + L.v.y * L.v.y + L.v.z * L.v.z;
    const double qg = q_gyro * dt, qa = q_accel * dt;
    P.set<3, 3>(N[0].x + (qg * (vv - L.v.x * L.v.x) + qa));
    P.set<3, 4>(N[1].x + (qg * (-L.v.x * L.v.y)));
    P.set<3, 5>(N[2].x + (qg * (-L.v.x * L.v.z)));
    P.set<4, 4>(N[1].y + (qg * (vv - L.v.y * L.v.y) + qa));
    P.set<4, 5>(N[2].y + (qg * (-L.v.y * L.v.z)));
    P.set<5, 5>(N[2].z + (qg * (vv - L.v.z * L.v.z) + qa));
    static_for<3>([&](auto k) { P.setcol3<3, 6 + k>(Zc[k]); });
    static_for<3>([&](auto k) { P.setcol3<3, 15 + k>(Zg[k]); });
    static_for<3>([&](auto k) { P.setcol3<3, 18 + k>(Za[k]); });
    // remaining columns: omega, p (9..11), a (12..14)
    static_for<3>([&](auto k) { P.setcol3<3, 0 + k>(ev_z<0 + k>(P, L)); RBIS_SCHED_FENCE(); });
    static_for<6>([&](auto k) { P.setcol3<3, 9 + k>(ev_z<9 + k>(P, L)); RBIS_SCHED_FENCE(); });
  }
  RBIS_SCHED_FENCE();
  // ---------------- E_chi : block row chi (6..8), sources chi, bg ----------------
  {
    V3 Zc[3], Zg[3], Zv[3];
    static_for<3>([&](auto k) { Zc[k] = ec_z<6 + k>(P, L); });
    static_for<3>([&](auto k) { Zg[k] = ec_z<15 + k>(P, L); });
    static_for<3>([&](auto k) { Zv[k] = ec_z<3 + k>(P, L); });  // not a source; done here to fuse Qd[chi,v]
    V3 A[3];
    mul_skew(Zc, L.wd, A);
    const double qg = q_gyro * dt;
    // P'[chi,chi] = Zc + Zc skew(wd) - dt Zg   (+ Qd[chi,chi] = dt q_gyro I)
    P.set<6, 6>(Zc[0].x + A[0].x - dt * Zg[0].x + qg);
    P.set<6, 7>(Zc[1].x + A[1].x - dt * Zg[1].x);
    P.set<6, 8>(Zc[2].x + A[2].x - dt * Zg[2].x);
    P.set<7, 7>(Zc[1].y + A[1].y - dt * Zg[1].y + qg);
    P.set<7, 8>(Zc[2].y + A[2].y - dt * Zg[2].y);
    P.set<8, 8>(Zc[2].z + A[2].z - dt * Zg[2].z + qg);
    static_for<3>([&](auto k) { P.setcol3<6, 15 + k>(Zg[k]); });
    // columns v: P'[chi, v_k] = z + Qd[chi, v_k];  Qd[v,chi] = dt q_gyro skew(v)  =>  Qd[chi_i, v_k] = qg*skew(v)[k][i]
    // skew(v) = [[0,-vz,vy],[vz,0,-vx],[-vy,vx,0]]
    P.setcol3<6, 3>({Zv[0].x, Zv[0].y + qg * (-L.v.z), Zv[0].z + qg * (L.v.y)});
    P.setcol3<6, 4>({Zv[1].x + qg * (L.v.z), Zv[1].y, Zv[1].z + qg * (-L.v.x)});
    P.setcol3<6, 5>({Zv[2].x + qg * (-L.v.y), Zv[2].y + qg * (L.v.x), Zv[2].z});
    // remaining columns: omega, p, a, ba
    static_for<3>([&](auto k) { P.setcol3<6, 0 + k>(ec_z<0 + k>(P, L)); RBIS_SCHED_FENCE(); });
    static_for<6>([&](auto k) { P.setcol3<6, 9 + k>(ec_z<9 + k>(P, L)); RBIS_SCHED_FENCE(); });
    static_for<3>([&](auto k) { P.setcol3<6, 18 + k>(ec_z<18 + k>(P, L)); RBIS_SCHED_FENCE(); });
  }
  // ---------------- rest of Qd and the overwrites, rbis.cpp:116,120-121 ----------------
  {
    const double qgb = q_gyro_bias * dt, qab = q_accel_bias * dt;
    static_for<3>([&](auto k) { P.set<15 + k, 15 + k>(P.get<15 + k, 15 + k>() + qgb); });
    static_for<3>([&](auto k) { P.set<18 + k, 18 + k>(P.get<18 + k, 18 + k>() + qab); });
    P.set<12, 12>(q_accel); P.set<13, 13>(q_accel); P.set<14, 14>(q_accel);
    P.set<12, 13>(0.0); P.set<12, 14>(0.0); P.set<13, 14>(0.0);
    P.set<0, 0>(q_gyro); P.set<1, 1>(q_gyro); P.set<2, 2>(q_gyro);
    P.set<0, 1>(0.0); P.set<0, 2>(0.0); P.set<1, 2>(0.0);
  }
}

// chiToQuat on the filter state: fold vec chi into the quaternion when its norm exceeds the tolerance
__device__ __forceinline__ void fold_chi(FilterState& s, double chi_tol) {
  const V3 c{s.x[6], s.x[7], s.x[8]};
  const double n = sqrt(c.x * c.x + c.y * c.y + c.z * c.z);
  if (n > chi_tol) {
    const Q4 q = qmul({s.qw, s.qx, s.qy, s.qz}, qexp(c, n));
    s.qw = q.w; s.qx = q.x; s.qy = q.y; s.qz = q.z;
    s.x[6] = 0; s.x[7] = 0; s.x[8] = 0;
  }
}

// addState(dstate) where dstate.vec = d (already added to s.x EXCEPT chi, passed separately) and
// dstate.quat = dq:  vec += d; chiToQuat(); quat *= dq
__device__ __forceinline__ void add_state_tail(FilterState& s, const V3& dchi, bool d_folded, const Q4& dq,
                                               double chi_tol, int renorm) {
  if (!d_folded) { s.x[6] += dchi.x; s.x[7] += dchi.y; s.x[8] += dchi.z; }
  fold_chi(s, chi_tol);
  if (d_folded) {
    const Q4 q = qmul({s.qw, s.qx, s.qy, s.qz}, dq);
    s.qw = q.w; s.qx = q.x; s.qy = q.y; s.qz = q.z;
  }
  if (renorm) {
    const double r = 1.0 / sqrt(s.qw * s.qw + s.qx * s.qx + s.qy * s.qy + s.qz * s.qz);
    s.qw *= r; s.qx *= r; s.qy *= r; s.qz *= r;
  }
}

// insUpdateState, rbis.cpp:37-75.  gb = q^-1 g_vec at the prior quaternion.
__device__ __forceinline__ void state_propagate(FilterState& s, const V3& gyro, const V3& acc, double dt, const V3& gb,
                                                double chi_tol, int renorm) {
  const V3 w{gyro.x - s.x[15], gyro.y - s.x[16], gyro.z - s.x[17]};
  const V3 a{acc.x - s.x[18], acc.y - s.x[19], acc.z - s.x[20]};
  s.x[0] = w.x; s.x[1] = w.y; s.x[2] = w.z;
  s.x[12] = a.x; s.x[13] = a.y; s.x[14] = a.z;
  const V3 v{s.x[3], s.x[4], s.x[5]};
  const V3 wxv = cross(w, v);
  const V3 dv{(-wxv.x + (gb.x + a.x)) * dt, (-wxv.y + (gb.y + a.y)) * dt, (-wxv.z + (gb.z + a.z)) * dt};
  V3 dchi{w.x * dt, w.y * dt, w.z * dt};
  const V3 rv = qrot({s.qw, s.qx, s.qy, s.qz}, v);
  const V3 dp{rv.x * dt, rv.y * dt, rv.z * dt};
  // dstate.chiToQuat()
  const double n = sqrt(dchi.x * dchi.x + dchi.y * dchi.y + dchi.z * dchi.z);
  Q4 dq{1, 0, 0, 0};
  const bool folded = n > chi_tol;
  if (folded) dq = qexp(dchi, n);
  // addState
  s.x[3] += dv.x; s.x[4] += dv.y; s.x[5] += dv.z;
  s.x[9] += dp.x; s.x[10] += dp.y; s.x[11] += dp.z;
  add_state_tail(s, dchi, folded, dq, chi_tol, renorm);
}

// x[idx] with a runtime (warp-uniform) index: select chain, registers cannot be indexed dynamically
__device__ __forceinline__ double pick_state(const double (&x)[NS], int idx) {
  double v = x[0];
#pragma unroll
  for (int k = 1; k < NS; k++) v = (idx == k) ? x[k] : v;
  return v;
}

// ------------------------------------------------------------------------------------------------
// One chunk of M (<= 3 fast, <= 9 general) measurement rows a0..a0+M-1 of a stream, processed as a
// standard EKF update on the CURRENT covariance:  S = R + P[idx,idx];  K^T = S^-1 P[idx,:];
// P -= P[:,idx] S^-1 P[idx,:];  x += K r;  loglik += -log det S - r^T S^-1 r   (rbis.cpp:134-142).
// The host splits a stream into chunks only along blocks where R is block diagonal, for which
// processing the chunks in sequence is algebraically identical to the reference's single batch
// update (chain rule of the Gaussian likelihood); residuals of later chunks are taken against the
// already-updated vec, chi residuals against the accumulated chi delta.
// ------------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void meas_chunk(Cov& P, FilterState& s, const StreamDesc& st, int a0, long long row,
                                           long long N, long long n, const V3& dquat, const V3& chi0) {
  // issue the measurement loads first; they are consumed after the covariance work
  double z[M], Rd[M];
#pragma unroll
  for (int a = 0; a < M; a++) z[a] = __ldg(st.z + (row * st.m + (a0 + a)) * N + n);
  if (st.r_mode == 1) {
#pragma unroll
    for (int a = 0; a < M; a++) Rd[a] = __ldg(st.R + (long long)(a0 + a) * N + n);
  }
  int idx[M];
#pragma unroll
  for (int a = 0; a < M; a++) idx[a] = st.idx[a0 + a];

  // HP = P[idx, :]
  double HP[M][NS];
#pragma unroll
  for (int a = 0; a < M; a++) static_for<NS>([&](auto c) { HP[a][c] = P.template getrc<c>(idx[a]); });

  // S (symmetric, full storage) = R + P[idx, idx]
  double S[M][M];
#pragma unroll
  for (int a = 0; a < M; a++)
#pragma unroll
    for (int b = a; b < M; b++) {
      const int i = idx[a] < idx[b] ? idx[a] : idx[b], j = idx[a] < idx[b] ? idx[b] : idx[a];
      double r;
      if (st.r_mode == 1) r = (a == b) ? Rd[a] : 0.0;
      else r = __ldg(st.R + (a0 + a) + (long long)st.m * (a0 + b));
      S[a][b] = r + P.getr(j * (j + 1) / 2 + i);
      S[b][a] = S[a][b];
    }
  // LDL^T (no pivoting; S is SPD), Sinv = L^-T D^-1 L^-1, logdet = sum log d
  double Lm[M][M], D[M];
  double logdet = 0;
#pragma unroll
  for (int k = 0; k < M; k++) {
    double d = S[k][k];
#pragma unroll
    for (int p = 0; p < k; p++) d -= Lm[k][p] * Lm[k][p] * D[p];
    D[k] = d;
    logdet += log(d);
    const double rd = 1.0 / d;
#pragma unroll
    for (int i = k + 1; i < M; i++) {
      double v = S[i][k];
#pragma unroll
      for (int p = 0; p < k; p++) v -= Lm[i][p] * Lm[k][p] * D[p];
      Lm[i][k] = v * rd;
    }
  }
  // Linv (unit lower): Li[i][j], i > j
  double Li[M][M];
#pragma unroll
  for (int j = 0; j < M; j++) {
    Li[j][j] = 1.0;
#pragma unroll
    for (int i = j + 1; i < M; i++) {
      double v = -Lm[i][j];
#pragma unroll
      for (int k = j + 1; k < i; k++) v -= Lm[i][k] * Li[k][j];
      Li[i][j] = v;
    }
  }
  double Sinv[M][M];
#pragma unroll
  for (int a = 0; a < M; a++)
#pragma unroll
    for (int b = a; b < M; b++) {
      double v = 0;
#pragma unroll
      for (int k = b; k < M; k++) v += Li[k][a] * Li[k][b] / D[k];
      Sinv[a][b] = v;
      Sinv[b][a] = v;
    }

  // covariance: for each column j, g_j = Sinv HP[:,j] (= row j of K); P[i,j] -= HP[:,i] . g_j, i <= j
  static_for<NS>([&](auto jc) {
    constexpr int j = jc;
    double g[M];
#pragma unroll
    for (int a = 0; a < M; a++) {
      double v = 0;
#pragma unroll
      for (int b = 0; b < M; b++) v += Sinv[a][b] * HP[b][j];
      g[a] = v;
    }
    static_for<j + 1>([&](auto ic) {
      constexpr int i = ic;
      double acc = P.template get<i, j>();
#pragma unroll
      for (int a = 0; a < M; a++) acc -= HP[a][i] * g[a];
      P.template set<i, j>(acc);
    });
    RBIS_SCHED_FENCE();
  });

  // residual against the current vec (prior + earlier chunks of this update)
  double r[M];
#pragma unroll
  for (int a = 0; a < M; a++) {
    const int k = idx[a] - 6;
    if (st.has_orient && k >= 0 && k <= 2) {
      const double dq = (k == 0) ? dquat.x : (k == 1) ? dquat.y : dquat.z;
      const double c0 = (k == 0) ? chi0.x : (k == 1) ? chi0.y : chi0.z;
      r[a] = dq - (pick_state(s.x, idx[a]) - c0);
    } else {
      r[a] = z[a] - pick_state(s.x, idx[a]);
    }
  }
  double y[M];
  double quad = 0;
#pragma unroll
  for (int a = 0; a < M; a++) {
    double v = 0;
#pragma unroll
    for (int b = 0; b < M; b++) v += Sinv[a][b] * r[b];
    y[a] = v;
    quad += r[a] * v;
  }
  // x += K r = HP^T (Sinv r)
#pragma unroll
  for (int c = 0; c < NS; c++) {
    double v = 0;
#pragma unroll
    for (int a = 0; a < M; a++) v += HP[a][c] * y[a];
    s.x[c] += v;
  }
  s.ll += -logdet - quad;
}

// General chunk (M = 4..9): same mathematics, compact loops, HP in local memory.  Only reached for
// measurement covariances that are not block diagonal in blocks of <= 3.
__device__ __noinline__ void meas_chunk_general(int M, Cov& P, FilterState& s, const StreamDesc& st, int a0,
                                                long long row, long long N, long long n, const V3& dquat,
                                                const V3& chi0) {
  double HP[MAX_MEAS][NS], S[MAX_MEAS][MAX_MEAS], Lm[MAX_MEAS][MAX_MEAS], D[MAX_MEAS], r[MAX_MEAS], y[MAX_MEAS];
  int idx[MAX_MEAS];
  for (int a = 0; a < M; a++) idx[a] = st.idx[a0 + a];
  for (int a = 0; a < M; a++)
    for (int c = 0; c < NS; c++) HP[a][c] = P.getr(slot(idx[a], c));
  for (int a = 0; a < M; a++)
    for (int b = 0; b < M; b++) {
      double rr;
      if (st.r_mode == 1) rr = (a == b) ? __ldg(st.R + (long long)(a0 + a) * N + n) : 0.0;
      else rr = __ldg(st.R + (a0 + a) + (long long)st.m * (a0 + b));
      S[a][b] = rr + P.getr(slot(idx[a], idx[b]));
    }
  double logdet = 0;
  for (int k = 0; k < M; k++) {
    double d = S[k][k];
    for (int p = 0; p < k; p++) d -= Lm[k][p] * Lm[k][p] * D[p];
    D[k] = d;
    logdet += log(d);
    for (int i = k + 1; i < M; i++) {
      double v = S[i][k];
      for (int p = 0; p < k; p++) v -= Lm[i][p] * Lm[k][p] * D[p];
      Lm[i][k] = v / d;
    }
  }
  // G = S^-1 HP by forward / diagonal / backward substitution, column by column (in place in W)
  double G[MAX_MEAS][NS];
  for (int c = 0; c < NS; c++) {
    double w[MAX_MEAS];
    for (int i = 0; i < M; i++) {
      double v = HP[i][c];
      for (int k = 0; k < i; k++) v -= Lm[i][k] * w[k];
      w[i] = v;
    }
    for (int i = 0; i < M; i++) w[i] /= D[i];
    for (int i = M - 1; i >= 0; i--) {
      double v = w[i];
      for (int k = i + 1; k < M; k++) v -= Lm[k][i] * w[k];
      w[i] = v;
    }
    for (int i = 0; i < M; i++) G[i][c] = w[i];
  }
  for (int j = 0; j < NS; j++)
    for (int i = 0; i <= j; i++) {
      const int sl = j * (j + 1) / 2 + i;
      double acc = P.getr(sl);
      for (int a = 0; a < M; a++) acc -= HP[a][i] * G[a][j];
      P.setr(sl, acc);
    }
  for (int a = 0; a < M; a++) {
    const int k = idx[a] - 6;
    const double xi = pick_state(s.x, idx[a]);
    if (st.has_orient && k >= 0 && k <= 2) {
      const double dq = (k == 0) ? dquat.x : (k == 1) ? dquat.y : dquat.z;
      const double c0 = (k == 0) ? chi0.x : (k == 1) ? chi0.y : chi0.z;
      r[a] = dq - (xi - c0);
    } else {
      r[a] = __ldg(st.z + (row * st.m + (a0 + a)) * N + n) - xi;
    }
  }
  // y = S^-1 r
  for (int i = 0; i < M; i++) {
    double v = r[i];
    for (int k = 0; k < i; k++) v -= Lm[i][k] * y[k];
    y[i] = v;
  }
  for (int i = 0; i < M; i++) y[i] /= D[i];
  for (int i = M - 1; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < M; k++) v -= Lm[k][i] * y[k];
    y[i] = v;
  }
  double quad = 0;
  for (int a = 0; a < M; a++) quad += r[a] * y[a];
#pragma unroll
  for (int c = 0; c < NS; c++) {
    double v = 0;
    for (int a = 0; a < M; a++) v += HP[a][c] * y[a];
    s.x[c] += v;
  }
  s.ll += -logdet - quad;
}

// state half of rbisApplyDelta for a whole measurement op: dstate = RBIS(K r) then addState.
// s.x already holds prior + K r (all chunks); chi0 is the prior vec chi.
__device__ __forceinline__ void meas_finish(FilterState& s, const V3& chi0, double chi_tol, int ctor_folds_chi,
                                            int renorm) {
  V3 dchi{s.x[6] - chi0.x, s.x[7] - chi0.y, s.x[8] - chi0.z};
  s.x[6] = chi0.x; s.x[7] = chi0.y; s.x[8] = chi0.z;
  const double n = sqrt(dchi.x * dchi.x + dchi.y * dchi.y + dchi.z * dchi.z);
  Q4 dq{1, 0, 0, 0};
  const bool folded = ctor_folds_chi && (n > chi_tol);
  if (folded) dq = qexp(dchi, n);
  add_state_tail(s, dchi, folded, dq, chi_tol, renorm);
}

// ------------------------------------------------------------------------------------------------
// The fused kernel: every lane loads its filter, runs the whole op program, stores it back.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB, 1) rbis_fused_kernel(const __grid_constant__ KParams p) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const long long N = p.N;
  long long n = (long long)blockIdx.x * TPB + tid;
  const bool active = n < N;
  if (!active) n = N - 1;  // idle lanes shadow the last filter and never store

  Cov P;
  P.Ps = smem + tid;
  FilterState s;
  static_for<NS>([&](auto i) { s.x[i] = p.vec[(long long)i * N + n]; });
  s.qw = p.quat[n]; s.qx = p.quat[N + n]; s.qy = p.quat[2 * N + n]; s.qz = p.quat[3 * N + n];
  s.ll = p.loglik[n];
  static_for<NP>([&](auto e) { P.template sets<e>(p.P[(long long)e * N + n]); });
  const double q_gyro = p.q_gyro[n], q_accel = p.q_accel[n], q_gyro_bias = p.q_gyro_bias[n],
               q_accel_bias = p.q_accel_bias[n];

  for (long long oi = 0; oi < p.n_ops; oi++) {
    const Op op = p.ops[oi];
    if (op.kind == 0) {
      // ---- IMU process step ----
      const double* base = p.imu + op.row * 6 * N + n;
      const V3 gyro{__ldg(base), __ldg(base + N), __ldg(base + 2 * N)};
      const V3 acc{__ldg(base + 3 * N), __ldg(base + 4 * N), __ldg(base + 5 * N)};
      const double dt = op.dt;
      const Q4 q{s.qw, s.qx, s.qy, s.qz};
      const V3 gb = qrot(qinv(q), V3{0.0, 0.0, -p.g_val});
      Lin L;
      L.v = {s.x[3], s.x[4], s.x[5]};
      L.wd = {s.x[0] * dt, s.x[1] * dt, s.x[2] * dt};  // omega of the PRIOR state (previous sample)
      L.vd = {L.v.x * dt, L.v.y * dt, L.v.z * dt};
      L.gd = {gb.x * dt, gb.y * dt, gb.z * dt};
      L.dt = dt;
      {
        const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
        const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
        const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
        const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
        L.Rd[0] = (1 - (tyy + tzz)) * dt; L.Rd[1] = (txy - twz) * dt;       L.Rd[2] = (txz + twy) * dt;
        L.Rd[3] = (txy + twz) * dt;       L.Rd[4] = (1 - (txx + tzz)) * dt; L.Rd[5] = (tyz - twx) * dt;
        L.Rd[6] = (txz - twy) * dt;       L.Rd[7] = (tyz + twx) * dt;       L.Rd[8] = (1 - (txx + tyy)) * dt;
      }
      cov_propagate(P, L, q_gyro, q_accel, q_gyro_bias, q_accel_bias);
      state_propagate(s, gyro, acc, dt, gb, p.chi_tol, p.renorm);
    } else if (op.kind == 1) {
      // ---- indexed / indexed-plus-orientation measurement ----
      const StreamDesc& st = p.streams[op.stream];
      V3 dquat{0, 0, 0};
      if (st.has_orient) {
        const double* qb = st.quat + op.row * 4 * N + n;
        const Q4 mq{__ldg(qb), __ldg(qb + N), __ldg(qb + 2 * N), __ldg(qb + 3 * N)};
        dquat = subtract_quats(mq, {s.qw, s.qx, s.qy, s.qz});  // rbis.cpp:199
      }
      const V3 chi0{s.x[6], s.x[7], s.x[8]};
      for (int ci = 0; ci < st.n_chunks; ci++) {
        const int a0 = st.chunk_start[ci];
        switch (st.chunk_len[ci]) {
          case 1: meas_chunk<1>(P, s, st, a0, op.row, N, n, dquat, chi0); break;
          case 2: meas_chunk<2>(P, s, st, a0, op.row, N, n, dquat, chi0); break;
          case 3: meas_chunk<3>(P, s, st, a0, op.row, N, n, dquat, chi0); break;
          default: meas_chunk_general(st.chunk_len[ci], P, s, st, a0, op.row, N, n, dquat, chi0); break;
        }
      }
      meas_finish(s, chi0, p.chi_tol, p.ctor_folds_chi, p.renorm);
    } else if (op.kind == 2) {
      // ---- snapshot into ring slot ----
      if (active) {
        double* d = p.snap + op.row * 257 * N + n;
        static_for<NS>([&](auto i) { d[(long long)i * N] = s.x[i]; });
        d[21 * N] = s.qw; d[22 * N] = s.qx; d[23 * N] = s.qy; d[24 * N] = s.qz;
        d[25 * N] = s.ll;
        static_for<NP>([&](auto e) { d[(long long)(26 + e) * N] = P.template gets<e>(); });
      }
    } else {
      // ---- restore from ring slot ----
      const double* d = p.snap + op.row * 257 * N + n;
      static_for<NS>([&](auto i) { s.x[i] = d[(long long)i * N]; });
      s.qw = d[21 * N]; s.qx = d[22 * N]; s.qy = d[23 * N]; s.qz = d[24 * N];
      s.ll = d[25 * N];
      static_for<NP>([&](auto e) { P.template sets<e>(d[(long long)(26 + e) * N]); });
    }
  }

  if (active) {
    static_for<NS>([&](auto i) { p.vec[(long long)i * N + n] = s.x[i]; });
    p.quat[n] = s.qw; p.quat[N + n] = s.qx; p.quat[2 * N + n] = s.qy; p.quat[3 * N + n] = s.qz;
    p.loglik[n] = s.ll;
    static_for<NP>([&](auto e) { p.P[(long long)e * N + n] = P.template gets<e>(); });
  }
}

// ------------------------------------------------------------------------------------------------
// layout conversion: full column-major 441 <-> packed upper 231 (both [k][N])
// ------------------------------------------------------------------------------------------------
__global__ void pack_cov_kernel(const double* __restrict__ full, double* __restrict__ packed, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  for (int j = 0; j < NS; j++)
    for (int i = 0; i <= j; i++) packed[(long long)slot(i, j) * N + n] = full[(long long)(i + NS * j) * N + n];
}
__global__ void unpack_cov_kernel(const double* __restrict__ packed, double* __restrict__ full, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  for (int j = 0; j < NS; j++)
    for (int i = 0; i < NS; i++) full[(long long)(i + NS * j) * N + n] = packed[(long long)slot(i, j) * N + n];
}
__global__ void fill_kernel(double* __restrict__ dst, double v, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = v;
}

}  // namespace rbisk
