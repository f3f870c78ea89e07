// rbis_ws.cuh -- the WARP-SPECIALISED mapping of the fused RBIS kernel for small / split ensembles (decoupled filters).
//
// Why it exists.  In the warp-group kernels (rbis_group.cuh) every lane of a group repeats the filter's
// serial arithmetic -- linearisation, insUpdateState, the 3x3 LDL^T, the log-likelihood, addState: ~1,100 of the ~1,400
// instructions a warp issues per step, whatever the group size -- so spreading a filter over more lanes buys almost nothing.
// Here the two kinds of work run in DIFFERENT WARPS of a team, each in the mapping that suits it:
//   * one STATE warp per 32 filters, one lane per filter, state in registers: exactly the serial code of the
//     lane-per-filter kernel (rbis_kernels.cuh), 32 filters per instruction instead of 4; it also reads (or, SYN, draws)
//     every input row, coalesced over the team's 32 consecutive filters;
//   * eight COVARIANCE warps, 8 lanes per filter (4 filters each), covariance as a full 15x15 matrix in shared memory:
//     the column / row passes and the measurement sweep of rbis_group.cuh (g_cov_propagate, g_sweep) and nothing else.
// They meet through a small exchange area in shared memory ([item][filter] with a padded row, conflict free from both
// sides) and named barriers (bar.arrive by the producer after a __threadfence_block, bar.sync by the consumer):
//   round = one IMU op, or one aligned-triple chunk of a measurement op
//   IMU:   state: Lin(x) -> PARAMS -> insUpdateState           cov: PARAMS -> Ad P Ad^T + Qd -> CONSUMED
//   chunk: cov: S = P[idx,idx] -> SREADY       state: SREADY -> + R, LDL^T, log det -> PARAMS -> residual, u
//          cov: PARAMS -> Y = L^-1 P[idx,:] -> YREADY -> sweep -> CONSUMED       state: YREADY -> x += Y^T u, log-likelihood
// The parameter slot and the PARAMS / CONSUMED barriers alternate with the round parity, so the state warp runs up to one
// round ahead of the covariance warps (its insUpdateState overlaps their congruences).
// EVERY ELEMENT IS COMPUTED BY THE SAME EXPRESSION AS IN THE OTHER MAPPINGS: same bits (tests/test_gpu_group.py).
// Programs it takes: IMU ops, measurement ops whose chunks are all aligned index triples, SNAPSHOT / RESTORE; decoupled
// ensembles only (the host falls back to the warp-group kernels otherwise).
//
// Included by rbis_fused_ws.cu after rbis_kernels.cuh and rbis_group.cuh.
#ifndef RBIS_WS_CUH_
#define RBIS_WS_CUH_

namespace rbisk {
namespace ws {

constexpr int G = 8;                 // lanes per filter in the covariance warps
using GE = grp::Geo<G, true>;
using CtxT = grp::Ctx<G, true>;
constexpr int COV_WARPS = 8, TEAM_WARPS = 1 + COV_WARPS, TEAM_THREADS = 32 * TEAM_WARPS, TEAM_FILTERS = 32;
constexpr int XROW = 33;             // doubles per exchange row: 32 filters + 1 (column and row accesses both conflict free)
constexpr int X_PAR = 0;             // two parameter slots of 22 rows: Lin of an IMU round, or the six LDL^T values of a chunk
constexpr int X_S = 44;              // S = P[idx,idx], 6 rows
constexpr int X_Y = 50;              // Y = L^-1 P[idx,:], 3 x 15 rows
constexpr int X_ROWS = X_Y + 3 * N_ACT;
constexpr int TEAM_DOUBLES = TEAM_FILTERS * GE::S + X_ROWS * XROW;
constexpr int MAX_TEAMS = 2;         // per CTA: 18 warps at <= 112 registers
constexpr int team_smem_bytes(int teams) { return teams * TEAM_DOUBLES * 8; }
static_assert(team_smem_bytes(MAX_TEAMS) <= 232448, "shared memory of two teams exceeds 227 KB");
// named barriers of a team: base = 1 + 6 * team
enum { B_PAR0 = 0, B_PAR1 = 1, B_CONS0 = 2, B_CONS1 = 3, B_SREADY = 4, B_YREADY = 5, B_PER_TEAM = 6 };

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TEAM_THREADS) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(TEAM_THREADS) : "memory"); }
// producer side of a hand-off: what this thread wrote is visible to the CTA before the arrival counts
__device__ __forceinline__ void publish(int id) {
  __threadfence_block();
  bar_arrive(id);
}

// x[I0 + A] for an aligned triple base I0 in {0, 3, .., 18} (run time, warp uniform)
template <int A>
__device__ __forceinline__ double pick_triple(const double (&x)[NS], int I0) {
  double v = x[A];
#pragma unroll
  for (int t = 1; t < NS / 3; t++) v = (I0 == 3 * t) ? x[3 * t + A] : v;
  return v;
}

template <bool SYN>
__global__ void __maxnreg__(104) rbis_ws_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int team = warp / TEAM_WARPS, wt = warp - team * TEAM_WARPS;  // wt 0: the state warp; 1..8: covariance warps
  const int teams = blockDim.x / TEAM_THREADS;
  double* const tbase = smem + (size_t)team * TEAM_DOUBLES;
  double* const X = tbase + TEAM_FILTERS * GE::S;
  const int bar0 = 1 + B_PER_TEAM * team;
  const long long N = p.N;
  const long long f0 = (((long long)blockIdx.x + p.block_offset) * teams + team) * TEAM_FILTERS;  // first filter of the team

  auto load_op = [&](long long i) {
    Op o;
    long long w0, w1;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(w0) : "l"(reinterpret_cast<const long long*>(p.ops + i)));
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(w1) : "l"(reinterpret_cast<const long long*>(p.ops + i) + 1));
    o.kind = (int)(w0 & 0xffffffffll); o.stream = (int)(w0 >> 32); o.row = w1;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(o.dt) : "l"(&p.ops[i].dt));
    return o;
  };

  // A round lasts ~1,000 cycles, about one trip to HBM, and nothing else hides the latency: the op table is pulled into L1 a
  // few lines ahead and ops are fetched two ahead (three in the state warp, which also prefetches the input rows two ops ahead).
  auto prefetch = [&](const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); };
  constexpr long long OPS_AHEAD = 16;  // ops (24 bytes each) of the table kept ahead in L1

  if (wt == 0) {
    // =============================== STATE WARP: one lane per filter ===============================
    long long n = f0 + lane;
    const bool active = n < N;
    if (!active) n = N - 1;  // idle lanes shadow the last filter and never store
    FilterState s;
    static_for<NS>([&](auto i) { s.x[i] = p.vec[(long long)i * N + n]; });
    s.qw = p.quat[n]; s.qx = p.quat[N + n]; s.qy = p.quat[2 * N + n]; s.qz = p.quat[3 * N + n];
    s.ll = p.loglik[n];
    const long long imu_n = p.imu_map ? (long long)__ldg(p.imu_map + n) : n;
    unsigned long long syn_kf = 0;
    double syn_sg = 0.0, syn_sa = 0.0;
    if constexpr (SYN) {
      syn_kf = p.syn.seed ^ ((unsigned long long)(p.syn.first_filter + n) * SYN_K1);
      syn_sg = p.syn.sigma_gyro >= 0 ? p.syn.sigma_gyro : sqrt(__ldg(p.q_gyro + n) / p.syn.dt);
      syn_sa = p.syn.sigma_accel >= 0 ? p.syn.sigma_accel : sqrt(__ldg(p.q_accel + n) / p.syn.dt);
    }
    double* const Xl = X + lane;  // this filter's column of the exchange area
    long long r = 0;              // round counter
    auto prefetch_inputs = [&](const Op& o) {
      if constexpr (SYN) {
        if (o.kind == 0) prefetch(p.syn.imu_mean + o.row * 6);
      } else {
        if (o.kind == 0) {
          const double* base = p.imu + o.row * 6 * p.imu_cols + imu_n;
#pragma unroll
          for (int k = 0; k < 6; k++) prefetch(base + k * p.imu_cols);
        } else if (o.kind == 1) {
          const StreamDesc& st = p.streams[o.stream];
          const long long sn = st.map ? (long long)__ldg(st.map + n) : n;
          const double* zb = st.z + o.row * st.m * st.cols + sn;
          for (int a = 0; a < st.m; a++) prefetch(zb + a * st.cols);
          if (st.has_orient) {
            const double* qb = st.quat + o.row * 4 * st.cols + sn;
#pragma unroll
            for (int k = 0; k < 4; k++) prefetch(qb + k * st.cols);
          }
        }
      }
    };
    for (long long k = 0; k < OPS_AHEAD && k < p.n_ops; k += 5) prefetch(p.ops + k);
    Op op1 = load_op(0);
    Op op2 = p.n_ops > 1 ? load_op(1) : op1;
    if (p.n_ops > 1) prefetch_inputs(op2);
    for (long long oi = 0; oi < p.n_ops; oi++) {
      const Op op = op1;
      op1 = op2;
      if (oi + 2 < p.n_ops) { op2 = load_op(oi + 2); prefetch_inputs(op2); }
      if (oi + OPS_AHEAD < p.n_ops) prefetch(p.ops + oi + OPS_AHEAD);
      if (op.kind == 0) {
        // ---- IMU process step: linearisation for the covariance warps, then insUpdateState (rbis.cpp:37-75) ----
        V3 gyro, acc;
        if constexpr (SYN) {
          const unsigned long long kfs = syn_kf ^ ((unsigned long long)__ldg(p.syn.imu_step + op.row) * SYN_K2);
          const double* mu = p.syn.imu_mean + op.row * 6;
          double nn[6];
          syn_normals_k<1, 6>(kfs, 0u, nn);
          gyro = {fma(syn_sg, nn[0], __ldg(mu)), fma(syn_sg, nn[1], __ldg(mu + 1)), fma(syn_sg, nn[2], __ldg(mu + 2))};
          acc = {fma(syn_sa, nn[3], __ldg(mu + 3)), fma(syn_sa, nn[4], __ldg(mu + 4)), fma(syn_sa, nn[5], __ldg(mu + 5))};
        } else {
          const double* base = p.imu + op.row * 6 * p.imu_cols + imu_n;
          const long long Ni = p.imu_cols;
          gyro = {ldg_early(base), ldg_early(base + Ni), ldg_early(base + 2 * Ni)};
          acc = {ldg_early(base + 3 * Ni), ldg_early(base + 4 * Ni), ldg_early(base + 5 * Ni)};
        }
        const double dt = op.dt;
        const Q4 q{s.qw, s.qx, s.qy, s.qz};
        const double tx_ = 2 * q.x, ty_ = 2 * q.y, tz_ = 2 * q.z;
        const double twx_ = tx_ * q.w, twy_ = ty_ * q.w, twz_ = tz_ * q.w;
        const double txx_ = tx_ * q.x, txy_ = ty_ * q.x, txz_ = tz_ * q.x;
        const double tyy_ = ty_ * q.y, tyz_ = tz_ * q.y, tzz_ = tz_ * q.z;
        const double R00 = 1 - (tyy_ + tzz_), R01 = txy_ - twz_, R02 = txz_ + twy_;
        const double R10 = txy_ + twz_, R11 = 1 - (txx_ + tzz_), R12 = tyz_ - twx_;
        const double R20 = txz_ - twy_, R21 = tyz_ + twx_, R22 = 1 - (txx_ + tyy_);
        const V3 gb{-p.g_val * R20, -p.g_val * R21, -p.g_val * R22};
        const V3 v{s.x[3], s.x[4], s.x[5]};
        if (r >= 2) bar_sync(bar0 + B_CONS0 + (int)(r & 1));  // the covariance warps are done with round r - 2: its slot is free
        {
          double* xp = Xl + (X_PAR + 22 * (int)(r & 1)) * XROW;
          xp[0] = v.x; xp[XROW] = v.y; xp[2 * XROW] = v.z;
          xp[3 * XROW] = s.x[0] * dt; xp[4 * XROW] = s.x[1] * dt; xp[5 * XROW] = s.x[2] * dt;  // omega of the PRIOR state
          xp[6 * XROW] = v.x * dt; xp[7 * XROW] = v.y * dt; xp[8 * XROW] = v.z * dt;
          xp[9 * XROW] = gb.x * dt; xp[10 * XROW] = gb.y * dt; xp[11 * XROW] = gb.z * dt;
          xp[12 * XROW] = R00 * dt; xp[13 * XROW] = R01 * dt; xp[14 * XROW] = R02 * dt;
          xp[15 * XROW] = R10 * dt; xp[16 * XROW] = R11 * dt; xp[17 * XROW] = R12 * dt;
          xp[18 * XROW] = R20 * dt; xp[19 * XROW] = R21 * dt; xp[20 * XROW] = R22 * dt;
          xp[21 * XROW] = dt;
        }
        publish(bar0 + B_PAR0 + (int)(r & 1));
        s.x[9] += fma(R02, v.z, fma(R01, v.y, R00 * v.x)) * dt;
        s.x[10] += fma(R12, v.z, fma(R11, v.y, R10 * v.x)) * dt;
        s.x[11] += fma(R22, v.z, fma(R21, v.y, R20 * v.x)) * dt;
        state_propagate<true>(s, gyro, acc, dt, gb, p.chi_tol, p.renorm);
        r++;
      } else if (op.kind == 1) {
        // ---- indexed / indexed-plus-orientation measurement: every chunk is an aligned index triple ----
        const StreamDesc& st = p.streams[op.stream];
        const long long sn = st.map ? (long long)__ldg(st.map + n) : n;
        MeasSrc<SYN> src{};
        if constexpr (SYN) {
          src.ss = &p.syn.st[op.stream];
          src.kfs = syn_kf ^ ((unsigned long long)__ldg(src.ss->step + op.row) * SYN_K2);
        }
        V3 dquat{0, 0, 0};
        if (st.has_orient) dquat = subtract_quats(src.quat(st, op.row, sn), {s.qw, s.qx, s.qy, s.qz});  // rbis.cpp:199
        const V3 chi0{s.x[6], s.x[7], s.x[8]};
        for (int ci = 0; ci < st.n_chunks; ci++) {
          const int a0 = st.chunk_start[ci];
          const int I0 = st.chunk_fast[ci];
          double z[3], Rdg[3];
          src.z3(st, op.row, a0, sn, z);
          if (st.r_mode == 1) {
#pragma unroll
            for (int a = 0; a < 3; a++) Rdg[a] = ldg_early(st.R + (long long)(a0 + a) * N + n);
          }
          bar_sync(bar0 + B_SREADY);
          const double* xs_ = Xl + X_S * XROW;
          double S00 = xs_[0], S10 = xs_[XROW], S20 = xs_[2 * XROW], S11 = xs_[3 * XROW], S21 = xs_[4 * XROW], S22 = xs_[5 * XROW];
          if (st.r_mode == 1) {
            S00 += Rdg[0]; S11 += Rdg[1]; S22 += Rdg[2];
          } else {
            const double* Rm = st.R + a0 + (long long)st.m * a0;
            S00 += __ldg(Rm); S11 += __ldg(Rm + st.m + 1); S22 += __ldg(Rm + 2 * st.m + 2);
            S10 += __ldg(Rm + 1); S20 += __ldg(Rm + 2); S21 += __ldg(Rm + st.m + 2);
          }
          const double d0 = S00, r0 = 1.0 / d0;
          const double l10 = S10 * r0, l20 = S20 * r0;
          const double d1 = fma(-l10 * l10, d0, S11), r1 = 1.0 / d1;
          const double l21 = fma(-l20 * l10, d0, S21) * r1;
          const double d2 = fma(-l21 * l21, d1, fma(-l20 * l20, d0, S22)), r2 = 1.0 / d2;
          if (r >= 2) bar_sync(bar0 + B_CONS0 + (int)(r & 1));
          {
            double* xp = Xl + (X_PAR + 22 * (int)(r & 1)) * XROW;
            xp[0] = l10; xp[XROW] = l20; xp[2 * XROW] = l21; xp[3 * XROW] = r0; xp[4 * XROW] = r1; xp[5 * XROW] = r2;
          }
          publish(bar0 + B_PAR0 + (int)(r & 1));
          const double pd = d0 * d1 * d2;
          const double logdet = (pd > 1e-290 && pd < 1e290) ? log(pd) : log(d0) + log(d1) + log(d2);
          double rr[3];
          if (I0 == 6 && st.has_orient) {
            rr[0] = dquat.x - (s.x[6] - chi0.x); rr[1] = dquat.y - (s.x[7] - chi0.y); rr[2] = dquat.z - (s.x[8] - chi0.z);
          } else {
            rr[0] = z[0] - pick_triple<0>(s.x, I0); rr[1] = z[1] - pick_triple<1>(s.x, I0); rr[2] = z[2] - pick_triple<2>(s.x, I0);
          }
          const double e0 = rr[0], e1 = fma(-l10, e0, rr[1]), e2 = fma(-l21, e1, fma(-l20, e0, rr[2]));
          const double u0 = e0 * r0, u1 = e1 * r1, u2 = e2 * r2;
          bar_sync(bar0 + B_YREADY);
          const double* xy = Xl + X_Y * XROW;
          static_for<N_ACT>([&](auto cc) {
            constexpr int c = cc;
            constexpr int xc = act_col(c);
            s.x[xc] = fma(xy[c * XROW], u0, fma(xy[(N_ACT + c) * XROW], u1, fma(xy[(2 * N_ACT + c) * XROW], u2, s.x[xc])));
          });
          s.ll += -logdet - fma(e2, u2, fma(e1, u1, e0 * u0));
          r++;
        }
        meas_finish(s, chi0, p.chi_tol, p.ctor_folds_chi, p.renorm);
      } else if (op.kind == 2) {
        double* d = p.snap + op.row * SNAP_ROWS * N + n;
        if (active) {
          static_for<NS>([&](auto i) { d[(long long)i * N] = s.x[i]; });
          d[21 * N] = s.qw; d[22 * N] = s.qx; d[23 * N] = s.qy; d[24 * N] = s.qz;
          d[25 * N] = s.ll;
        }
      } else {
        const double* d = p.snap + op.row * SNAP_ROWS * N + n;
        static_for<NS>([&](auto i) { s.x[i] = d[(long long)i * N]; });
        s.qw = d[21 * N]; s.qx = d[22 * N]; s.qy = d[23 * N]; s.qz = d[24 * N];
        s.ll = d[25 * N];
      }
    }
    // the last two rounds' CONSUMED arrivals are still open: complete those barrier phases before leaving
    if (r >= 2) bar_sync(bar0 + B_CONS0 + (int)(r & 1));
    if (r >= 1) bar_sync(bar0 + B_CONS0 + (int)((r + 1) & 1));
    if (active) {
      static_for<NS>([&](auto i) { p.vec[(long long)i * N + n] = s.x[i]; });
      p.quat[n] = s.qw; p.quat[N + n] = s.qx; p.quat[2 * N + n] = s.qy; p.quat[3 * N + n] = s.qz;
      p.loglik[n] = s.ll;
    }
  } else {
    // =============================== COVARIANCE WARPS: 8 lanes per filter ===============================
    const int l = lane % G;
    const int fl = (wt - 1) * GE::FPW + lane / G;  // filter within the team
    long long n = f0 + fl;
    const bool active = n < N;
    if (!active) n = N - 1;
    const CtxT cx(tbase + (size_t)fl * GE::S, l);
    const QNoise qn{__ldg(p.q_gyro + n), __ldg(p.q_accel + n), __ldg(p.q_gyro_bias + n), __ldg(p.q_accel_bias + n)};
    grp::g_cov_load<G, true>(cx.Pf, l, p.P + n, N);
    bool imu_seen = false;
    const double* const Xf = X + fl;
    long long r = 0;
    for (long long k = 0; k < OPS_AHEAD && k < p.n_ops; k += 5) prefetch(p.ops + k);
    Op op1 = load_op(0);
    Op op2 = p.n_ops > 1 ? load_op(1) : op1;
    for (long long oi = 0; oi < p.n_ops; oi++) {
      const Op op = op1;
      op1 = op2;
      if (oi + 2 < p.n_ops) op2 = load_op(oi + 2);
      if (oi + OPS_AHEAD < p.n_ops) prefetch(p.ops + oi + OPS_AHEAD);
      if (op.kind == 0) {
        bar_sync(bar0 + B_PAR0 + (int)(r & 1));
        const double* xp = Xf + (X_PAR + 22 * (int)(r & 1)) * XROW;
        Lin L;
        L.v = {xp[0], xp[XROW], xp[2 * XROW]};
        L.wd = {xp[3 * XROW], xp[4 * XROW], xp[5 * XROW]};
        L.vd = {xp[6 * XROW], xp[7 * XROW], xp[8 * XROW]};
        L.gd = {xp[9 * XROW], xp[10 * XROW], xp[11 * XROW]};
#pragma unroll
        for (int k = 0; k < 9; k++) L.Rd[k] = xp[(12 + k) * XROW];
        L.dt = xp[21 * XROW];
        imu_seen = true;
        grp::g_cov_propagate<G, true>(cx, L, qn);
        bar_arrive(bar0 + B_CONS0 + (int)(r & 1));
        r++;
      } else if (op.kind == 1) {
        const StreamDesc& st = p.streams[op.stream];
        for (int ci = 0; ci < st.n_chunks; ci++) {
          const int p0 = grp::pos_of<true>(st.chunk_fast[ci]);
          constexpr int LD = GE::LD, NJ = GE::NJ;
          double yo[NJ][3];
#pragma unroll
          for (int j = 0; j < NJ; j++) {
            const double* b = cx.col[j] + p0 * LD;
            yo[j][0] = b[0]; yo[j][1] = b[LD]; yo[j][2] = b[2 * LD];
          }
          if (l == 0) {
            const double* sp = cx.Pf + p0 * LD + p0;
            double* xs_ = X + fl + X_S * XROW;
            xs_[0] = sp[0]; xs_[XROW] = sp[LD]; xs_[2 * XROW] = sp[2 * LD];
            xs_[3 * XROW] = sp[LD + 1]; xs_[4 * XROW] = sp[2 * LD + 1]; xs_[5 * XROW] = sp[2 * LD + 2];
          }
          publish(bar0 + B_SREADY);
          bar_sync(bar0 + B_PAR0 + (int)(r & 1));
          const double* xp = Xf + (X_PAR + 22 * (int)(r & 1)) * XROW;
          const double l10 = xp[0], l20 = xp[XROW], l21 = xp[2 * XROW], r0 = xp[3 * XROW], r1 = xp[4 * XROW], r2 = xp[5 * XROW];
          double wj[NJ][3];
          static_for<NJ>([&](auto jc) {
            constexpr int j = jc;
            yo[j][1] = fma(-l10, yo[j][0], yo[j][1]);
            yo[j][2] = fma(-l21, yo[j][1], fma(-l20, yo[j][0], yo[j][2]));
            wj[j][0] = yo[j][0] * r0; wj[j][1] = yo[j][1] * r1; wj[j][2] = yo[j][2] * r2;
            if (cx.template own<j>()) {
              const int c = cx.c[j];
              double* y = cx.Ys + 4 * c;
              *reinterpret_cast<double2*>(y) = make_double2(yo[j][0], yo[j][1]);
              y[2] = yo[j][2];
              double* xy = X + fl + (X_Y + c) * XROW;
              xy[0] = yo[j][0]; xy[N_ACT * XROW] = yo[j][1]; xy[2 * N_ACT * XROW] = yo[j][2];
            }
          });
          __syncwarp();  // Y is complete, and every lane has read its part of the three rows before any lane overwrites them
          publish(bar0 + B_YREADY);
          grp::g_sweep<G, true, 3>(cx, wj);
          __syncwarp();
          bar_arrive(bar0 + B_CONS0 + (int)(r & 1));
          r++;
        }
      } else if (op.kind == 2) {
        double* dc = p.snap + op.row * SNAP_ROWS * N + n + 26 * N;
        grp::g_cov_store<G, true>(cx.Pf, l, dc, N, active);
        if (active) {
          for (int k = l; k < N_REST; k += G) dc[(long long)c_act.rest[k] * N] = 0.0;
          for (int k = l; k < 12; k += G) {
            const int s_ = c_act.blk[k];
            const int jc = col_of_slot(s_), ir = s_ - jc * (jc + 1) / 2;
            dc[(long long)s_ * N] = imu_seen ? ((ir != jc) ? 0.0 : (jc < 3 ? qn.q_gyro : qn.q_accel)) : p.P[(long long)s_ * N + n];
          }
        }
      } else {
        const double* dc = p.snap + op.row * SNAP_ROWS * N + n + 26 * N;
        __syncwarp();
        grp::g_cov_load<G, true>(cx.Pf, l, dc, N);
        if (active)
          for (int k = l; k < 12; k += G) {
            const int s_ = c_act.blk[k];
            p.P[(long long)s_ * N + n] = dc[(long long)s_ * N];
          }
        imu_seen = false;
      }
    }
    grp::g_cov_store<G, true>(cx.Pf, l, p.P + n, N, active);
    if (active && imu_seen)
      for (int k = l; k < 12; k += G) {
        const int s_ = c_act.blk[k];
        const int jc = col_of_slot(s_), ir = s_ - jc * (jc + 1) / 2;
        p.P[(long long)s_ * N + n] = (ir != jc) ? 0.0 : (jc < 3 ? qn.q_gyro : qn.q_accel);
      }
  }
}

}  // namespace ws
}  // namespace rbisk
#endif  // RBIS_WS_CUH_
