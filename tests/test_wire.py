"""Wire structs and the KVH batch decode, pinned to the reference's own compiled sources (oracle/_ref):
rbisCreateFilterStateMessage / RBIS(const pronto_filter_state_t*) (MSE/rbis.cpp:268-285, rbis.hpp:58-67) and
IMUStream::convertFromLCMBatch (estimate_tools/src/estimate_tools/imu_stream.cpp:62-97)."""
import ctypes as C

import numpy as np
import pytest

from common import random_ensemble


@pytest.fixture(scope="module")
def ref_lib():
    from oracle import oracle_api

    path = oracle_api.build_ref()
    if path is None:
        pytest.skip("oracle/_ref/librbis_ref.so is not built and /root/reference is not mounted")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.orc_create_filter_state_message.argtypes = [vp, vp, C.c_int64, vp, vp]
    lib.orc_rbis_from_filter_state.argtypes = [vp, vp, vp, C.POINTER(C.c_int64)]
    lib.orc_kvh_decode.argtypes = [C.c_int64, C.c_int, vp, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int)]
    lib.orc_kvh_decode.restype = C.c_int
    return lib


def _kvh_log(rng, n_batches, with_skip):
    """A KVH batch log: every batch repeats the newest 15 packets (newest first), 1-4 of them new; optionally the packet
    counter restarts in the middle (driver restart)."""
    packets, count, ut = [], 100, 1_000_000
    batches = []
    for b in range(n_batches):
        if with_skip and b == n_batches // 2:
            count, packets = 3, []
        for _ in range(int(rng.integers(0, 5))):
            count += 1
            ut += int(rng.integers(900, 1100))
            packets.append((ut, count, rng.normal(size=3) * 1e-3, rng.normal(size=3) + [0, 0, 9.8]))
        if not packets:
            count += 1; ut += 1000
            packets.append((ut, count, rng.normal(size=3) * 1e-3, rng.normal(size=3)))
        batches.append((ut + 250, list(reversed(packets[-15:]))))
    return batches


@pytest.mark.parametrize("with_skip", [False, True])
def test_kvh_decode_matches_reference_imu_stream(rbis_lib, ref_lib, with_skip):
    from pronto_b200 import capi

    lib = rbis_lib
    rng = np.random.default_rng(5 + with_skip)
    s = C.c_void_p()
    capi.check(lib.rbis_kvh_stream_create(C.byref(s)))
    ref_lib.orc_kvh_reset()
    total_new = 0
    for batch_utime, pk in _kvh_log(rng, 60, with_skip):
        n = len(pk)
        raw = (capi.KvhPacket * n)()
        flat = np.zeros((n, 8))
        for i, (ut, cnt, dr, la) in enumerate(pk):
            raw[i].utime, raw[i].packet_count = ut, cnt
            for k in range(3):
                raw[i].delta_rotation[k], raw[i].linear_acceleration[k] = dr[k], la[k]
            flat[i] = [ut, cnt, *dr, *la]
        on, oo = (capi.ImuPacket * n)(), (capi.ImuPacket * n)()
        nn, no = C.c_int32(0), C.c_int32(0)
        capi.check(lib.rbis_kvh_decode_batch(s, batch_utime, n, raw, on, C.byref(nn), oo, C.byref(no)))
        rn_, ro_ = np.zeros((n, 11)), np.zeros((n, 11))
        cn, co = C.c_int(0), C.c_int(0)
        ref_lib.orc_kvh_decode(batch_utime, n, flat.ctypes.data, rn_.ctypes.data, C.byref(cn), ro_.ctypes.data, C.byref(co))
        assert (nn.value, no.value) == (cn.value, co.value)
        for got, ref, cnt in ((on, rn_, nn.value), (oo, ro_, no.value)):
            for i in range(cnt):
                g = got[i]
                row = [g.utime_raw, g.utime_batch, g.utime, g.utime_delta, g.packet_count, *g.delta_rotation, *g.linear_acceleration]
                assert np.array_equal(np.array(row, dtype=np.float64), ref[i]), (i, row, ref[i])
        total_new += nn.value
    assert total_new > 60
    lib.rbis_kvh_stream_destroy(s)


def test_kvh_imu_step_follows_process_message_atlas(rbis_lib):
    """gyro = R (delta_rotation / raw_dt), accel = R a + t, dt = default for the first message then the batch utime difference
    (MSE/sensor_handlers.cpp:209-246)."""
    from pronto_b200 import capi

    lib = rbis_lib
    p = capi.ImuPacket()
    p.utime_delta = 1000
    for k, (dr, la) in enumerate(zip((1e-3, -2e-3, 5e-4), (0.1, 0.2, 9.8))):
        p.delta_rotation[k], p.linear_acceleration[k] = dr, la
    q = np.array([np.cos(0.3), 0.0, 0.0, np.sin(0.3)])       # yaw 0.6 rad
    t = np.array([0.01, -0.02, 0.3])
    gyro, accel = np.zeros(3), np.zeros(3)
    dt, prev = C.c_double(0), C.c_int64(0)
    dp = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(lib.rbis_kvh_imu_step(C.byref(p), 5_000_000, dp(q), dp(t), 1e-3, C.byref(prev), dp(gyro), dp(accel), C.byref(dt)))
    c, s_ = np.cos(0.6), np.sin(0.6)
    R = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
    assert np.allclose(gyro, R @ (np.array([1e-3, -2e-3, 5e-4]) / 1e-3), rtol=0, atol=1e-12)
    assert np.allclose(accel, R @ np.array([0.1, 0.2, 9.8]) + t, rtol=0, atol=1e-12)
    assert dt.value == 1e-3 and prev.value == 5_000_000
    capi.check(lib.rbis_kvh_imu_step(C.byref(p), 5_001_030, dp(q), dp(t), 1e-3, C.byref(prev), dp(gyro), dp(accel), C.byref(dt)))
    assert abs(dt.value - 1.03e-3) < 1e-15 and prev.value == 5_001_030


def test_indexed_measurement_message_becomes_a_stream(rbis_lib):
    from pronto_b200 import capi

    lib = rbis_lib
    msg = capi.IndexedMeasurement()
    idx = [8, 9, 10, 11]                       # quick-lock (quick_lock.cpp:132)
    R = np.arange(16, dtype=np.float64).reshape(4, 4)
    R = R @ R.T + np.eye(4)
    msg.measured_dim, msg.measured_cov_dim = 4, 16
    for a in range(4):
        msg.z_indices[a], msg.z_effective[a] = idx[a], 0.5 * a
    for e, v in enumerate(R.T.reshape(-1)):     # Map<MatrixXd>(R_effective, m, m): column-major
        msg.R_effective[e] = v
    st, Rout = capi.Stream(), np.zeros(81)
    capi.check(lib.rbis_stream_from_indexed_measurement(C.byref(msg), C.byref(st), Rout.ctypes.data))
    assert (st.m, st.has_orientation, st.r_mode, st.sensor_id) == (4, 0, capi.R_SHARED_FULL, 5)
    assert list(st.idx)[:4] == idx and st.R == Rout.ctypes.data
    assert np.array_equal(Rout[:16].reshape(4, 4).T, R)
    msg.measured_cov_dim = 15
    assert lib.rbis_stream_from_indexed_measurement(C.byref(msg), C.byref(st), Rout.ctypes.data) == -1


@pytest.mark.gpu
def test_filter_state_messages_match_reference_pack_and_unpack(ref_lib):
    from pronto_b200 import RBISBatch, capi

    N = 40
    vec, quat, cov = random_ensemble(N, seed=12)
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov, utime=123_456)
        msgs = (capi.FilterState * 7)()
        capi.check(b.lib.rbis_batch_get_filter_states(b.h, 5, 7, msgs))
        for k in range(7):
            n = 5 + k
            ref = np.zeros(469)
            # the device keeps the covariance symmetric-packed: the reference packs the same symmetric matrix
            P = cov[:, n].reshape(21, 21)
            Ps = np.triu(P.T) + np.triu(P.T, 1).T
            # (named arrays: a temporary's buffer may be gone by the time the call reads it)
            v_n, q_n, P_n = np.ascontiguousarray(vec[:, n]), np.ascontiguousarray(quat[:, n]), np.ascontiguousarray(Ps.T.reshape(-1))
            ref_lib.orc_create_filter_state_message(v_n.ctypes.data, q_n.ctypes.data, 123_456, P_n.ctypes.data, ref.ctypes.data)
            m = msgs[k]
            got = np.array([m.utime, *m.quat, m.num_states, *m.state, m.num_cov_elements, *m.cov], dtype=np.float64)
            bad = np.flatnonzero(got != ref)
            assert bad.size == 0, (k, bad[:8], got[bad[:8]], ref[bad[:8]])
        # and back: RBIS(msg) of the reference == what set_filter_states stores
        msgs[2].state[4] = 9.25
        msgs[2].quat[0], msgs[2].quat[1] = 0.6, 0.8
        msgs[2].quat[2] = msgs[2].quat[3] = 0.0
        capi.check(b.lib.rbis_batch_set_filter_states(b.h, 20, 7, msgs))
        gv, gq, gP, _ = b.get_filter(22)
        flat = np.array([msgs[2].utime, *msgs[2].quat, 21, *msgs[2].state, 441, *msgs[2].cov], dtype=np.float64)
        rv, rq, ut = np.zeros(21), np.zeros(4), C.c_int64(0)
        ref_lib.orc_rbis_from_filter_state(flat.ctypes.data, rv.ctypes.data, rq.ctypes.data, C.byref(ut))
        assert np.array_equal(gv, rv) and np.array_equal(gq, rq) and ut.value == 123_456
        assert np.array_equal(gP, np.array(msgs[2].cov))
