"""IMU conditioning ("next" row 4 of SURVEY.md 8f): the accelerometer notch cascade of InsHandler::doFilter
(MSE/sensor_handlers.cpp:29-41,155-162) over IIRNotch (estimate_tools/src/estimate_tools/iir_notch.cpp:3-60).

CPU: the oracle's restatement against the reference's own compiled iir_notch.cpp (oracle/_ref) and against
scipy.signal.lfilter with the same coefficients.  GPU: rbis_batch_notch_filter against the oracle, chunk by chunk.
"""
import numpy as np
import pytest

from pronto_b200 import RBISBatch, capi


def _signal(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 1000.0
    return rng.normal(size=n) * 0.3 + 2.0 * np.sin(2 * np.pi * 87 * t) + 0.7 * np.sin(2 * np.pi * 174 * t + 0.3) + 9.8


def test_oracle_cascade_vs_scipy_and_reference(oracle):
    from scipy.signal import lfilter

    x = _signal(4000)
    y, st, co = oracle.notch_cascade(x, 87.0, 1000.0, 3)
    z = x
    for i in range(3):
        z = lfilter(co[i, :3], co[i, 3:], z)
    assert np.max(np.abs(y - z)) < 1e-11
    # chaining two halves through the carried state equals one pass
    y1, st1, _ = oracle.notch_cascade(x[:1777], 87.0)
    y2, st2, _ = oracle.notch_cascade(x[1777:], 87.0, state=st1)
    assert np.array_equal(np.concatenate([y1, y2]), y) and np.array_equal(st2, st)
    # the notch does what it says: the 87 Hz line is gone, the mean (gravity) passes
    spec = lambda s: np.abs(np.fft.rfft(s[1000:] - np.mean(s[1000:])))[int(87 * 3)]
    assert spec(y) < 0.02 * spec(x) and abs(np.mean(y[2000:]) - 9.8) < 0.05
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not available")
    with oracle.reference():
        ry, rst, rco = oracle.notch_cascade(x, 87.0, 1000.0, 3)
    assert np.array_equal(co, rco)
    assert np.max(np.abs(y - ry)) < 1e-12 and np.max(np.abs(st - rst)) < 1e-12


def _ref_next():
    import os

    return np.load(os.path.join(os.path.dirname(__file__), "golden", "rbis_reference_golden_next.npz"))


def _golden_signal(n=3000, seed=11):  # as tests/golden/make_reference_golden_next.py:notch_signal
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 1000.0
    return rng.normal(size=n) * 0.3 + 2.0 * np.sin(2 * np.pi * 85 * t) + 9.8


def test_oracle_cascade_matches_reference_golden(oracle):
    g = _ref_next()
    y, st, co = oracle.notch_cascade(_golden_signal(), 85.0, 1000.0, 3)
    assert np.array_equal(co, g["notch_coeffs"])
    assert np.max(np.abs(y - g["notch_y"])) < 1e-12 and np.max(np.abs(st - g["notch_state"])) < 1e-12


@pytest.mark.gpu
def test_gpu_notch_matches_reference_golden():
    g = _ref_next()
    x = _golden_signal()
    imu = np.zeros((len(x), 6, 5))
    imu[:, 3:6, :] = x[:, None, None]
    with RBISBatch(5) as b:
        b.notch_configure(85.0, 1000.0, 3)
        b.notch_filter(imu)
    assert np.max(np.abs(imu[:, 3:6, :] - g["notch_y"][:, None, None])) < 1e-12 and np.all(imu[:, :3] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("device_resident", [False, True])
def test_gpu_notch_matches_oracle_chunk_by_chunk(oracle, device_resident):
    N, rows = 200, 700
    rng = np.random.default_rng(5)
    imu = rng.normal(size=(rows, 6, N))
    for c in range(0, N, 7):
        for ch in range(3):
            imu[:, 3 + ch, c] = _signal(rows, seed=c * 3 + ch)
    ref = imu.copy()
    for c in range(N):
        for ch in range(3):
            ref[:, 3 + ch, c] = oracle.notch_cascade(np.ascontiguousarray(imu[:, 3 + ch, c]), 85.0, 1000.0, 3)[0]
    got = imu.copy()
    with RBISBatch(N) as b:
        with pytest.raises(capi.RBISError):
            b._notch_cols = N
            b.notch_filter(got[:10].copy())  # not configured
        b.notch_configure(85.0, 1000.0, 3)
        for r0, r1 in ((0, 256), (256, 257), (257, 700)):
            chunk = np.ascontiguousarray(got[r0:r1])
            if device_resident:
                import torch

                t = torch.from_numpy(chunk).cuda()
                b.notch_filter(t)
                b.synchronize()
                chunk = t.cpu().numpy()
            else:
                b.notch_filter(chunk)
            got[r0:r1] = chunk
    assert np.array_equal(got[:, :3], imu[:, :3])  # gyro rows untouched
    assert np.max(np.abs(got - ref)) < 1e-12
