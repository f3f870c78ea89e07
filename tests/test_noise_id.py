"""IMU-noise identification (state-estimator/src/noise_id/noise_id.cpp:9-65): the C++ oracle
restatement against the independent numpy one (CPU), and pronto_b200.noise_id on the GPU against the oracle."""
import numpy as np
import pytest

from common import nominal_q, scenario


def filter_history(T, seed=3):
    """A filter-state history as the reference loads from a log (loadFilterHistory): the head state and covariance of
    one oracle filter after every IMU step of an IMU + leg-odometry run, subsampled to the IMU steps."""
    from oracle import oracle_api

    sc = scenario(1, T, with_pose=False, seed=seed)
    st = sc["st"]
    out = oracle_api.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"],
                                  [dict(idx=[3, 4, 5], z=st["legodo"], R=st["R_legodo"])], st["events"], trace=True)
    ev = st["events"]
    last_of_step = [i for i, e in enumerate(ev) if i + 1 == len(ev) or ev[i + 1][0] == 0]  # head after each step's updates
    vec = np.concatenate([sc["vec"].T, out["trace_vec"][last_of_step, :, 0]])
    quat = np.concatenate([sc["quat"].T, out["trace_quat"][last_of_step, :, 0]])
    cov = np.concatenate([sc["cov"].T, out["trace_cov"][last_of_step, :, 0]])
    return np.ascontiguousarray(vec), np.ascontiguousarray(quat), np.ascontiguousarray(cov)


def test_oracle_noise_id_matches_numpy_restatement():
    from oracle import oracle_api, rbis_numpy as rn

    vec, quat, cov = filter_history(45)
    states = [rn.State(vec[t].copy(), quat[t].copy()) for t in range(vec.shape[0])]
    covs = [cov[t].reshape(21, 21).T.copy() for t in range(vec.shape[0])]
    q = nominal_q()
    for (qg, qa, nw) in ((q[0], q[1], 10), (3 * q[0], 0.5 * q[1], 7), (q[0], q[1], 44)):
        ref, errs = oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, qg, qa, nw)
        got, errs_np = rn.noise_id_neg_loglik(states, covs, 1e-3, qg, qa, nw)
        assert len(errs) == len(errs_np) == (vec.shape[0] - 1) // nw
        assert abs(ref - got) <= 1e-9 * max(1.0, abs(ref))
        assert np.max(np.abs(errs - np.array(errs_np))) < 1e-12
    # a window longer than the history yields no complete window and a zero likelihood (noise_id.cpp:33-34)
    ref, errs = oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, q[0], q[1], 100)
    assert ref == 0.0 and len(errs) == 0


def test_oracle_noise_id_matches_the_reference_compiled():
    """oracle restatement == the reference's own state-estimator/src/noise_id/noise_id.cpp (sampleProcessForward,
    negLogLikelihood) compiled unmodified into oracle/_ref: per-window errors and the likelihood."""
    from oracle import oracle_api

    if oracle_api.build_ref() is None:
        pytest.skip("oracle/_ref/librbis_ref.so is not built and /root/reference is not mounted")
    vec, quat, cov = filter_history(45)
    q = nominal_q()
    for (qg, qa, nw) in ((q[0], q[1], 10), (3 * q[0], 0.5 * q[1], 7), (q[0], q[1], 44), (q[0], q[1], 100)):
        got, errs = oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, qg, qa, nw)
        with oracle_api.reference():
            ref, errs_ref = oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, qg, qa, nw)
        assert len(errs) == len(errs_ref)
        assert abs(ref - got) <= 1e-10 * max(1.0, abs(ref)), (ref, got)
        if len(errs):
            assert np.max(np.abs(errs - errs_ref)) < 1e-12


def test_oracle_noise_id_prefers_the_generating_noise_level():
    """Sanity of the restated likelihood convention: rolling forward NOISE-FREE inputs, the error is tiny, so the
    likelihood must prefer smaller process noise (log det term dominates)."""
    from oracle import oracle_api

    vec, quat, cov = filter_history(60)
    q = nominal_q()
    vals = [oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, s * q[0], s * q[1], 20)[0] for s in (0.25, 1.0, 4.0)]
    assert np.all(np.isfinite(vals))
    assert vals[0] != vals[1] != vals[2]


@pytest.mark.gpu
def test_gpu_noise_id_grid_matches_oracle():
    from oracle import oracle_api
    from pronto_b200 import noise_id

    vec, quat, cov = filter_history(330)
    q = nominal_q()
    qg = q[0] * np.array([0.3, 1.0, 1.0, 3.0, 0.7])
    qa = q[1] * np.array([1.0, 0.3, 1.0, 3.0, 1.9])
    nll, terms = noise_id.neg_log_likelihood(vec, quat, cov, 1e-3, qg, qa, 50, return_terms=True)
    assert terms.shape == (5, 6)
    for g in range(5):
        ref, _ = oracle_api.noise_id_neg_loglik(vec, quat, cov, 1e-3, qg[g], qa[g], 50)
        assert abs(nll[g] - ref) <= 1e-8 * max(1.0, abs(ref)), (g, nll[g], ref)
