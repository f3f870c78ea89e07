"""Parity norms of SURVEY.md 8d (numpy, host side; no filter arithmetic).

vec   : max_n ||dvec_n||_inf / max(1, ||vec_n||_inf)
quat  : max |dquat| after aligning the sign of each quaternion pair
cov   : max_n max_ij |dP_ij| / sqrt(P_ii P_jj)   (reference diagonal; zero diagonals fall back to 1)
Arrays are structure-of-arrays: vec [21][N], quat [4][N], cov [441][N] (element r + 21 c).
"""
import numpy as np


def max_errors(vec, quat, cov, ref_vec, ref_quat, ref_cov, per_filter=False):
    vec, quat, ref_vec, ref_quat = (np.asarray(a, dtype=np.float64) for a in (vec, quat, ref_vec, ref_quat))
    ev = np.max(np.abs(vec - ref_vec), axis=0) / np.maximum(1.0, np.max(np.abs(ref_vec), axis=0))
    sign = np.where(np.sum(quat * ref_quat, axis=0) < 0, -1.0, 1.0)
    eq = np.max(np.abs(quat * sign - ref_quat), axis=0)
    out = dict(vec=ev, quat=eq)
    if cov is not None and ref_cov is not None:
        cov, ref_cov = np.asarray(cov, dtype=np.float64), np.asarray(ref_cov, dtype=np.float64)
        N = cov.shape[1]
        d = np.stack([ref_cov[i + 21 * i] for i in range(21)])  # [21][N]
        d = np.where(d > 0, d, 1.0)
        s = np.sqrt(d)
        scale = (s[:, None, :] * s[None, :, :]).reshape(441, N)  # element (r,c) -> r*21+c; symmetric so order is moot
        out["cov"] = np.max(np.abs(cov - ref_cov) / scale, axis=0)
    if per_filter:
        return out
    return {k: float(np.max(v)) for k, v in out.items()}
