"""Shared helpers for the parity tests (tests only)."""
import numpy as np


def vel(a):
    return np.stack([a["vx"], a["vy"], a["vz"]], 1).astype(np.float64)


def pos(a):
    return np.stack([a["px"], a["py"], a["pz"]], 1).astype(np.float64)


# SURVEY.md §8d parity metric (north_star: relative tolerance 1e-5 after one step):
#   |dp_i| <= tol * max(|p_i|, W/2)   and   |dv_i| <= tol * max(|v_i|, v_rms)
# The floors exist because a perfect f32 implementation already differs from f64 truth by up to
# 7.7e-5 of |v_i| for particles whose forces nearly cancel.
def parity_errors(out, ref, world_size):
    v, vr = vel(out), vel(ref)
    p, pr = pos(out), pos(ref)
    vrms = float(np.sqrt((vr ** 2).sum(1).mean())) if len(vr) else 0.0
    vden = np.maximum(np.linalg.norm(vr, axis=1), max(vrms, 1e-30))
    pden = np.maximum(np.linalg.norm(pr, axis=1), world_size / 2)
    dv = np.linalg.norm(v - vr, axis=1) / vden
    dp = np.linalg.norm(p - pr, axis=1) / pden
    return dv, dp


def assert_parity(out, ref, world_size, tol=1e-5, mask=None, what=""):
    assert np.array_equal(out["id"], ref["id"]), "ids / index order changed"
    dv, dp = parity_errors(out, ref, world_size)
    if mask is not None:
        dv, dp = dv[mask], dp[mask]
    assert dv.size == 0 or dv.max() <= tol, f"{what}: velocity parity {dv.max():.3e} > {tol} ({(dv > tol).sum()} particles)"
    assert dp.size == 0 or dp.max() <= tol, f"{what}: position parity {dp.max():.3e} > {tol}"
