"""Long-run drift (north_star: "over 1,000 steps the total-energy/momentum drift must match the reference's
drift within a stated bound").

The system is dissipative, driven and chaotic (asymmetric forces, src/bin/main.rs:133-139): two f32 runs that
differ only in summation order decorrelate after ~150 steps (measured at BASELINE.json config 3 scale in
profiles/r01_drift_config3.json: pair kernel, cell list and CPU oracle all drift apart from each other at the
same rate).  The stated bound is therefore calibrated on the oracle's own sensitivity: the CPU oracle run twice,
with f32 and with f64 force accumulation, gives the envelope E(t) = |KE_f32 - KE_f64| / KE; the GPU must stay
within max(20 * E(t), floor(t)) of the f32 oracle, floor = 1e-6 up to 50 steps and 1e-4 * (t/100)^2 later, and the
time-averaged kinetic energy over each 100-step window must agree within the window's own fluctuation (three
fluctuations for the last window, which lies past the decorrelation time: see the comment at the assertion).
"""
import numpy as np
import pytest

import particle_3d as p3
from particle_3d import _abi
from oracle import oracle as O

from helpers import vel

pytestmark = pytest.mark.gpu
TS = float(np.float32(1.0 / 60.0))
N, W, STEPS = 4096, 24.0, 400


def _diag(a):
    v = vel(a)
    return 0.5 * (v ** 2).sum(), v.sum(0)


@pytest.fixture(scope="module")
def oracle_curves(default_params):
    prm = dict(default_params, world_size=W)
    start = p3.generate_plummer(W, N, W / 6, seed=42)
    out = {}
    for name, acc64 in (("f32", False), ("f64acc", True)):
        cur, ke, mom = start.copy(), [], []
        for _ in range(STEPS):
            cur = O.update(prm, TS, cur, mode=O.IDEAL, acc64=acc64)["out"]
            k, m = _diag(cur)
            ke.append(k)
            mom.append(m)
        out[name] = (np.array(ke), np.array(mom))
    return prm, start, out


@pytest.mark.parametrize("kernel", [_abi.FORCE_PAIR, _abi.FORCE_CELLS, _abi.FORCE_REFERENCE_ORDER],
                         ids=["pair", "cells", "reference_order"])
def test_energy_and_momentum_drift_match_the_oracle(oracle_curves, kernel):
    prm, start, ref = oracle_curves
    ke_ref, p_ref = ref["f32"]
    ke_64, _ = ref["f64acc"]
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    eng.upload(start, 5)
    P = p3.Engine.make_params(**prm)
    ke, mom = [], []
    for _ in range(STEPS):
        eng.step(P, TS, 1)
        d = eng.diagnostics()
        ke.append(d["ke"])
        mom.append(d["p"])
    eng.close()
    ke, mom = np.array(ke), np.array(mom)
    env = np.abs(ke_ref - ke_64) / ke_ref
    env = np.maximum.accumulate(env)  # the envelope only grows
    rel = np.abs(ke - ke_ref) / ke_ref
    vrms = np.sqrt(2 * ke_ref / N)
    prel = np.abs(mom - p_ref).max(1) / (N * vrms)
    for t in (1, 10, 50, 100, 200, 300, 400):
        floor = 1e-6 if t <= 50 else 1e-4 * (t / 100.0) ** 2
        bound = max(20 * env[t - 1], floor)
        assert rel[t - 1] <= bound, f"KE drift at step {t}: {rel[t-1]:.3e} > {bound:.3e} (oracle envelope {env[t-1]:.3e})"
        assert prel[t - 1] <= bound, f"momentum drift at step {t}: {prel[t-1]:.3e} > {bound:.3e}"
    # time-averaged energy per 100-step window agrees within the window's own fluctuation.  The pair kernel adds
    # forces with floating-point atomics, so two runs of the SAME binary differ in summation order and decorrelate
    # like any other pair of implementations: over 16 repeated runs the first three windows agreed with the oracle
    # to < 0.005 fluctuations, the last one (steps 300-400, past the decorrelation time) scattered between 0.17
    # and 1.2 fluctuations — hence three fluctuations there.
    for a in range(0, STEPS, 100):
        m_gpu, m_ref, s_ref = ke[a:a + 100].mean(), ke_ref[a:a + 100].mean(), ke_ref[a:a + 100].std()
        slack = 3.0 if a >= 300 else 1.0
        assert abs(m_gpu - m_ref) <= slack * max(s_ref, 0.02 * m_ref), (a, m_gpu, m_ref, s_ref)
    print(f"\nkernel {kernel}: KE rel diff at 10/50/100/200/400 = "
          + ", ".join(f"{rel[t-1]:.2e}" for t in (10, 50, 100, 200, 400))
          + " | oracle f32-vs-f64 envelope = " + ", ".join(f"{env[t-1]:.2e}" for t in (10, 50, 100, 200, 400)))


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json config 3 at full size: N = 262,144 Plummer-like cloud (a = W/6, W = 64, seed 42), 1,000 steps.
# The CPU side is cached (an hour of host time): tests/golden/drift_config3_oracle_n262144.npz holds the oracle's
# KE(t) and sum v(t) in the reference's arithmetic (f32 force sums), ..._acc64.npz the same run with f64 force sums;
# both made by tools/make_drift_reference.py.  Stated bound, asserted at t = 1, 10, 60, 100, 300, 1000:
#     |KE_gpu(t) - KE_ref(t)| / KE_ref(t)  <=  max(K * E_ke(t), 1e-6)          K = 8
#     |P_gpu(t) - P_ref(t)|_inf / (N v_rms) <=  max(K * E_p(t),  1e-6)
# where E(t) is the running maximum of the oracle's OWN deviation between its f32 and f64-accumulate runs: the
# reference's sensitivity to nothing but summation order.  The divergence grows like exp(0.084 t) until it
# saturates near step 300, so K = 8 is a shift of 25 steps along that exponential; the 1e-6 floor covers the first
# ~60 steps, where the envelope (1e-9) is far below what ANY second f32 implementation can reach (the kernels use
# rsqrt and a different summation order: 6e-8 measured).  E(t) saturates (0.24 at step 300, 0.46 at 600: two runs of
# the ORACLE differ that much), so at t = 300 and 1000 the pointwise bound only says "no blow-up"; what is comparable
# past the decorrelation time is statistics: the mean KE of every 100-step window is held to
# max(K * the oracle's own window deviation, 1e-5), and to a factor 2.5 of the oracle's window mean throughout.
CFG3_N, CFG3_W, CFG3_STEPS, CFG3_K = 262144, 64.0, 1000, 8.0
CFG3_MARKS = (1, 10, 60, 100, 300, 1000)


def _running_max(x):
    return np.maximum.accumulate(np.asarray(x, dtype=np.float64))


@pytest.mark.parametrize("kernel", [_abi.FORCE_PAIR, _abi.FORCE_CELLS], ids=["pair", "cells"])
def test_config3_1000_step_drift_matches_the_oracle(default_params, kernel):
    import os

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    f32 = np.load(os.path.join(gold, f"drift_config3_oracle_n{CFG3_N}.npz"))
    f64 = np.load(os.path.join(gold, f"drift_config3_oracle_n{CFG3_N}_acc64.npz"))
    assert int(f32["steps_done"]) == CFG3_STEPS and int(f64["steps_done"]) == CFG3_STEPS, "golden curves are incomplete"
    ke_ref, p_ref, ke_64, p_64 = f32["ke"], f32["mom"], f64["ke"], f64["mom"]
    vrms = np.sqrt(2.0 * ke_ref / CFG3_N)
    env_ke = _running_max(np.abs(ke_ref - ke_64) / ke_ref)
    env_p = _running_max(np.abs(p_ref - p_64).max(1) / (CFG3_N * vrms))

    prm = dict(default_params, world_size=CFG3_W)
    P = p3.Engine.make_params(**prm)
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    eng.upload(p3.generate_plummer(CFG3_W, CFG3_N, CFG3_W / 6, seed=42), 5)
    ke, mom = np.zeros(CFG3_STEPS), np.zeros((CFG3_STEPS, 3))
    for s in range(CFG3_STEPS):
        eng.step(P, TS, 1)
        d = eng.diagnostics()
        ke[s], mom[s] = d["ke"], d["p"]
    assert d["count"] == CFG3_N
    eng.close()

    rel_ke = np.abs(ke - ke_ref) / ke_ref
    rel_p = np.abs(mom - p_ref).max(1) / (CFG3_N * vrms)
    report = []
    for t in CFG3_MARKS:
        b_ke, b_p = max(CFG3_K * env_ke[t - 1], 1e-6), max(CFG3_K * env_p[t - 1], 1e-6)
        report.append(f"t={t}: KE {rel_ke[t-1]:.2e} (bound {b_ke:.2e}), P {rel_p[t-1]:.2e} (bound {b_p:.2e})")
        assert rel_ke[t - 1] <= b_ke, f"KE drift at step {t}: {rel_ke[t-1]:.3e} > {b_ke:.3e} (oracle envelope {env_ke[t-1]:.3e})"
        assert rel_p[t - 1] <= b_p, f"momentum drift at step {t}: {rel_p[t-1]:.3e} > {b_p:.3e} (oracle envelope {env_p[t-1]:.3e})"
    w = 100
    means = lambda x: np.asarray(x).reshape(-1, w).mean(1)
    m_gpu, m_ref, m_64 = means(ke), means(ke_ref), means(ke_64)
    env_w = _running_max(np.abs(m_ref - m_64) / m_ref)
    for k in range(CFG3_STEPS // w):
        dev, bound = abs(m_gpu[k] - m_ref[k]) / m_ref[k], max(CFG3_K * env_w[k], 1e-5)
        report.append(f"window {k*w}-{(k+1)*w}: mean KE dev {dev:.2e} (bound {bound:.2e})")
        assert dev <= bound, f"window {k*w}-{(k+1)*w}: mean KE deviates by {dev:.3e} > {bound:.3e}"
        assert 0.4 <= m_gpu[k] / m_ref[k] <= 2.5, f"window {k*w}-{(k+1)*w}: mean KE {m_gpu[k]:.4e} vs oracle {m_ref[k]:.4e}"
    print("\n" + "\n".join(report))
