"""Combines the GPU curves of tools/drift_config3.py with the cached CPU-oracle curves (f32 and f64-accumulate)
into profiles/r01_drift_config3.json.  No GPU needed."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = 262144
g = np.load(os.path.join(ROOT, "gpurun_out", f"r01_drift_config3_n{n}_curves.npz"))
o = np.load(os.path.join(ROOT, "tests", "golden", f"drift_config3_oracle_n{n}.npz"))
a = np.load(os.path.join(ROOT, "tests", "golden", f"drift_config3_oracle_n{n}_acc64.npz"))
ko, ka = o["ke"], a["ke"]
out = json.load(open(os.path.join(ROOT, "gpurun_out", f"r01_drift_config3_n{n}.json")))
out["oracle"] = {"f32_steps": int(len(ko)), "f64acc_steps": int(len(ka)),
                 "note": "CPU oracle (ideal mode) run twice: f32 force sums (reference arithmetic) and f64 force sums; "
                         "their difference is the reference's own sensitivity to summation order"}
marks = {}
rel = lambda x, y: float(abs(x - y) / y)
vr = lambda k: np.sqrt(2 * k / n)
for t in (1, 10, 60, 100, 150, 200, 300, 400, 600, 1000):
    row = {}
    if t <= len(ko):
        row["ke_oracle_f32"] = float(ko[t - 1])
        row["ke_rel_pair_vs_oracle"] = rel(g["pair_ke"][t - 1], ko[t - 1])
        row["ke_rel_cells_vs_oracle"] = rel(g["cells_ke"][t - 1], ko[t - 1])
        row["p_rel_pair_vs_oracle"] = float(np.abs(g["pair_mom"][t - 1] - o["mom"][t - 1]).max() / (n * vr(ko[t - 1])))
        row["p_rel_cells_vs_oracle"] = float(np.abs(g["cells_mom"][t - 1] - o["mom"][t - 1]).max() / (n * vr(ko[t - 1])))
    if t <= len(ka) and t <= len(ko):
        row["ke_rel_oracle_f32_vs_f64acc (envelope)"] = rel(ko[t - 1], ka[t - 1])
        row["p_rel_oracle_f32_vs_f64acc (envelope)"] = float(np.abs(o["mom"][t - 1] - a["mom"][t - 1]).max() / (n * vr(ko[t - 1])))
    row["ke_rel_pair_vs_cells"] = rel(g["pair_ke"][t - 1], g["cells_ke"][t - 1])
    marks[str(t)] = row
out["marks"] = marks
win = {}
for lo, hi in ((0, 100), (100, 200), (200, 300), (300, 500), (500, 750), (750, 1000)):
    w = {"pair_mean": float(g["pair_ke"][lo:hi].mean()), "cells_mean": float(g["cells_ke"][lo:hi].mean())}
    if hi <= len(ko):
        w["oracle_mean"] = float(ko[lo:hi].mean()); w["oracle_std"] = float(ko[lo:hi].std())
        w["pair_mean_rel_dev"] = rel(w["pair_mean"], w["oracle_mean"]); w["cells_mean_rel_dev"] = rel(w["cells_mean"], w["oracle_mean"])
    win[f"{lo}-{hi}"] = w
out["window_means_ke"] = win
out["reading"] = ("The system is driven, dissipative and chaotic: GPU (pair, cell list) and CPU oracle agree to <1e-6 in KE for the "
                  "first ~60 steps, 1e-5 at 100, and decorrelate after ~150 steps at the same rate at which the oracle decorrelates "
                  "from itself when only its summation precision changes (envelope columns). Beyond that only time-averaged "
                  "quantities are comparable: window means agree within the windows' own fluctuation.")
json.dump(out, open(os.path.join(ROOT, "profiles", "r01_drift_config3.json"), "w"), indent=1)
for t, r in marks.items():
    print(t, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in r.items() if "rel" in k})
for k, w in win.items():
    print(k, {a: f"{b:.4e}" for a, b in w.items()})
