"""BASELINE.json config 3 on the GPU: N = 262,144 Plummer-like cloud, W = 64, a = W/6, seed 42, 1,000 steps.
Records KE(t) = 1/2 sum|v|^2 and P(t) = sum v per step for each force kernel and compares with the cached CPU
oracle curve (tests/golden/drift_config3_oracle_n262144.npz, made by tools/make_drift_reference.py).
Writes profiles/r01_drift_config3.json.  Usage: python tools/drift_config3.py [n=262144] [steps=1000]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
import particle_3d as p3
from particle_3d import _abi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
W = 64.0
TS = float(np.float32(1.0 / 60.0))
prm = dict(p3.default_params_dict(), world_size=W)
P = p3.Engine.make_params(**prm)
start = p3.generate_plummer(W, n, W / 6, seed=42)

runs = {}
for name, kernel, nsteps in (("pair", _abi.FORCE_PAIR, steps), ("cells", _abi.FORCE_CELLS, steps),
                             ("reference_order", _abi.FORCE_REFERENCE_ORDER, min(steps, 60))):
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    eng.upload(start, 5)
    ke, mom = [], []
    t0 = time.time()
    for s in range(nsteps):
        eng.step(P, TS, 1)
        d = eng.diagnostics()
        ke.append(d["ke"])
        mom.append(d["p"])
    runs[name] = {"ke": np.array(ke), "mom": np.array(mom), "seconds": time.time() - t0}
    print(f"{name}: {nsteps} steps in {runs[name]['seconds']:.1f}s  KE[-1]={ke[-1]:.6e}", flush=True)
    eng.close()

out = {"config": {"n": n, "world_size": W, "scale_a": W / 6, "seed": 42, "steps": steps, "ts": TS},
       "marks": {}, "kernels": {k: {"steps": len(v["ke"]), "seconds": v["seconds"]} for k, v in runs.items()}}
ref_path = os.path.join(ROOT, "tests", "golden", f"drift_config3_oracle_n{n}.npz")
ref = np.load(ref_path) if os.path.exists(ref_path) else None
vrms = lambda ke_t: np.sqrt(2.0 * ke_t / n)
for t in (1, 10, 60, 100, 300, 1000):
    if t > steps:
        continue
    row = {"ke_pair": float(runs["pair"]["ke"][t - 1]), "ke_cells": float(runs["cells"]["ke"][t - 1])}
    row["ke_rel_pair_vs_cells"] = abs(row["ke_pair"] - row["ke_cells"]) / row["ke_cells"]
    dp = np.abs(runs["pair"]["mom"][t - 1] - runs["cells"]["mom"][t - 1]).max()
    row["p_rel_pair_vs_cells"] = float(dp / (n * vrms(row["ke_cells"])))
    if t <= len(runs["reference_order"]["ke"]):
        k0 = float(runs["reference_order"]["ke"][t - 1])
        row["ke_rel_pair_vs_reference_order"] = abs(row["ke_pair"] - k0) / k0
    if ref is not None and t <= int(ref["steps_done"]):
        k0 = float(ref["ke"][t - 1])
        row["ke_oracle"] = k0
        row["ke_rel_pair_vs_oracle"] = abs(row["ke_pair"] - k0) / k0
        row["ke_rel_cells_vs_oracle"] = abs(row["ke_cells"] - k0) / k0
        row["p_rel_pair_vs_oracle"] = float(np.abs(runs["pair"]["mom"][t - 1] - ref["mom"][t - 1]).max() / (n * vrms(k0)))
    out["marks"][str(t)] = row
    print(t, json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r01_drift_config3_n{n}.json"), "w"), indent=1)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"r01_drift_config3_n{n}_curves.npz"),
                    **{f"{k}_{q}": v[q] for k, v in runs.items() for q in ("ke", "mom")})
