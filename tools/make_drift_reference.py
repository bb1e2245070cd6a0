"""CPU-oracle reference curves for BASELINE.json config 3: N = 262,144 Plummer-like cloud, W = 64,
scale a = W/6, seed 42, 1,000 steps; records KE(t) = 1/2 sum |v|^2 and P(t) = sum v every step.

Runs for about an hour on 8 cores; the result is cached in tests/golden/drift_config3_oracle.npz so
that the GPU drift check never has to re-run the CPU side.  Usage:
    python tools/make_drift_reference.py [n=262144] [steps=1000] [threads=6] [acc64=0]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
import particle_3d as p3
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 6
acc64 = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
W = 64.0
prm = dict(p3.default_params_dict(), world_size=W)
cur = p3.generate_plummer(W, n, W / 6, seed=42)
TS = float(np.float32(1.0 / 60.0))
tag = f"n{n}" + ("_acc64" if acc64 else "")
out = os.path.join(ROOT, "tests", "golden", f"drift_config3_oracle_{tag}.npz")
ke, mom, t0 = [], [], time.time()
for s in range(steps):
    cur = O.update(prm, TS, cur, mode=O.IDEAL, acc64=acc64, nthreads=threads)["out"]
    v = np.stack([cur["vx"], cur["vy"], cur["vz"]], 1).astype(np.float64)
    ke.append(0.5 * (v ** 2).sum())
    mom.append(v.sum(0))
    if (s + 1) % 25 == 0 or s + 1 == steps:
        np.savez_compressed(out, ke=np.array(ke), mom=np.array(mom), n=n, world_size=W, scale_a=W / 6, seed=42,
                            steps_done=s + 1, mode="ideal", acc64=acc64)
        print(f"step {s+1}/{steps}  ke={ke[-1]:.6e}  elapsed {time.time()-t0:.0f}s", flush=True)
