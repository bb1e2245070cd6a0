"""KE(t) curves for a small Plummer cloud from any backend, to separate chaos from systematic differences.
usage: drift_small.py <backend: oracle|oracle64|pair|cells|ref> <seed> [n=16384] [W=32] [steps=1000] [outdir]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
import particle_3d as p3
backend, seed = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
W = float(sys.argv[4]) if len(sys.argv) > 4 else 32.0
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
outdir = sys.argv[6] if len(sys.argv) > 6 else os.path.join(ROOT, "build")
TS = float(np.float32(1 / 60))
prm = dict(p3.default_params_dict(), world_size=W)
cur = p3.generate_plummer(W, n, W / 6, seed=seed)
ke = []
t0 = time.time()
if backend.startswith("oracle"):
    from oracle import oracle as O
    for s in range(steps):
        cur = O.update(prm, TS, cur, mode=O.IDEAL, acc64=(backend == "oracle64"))["out"]
        v = np.stack([cur["vx"], cur["vy"], cur["vz"]], 1).astype(np.float64)
        ke.append(0.5 * (v ** 2).sum())
else:
    from particle_3d import _abi
    k = {"pair": _abi.FORCE_PAIR, "cells": _abi.FORCE_CELLS, "ref": _abi.FORCE_REFERENCE_ORDER}[backend]
    eng = p3.Engine(0); eng.set_option(_abi.OPT_FORCE_KERNEL, k); eng.upload(cur, 5)
    P = p3.Engine.make_params(**prm)
    for s in range(steps):
        eng.step(P, TS, 1); ke.append(eng.diagnostics()["ke"])
ke = np.array(ke)
np.save(os.path.join(outdir, f"drift_small_{backend}_s{seed}_n{n}.npy"), ke)
w = [ke[a:a + 250].mean() for a in range(0, steps, 250)]
print(f"{backend} seed {seed}: {time.time()-t0:.0f}s window means " + " ".join(f"{x:.4e}" for x in w), flush=True)
