"""Latency of the drop-in call p3d_update (host arrays in, host arrays out) at small particle counts, default kernel
selection (P3D_FORCE_AUTO), pageable and pinned host memory.  Usage: python tools/update_latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
sys.path.insert(0, ROOT)
import torch

import particle_3d as p3
from particle_3d import _abi

TS = 1.0 / 60.0
for n, W in ((1000, 10.0), (4096, 16.0), (16384, 25.4), (65536, 40.3), (262144, 64.0), (1048576, 101.6)):
    prm = dict(p3.default_params_dict(), world_size=W)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(W, n, seed=42)
    eng = p3.Engine(0)
    row = {}
    for name, pinned in (("pageable", False), ("pinned", True)):
        if pinned:
            hin = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
            hout = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
            a, b = hin.numpy().view(_abi.PARTICLE), hout.numpy().view(_abi.PARTICLE)
        else:
            a, b = np.empty(n, _abi.PARTICLE), np.empty(n, _abi.PARTICLE)
        a[:] = parts
        for _ in range(5):
            eng.update_into(P, TS, a, b)
        reps = 200 if n <= 65536 else 30
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.update_into(P, TS, a, b)
            a, b = b, a
        row[name] = (time.perf_counter() - t0) / reps * 1e6
    print(f"p3d_update n={n}: pageable {row['pageable']:.1f} us, pinned {row['pinned']:.1f} us per call", flush=True)
    eng.close()
