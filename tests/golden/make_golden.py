"""Generates tests/golden/*.npz from the CPU oracle (oracle/p3d_oracle.c).

The reference ships no tests or fixtures and cannot be built here (no Rust
toolchain), so these vectors are outputs of OUR restatement of src/lib.rs on seeded inputs.  They
pin the oracle against regressions and give the GPU tests a fixed target.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
import particle_3d as p3  # host-only use: seeded scene generator
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
TS = np.float32(1.0 / 60.0)  # src/bin/main.rs:164,194


def diag(a):
    v = np.stack([a["vx"], a["vy"], a["vz"]], 1).astype(np.float64)
    return 0.5 * (v ** 2).sum(), v.sum(0)


def main():
    prm = p3.default_params_dict()
    # config 1: default scene, N = 1000, 100 steps, both oracle modes
    start = p3.generate_particles(prm["world_size"], 1000, seed=42)
    out = {"start": start}
    for mode, name in ((O.IDEAL, "ideal"), (O.FAITHFUL, "faithful")):
        cur = start.copy()
        ke, mom = [], []
        for step in range(100):
            r = O.update(prm, float(TS), cur, mode=mode, want_force=(step == 0), want_affected=(step == 0))
            if step == 0:
                out[f"{name}_step1"] = r["out"]
                out[f"{name}_force1"] = r["force"]
                if mode == O.FAITHFUL:
                    out["affected1"] = r["affected"]
                    out["stats1"] = np.array([r["stats"][k] for k in ("candidates", "in_radius", "nonzero", "dup_bucket_queries", "affected")], dtype=np.int64)
            cur = r["out"]
            k, m = diag(cur)
            ke.append(k)
            mom.append(m)
        out[f"{name}_step100"] = cur
        out[f"{name}_ke"] = np.array(ke)
        out[f"{name}_mom"] = np.array(mom)
    np.savez_compressed(os.path.join(HERE, "default_scene_n1000_seed42.npz"), **out)

    # config 2 (reduced): N = 4096 uniform at density 1, one step, walls + gravity variant too
    W = 16.0
    start = p3.generate_particles(W, 4096, seed=7)
    p2 = dict(prm, world_size=W)
    a = O.update(p2, float(TS), start, mode=O.IDEAL)["out"]
    p2w = dict(p2, walls=True, acceleration=(0.0, -9.8, 0.0))
    b = O.update(p2w, float(TS), start, mode=O.IDEAL)["out"]
    np.savez_compressed(os.path.join(HERE, "uniform_n4096_seed7.npz"), start=start, ideal_step1=a, ideal_walls_gravity_step1=b)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
