"""Small end-to-end case for compute-sanitizer: every kernel of the engine once, N = 6,000."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200")); sys.path.insert(0, ROOT)
import particle_3d as p3
from particle_3d import _abi

W, n = 18.2, 6000
prm = dict(p3.default_params_dict(), world_size=W)
P = p3.Engine.make_params(**prm)
parts = p3.generate_particles(W, n, seed=42)
eng = p3.Engine(0)
for kernel in (_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS):
    for faithful in (0, 1):
        eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        eng.set_option(_abi.OPT_FAITHFUL, faithful)
        out = eng.update(P, 1 / 60, parts)
        eng.upload(parts, 5)
        eng.step(P, 1 / 60, 2)
        d = eng.diagnostics()
        f = eng.download_forces()
        assert np.isfinite(f).all() and d["count"] == n
eng.set_option(_abi.OPT_FAITHFUL, 0)
eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)
eng.set_shard(0, 2); eng.upload(parts, 5); eng.shard_force(P); eng.shard_integrate(P, 1 / 60); eng.shard_commit(); eng.sync()
far = parts.copy(); far["px"][::5] += 40.0
eng.set_shard(0, 1)
eng.update(P, 1 / 60, far)  # out-of-box fallback path
eng.close()
print("sanitize case done")
