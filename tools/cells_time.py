"""Device-resident cell-list stepping (P3D_FORCE_CELLS through p3d_step: CUDA-graph replay, slots kept in cell
order), wall clock per step.  Usage: python tools/cells_time.py [n W plummer steps]..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200"))
sys.path.insert(0, ROOT)
import particle_3d as p3
from particle_3d import _abi

cases = ((1048576, 101.6, False, 200), (262144, 64.0, True, 200), (262144, 64.0, False, 200), (4194304, 161.3, False, 100))
if len(sys.argv) > 1:
    a = sys.argv[1:]
    cases = [(int(a[i]), float(a[i + 1]), bool(int(a[i + 2])), int(a[i + 3])) for i in range(0, len(a), 4)]
for n, W, pl, steps in cases:
    prm = dict(p3.default_params_dict(), world_size=W)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_plummer(W, n, W / 6, 42) if pl else p3.generate_particles(W, n, 42)
    e = p3.Engine(0)
    e.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
    e.upload(parts, 5)
    e.step(P, 1 / 60, 10)
    e.sync()
    t0 = time.perf_counter()
    e.step(P, 1 / 60, steps)
    e.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"resident cells n={n} plummer={pl}: {dt*1e3:.4f} ms/step", flush=True)
    e.close()
