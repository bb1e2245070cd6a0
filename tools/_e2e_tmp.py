import sys, os, time
sys.path.insert(0, "3d-particle-simulation-_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import particle_3d as p3
from particle_3d import _abi
n, W = 1048576, 101.6
prm = p3.default_params_dict(); prm["world_size"] = W
parts = p3.generate_particles(W, n, 42)
P = p3.Engine.make_params(**prm)
for kernel in (_abi.FORCE_PAIR, _abi.FORCE_CELLS):
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    hin = torch.empty(n * 28, dtype=torch.uint8).pin_memory(); hout = torch.empty(n * 28, dtype=torch.uint8).pin_memory()
    a_in = hin.numpy().view(_abi.PARTICLE); a_out = hout.numpy().view(_abi.PARTICLE)
    a_in[:] = parts
    for rep in range(3):
        t0 = time.perf_counter(); eng.upload(a_in, 5); eng.sync(); t1 = time.perf_counter()
        eng.step(P, 1/60, 1); eng.sync(); t2 = time.perf_counter()
        out = eng.download(); t3 = time.perf_counter()
    print(f"kernel {kernel}: upload {1e3*(t1-t0):.3f} ms  step {1e3*(t2-t1):.3f} ms  download(new array) {1e3*(t3-t2):.3f} ms", flush=True)
    eng.set_option(_abi.OPT_TIMING, 1)
    eng.update_into(P, 1/60, a_in, a_out)
    t0 = time.perf_counter(); eng.update_into(P, 1/60, a_in, a_out); t1 = time.perf_counter()
    print("  update_into wall", 1e3*(t1-t0), "timing", {k: round(v, 3) for k, v in eng.timing().items()}, flush=True)
    eng.close()
