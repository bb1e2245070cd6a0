"""Memory safety without compute-sanitizer (closed on the target pool): the engine is rebuilt with
-DP3D_BOUNDS_CHECK (`__graft_entry__.build()` -> build/checked/libp3d.so), where every data-dependent slot /
cell index of the kernels is compared with its extent and violations are counted, and a run through every
kernel at awkward sizes must count none.  Runs in a child process because the library path is fixed at import."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "build", "checked", "libp3d.so")

CHILD = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1] + "/3d-particle-simulation-_b200"); sys.path.insert(0, sys.argv[1])
import particle_3d as p3
from particle_3d import _abi
assert _abi.LIB_PATH.endswith("build/checked/libp3d.so"), _abi.LIB_PATH
rng = np.random.default_rng(11)
ts = 1.0 / 60.0
total = 0
for n, W, T in ((1, 8.0, 1), (127, 8.0, 3), (129, 8.0, 5), (257, 9.0, 64), (4097, 12.0, 5), (6000, 18.2, 7), (70001, 41.0, 5)):
    A = rng.uniform(-1, 1, T * T).astype(np.float32)
    prm = dict(p3.default_params_dict(), world_size=W, id_count=T, attraction_matrix=list(A))
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(W, n, seed=n, id_count=T)
    eng = p3.Engine(0)
    assert eng.debug_bounds_violations() is not None, "not a self-checking build"
    for kernel in (_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS):
        if kernel == _abi.FORCE_REFERENCE_ORDER and n > 10000:
            continue
        for block in (128, 256):
            for faithful in (0, 1):
                eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
                eng.set_option(_abi.OPT_BLOCK_SIZE, block)
                eng.set_option(_abi.OPT_FAITHFUL, faithful)
                out = eng.update(P, ts, parts)
                eng.upload(out, T)
                eng.step(P, ts, 7)          # graph replay for >= 6 steps
                eng.download(); eng.download_forces(); eng.download_render(W); eng.diagnostics()
    eng.set_option(_abi.OPT_FAITHFUL, 0)
    for world in (2, 3, 8):                  # emulated shards on one GPU
        for rank in range(world):
            eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)
            eng.set_shard(rank, world); eng.upload(parts, T)
            eng.shard_force(P); eng.shard_integrate(P, ts); eng.shard_commit(); eng.sync()
    eng.set_shard(0, 1)
    far = parts.copy(); far["px"][::5] += 3 * W; far["py"][::7] -= 2 * W   # out-of-box input: fallback paths
    for kernel in (_abi.FORCE_PAIR, _abi.FORCE_CELLS):
        if n > 10000 and kernel == _abi.FORCE_PAIR:
            continue
        eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        eng.update(P, ts, far)
    v = eng.debug_bounds_violations()
    print(f"n={n} W={W} T={T}: violations so far {v}", flush=True)
    total = v
    eng.close()
print("VIOLATIONS", total)
"""


def test_no_out_of_range_index_in_any_kernel():
    if not os.path.exists(CHECKED):
        pytest.skip("build/checked/libp3d.so missing: run __graft_entry__.build()")
    env = dict(os.environ, P3D_LIB=CHECKED)
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "VIOLATIONS 0" in r.stdout, r.stdout[-2000:]
