"""Multi-GPU sharded stepping on real GPUs (skipped when fewer than 2 are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT, "3d-particle-simulation-_b200")); sys.path.insert(0, ROOT)
import particle_3d as p3
from particle_3d import _abi
from particle_3d.sharded import ShardedStepper, engine_tensors, exchange_peer_handles, sharded_upload, part_range
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
W, n, steps = 30.0, 27000, 3
prm = dict(p3.default_params_dict(), world_size=W)
parts = p3.generate_particles(W, n, seed=42)
P = p3.Engine.make_params(**prm)
eng = p3.Engine(local)
eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)
s = torch.cuda.Stream(device=local); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
eng.set_shard(rank, world)
eng.upload(parts, 5)
st = ShardedStepper(eng, dist, rank, world, lambda: engine_tensors(eng, local))
st.step(P, 1/60, steps)
torch.cuda.synchronize()
np.save(os.path.join(OUT, f"shard_{rank}.npy"), eng.download())
# the same run through the fused P2P kernel (no NCCL on the data path)
eng.upload(parts, 5)
exchange_peer_handles(eng, dist, world)
bar = torch.zeros(1, device=f"cuda:{local}")
st = ShardedStepper(eng, dist, rank, world, lambda: engine_tensors(eng, local), fused=True, barrier_tensor=bar)
st.step(P, 1/60, steps)
torch.cuda.synchronize()
np.save(os.path.join(OUT, f"fused_{rank}.npy"), eng.download())
# sharded host traffic: every rank uploads 1/world of the array, NCCL all-gathers the staging array, every rank
# reads its own part of the result back (same device buffers: the IPC mappings stay valid)
sharded_upload(eng, dist, rank, world, local, parts, 5)
st.reset()
st.step(P, 1/60, steps)
torch.cuda.synchronize()
c0, c1 = part_range(n, rank, world)
mine = np.empty(c1 - c0, dtype=_abi.PARTICLE)
eng.download_part_into(mine, c0)
np.save(os.path.join(OUT, f"part_{rank}.npy"), mine)
dist.barrier(); eng.ipc_close(); dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpu_sharded_step_matches_single_gpu_and_oracle(tmp_path, default_params):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(8, torch.cuda.device_count())  # every visible GPU of the box
    import particle_3d as p3
    from particle_3d import _abi
    from oracle import oracle as O
    from helpers import parity_errors

    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\nOUT = {str(tmp_path)!r}\n" + WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    W, n = 30.0, 27000
    prm = dict(default_params, world_size=W)
    ref = p3.generate_particles(W, n, seed=42)
    for _ in range(3):
        ref = O.update(prm, 1 / 60, ref, mode=O.IDEAL)["out"]
    for prefix in ("shard", "fused"):
        _check([np.load(tmp_path / f"{prefix}_{r}.npy") for r in range(world)], ref, W, prefix)
    # the parts the ranks read back, concatenated, are the whole updated array in the caller's order
    whole = np.concatenate([np.load(tmp_path / f"part_{r}.npy") for r in range(world)])
    assert whole.shape == ref.shape
    _check([whole], ref, W, "sharded upload + part download")


def _check(outs, ref, W, what):
    from helpers import parity_errors
    # both variants gather positions AND velocities: every rank holds the whole state after a step
    for r, got in enumerate(outs):
        assert np.array_equal(got["id"], ref["id"]), what
        dv, dp = parity_errors(got, ref, W)
        assert dp.max() < 5e-5 and dv.max() < 5e-5, (what, r, dv.max(), dp.max())
    for got in outs[1:]:
        assert got.tobytes() == outs[0].tobytes(), f"{what}: the ranks' copies of the state differ"
