"""The N>1 host path on CPU: ShardedStepper's collective sequence over gloo, world_size 2 and 3.

The GPU engine is replaced by a stand-in with the same shard API whose arithmetic is the CPU oracle
(tests may use the oracle; the product never does).  What is under test is the driver: which ranks
integrate which slots, the force all-reduce, the position all-gather into the right offsets, and the
buffer swap at commit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import particle_3d as p3
from particle_3d.sharded import ShardedStepper, shard_rows, shard_slot_range
from oracle import oracle as O

TS = float(np.float32(1.0 / 60.0))
BLOCK = 128


class HostShardEngine:
    """Same shard API as particle_3d.Engine; state in torch CPU tensors; arithmetic = oracle."""

    def __init__(self, prm, parts, rank, world):
        self.prm, self.rank, self.world = prm, rank, world
        n = len(parts)
        self.n = n
        self.n_blocks = -(-n // (BLOCK * world)) * world  # padded like build_layout
        ns = self.n_blocks * BLOCK
        self.ids = parts["id"].copy()
        self.pos = [torch.zeros(ns, 4), torch.zeros(ns, 4)]
        self.vel_t = torch.zeros(ns, 4)          # like P3D_BUF_VEL: float4 per slot
        self.vel = self.vel_t.numpy()[:, :3]     # view: the oracle-side code below writes through it
        self.force = torch.zeros(ns, 4)
        self.cur = 0
        self.pos[0][:n, 0] = torch.from_numpy(parts["px"].copy())
        self.pos[0][:n, 1] = torch.from_numpy(parts["py"].copy())
        self.pos[0][:n, 2] = torch.from_numpy(parts["pz"].copy())
        self.vel[:n] = np.stack([parts["vx"], parts["vy"], parts["vz"]], 1)

    def tensors(self):
        return {"pos": self.pos[self.cur], "pos_next": self.pos[self.cur ^ 1], "force": self.force, "vel": self.vel_t}

    def _particles(self):
        a = np.zeros(self.n, O.PARTICLE)
        p = self.pos[self.cur].numpy()
        a["px"], a["py"], a["pz"] = p[: self.n, 0], p[: self.n, 1], p[: self.n, 2]
        a["vx"], a["vy"], a["vz"] = self.vel[: self.n, 0], self.vel[: self.n, 1], self.vel[: self.n, 2]
        a["id"] = self.ids
        return a

    def shard_force(self, params):
        f = O.update(self.prm, TS, self._particles(), mode=O.IDEAL, want_force=True)["force"]
        mine = np.zeros(self.n_blocks * BLOCK, bool)
        for row in shard_rows(self.n_blocks, self.rank, self.world):
            mine[row * BLOCK:(row + 1) * BLOCK] = True
        self.force.zero_()
        part = np.where(mine[: self.n, None], f, 0.0).astype(np.float32)  # partial forces: own rows only
        self.force[: self.n, :3] = torch.from_numpy(part)

    def shard_range(self):
        return shard_slot_range(self.n_blocks, BLOCK, self.rank, self.world)

    def shard_integrate(self, params, ts):
        s0, s1 = self.shard_range()
        e = min(s1, self.n)
        nxt = self.pos[self.cur ^ 1]
        if e > s0:
            a = self._particles()[s0:e]
            out = O.integrate(self.prm, ts, a, self.force.numpy()[s0:e, :3])
            nxt[s0:e, 0] = torch.from_numpy(out["px"].copy())
            nxt[s0:e, 1] = torch.from_numpy(out["py"].copy())
            nxt[s0:e, 2] = torch.from_numpy(out["pz"].copy())
            self.vel[s0:e] = np.stack([out["vx"], out["vy"], out["vz"]], 1)

    def shard_commit(self):
        self.cur ^= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, steps, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = dict(p3.default_params_dict(), world_size=12.0)
    parts = p3.generate_particles(12.0, n, seed=4)
    eng = HostShardEngine(prm, parts, rank, world)
    st = ShardedStepper(eng, dist, rank, world, eng.tensors)
    st.step(None, TS, steps)
    s0, s1 = eng.shard_range()
    np.save(os.path.join(outdir, f"pos_{rank}.npy"), eng.pos[eng.cur].numpy()[:n, :3])
    np.save(os.path.join(outdir, f"vel_{rank}.npy"), eng.vel[:n])
    np.save(os.path.join(outdir, f"rng_{rank}.npy"), np.array([s0, s1, st.collectives]))
    # a re-upload puts the engine back on its first position buffer: after an odd number of steps the stepper's
    # cached buffer views are for the other parity and reset() must drop them
    eng.__init__(prm, parts, rank, world)
    st.reset()
    st.step(None, TS, 2)
    np.save(os.path.join(outdir, f"pos2_{rank}.npy"), eng.pos[eng.cur].numpy()[:n, :3])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stepper_over_gloo(tmp_path, world):
    n, steps = 1500, 3
    mp.spawn(_worker, args=(world, _free_port(), n, steps, str(tmp_path)), nprocs=world, join=True)
    prm = dict(p3.default_params_dict(), world_size=12.0)
    ref = p3.generate_particles(12.0, n, seed=4)
    for _ in range(steps):
        ref = O.update(prm, TS, ref, mode=O.IDEAL)["out"]
    refp = np.stack([ref["px"], ref["py"], ref["pz"]], 1)
    refv = np.stack([ref["vx"], ref["vy"], ref["vz"]], 1)
    covered = np.zeros(n, bool)
    for r in range(world):
        pos = np.load(tmp_path / f"pos_{r}.npy")
        vel = np.load(tmp_path / f"vel_{r}.npy")
        s0, s1, ncoll = np.load(tmp_path / f"rng_{r}.npy")
        assert ncoll == 3 * steps  # all-reduce(forces) + all-gather(positions) + all-gather(velocities) per step
        # every rank holds every position after the all-gather, bit-identical to the 1-rank oracle run
        assert np.array_equal(pos, refp)
        e = min(s1, n)
        # velocities are gathered too: every rank holds the whole state (any rank can serve any part of the array)
        assert np.array_equal(vel, refv)
        covered[s0:e] = True
    assert covered.all()
    ref2 = p3.generate_particles(12.0, n, seed=4)
    for _ in range(2):
        ref2 = O.update(prm, TS, ref2, mode=O.IDEAL)["out"]
    for r in range(world):  # the run after re-upload + reset()
        assert np.array_equal(np.load(tmp_path / f"pos2_{r}.npy"), np.stack([ref2["px"], ref2["py"], ref2["pz"]], 1))


def test_shard_arithmetic():
    for n_blocks, world in ((8, 2), (9, 3), (16, 8), (5, 5)):
        ranges = [shard_slot_range(n_blocks, BLOCK, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n_blocks * BLOCK
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        rows = np.concatenate([shard_rows(n_blocks, r, world) for r in range(world)])
        assert sorted(rows) == list(range(n_blocks))


def test_part_range_tiles_the_callers_array():
    """The split of the caller's array over ranks (sharded upload / download, multi-device handle): equal parts of
    ceil(n / world), the last ones short or empty, together exactly [0, n)."""
    from particle_3d.sharded import part_range

    for n in (0, 1, 7, 1000, 1048576, 1048577):
        for world in (1, 2, 3, 8):
            at = 0
            per = -(-n // world)
            for r in range(world):
                c0, c1 = part_range(n, r, world)
                assert c0 == at and c0 <= c1 <= n and c1 - c0 <= per
                assert (c1 - c0 == per) or c1 == n  # only the tail may be short
                at = c1
            assert at == n
