"""One-process-per-GPU stepping over torch.distributed (NCCL on GPUs; gloo in the CPU tests).

The path shards with ONE real exchange per step: every rank needs all positions
to evaluate its share of the pair interactions.  Because the pair kernel evaluates each unordered
block pair once and updates both particles, a rank's force pass yields partial forces for
particles owned by other ranks, so the step is

    shard_force      : this rank's block rows -> partial forces for all slots
    all_reduce(sum)  : forces                               (16 B per slot)
    shard_integrate  : this rank's slot range -> new positions of that range
    all_gather       : new positions                        (16 B per slot)
    shard_commit     : swap position buffers

Fused mode (GPUs of one node): both collectives move into ONE kernel, k_integrate_fused, which
sums the partial forces straight out of every peer's force buffer (P2P loads over NVLink),
integrates, and stores the new positions into every peer's position buffer (P2P stores).  NCCL is
then only used for two one-element all-reduces per step that act as cross-rank barriers.

The engine is anything with the `Engine` shard API (particle_3d.Engine on GPUs; the tests
substitute a CPU stand-in built on the oracle to cover the collective plumbing with gloo).
"""
from __future__ import annotations

import numpy as np

from . import _abi


class _CudaView:
    """Exposes an engine-owned device buffer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n_slots: int):
        self.__cuda_array_interface__ = {
            "shape": (n_slots, 4),
            "typestr": "<f4",
            "data": (ptr, False),
            "version": 2,
            "strides": None,
        }


def engine_tensors(engine, device_index: int):
    """torch views (n_slots x 4 float32) of the engine's POS, POS_NEXT and FORCE buffers."""
    import torch

    out = {}
    for name, which in (("pos", _abi.BUF_POS), ("pos_next", _abi.BUF_POS_NEXT), ("force", _abi.BUF_FORCE),
                        ("vel", _abi.BUF_VEL)):
        ptr, n = engine.device_buffer(which)
        out[name] = torch.as_tensor(_CudaView(ptr, n), device=f"cuda:{device_index}")
    return out


def exchange_peer_handles(engine, dist, world: int):
    """All-gathers every rank's CUDA IPC handles and opens the peers' buffers (fused mode)."""
    mine = engine.ipc_export()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    engine.ipc_import(world, b"".join(gathered))


class ShardedStepper:
    def __init__(self, engine, dist, rank: int, world: int, tensors_fn, fused: bool = False, barrier_tensor=None):
        """tensors_fn() -> dict(pos=, pos_next=, force=) of torch tensors aliasing the engine's
        CURRENT buffers (POS/POS_NEXT swap at every commit, so it is called once per parity)."""
        self.engine, self.dist, self.rank, self.world = engine, dist, rank, world
        self._tensors_fn = tensors_fn
        self._views = [None, None]
        self._parity = 0
        self.collectives = 0
        # fused mode: forces and positions travel through k_integrate_fused over peer memory; the only
        # collective left is a one-element all-reduce used as a stream-ordered cross-rank barrier
        self.fused = fused and world > 1
        self._bar = barrier_tensor
        # The collectives run on torch's current stream: the engine's kernels must be on the same stream, otherwise
        # the force all-reduce / the fused kernel's peer reads race with them.  (CPU stand-ins have no streams.)
        if hasattr(engine, "set_stream"):
            try:
                import torch

                if torch.cuda.is_available():
                    engine.set_stream(torch.cuda.current_stream().cuda_stream)
            except ImportError:
                pass

    def reset(self):
        """Call after a re-upload: the engine starts again from its first position buffer."""
        self._views = [None, None]
        self._parity = 0

    def _tensors(self):
        if self._views[self._parity] is None:
            self._views[self._parity] = self._tensors_fn()
        return self._views[self._parity]

    def step(self, params, ts: float, n_steps: int = 1):
        eng, dist = self.engine, self.dist
        for _ in range(n_steps):
            if self.fused:
                eng.shard_force(params)
                dist.all_reduce(self._bar)   # every rank's partial forces are complete
                eng.shard_integrate_fused(params, ts)
                dist.all_reduce(self._bar)   # every rank's position stores have landed
                eng.shard_commit()
                self.collectives += 2
                continue
            t = self._tensors()
            eng.shard_force(params)
            if self.world > 1:
                dist.all_reduce(t["force"], op=dist.ReduceOp.SUM)
                self.collectives += 1
            eng.shard_integrate(params, ts)
            if self.world > 1:
                s0, s1 = eng.shard_range()
                # the send shard is copied first: not every backend accepts an input that aliases the output
                dist.all_gather_into_tensor(t["pos_next"], t["pos_next"][s0:s1].clone())
                # velocities too, so that every rank holds the whole state (any rank can download any part)
                if "vel" in t:
                    dist.all_gather_into_tensor(t["vel"], t["vel"][s0:s1].clone())
                    self.collectives += 1
                self.collectives += 1
            eng.shard_commit()
            self._parity ^= 1


def part_range(n: int, rank: int, world: int):
    """Callers [begin, end) rank `rank` moves over its own PCIe link (equal parts of ceil(n / world))."""
    per = (n + world - 1) // world
    c0 = min(n, per * rank)
    return c0, min(n, c0 + per)


def sharded_upload(engine, dist, rank: int, world: int, device_index: int, particles, id_count: int):
    """Every rank copies only its 1/world of `particles` (a host array all ranks hold, or hold their part of) to its
    GPU, the staging arrays are all-gathered over NVLink, then every rank builds the layout.  Replaces `world` full
    host-to-device copies of the same array."""
    import torch

    n = particles.shape[0]
    c0, c1 = part_range(n, rank, world)
    engine.upload_part(particles[c0:c1], c0, n, id_count)
    if world > 1:
        ptr, cap = engine.device_buffer(_abi.BUF_AOS)
        per = cap // world

        class _V:
            __cuda_array_interface__ = {"shape": (cap * 7,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                        "strides": None}

        t = torch.as_tensor(_V(), device=f"cuda:{device_index}")
        dist.all_gather_into_tensor(t, t[rank * per * 7:(rank + 1) * per * 7].clone())
        torch.cuda.current_stream().synchronize()  # the layout kernels may run on another stream than the collective
    engine.upload_commit(n)


def shard_slot_range(n_blocks: int, block: int, rank: int, world: int):
    """Slot range [begin, end) rank `rank` integrates (same arithmetic as p3d_shard_range)."""
    per = (n_blocks + world - 1) // world * block
    n_slots = n_blocks * block
    s0 = min(n_slots, rank * per)
    return s0, min(n_slots, s0 + per)


def shard_rows(n_blocks: int, rank: int, world: int) -> np.ndarray:
    """Block rows whose pairs rank `rank` evaluates: rank, rank+world, ... (k_force_pair's row_stride)."""
    return np.arange(rank, n_blocks, world)
