"""Host-side mirror of the reference crate `particle_3d` (src/lib.rs) over the C ABI in include/p3d.h.

Same names, same argument meaning and the same error behaviour as the reference:

    Particle   <- src/lib.rs:12-17   (position, velocity, id)
    Particles  <- src/lib.rs:20-33   (all fields public and freely mutable between steps)
    Particles.update(ts) -> array of Particle   <- src/lib.rs:130

`update` runs on the B200 through libp3d.so; there is no CPU path.  Particles are held as a
numpy structured array with the 28-byte layout of `p3d_particle`, so `active_particles[i]`
reads like the Rust `Vec<Particle>`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import PARTICLE, P3DError  # noqa: F401

__all__ = ["Particle", "Particles", "Engine", "PARTICLE", "generate_particles", "default_scene",
           "generate_plummer", "P3DError"]


def Particle(position=(0.0, 0.0, 0.0), velocity=(0.0, 0.0, 0.0), id=0):  # noqa: A002 - field name of the reference
    """One particle record (src/lib.rs:12-17) as a numpy scalar of dtype PARTICLE."""
    a = np.zeros((), dtype=PARTICLE)
    a["px"], a["py"], a["pz"] = position
    a["vx"], a["vy"], a["vz"] = velocity
    a["id"] = id
    return a


class Engine:
    """Thin RAII wrapper over p3d_engine* (device-resident stepping, options, timing)."""

    def __init__(self, device=0):
        """device: one CUDA device index, or a list of 2..8 indices of one node (p3d_create_multi: one handle drives
        them all from this thread; p3d_update / step / download work unchanged)."""
        self._lib = _abi.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            _abi.check(self._lib.p3d_create_multi(devs, len(device), C.byref(h)))
            self.devices = [int(d) for d in device]
            device = self.devices[0]
        else:
            _abi.check(self._lib.p3d_create(device, C.byref(h)))
            self.devices = [int(device)]
        self._h = h
        self.device = device
        self._n = 0
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.p3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameter marshalling -------------------------------------------------------------
    @staticmethod
    def make_params(world_size, coefficient, interaction_force, min_pull_ratio, particle_effect_radius,
                    id_count, attraction_matrix, walls=False, acceleration=(0.0, 0.0, 0.0)):
        A = np.ascontiguousarray(np.asarray(attraction_matrix, dtype=np.float32).ravel())
        # the reference indexes `attraction_matrix[id * id_count + other_id]` (src/lib.rs:225-228): a longer Vec is
        # fine (only the first id_count^2 entries can be reached by valid ids), a shorter one panics
        if A.size < int(id_count) * int(id_count):
            raise IndexError("attraction_matrix is shorter than id_count*id_count entries (src/lib.rs:225-228)")
        p = _abi.Params()
        p.world_size = world_size
        p.coefficient = coefficient
        p.interaction_force = interaction_force
        p.min_pull_ratio = min_pull_ratio
        p.particle_effect_radius = particle_effect_radius
        p.accel = (C.c_float * 3)(*[float(a) for a in acceleration])
        p.walls = 1 if walls else 0
        p.id_count = int(id_count)
        p.attraction_matrix = A.ctypes.data_as(C.POINTER(C.c_float))
        p._keepalive = A
        return p

    # -- ABI calls ---------------------------------------------------------------------------
    def update(self, params, ts: float, particles: np.ndarray) -> np.ndarray:
        inp = np.ascontiguousarray(particles, dtype=PARTICLE)
        out = np.empty_like(inp)
        _abi.check(self._lib.p3d_update(self._h, C.byref(params), ts, inp.ctypes.data, out.ctypes.data, inp.shape[0]))
        self._n = inp.shape[0]
        return out

    def update_into(self, params, ts: float, inp: np.ndarray, out: np.ndarray):
        """p3d_update on caller-owned (e.g. pinned) buffers."""
        _abi.check(self._lib.p3d_update(self._h, C.byref(params), ts, inp.ctypes.data, out.ctypes.data, inp.shape[0]))
        self._n = inp.shape[0]

    def upload(self, particles: np.ndarray, id_count: int):
        inp = np.ascontiguousarray(particles, dtype=PARTICLE)
        _abi.check(self._lib.p3d_upload(self._h, inp.ctypes.data, inp.shape[0], id_count))
        self._n = inp.shape[0]

    def step(self, params, ts: float, n_steps: int = 1):
        _abi.check(self._lib.p3d_step(self._h, C.byref(params), ts, n_steps))

    def sync(self):
        _abi.check(self._lib.p3d_sync(self._h))

    def download(self) -> np.ndarray:
        out = np.empty(self._n, dtype=PARTICLE)
        _abi.check(self._lib.p3d_download(self._h, out.ctypes.data, self._n))
        return out

    def download_into(self, out: np.ndarray):
        """p3d_download into a caller-owned (e.g. pinned) buffer of self._n particles."""
        if out.dtype != PARTICLE or out.shape[0] != self._n or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a contiguous PARTICLE array of {self._n} entries")
        _abi.check(self._lib.p3d_download(self._h, out.ctypes.data, self._n))

    def upload_part(self, part: np.ndarray, i_begin: int, n: int, id_count: int):
        """Sharded upload, phase 1: this rank's callers [i_begin, i_begin + len(part)) of n -> the staging array."""
        inp = np.ascontiguousarray(part, dtype=PARTICLE)
        _abi.check(self._lib.p3d_upload_part(self._h, inp.ctypes.data, i_begin, i_begin + inp.shape[0], n, id_count))

    def upload_commit(self, n: int):
        """Sharded upload, phase 2 (after the driver all-gathered BUF_AOS): layout + pack."""
        _abi.check(self._lib.p3d_upload_commit(self._h))
        self._n = n

    def download_part_into(self, out: np.ndarray, i_begin: int):
        if out.dtype != PARTICLE or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a contiguous PARTICLE array")
        _abi.check(self._lib.p3d_download_part(self._h, out.ctypes.data, i_begin, i_begin + out.shape[0]))

    def slot_of(self) -> np.ndarray:
        """Caller index -> slot of the resident layout."""
        out = np.empty(self._n, dtype=np.uint32)
        _abi.check(self._lib.p3d_slot_of(self._h, out.ctypes.data, self._n))
        return out

    def download_forces(self) -> np.ndarray:
        out = np.zeros((self._n, 3), dtype=np.float32)
        _abi.check(self._lib.p3d_download_forces(self._h, out.ctypes.data, self._n))
        return out

    def download_render(self, world_size: float) -> np.ndarray:
        """The state as the reference app's render storage buffer (16-byte header + 32-byte particles)."""
        out = np.zeros(16 + 32 * self._n, dtype=np.uint8)
        _abi.check(self._lib.p3d_download_render(self._h, world_size, out.ctypes.data, out.nbytes, self._n))
        return out

    def diagnostics(self) -> dict:
        d = (C.c_double * 8)()
        _abi.check(self._lib.p3d_diagnostics(self._h, d))
        return {"ke": d[0], "p": (d[1], d[2], d[3]), "max_v2": d[4], "count": int(d[5]), "sum_p2": d[6]}

    def debug_bounds_violations(self):
        """Out-of-range slot / cell indices caught by a self-checking build (-DP3D_BOUNDS_CHECK); None for a product build."""
        c = C.c_ulonglong(0)
        _abi.check(self._lib.p3d_debug_bounds_violations(self._h, C.byref(c)))
        return None if c.value == 2 ** 64 - 1 else int(c.value)

    def set_option(self, option: int, value: int):
        _abi.check(self._lib.p3d_set_option(self._h, option, value))

    def get_option(self, option: int) -> int:
        v = C.c_int()
        _abi.check(self._lib.p3d_get_option(self._h, option, C.byref(v)))
        return v.value

    def timing(self) -> dict:
        ms = (C.c_float * 12)()
        _abi.check(self._lib.p3d_get_timing(self._h, ms))
        keys = ["force", "integrate", "pack", "unpack", "partition", "h2d", "d2h", "total", "pair", "bxb", "steps"]
        return dict(zip(keys, [float(x) for x in ms]))

    def counters(self) -> dict:
        c = (C.c_uint64 * 4)()
        _abi.check(self._lib.p3d_get_counters(self._h, c))
        return {"kernels": int(c[0]), "force": int(c[1]), "integrate": int(c[2])}

    def set_stream(self, cuda_stream_ptr):
        """Run on the given cudaStream_t.  torch reports its default stream as handle 0, which the ABI
        reads as "engine's own stream"; 0 is therefore passed as cudaStreamLegacy (0x1).  None = own stream."""
        if cuda_stream_ptr is None:
            ptr = 0
        else:
            ptr = int(cuda_stream_ptr) or 1
        _abi.check(self._lib.p3d_set_stream(self._h, C.c_void_p(ptr)))

    def device_buffer(self, which: int):
        p, n = C.c_void_p(), C.c_size_t()
        _abi.check(self._lib.p3d_device_buffer(self._h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_shard(self, rank: int, world: int):
        _abi.check(self._lib.p3d_set_shard(self._h, rank, world))

    def shard_range(self):
        a, b = C.c_size_t(), C.c_size_t()
        _abi.check(self._lib.p3d_shard_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def shard_force(self, params):
        _abi.check(self._lib.p3d_shard_force(self._h, C.byref(params)))

    def shard_integrate(self, params, ts: float):
        _abi.check(self._lib.p3d_shard_integrate(self._h, C.byref(params), ts))

    def shard_commit(self):
        _abi.check(self._lib.p3d_shard_commit(self._h))

    def shard_integrate_fused(self, params, ts: float):
        _abi.check(self._lib.p3d_shard_integrate_fused(self._h, C.byref(params), ts))

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(_abi.IPC_HANDLES * 64)
        _abi.check(self._lib.p3d_ipc_export(self._h, buf))
        return buf.raw

    def ipc_import(self, world: int, all_handles: bytes):
        assert len(all_handles) == world * _abi.IPC_HANDLES * 64
        _abi.check(self._lib.p3d_ipc_import(self._h, world, all_handles))

    def ipc_close(self):
        _abi.check(self._lib.p3d_ipc_close(self._h))


class Particles:
    """Mirror of `pub struct Particles` (src/lib.rs:20-33); every field is public and mutable."""

    def __init__(self, world_size, active_particles, id_count, attraction_matrix, colors=None,
                 coefficient=0.97, interaction_force=1.0, min_pull_ratio=0.3, particle_effect_radius=2.0,
                 walls=False, acceleration=(0.0, 0.0, 0.0), past_particles=None, device: int = 0):
        self.world_size = world_size
        self.active_particles = np.ascontiguousarray(active_particles, dtype=PARTICLE)
        self.past_particles = (np.zeros(0, dtype=PARTICLE) if past_particles is None
                               else np.ascontiguousarray(past_particles, dtype=PARTICLE))
        self.id_count = id_count
        self.attraction_matrix = list(np.asarray(attraction_matrix, dtype=np.float32).ravel())
        self.colors = colors if colors is not None else []  # render only (src/lib.rs:26)
        self.coefficient = coefficient
        self.interaction_force = interaction_force
        self.min_pull_ratio = min_pull_ratio
        self.particle_effect_radius = particle_effect_radius
        self.walls = walls
        self.acceleration = tuple(acceleration)
        # P3D_DEVICES="0,1,2,3": one engine handle over several GPUs of the node (p3d_create_multi)
        env = os.environ.get("P3D_DEVICES", "").strip()
        self._device = [int(d) for d in env.split(",")] if env and device == 0 else device
        if isinstance(self._device, list) and len(self._device) == 1:
            self._device = self._device[0]
        # faithful = True reproduces the reference's bucket double-visit quirk (P3D_OPT_FAITHFUL).
        # The default (False, or P3D_FAITHFUL=1 in the environment to flip it) evaluates every in-range pair exactly
        # once, which DEVIATES from src/lib.rs:195-206 for the ~26*k/N of the particles whose 27 hashed cells collide.
        self.faithful = os.environ.get("P3D_FAITHFUL", "0") not in ("", "0")
        self._engine = None

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._device)
        return self._engine

    def _params(self):
        return Engine.make_params(self.world_size, self.coefficient, self.interaction_force, self.min_pull_ratio,
                                  self.particle_effect_radius, self.id_count, self.attraction_matrix, self.walls,
                                  self.acceleration)

    def update(self, ts: float) -> np.ndarray:
        """`pub fn update(&mut self, ts: f32) -> Vec<Particle>` (src/lib.rs:130).

        Raises AssertionError when world_size < 2*particle_effect_radius (src/lib.rs:132) and
        IndexError for an id >= id_count (src/lib.rs:225-228).  Afterwards `past_particles` is the
        pre-step state (src/lib.rs:167) and `active_particles` the post-step state in the same
        index order; the return value is a copy of it (src/lib.rs:271).
        """
        self.engine.set_option(_abi.OPT_FAITHFUL, 1 if self.faithful else 0)
        new = self.engine.update(self._params(), ts, self.active_particles)
        self.past_particles = self.active_particles  # src/lib.rs:167 swap
        self.active_particles = new
        return new.copy()  # src/lib.rs:271 clone


def _particles_run(self, ts: float, n_steps: int) -> np.ndarray:
    """n_steps x update(ts) with the state resident in HBM (one upload, one download): what a headless run
    wants.  Equivalent to calling update() n_steps times; past_particles is the state before the LAST step
    only when n_steps == 1, otherwise it is the state before the run."""
    eng = self.engine
    eng.set_option(_abi.OPT_FAITHFUL, 1 if self.faithful else 0)
    before = self.active_particles
    eng.upload(before, self.id_count)
    eng.step(self._params(), ts, n_steps)
    self.active_particles = eng.download()
    self.past_particles = before
    return self.active_particles.copy()


Particles.run = _particles_run


def default_scene(n: int = 1000, seed: int = 42, device: int = 0) -> Particles:
    """The default scene of src/bin/main.rs:123-148 with a seeded generator."""
    lib = _abi.load()
    prm = _abi.Params()
    mat = (C.c_float * 25)()
    lib.p3d_scene_default_params(C.byref(prm), mat)
    parts = generate_particles(prm.world_size, n, seed=seed, id_count=prm.id_count)
    colors = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0), (1.0, 0.0, 1.0)]  # main.rs:126-132
    return Particles(world_size=prm.world_size, active_particles=parts, id_count=prm.id_count,
                     attraction_matrix=list(mat), colors=colors, coefficient=prm.coefficient,
                     interaction_force=prm.interaction_force, min_pull_ratio=prm.min_pull_ratio,
                     particle_effect_radius=prm.particle_effect_radius, walls=bool(prm.walls),
                     acceleration=tuple(prm.accel), device=device)


def default_params_dict() -> dict:
    lib = _abi.load()
    prm = _abi.Params()
    mat = (C.c_float * 25)()
    lib.p3d_scene_default_params(C.byref(prm), mat)
    return dict(world_size=prm.world_size, coefficient=prm.coefficient, interaction_force=prm.interaction_force,
                min_pull_ratio=prm.min_pull_ratio, particle_effect_radius=prm.particle_effect_radius,
                id_count=int(prm.id_count), attraction_matrix=[float(x) for x in mat], walls=bool(prm.walls),
                acceleration=tuple(float(x) for x in prm.accel))


def generate_particles(world_size: float, count: int, seed: int = 42, id_count: int = 5) -> np.ndarray:
    """Seeded restatement of `generate_particles` (src/bin/main.rs:60-87): uniform box, v = 0."""
    out = np.zeros(count, dtype=PARTICLE)
    _abi.load().p3d_scene_uniform(seed, count, world_size, id_count, out.ctypes.data)
    return out


def generate_plummer(world_size: float, count: int, scale_a: float, seed: int = 42, id_count: int = 5) -> np.ndarray:
    """Plummer-like clustered cloud truncated to the box (BASELINE.json config 3)."""
    out = np.zeros(count, dtype=PARTICLE)
    _abi.load().p3d_scene_plummer(seed, count, world_size, scale_a, id_count, out.ctypes.data)
    return out
