"""ctypes binding of the CPU oracle (oracle/p3d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (libp3d.so and the particle_3d host mirror) never does.
Parity status: unpinned by the reference (it has no tests); see p3d_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libp3d_oracle.so")

# 28-byte particle: lib.rs:12-17
PARTICLE = np.dtype(
    [("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4"), ("id", "<u4")]
)
assert PARTICLE.itemsize == 28

FAITHFUL, IDEAL = 0, 1


class _Params(C.Structure):
    _fields_ = [
        ("world_size", C.c_float),
        ("coefficient", C.c_float),
        ("interaction_force", C.c_float),
        ("min_pull_ratio", C.c_float),
        ("particle_effect_radius", C.c_float),
        ("accel", C.c_float * 3),
        ("walls", C.c_uint32),
        ("id_count", C.c_uint32),
        ("attraction_matrix", C.POINTER(C.c_float)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("candidates", C.c_uint64),
        ("in_radius", C.c_uint64),
        ("nonzero", C.c_uint64),
        ("dup_bucket_queries", C.c_uint64),
        ("affected", C.c_uint64),
        ("t_build_s", C.c_double),
        ("t_force_s", C.c_double),
    ]

    def asdict(self):
        return {k: (float(getattr(self, k)) if k.startswith("t_") else int(getattr(self, k))) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile the oracle with the system gcc (the recipe is oracle/Makefile)."""
    src = os.path.join(_HERE, "p3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.ora_siphash.restype = C.c_uint64
        L.ora_siphash.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t]
        L.ora_hash_cell.restype = C.c_uint64
        L.ora_hash_cell.argtypes = [C.c_int64] * 3
        L.ora_cell_coord.restype = None
        L.ora_cell_coord.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
        L.ora_calculate_force.restype = C.c_float
        L.ora_calculate_force.argtypes = [C.c_float, C.c_float, C.c_float]
        L.ora_handle_wall_collision.restype = None
        L.ora_handle_wall_collision.argtypes = [C.c_float, C.c_uint32, C.c_void_p]
        L.ora_update.restype = C.c_int
        L.ora_update.argtypes = [
            C.POINTER(_Params), C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.POINTER(Stats), C.c_int,
        ]
        L.ora_update_sample.restype = C.c_int
        L.ora_update_sample.argtypes = [C.POINTER(_Params), C.c_float, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_size_t, C.c_size_t, C.c_int, C.POINTER(Stats), C.c_int]
        L.ora_bruteforce_forces.restype = C.c_int
        L.ora_bruteforce_forces.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.ora_integrate.restype = None
        L.ora_integrate.argtypes = [C.POINTER(_Params), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.ora_num_threads.restype = C.c_int
        L.ora_update_indices.restype = C.c_int
        L.ora_update_indices.argtypes = [C.POINTER(_Params), C.c_float, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Stats), C.c_int]
        L.ora_scene_default_params.restype = None
        L.ora_scene_default_params.argtypes = [C.POINTER(_Params), C.POINTER(C.c_float)]
        L.ora_scene_uniform.restype = None
        L.ora_scene_uniform.argtypes = [C.c_uint64, C.c_size_t, C.c_float, C.c_uint32, C.c_void_p]
        L.ora_scene_plummer.restype = None
        L.ora_scene_plummer.argtypes = [C.c_uint64, C.c_size_t, C.c_float, C.c_float, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _mk_params(p: dict):
    """p: dict with the scalar fields of `Particles` (lib.rs:20-33) + attraction_matrix."""
    A = np.ascontiguousarray(np.asarray(p["attraction_matrix"], dtype=np.float32).ravel())
    T = int(p["id_count"])
    assert A.size == T * T, "attraction_matrix must be id_count^2"
    prm = _Params()
    prm.world_size = p["world_size"]
    prm.coefficient = p["coefficient"]
    prm.interaction_force = p["interaction_force"]
    prm.min_pull_ratio = p["min_pull_ratio"]
    prm.particle_effect_radius = p["particle_effect_radius"]
    acc = p.get("acceleration", (0.0, 0.0, 0.0))
    prm.accel = (C.c_float * 3)(*[float(a) for a in acc])
    prm.walls = 1 if p.get("walls", False) else 0
    prm.id_count = T
    prm.attraction_matrix = A.ctypes.data_as(C.POINTER(C.c_float))
    return prm, A  # keep A alive


def siphash(c, d, k0, k1, msg: bytes) -> int:
    return int(lib().ora_siphash(c, d, k0, k1, msg, len(msg)))


def hash_cell(x, y, z) -> int:
    return int(lib().ora_hash_cell(x, y, z))


def cell_coord(radius, v):
    vin = (C.c_float * 3)(*[float(a) for a in v])
    out = (C.c_int64 * 3)()
    lib().ora_cell_coord(radius, vin, out)
    return tuple(int(a) for a in out)


def calculate_force(m, distance, attraction) -> float:
    return float(lib().ora_calculate_force(m, distance, attraction))


def handle_wall_collision(world_size, walls, particle):
    a = np.array([particle], dtype=PARTICLE)
    lib().ora_handle_wall_collision(world_size, 1 if walls else 0, a.ctypes.data)
    return a[0]


def update(params: dict, ts: float, particles: np.ndarray, mode: int = IDEAL, acc64: bool = False,
           want_force: bool = False, want_affected: bool = False, nthreads: int = 0):
    """One `Particles::update(ts)` (lib.rs:130).  Returns dict(out, force, affected, stats)."""
    assert particles.dtype == PARTICLE
    inp = np.ascontiguousarray(particles)
    n = inp.shape[0]
    out = np.empty_like(inp)
    force = np.zeros((n, 3), np.float32) if want_force else None
    aff = np.zeros(n, np.uint8) if want_affected else None
    st = Stats()
    prm, _keep = _mk_params(params)
    rc = lib().ora_update(
        C.byref(prm), ts, inp.ctypes.data, out.ctypes.data, n, mode, 1 if acc64 else 0,
        force.ctypes.data if want_force else None, aff.ctypes.data if want_affected else None,
        C.byref(st), nthreads,
    )
    if rc == 1:
        raise AssertionError("world_size >= 2.0 * particle_effect_radius (lib.rs:132)")
    if rc == 2:
        raise IndexError("particle id >= id_count (lib.rs:225-228)")
    if rc:
        raise MemoryError("oracle allocation failed")
    return {"out": out, "force": force, "affected": aff, "stats": st.asdict()}


def update_sample(params: dict, ts: float, particles: np.ndarray, i_begin: int, i_end: int, mode: int = FAITHFUL,
                  nthreads: int = 0):
    """Advance only particles [i_begin, i_end) of one step; returns (out, stats incl. t_build_s/t_force_s)."""
    inp = np.ascontiguousarray(particles)
    n = inp.shape[0]
    i_end = min(i_end, n)
    out = np.empty(max(i_end - i_begin, 0), dtype=PARTICLE)
    st = Stats()
    prm, _keep = _mk_params(params)
    rc = lib().ora_update_sample(C.byref(prm), ts, inp.ctypes.data, out.ctypes.data, n, i_begin, i_end, mode,
                                 C.byref(st), nthreads)
    if rc:
        raise AssertionError(f"oracle rc={rc}")
    return out, st.asdict()


def update_indices(params: dict, ts: float, particles: np.ndarray, indices, mode: int = IDEAL, nthreads: int = 0):
    """Advance only particles `indices` of one step (hash table over all n); returns (out, stats), out[k] = particle indices[k]."""
    inp = np.ascontiguousarray(particles)
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    out = np.empty(idx.shape[0], dtype=PARTICLE)
    st = Stats()
    prm, _keep = _mk_params(params)
    rc = lib().ora_update_indices(C.byref(prm), ts, inp.ctypes.data, out.ctypes.data, inp.shape[0],
                                  idx.ctypes.data, idx.shape[0], mode, C.byref(st), nthreads)
    if rc:
        raise AssertionError(f"oracle rc={rc}")
    return out, st.asdict()


def default_params_dict() -> dict:
    """Default scene constants (src/bin/main.rs:123-148) as the dict the oracle entry points take."""
    prm = _Params()
    mat = (C.c_float * 25)()
    lib().ora_scene_default_params(C.byref(prm), mat)
    return dict(world_size=prm.world_size, coefficient=prm.coefficient, interaction_force=prm.interaction_force,
                min_pull_ratio=prm.min_pull_ratio, particle_effect_radius=prm.particle_effect_radius,
                id_count=int(prm.id_count), attraction_matrix=[float(x) for x in mat], walls=bool(prm.walls),
                acceleration=tuple(float(x) for x in prm.accel))


def scene_uniform(world_size: float, count: int, seed: int = 42, id_count: int = 5) -> np.ndarray:
    out = np.zeros(count, dtype=PARTICLE)
    lib().ora_scene_uniform(seed, count, world_size, id_count, out.ctypes.data)
    return out


def scene_plummer(world_size: float, count: int, scale_a: float, seed: int = 42, id_count: int = 5) -> np.ndarray:
    out = np.zeros(count, dtype=PARTICLE)
    lib().ora_scene_plummer(seed, count, world_size, scale_a, id_count, out.ctypes.data)
    return out


def bruteforce_forces(params: dict, particles: np.ndarray, nthreads: int = 0) -> np.ndarray:
    inp = np.ascontiguousarray(particles)
    n = inp.shape[0]
    f = np.zeros((n, 3), np.float64)
    prm, _keep = _mk_params(params)
    rc = lib().ora_bruteforce_forces(C.byref(prm), inp.ctypes.data, n, f.ctypes.data, nthreads)
    if rc:
        raise AssertionError(f"oracle bruteforce rc={rc}")
    return f


def integrate(params: dict, ts: float, particles: np.ndarray, force: np.ndarray) -> np.ndarray:
    inp = np.ascontiguousarray(particles)
    f = np.ascontiguousarray(force, dtype=np.float32)
    out = np.empty_like(inp)
    prm, _keep = _mk_params(params)
    lib().ora_integrate(C.byref(prm), ts, inp.ctypes.data, f.ctypes.data, out.ctypes.data, inp.shape[0])
    return out


def num_threads() -> int:
    return int(lib().ora_num_threads())
