"""ctypes binding of include/p3d.h (libp3d.so).  Mirrors the header one to one.

The library is the product: importing this module never falls back to a CPU implementation.
If libp3d.so is missing, `load()` raises; if no sm_100 device is present, `p3d_create` returns
P3D_ERR_NO_DEVICE and `check()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# P3D_LIB: load another build of the same library instead (the self-checking build of tests/test_gpu_bounds.py)
LIB_PATH = os.environ.get("P3D_LIB") or os.path.join(_PKG_ROOT, "libp3d.so")

# error codes (include/p3d.h)
OK, ERR_WORLD_TOO_SMALL, ERR_BAD_ID, ERR_CUDA, ERR_INVALID, ERR_NO_DEVICE = 0, 1, 2, 3, 4, 5
MAX_TYPES = 64
# options
OPT_FORCE_KERNEL, OPT_TIMING, OPT_GRAPH, OPT_BLOCK_SORT, OPT_BLOCK_SIZE, OPT_FAITHFUL = 0, 1, 2, 3, 4, 5
FORCE_AUTO, FORCE_REFERENCE_ORDER, FORCE_PAIR, FORCE_CELLS = 0, 1, 2, 3
BUF_POS, BUF_POS_NEXT, BUF_VEL, BUF_FORCE, BUF_AOS = 0, 1, 2, 3, 4
IPC_HANDLES = 4  # P3D_IPC_HANDLES: force, both position buffers, velocities

# p3d_particle: 28 bytes (src/lib.rs:12-17)
PARTICLE = np.dtype(
    [("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("vz", "<f4"), ("id", "<u4")]
)
assert PARTICLE.itemsize == 28


class Params(C.Structure):
    """p3d_params (include/p3d.h) = scalar fields of `Particles`, src/lib.rs:20-33."""

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


EXPORTS = [
    "p3d_abi_version", "p3d_create", "p3d_destroy", "p3d_last_error", "p3d_update", "p3d_upload", "p3d_step",
    "p3d_download", "p3d_sync", "p3d_download_forces", "p3d_download_render", "p3d_diagnostics", "p3d_set_option", "p3d_get_option",
    "p3d_get_timing", "p3d_get_counters", "p3d_set_stream", "p3d_device_buffer", "p3d_set_shard",
    "p3d_shard_range", "p3d_shard_force", "p3d_shard_integrate", "p3d_shard_commit",
    "p3d_ipc_export", "p3d_ipc_import", "p3d_ipc_close", "p3d_shard_integrate_fused",
    "p3d_scene_default_params", "p3d_scene_uniform", "p3d_scene_plummer",
    "p3d_debug_bounds_violations",
    "p3d_create_multi", "p3d_slot_of", "p3d_upload_part", "p3d_upload_commit", "p3d_download_part",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    PP = C.POINTER(Params)
    L.p3d_abi_version.restype = i32
    L.p3d_create.restype = i32
    L.p3d_create.argtypes = [i32, C.POINTER(vp)]
    L.p3d_destroy.restype = None
    L.p3d_destroy.argtypes = [vp]
    L.p3d_last_error.restype = C.c_char_p
    L.p3d_update.restype = i32
    L.p3d_update.argtypes = [vp, PP, f32, vp, vp, sz]
    L.p3d_upload.restype = i32
    L.p3d_upload.argtypes = [vp, vp, sz, C.c_uint32]
    L.p3d_step.restype = i32
    L.p3d_step.argtypes = [vp, PP, f32, i32]
    L.p3d_download.restype = i32
    L.p3d_download.argtypes = [vp, vp, sz]
    L.p3d_sync.restype = i32
    L.p3d_sync.argtypes = [vp]
    L.p3d_download_forces.restype = i32
    L.p3d_download_forces.argtypes = [vp, vp, sz]
    L.p3d_download_render.restype = i32
    L.p3d_download_render.argtypes = [vp, f32, vp, sz, sz]
    L.p3d_diagnostics.restype = i32
    L.p3d_diagnostics.argtypes = [vp, C.POINTER(C.c_double)]
    L.p3d_set_option.restype = i32
    L.p3d_set_option.argtypes = [vp, i32, i32]
    L.p3d_get_option.restype = i32
    L.p3d_get_option.argtypes = [vp, i32, C.POINTER(i32)]
    L.p3d_get_timing.restype = i32
    L.p3d_get_timing.argtypes = [vp, C.POINTER(f32)]
    L.p3d_get_counters.restype = i32
    L.p3d_get_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.p3d_set_stream.restype = i32
    L.p3d_set_stream.argtypes = [vp, vp]
    L.p3d_device_buffer.restype = i32
    L.p3d_device_buffer.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(sz)]
    L.p3d_set_shard.restype = i32
    L.p3d_set_shard.argtypes = [vp, i32, i32]
    L.p3d_shard_range.restype = i32
    L.p3d_shard_range.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.p3d_shard_force.restype = i32
    L.p3d_shard_force.argtypes = [vp, PP]
    L.p3d_shard_integrate.restype = i32
    L.p3d_shard_integrate.argtypes = [vp, PP, f32]
    L.p3d_shard_commit.restype = i32
    L.p3d_shard_commit.argtypes = [vp]
    L.p3d_ipc_export.restype = i32
    L.p3d_ipc_export.argtypes = [vp, C.c_char_p]
    L.p3d_ipc_import.restype = i32
    L.p3d_ipc_import.argtypes = [vp, i32, C.c_char_p]
    L.p3d_ipc_close.restype = i32
    L.p3d_ipc_close.argtypes = [vp]
    L.p3d_shard_integrate_fused.restype = i32
    L.p3d_shard_integrate_fused.argtypes = [vp, PP, f32]
    L.p3d_scene_default_params.restype = None
    L.p3d_scene_default_params.argtypes = [PP, C.POINTER(f32)]
    L.p3d_scene_uniform.restype = None
    L.p3d_scene_uniform.argtypes = [C.c_uint64, sz, f32, C.c_uint32, vp]
    L.p3d_scene_plummer.restype = None
    L.p3d_scene_plummer.argtypes = [C.c_uint64, sz, f32, f32, C.c_uint32, vp]
    L.p3d_create_multi.restype = i32
    L.p3d_create_multi.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    L.p3d_slot_of.restype = i32
    L.p3d_slot_of.argtypes = [vp, vp, sz]
    L.p3d_upload_part.restype = i32
    L.p3d_upload_part.argtypes = [vp, vp, sz, sz, sz, C.c_uint32]
    L.p3d_upload_commit.restype = i32
    L.p3d_upload_commit.argtypes = [vp]
    L.p3d_download_part.restype = i32
    L.p3d_download_part.argtypes = [vp, vp, sz, sz]
    L.p3d_debug_bounds_violations.restype = i32
    L.p3d_debug_bounds_violations.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    _lib = L
    return L


class P3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"p3d error {code}: {msg}")
        self.code = code


def check(rc: int):
    """Maps a non-zero return onto the exception type the reference's panic corresponds to."""
    if rc == OK:
        return
    msg = load().p3d_last_error().decode("utf-8", "replace")
    if rc == ERR_WORLD_TOO_SMALL:
        raise AssertionError(msg)  # assert! at src/lib.rs:132
    if rc == ERR_BAD_ID:
        raise IndexError(msg)  # slice index panic at src/lib.rs:225-228
    raise P3DError(rc, msg)
