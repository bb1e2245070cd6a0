"""ctypes loader of libp3d_microbench.so (include/p3d_microbench.h): FP32-pipe probes behind the roofline
denominator.  Measurement infrastructure: used by bench.py, tools/gpu_dev.py and the tests, never by the product."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "3d-particle-simulation-_b200", "libp3d_microbench.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.p3d_microbench.restype = C.c_int
        L.p3d_microbench.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def run(device: int, kind: int, iters: int):
    """-> (rc, [lane-FMAs/s, kernel ms, SM count, max SM MHz])"""
    out = (C.c_double * 4)()
    rc = load().p3d_microbench(device, kind, iters, out)
    return rc, [float(x) for x in out]
