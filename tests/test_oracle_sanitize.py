"""Host sanitizers on the CPU oracle: every entry point of oracle/p3d_oracle.c driven by
tests/c/oracle_sanitize.c under AddressSanitizer + UndefinedBehaviorSanitizer (+ float-cast-overflow, leak check),
on ordinary and hostile inputs.  compute-sanitizer is closed on the GPU pool, so for the device side the
self-checking build (tests/test_gpu_bounds.py) stands in; the checker itself is held to the real tools here."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_and_ubsan(tmp_path):
    exe = tmp_path / "oracle_sanitize"
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    r = subprocess.run([cc, "-std=c11", "-g", "-O1", "-fno-omit-frame-pointer", "-ffp-contract=off",
                        "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize-recover=all", "-fopenmp",
                        "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "oracle"), "-o", str(exe),
                        os.path.join(ROOT, "tests", "c", "oracle_sanitize.c"), os.path.join(ROOT, "oracle", "p3d_oracle.c"),
                        "-lm"], capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr
                              or "libasan" in r.stderr or "libubsan" in r.stderr):
        pytest.skip("this toolchain ships no sanitizer runtimes")
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1",
               OMP_NUM_THREADS="2")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout + r.stderr
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr


def test_host_only_product_code_is_clean_under_asan_and_ubsan(tmp_path):
    """csrc/p3d_scene.cpp (seeded generators, default scene) is the one part of the product that runs on the host
    alone; it gets the same treatment (the rest of libp3d.so needs the CUDA runtime and a device)."""
    san = ["-g", "-O1", "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize-recover=all"]
    inc = ["-I", os.path.join(ROOT, "include")]
    steps = [
        ["/usr/bin/g++", "-std=c++17", *san, *inc, "-c", "-o", str(tmp_path / "scene.o"),
         os.path.join(ROOT, "3d-particle-simulation-_b200", "csrc", "p3d_scene.cpp")],
        ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", *san, *inc, "-c", "-o", str(tmp_path / "drv.o"),
         os.path.join(ROOT, "tests", "c", "scene_sanitize.c")],
        ["/usr/bin/g++", "-fsanitize=address,undefined", "-o", str(tmp_path / "scene_sanitize"), str(tmp_path / "drv.o"),
         str(tmp_path / "scene.o"), "-lm"],
    ]
    for cmd in steps:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 and ("libasan" in r.stderr or "libubsan" in r.stderr or "-lasan" in r.stderr):
            pytest.skip("this toolchain ships no sanitizer runtimes")
        assert r.returncode == 0, r.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([str(tmp_path / "scene_sanitize")], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout + r.stderr
