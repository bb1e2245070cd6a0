"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm runs on the CPU
and prints exactly one JSON line with the agreed keys; the engine arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "3", "--particles", "20000")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "pair_interactions_per_s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] == 20000.0 ** 2 / (d["ms_per_step"] * 1e-3)
    assert "workload" in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly_and_ignore_omp_threads():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3", "--particles", "5000",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "3", "--particles", "5000", env={"OMP_NUM_THREADS": "1"})
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))  # torchrun's OMP_NUM_THREADS=1 must not throttle it


def test_engine_arm_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("GPU present")
    except ImportError:
        pass
    r = _run("--steps", "1", "--warmup", "3", "--particles", "4096", "--no-cpu")
    assert r.returncode != 0  # no CPU fallback


def test_graft_entry_points_exist():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    assert callable(g.build) and callable(g.smoke)


def test_weak_scaling_particle_counts():
    """--weak holds the all-pairs work per GPU at the 1-GPU value: N_G = 1,048,576 * sqrt(G), whole blocks per GPU."""
    sys.path.insert(0, ROOT)
    import bench

    assert bench.weak_particles(1) == 1048576
    for g in (2, 4, 8):
        n = bench.weak_particles(g)
        assert n % (256 * g) == 0
        assert abs(n * n / g / 1048576.0 ** 2 - 1.0) < 2e-3
    assert bench.weak_particles(8) == 2965504  # the run recorded in DESIGN.md


def test_multi_gpu_roofline_arithmetic():
    """The sharded run's roofline object: N^2 x 20 flop over the whole step against world x the per-GPU FP32 peak,
    with the measured round-1 numbers (8 GPUs, N = 1M, 41.55 ms/step, force pass 41.31 ms on the slowest rank)."""
    import bench

    r = bench.multi_gpu_roofline(1048576, 8, 41.55, 148, 1965.0,
                                 {"force": [41.19, 41.31, 41.2, 41.25, 41.3, 41.22, 41.28, 41.21]})
    assert r["bound"] == "fp32_fma" and r["unit"] == "TFLOP/s"
    assert abs(r["peak"] - 8 * 74.44992) < 1e-6
    assert abs(r["achieved"] - 1048576.0 ** 2 * 20 / 41.55e-3 / 1e12) < 1e-6
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.88 < r["frac"] < 0.90
    assert abs(r["per_gpu"]["achieved"] * 8 - r["achieved"]) < 1e-9
    assert r["force_pass"]["ms_slowest_rank"] == 41.31 and r["force_pass"]["frac"] > r["frac"]
    # no per-rank diagnostic (--no-fused): the whole-step figures alone
    assert "force_pass" not in bench.multi_gpu_roofline(1048576, 2, 164.3, 148, 1965.0, None)


def test_reference_arm_does_not_load_the_product_library_and_configs_match():
    """the reference arm must build its inputs without libp3d.so (the oracle has its own seeded scenes),
    and both arms must print the same `config` dict (the driver compares them)."""
    code = ("import sys, json, io, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '3', "
            "'--particles', '8000']; sys.path.insert(0, %r); import bench; bench.main(); "
            "maps = open('/proc/self/maps').read(); "
            "sys.stderr.write('MAPPED:' + ','.join(sorted({l.split('/')[-1] for l in maps.splitlines() if 'libp3d' in l})) + '\\n')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    mapped = [l for l in r.stderr.splitlines() if l.startswith("MAPPED:")][-1]
    assert "libp3d_oracle.so" in mapped and "libp3d.so" not in mapped.replace("libp3d_oracle.so", ""), mapped
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    sys.path.insert(0, ROOT)
    import bench

    W = round(8000.0 ** (1.0 / 3.0), 1)
    assert d["config"] == bench.config_dict(8000, W, "uniform", False)  # what the engine arm prints for the same args
    assert d["cpu_cores"] == d["cpu_baseline"]["cores"]
