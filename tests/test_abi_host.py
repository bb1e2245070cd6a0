"""C-ABI surface and host-side logic that need no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import particle_3d as p3
from particle_3d import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "p3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/p3d.h but not exported by libp3d.so"
    assert sorted(_abi.EXPORTS) == declared, "particle_3d/_abi.py must mirror include/p3d.h one to one"


def test_abi_version_and_struct_layout():
    lib = _abi.load()
    assert lib.p3d_abi_version() == 1
    assert _abi.PARTICLE.itemsize == 28  # src/lib.rs:12-17: 2 x Vector3<f32> + u32
    assert [_abi.PARTICLE.fields[k][1] for k in ("px", "py", "pz", "vx", "vy", "vz", "id")] == [0, 4, 8, 12, 16, 20, 24]
    assert C.sizeof(_abi.Params) == 48 and _abi.Params.attraction_matrix.offset == 40


def test_default_scene_constants(default_params):
    # src/bin/main.rs:123-148
    d = default_params
    assert (d["world_size"], d["id_count"], d["particle_effect_radius"]) == (10.0, 5, 2.0)
    assert d["coefficient"] == pytest.approx(0.97) and d["interaction_force"] == 1.0
    assert d["min_pull_ratio"] == pytest.approx(0.3) and d["walls"] is False and d["acceleration"] == (0.0, 0.0, 0.0)
    assert d["attraction_matrix"] == [0.5, 1.0, -0.5, 0.0, -1.0, 1.0, 1.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.5, 1.5, -1.0,
                                      0.0, 0.0, 0.0, 0.0, -1.0, 1.0, 1.0, 1.0, 1.0, 0.5]


def test_uniform_generator_distribution():
    a = p3.generate_particles(25.4, 20000, seed=1)
    b = p3.generate_particles(25.4, 20000, seed=1)
    c = p3.generate_particles(25.4, 20000, seed=2)
    assert a.tobytes() == b.tobytes() and a.tobytes() != c.tobytes()
    for k in ("px", "py", "pz"):
        assert -12.7 <= a[k].min() and a[k].max() <= 12.7 and abs(a[k].mean()) < 0.3
    counts = np.bincount(a["id"], minlength=5)
    assert counts.min() > 3600 and counts.max() < 4400


def test_plummer_generator_is_clustered_and_in_box():
    a = p3.generate_plummer(64.0, 20000, 64.0 / 6, seed=3)
    r = np.sqrt(a["px"].astype(np.float64) ** 2 + a["py"] ** 2 + a["pz"] ** 2)
    assert np.abs(np.stack([a["px"], a["py"], a["pz"]])).max() < 32.0
    # Plummer: half of the (untruncated) mass lies inside ~1.3 a
    assert 0.35 < (r < 1.305 * 64.0 / 6).mean() < 0.65


def test_particle_record_and_particles_fields():
    p = p3.Particle(position=(1, 2, 3), velocity=(4, 5, 6), id=2)
    assert (p["px"], p["vz"], p["id"]) == (1.0, 6.0, 2)
    sim = p3.default_scene(n=10, seed=5)
    for field in ("world_size", "active_particles", "past_particles", "id_count", "attraction_matrix", "colors",
                  "coefficient", "interaction_force", "min_pull_ratio", "particle_effect_radius", "walls",
                  "acceleration"):  # src/lib.rs:20-33, all pub
        assert hasattr(sim, field)
    assert len(sim.active_particles) == 10 and len(sim.past_particles) == 0 and len(sim.colors) == 5


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the engine must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(p3.P3DError) as ei:
        p3.Engine(0)
    assert ei.value.code == _abi.ERR_NO_DEVICE
    sim = p3.default_scene(n=4)
    with pytest.raises(p3.P3DError):
        sim.update(1 / 60)
    out = (C.c_double * 4)()
    assert _abi.load().p3d_microbench(0, 1, 10, out) == _abi.ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "3d-particle-simulation-_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".rs", ".toml", "Makefile")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "p3d_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_header_is_plain_c(tmp_path):
    """include/p3d.h must be consumable from C (cgo / bindgen / ctypes generators): C99, no C++ or torch types."""
    import subprocess

    src = tmp_path / "h.c"
    src.write_text('#include "p3d.h"\nint main(void) { p3d_params p; (void)p; return sizeof(p3d_particle) == 28 ? 0 : 1; }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-L", os.path.join(ROOT, "3d-particle-simulation-_b200"), "-o", str(tmp_path / "h"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(tmp_path / "h")]).returncode == 0
