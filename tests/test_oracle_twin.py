"""The C oracle against a second, independently written restatement of src/lib.rs (oracle/lib_rs_twin.py):
numpy float32 scalars, Python integers, a byte-message SipHash.  Both follow the reference statement by
statement in a sequential execution order, so they must agree bit for bit — positions, velocities, total forces,
bucket double visits included (FAITHFUL mode is the literal reading)."""
import numpy as np
import pytest

import particle_3d as p3
from oracle import lib_rs_twin as twin
from oracle import oracle as O

TS = 1.0 / 60.0


def _cloud(W, n, seed, speed=0.0):
    a = p3.generate_particles(W, n, seed=seed)
    if speed:
        rng = np.random.default_rng(seed)
        for k in ("vx", "vy", "vz"):
            a[k] = rng.normal(0.0, speed, n).astype(np.float32)
    return a


def _same(a, b):
    for k in ("px", "py", "pz", "vx", "vy", "vz"):
        assert np.array_equal(a[k], b[k]), k  # == on floats: +0 and -0 compare equal, everything else bit for bit
    assert np.array_equal(a["id"], b["id"])


def test_twin_siphash_vectors_and_anchor():
    key = bytes(range(16))
    k0, k1 = int.from_bytes(key[:8], "little"), int.from_bytes(key[8:], "little")
    # SipHash-2-4 reference vectors (Aumasson & Bernstein, Appendix A and the reference implementation's table)
    assert twin.siphash(bytes(range(15)), 2, 4, k0, k1) == 0xA129CA6149BE45E5
    assert twin.siphash(b"", 2, 4, k0, k1) == 0x726FDB47DD0E0E31
    assert twin.siphash(b"") == 0xD1FBA762150C532C  # DefaultHasher::new().finish()
    for cell in [(0, 0, 0), (1, -1, 2), (-7, 3, 123456789), (2 ** 63 - 1, -(2 ** 63), 5)]:
        assert twin.hash_cell(cell) == O.hash_cell(*cell)


@pytest.mark.parametrize("d", [0.0, 1e-6, 0.1, 0.29999998, 0.3, 0.30000004, 0.5, 0.65, 0.99999994, 1.0, 1.7])
@pytest.mark.parametrize("m", [0.0, 0.3, 0.999, 1.0, 1.5])
def test_twin_calculate_force(m, d):
    with np.errstate(all="ignore"):
        a, b = float(twin.calculate_force(m, d, -0.75)), O.calculate_force(m, d, -0.75)
    assert a == b or (a != a and b != b)


@pytest.mark.parametrize("v", [0.0, 1.9999999, 2.0, -1.9999999, -2.0, 5.3, -5.3, 1e30, -1e30, float("nan"), float("inf")])
def test_twin_cell_coord(v):
    assert twin.cell_coord(2.0, (v, -v, 0.5 * v)) == O.cell_coord(2.0, (v, -v, 0.5 * v))


CASES = [
    # name, n, W, overrides, speed, steps
    ("default_scene_density", 160, 10.0, {}, 0.0, 2),
    ("moving_walls_gravity", 120, 8.0, {"walls": True, "acceleration": (0.0, -9.8, 0.3)}, 3.0, 2),
    ("box_is_two_radii", 90, 4.0, {}, 2.0, 1),
    ("short_cutoff", 150, 5.0, {"particle_effect_radius": 0.7}, 1.0, 1),
    ("repulsion_outlives_attraction", 100, 6.0, {"min_pull_ratio": 1.5, "particle_effect_radius": 2.5}, 0.5, 1),
    ("strong_drag_clamp", 80, 6.0, {"coefficient": 70.0}, 4.0, 1),
    ("fast_particles_single_wrap", 80, 6.0, {}, 500.0, 1),
    ("tiny_hash_table", 7, 4.0, {}, 1.0, 2),
]


@pytest.mark.parametrize("name,n,W,over,speed,steps", CASES, ids=[c[0] for c in CASES])
def test_c_oracle_equals_twin_bit_for_bit(default_params, name, n, W, over, speed, steps):
    prm = dict(default_params, world_size=W, **over)
    state = _cloud(W, n, seed=len(name), speed=speed)
    for _ in range(steps):
        ref = O.update(prm, TS, state, mode=O.FAITHFUL, want_force=True, nthreads=1)
        out, force = twin.update(prm, TS, state)
        assert np.array_equal(ref["force"], force)
        _same(ref["out"], out)
        state = out


def test_twin_sees_the_bucket_double_visit(default_params):
    """At least one configuration here exercises the quirk itself: FAITHFUL differs from IDEAL, and the twin —
    which knows nothing about modes — sides with FAITHFUL."""
    prm = dict(default_params, world_size=10.0)
    for seed in range(1, 40):
        state = _cloud(10.0, 120, seed=seed)
        r = O.update(prm, TS, state, mode=O.FAITHFUL, want_force=True, want_affected=True, nthreads=1)
        if r["affected"].any():
            ideal = O.update(prm, TS, state, mode=O.IDEAL, want_force=True, nthreads=1)
            assert not np.array_equal(ideal["force"], r["force"])
            _, force = twin.update(prm, TS, state)
            assert np.array_equal(force, r["force"])
            return
    pytest.fail("no seed produced a bucket double visit")


def test_twin_errors(default_params):
    state = _cloud(10.0, 10, seed=1)
    with pytest.raises(AssertionError):
        twin.update(dict(default_params, world_size=3.9), TS, state)
    close = state.copy()
    close["px"][1], close["py"][1], close["pz"][1] = close["px"][0] + 0.5, close["py"][0], close["pz"][0]
    close["id"][1] = 9  # row 0, column 9 is still inside the 25-entry matrix; row 9 is not
    close["id"][0] = 9
    with pytest.raises(IndexError):
        twin.update(default_params, TS, close)


def test_twin_reproduces_the_committed_default_scene_fixture(default_params):
    """BASELINE config 1 (default scene, N = 1000): the fixture the GPU tests are held against is the C oracle's
    output; the twin must land on the same bits — state and total forces, the 28 double-visited particles included."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "default_scene_n1000_seed42.npz"))
    assert int(g["affected1"].sum()) > 0
    out, force = twin.update(default_params, TS, g["start"])
    assert np.array_equal(force, g["faithful_force1"])
    _same(out, g["faithful_step1"])


def _numpy_all_pairs_forces(prm, parts):
    """All N x N x 27 (particle, neighbour, image) triples, vectorised, no cells and no hash: what the reference's
    neighbour search is meant to find.  Image positions and the cutoff test use the reference's float32 arithmetic
    (src/lib.rs:190-192,211-220); the force sum itself is accumulated in float64."""
    W, r, m = np.float32(prm["world_size"]), np.float32(prm["particle_effect_radius"]), np.float32(prm["min_pull_ratio"])
    T = int(prm["id_count"])
    A = np.asarray(prm["attraction_matrix"], np.float32).reshape(-1)[: T * T].reshape(T, T)
    pos = np.stack([parts["px"], parts["py"], parts["pz"]], 1).astype(np.float32)
    a = A[parts["id"][:, None], parts["id"][None, :]].astype(np.float32)
    total = np.zeros((len(parts), 3))
    with np.errstate(all="ignore"):
        for ox in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for oz in (-1, 0, 1):
                    shifted = pos + np.array([ox, oy, oz], np.float32) * W            # float32, like the reference
                    rel = pos[None, :, :] - shifted[:, None, :]                       # [i, j] = other - (self + offset)
                    d2 = (rel[..., 0] * rel[..., 0] + rel[..., 1] * rel[..., 1]) + rel[..., 2] * rel[..., 2]
                    hit = (d2 > 0) & (d2 < r * r)
                    d = np.sqrt(d2)
                    f = np.where(d < m, d / m - np.float32(1),
                                 np.where((m < d) & (d < 1), a * (np.float32(1) - np.abs(np.float32(2) * d - np.float32(1) - m)
                                                                   / (np.float32(1) - m)), np.float32(0)))
                    contrib = np.where(hit[..., None], rel / d[..., None] * f[..., None], 0).astype(np.float64)
                    total += contrib.sum(1)
    return total


@pytest.mark.parametrize("name,n,W,over", [("default", 600, 8.4, {}), ("two_radii", 300, 4.0, {}),
                                            ("short_cutoff", 500, 6.0, {"particle_effect_radius": 0.7}),
                                            ("m_above_one", 400, 7.0, {"min_pull_ratio": 1.5, "particle_effect_radius": 2.5})],
                         ids=lambda v: v if isinstance(v, str) else None)
def test_ideal_mode_equals_a_vectorised_all_pairs_sum(default_params, name, n, W, over):
    """The oracle's IDEAL mode (the physics the GPU engine implements by default) against a numpy evaluation of every
    (particle, neighbour, image) triple written here: the spatial hash must find each in-range triple exactly once."""
    prm = dict(default_params, world_size=W, **over)
    parts = _cloud(W, n, seed=31 + n)
    want = _numpy_all_pairs_forces(prm, parts)
    got = O.update(prm, TS, parts, mode=O.IDEAL, want_force=True, nthreads=1)["force"].astype(np.float64)
    frms = np.sqrt((want ** 2).sum(1).mean())
    err = np.linalg.norm(got - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), frms)
    assert frms > 0 and err.max() < 5e-6, err.max()   # float32 summation of ~30 terms against float64
