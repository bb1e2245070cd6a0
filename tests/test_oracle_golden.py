"""Oracle vs committed golden vectors and vs the independent brute-force evaluation."""
import os

import numpy as np
import pytest

import particle_3d as p3
from oracle import oracle as O

from helpers import vel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TS = float(np.float32(1.0 / 60.0))


@pytest.fixture(scope="module")
def g1():
    return np.load(os.path.join(GOLD, "default_scene_n1000_seed42.npz"))


def test_generator_is_reproducible(g1, default_params):
    again = p3.generate_particles(default_params["world_size"], 1000, seed=42)
    assert again.tobytes() == g1["start"].tobytes()
    assert np.abs(again["px"]).max() <= 5.0 and set(np.unique(again["id"])) == {0, 1, 2, 3, 4}
    assert not vel(again).any()  # main.rs:73


@pytest.mark.parametrize("mode,name", [(O.IDEAL, "ideal"), (O.FAITHFUL, "faithful")])
def test_step1_matches_golden_bitwise(g1, default_params, mode, name):
    r = O.update(default_params, TS, g1["start"], mode=mode, want_force=True, want_affected=True)
    assert r["out"].tobytes() == g1[f"{name}_step1"].tobytes()
    assert r["force"].tobytes() == g1[f"{name}_force1"].tobytes()
    if mode == O.FAITHFUL:
        assert np.array_equal(r["affected"], g1["affected1"])
        st = r["stats"]
        assert [st[k] for k in ("candidates", "in_radius", "nonzero", "dup_bucket_queries", "affected")] == list(g1["stats1"])


def test_thread_count_does_not_change_results(g1, default_params):
    a = O.update(default_params, TS, g1["start"], mode=O.FAITHFUL, nthreads=1)["out"]
    b = O.update(default_params, TS, g1["start"], mode=O.FAITHFUL, nthreads=4)["out"]
    assert a.tobytes() == b.tobytes()


def test_100_steps_match_golden(g1, default_params):
    cur = g1["start"].copy()
    for _ in range(100):
        cur = O.update(default_params, TS, cur, mode=O.IDEAL)["out"]
    assert cur.tobytes() == g1["ideal_step100"].tobytes()
    ke = 0.5 * (vel(cur) ** 2).sum()
    assert ke == pytest.approx(g1["ideal_ke"][-1], rel=1e-12)


def test_ideal_mode_equals_bruteforce_all_pairs(g1, default_params):
    # The cell-list walk (lib.rs:177-236, each bucket once) must find exactly the pairs an
    # O(27 N^2) scan finds: same set of interactions, f32 vs f64 rounding only.
    f = O.update(default_params, TS, g1["start"], mode=O.IDEAL, want_force=True)["force"]
    bf = O.bruteforce_forces(default_params, g1["start"])
    assert np.abs(f - bf).max() < 5e-6


def test_faithful_quirk_is_confined_to_the_affected_mask(g1, default_params):
    # Appendix B.1: duplicate buckets double-count neighbours for ~3% of particles at N = 1000
    fi, ff = g1["ideal_force1"], g1["faithful_force1"]
    differs = np.abs(fi - ff).max(1) > 0
    assert differs.sum() > 0
    assert not (differs & (g1["affected1"] == 0)).any()
    assert 10 <= int(g1["affected1"].sum()) <= 60


def test_survey_work_statistics(g1):
    cand, inr, nz, dupq, aff = [int(x) for x in g1["stats1"]]
    # SURVEY.md Appendix D (seed 7 there): ~1140 candidates, ~34 in-radius, ~3 non-zero per particle
    assert 900 <= cand / 1000 <= 1300 and 25 <= inr / 1000 <= 45 and 2 <= nz / 1000 <= 4


def test_acc64_variant_is_close_to_f32(g1, default_params):
    a = O.update(default_params, TS, g1["start"], mode=O.IDEAL, want_force=True)["force"]
    b = O.update(default_params, TS, g1["start"], mode=O.IDEAL, acc64=True, want_force=True)["force"]
    assert np.abs(a - b).max() < 2e-6


def test_uniform_4096_golden(default_params):
    g = np.load(os.path.join(GOLD, "uniform_n4096_seed7.npz"))
    prm = dict(default_params, world_size=16.0)
    assert O.update(prm, TS, g["start"], mode=O.IDEAL)["out"].tobytes() == g["ideal_step1"].tobytes()
    prm = dict(prm, walls=True, acceleration=(0.0, -9.8, 0.0))
    assert O.update(prm, TS, g["start"], mode=O.IDEAL)["out"].tobytes() == g["ideal_walls_gravity_step1"].tobytes()


def test_integrate_only_matches_update(g1, default_params):
    r = O.update(default_params, TS, g1["start"], mode=O.IDEAL, want_force=True)
    again = O.integrate(default_params, TS, g1["start"], r["force"])
    assert again.tobytes() == r["out"].tobytes()


@pytest.mark.parametrize("seed", range(16))
def test_cell_walk_equals_bruteforce_on_random_configurations(seed):
    """The restated spatial-hash walk (lib.rs:135-236, each bucket once) against the independent O(27 N^2) f64
    scan, over random parameters incl. r < 1, m > 1, W == 2r, many types and positions outside the box."""
    rng = np.random.default_rng(400 + seed)
    T = int(rng.integers(1, 7))
    W = float(rng.uniform(4.0, 30.0))
    r = float(rng.choice([rng.uniform(0.2, 1.0), rng.uniform(1.0, min(W / 2, 5.0)), W / 2]))
    prm = dict(world_size=W, coefficient=0.5, interaction_force=1.0, min_pull_ratio=float(rng.choice([0.0, 0.3, 1.0, 1.2])),
               particle_effect_radius=r, id_count=T, attraction_matrix=[float(x) for x in rng.uniform(-1.5, 1.5, T * T)])
    n = int(rng.choice([3, 50, 400, 900]))
    p = np.zeros(n, O.PARTICLE)
    spread = W / 2 * (1.4 if seed % 4 == 0 else 1.0)
    for k in ("px", "py", "pz"):
        p[k] = rng.uniform(-spread, spread, n).astype(np.float32)
    p["id"] = rng.integers(0, T, n)
    f = O.update(prm, TS, p, mode=O.IDEAL, want_force=True)["force"].astype(np.float64)
    bf = O.bruteforce_forces(prm, p)
    scale = max(np.abs(bf).max(), 1.0)
    assert np.abs(f - bf).max() / scale < 5e-6
