"""The oracle's own seeded scenes (used by bench.py --impl reference so that the reference arm never loads
libp3d.so) are byte-identical to the product's generators, and ora_update_indices agrees with ora_update."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.mark.parametrize("n,W,seed", [(0, 10.0, 42), (1, 10.0, 42), (1000, 10.0, 42), (50000, 101.6, 42), (4097, 25.4, 7)])
def test_uniform_scene_matches_the_product_generator(n, W, seed):
    import particle_3d as p3

    a, b = O.scene_uniform(W, n, seed), p3.generate_particles(W, n, seed=seed)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n,W,seed", [(1000, 10.0, 42), (30000, 64.0, 42), (5000, 64.0, 3)])
def test_plummer_scene_matches_the_product_generator(n, W, seed):
    import particle_3d as p3

    a, b = O.scene_plummer(W, n, W / 6, seed), p3.generate_plummer(W, n, W / 6, seed=seed)
    assert a.tobytes() == b.tobytes()


def test_default_params_match_the_product_table(default_params):
    assert O.default_params_dict() == default_params


def test_update_indices_equals_the_full_step():
    prm = dict(O.default_params_dict(), world_size=25.4)
    parts = O.scene_uniform(25.4, 16384, 42)
    full = O.update(prm, 1 / 60, parts, mode=O.IDEAL)["out"]
    rng = np.random.default_rng(5)
    idx = rng.choice(16384, size=777, replace=False)
    out, st = O.update_indices(prm, 1 / 60, parts, idx, mode=O.IDEAL)
    assert out.tobytes() == full[idx].tobytes()
    assert st["candidates"] > 0
    out_f, _ = O.update_indices(prm, 1 / 60, parts, idx, mode=O.FAITHFUL)
    assert out_f.tobytes() == O.update(prm, 1 / 60, parts, mode=O.FAITHFUL)["out"][idx].tobytes()
    empty, _ = O.update_indices(prm, 1 / 60, parts, np.zeros(0, np.uint64))
    assert empty.shape == (0,)
    with pytest.raises(AssertionError):
        O.update_indices(prm, 1 / 60, parts, [16384])
