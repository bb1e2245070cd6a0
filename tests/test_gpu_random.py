"""Randomised parity sweep: random parameters (the whole range the reference UI allows and beyond), random sizes and
distributions, every force kernel and the faithful mode, against the CPU oracle.  Seeded, so failures reproduce."""
import numpy as np
import pytest

import particle_3d as p3
from particle_3d import _abi
from oracle import oracle as O

from helpers import assert_parity

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    T = int(rng.integers(1, 9))
    W = float(rng.uniform(4.0, 40.0))
    r = float(rng.choice([rng.uniform(0.1, 1.0), rng.uniform(1.0, min(W / 2, 6.0)), W / 2]))
    m = float(rng.choice([0.0, 1.0, rng.uniform(0.01, 0.99), rng.uniform(1.0, 1.3)]))
    prm = dict(world_size=W, coefficient=float(rng.uniform(0, 1.5)), interaction_force=float(rng.uniform(0, 10)),
               min_pull_ratio=m, particle_effect_radius=r, id_count=T,
               attraction_matrix=[float(x) for x in rng.uniform(-1.5, 1.5, T * T)], walls=bool(rng.integers(0, 2)),
               acceleration=tuple(float(x) for x in rng.uniform(-3, 3, 3)) if rng.integers(0, 2) else (0.0, 0.0, 0.0))
    n = int(rng.choice([1, 2, 37, 500, 1500, 3000]))
    parts = np.zeros(n, _abi.PARTICLE)
    kind = int(rng.integers(0, 3))
    scale = W / 2 if kind != 1 else W / 10          # uniform box / tight cluster / spilling outside the box
    for k in ("px", "py", "pz"):
        parts[k] = rng.uniform(-scale, scale, n).astype(np.float32) * (1.3 if kind == 2 else 1.0)
    for k in ("vx", "vy", "vz"):
        parts[k] = rng.normal(0, 1.0, n).astype(np.float32)
    parts["id"] = rng.integers(0, T, n)
    ts = float(np.float32(rng.choice([1 / 60, 1 / 1000, 0.1])))
    return prm, parts, ts


@pytest.mark.parametrize("seed", range(60))
def test_random_configuration(seed):
    prm, parts, ts = _case(seed)
    W = prm["world_size"]
    ideal = O.update(prm, ts, parts, mode=O.IDEAL)["out"]
    e = p3.Engine(0)
    P = p3.Engine.make_params(**prm)
    for kernel in (_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS):
        e.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        e.set_option(_abi.OPT_FAITHFUL, 0)
        assert_parity(e.update(P, ts, parts), ideal, W, what=f"seed {seed} kernel {kernel} {prm}")
    # faithful mode: only defined for in-box inputs and boxes of at least three cells
    inbox = all(np.abs(parts[k]).max() <= W / 2 for k in ("px", "py", "pz")) if len(parts) else True
    reach = min(prm["particle_effect_radius"], max(1.0, prm["min_pull_ratio"]))
    if inbox and W / reach >= 3.2:
        faithful = O.update(prm, ts, parts, mode=O.FAITHFUL)["out"]
        e.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
        e.set_option(_abi.OPT_FAITHFUL, 1)
        assert_parity(e.update(P, ts, parts), faithful, W, what=f"seed {seed} faithful {prm}")
    e.close()


@pytest.mark.parametrize("seed", range(4))
def test_random_configuration_large(seed):
    """The same sweep at a size that takes the 256-particle block layout (n >= 65,536)."""
    rng = np.random.default_rng(5000 + seed)
    prm, _, ts = _case(100 + seed)
    n = 70000
    W = float(rng.uniform(30.0, 60.0))
    prm["world_size"] = W
    prm["particle_effect_radius"] = float(rng.uniform(0.5, 3.0))
    T = prm["id_count"]
    parts = p3.generate_plummer(W, n, W / 5, seed=seed, id_count=T) if seed % 2 else p3.generate_particles(W, n, seed=seed, id_count=T)
    parts["vx"] = rng.normal(0, 1, n).astype(np.float32)
    ideal = O.update(prm, ts, parts, mode=O.IDEAL)["out"]
    e = p3.Engine(0)
    P = p3.Engine.make_params(**prm)
    for kernel in (_abi.FORCE_PAIR, _abi.FORCE_CELLS):
        e.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        assert_parity(e.update(P, ts, parts), ideal, W, what=f"large seed {seed} kernel {kernel} {prm}")
    e.close()
