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


@pytest.mark.parametrize("seed", range(8))
def test_random_device_resident_session(seed):
    """A device-resident session: parameters, world size, walls, kernel and step counts change between
    p3d_step calls (as the reference UI does between frames, src/bin/main.rs:263-359); the oracle follows."""
    rng = np.random.default_rng(9000 + seed)
    prm, parts, _ = _case(200 + seed)
    n = 2500
    W = prm["world_size"]
    T = prm["id_count"]
    parts = p3.generate_particles(W, n, seed=seed, id_count=T)
    parts["vy"] = rng.normal(0, 0.5, n).astype(np.float32)
    e = p3.Engine(0)
    e.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)  # type-grouped layout: every kernel can run on it
    e.upload(parts, T)
    ref = parts
    ts = float(np.float32(1 / 60))
    for it in range(6):
        if rng.integers(0, 2):
            prm["walls"] = not prm["walls"]
        if rng.integers(0, 2):  # the UI keeps world_size >= 2r (main.rs:286-290); shrinking can leave particles outside
            prm["world_size"] = float(max(2 * prm["particle_effect_radius"], W * rng.uniform(0.8, 1.3)))
        prm["min_pull_ratio"] = float(rng.uniform(0, 1))
        prm["attraction_matrix"] = [float(x) for x in rng.uniform(-1, 1, T * T)]
        e.set_option(_abi.OPT_FORCE_KERNEL, int(rng.choice([_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS])))
        steps = int(rng.choice([1, 2, 7]))
        e.step(p3.Engine.make_params(**prm), ts, steps)
        for _ in range(steps):
            ref = O.update(prm, ts, ref, mode=O.IDEAL)["out"]
        out = e.download()
        # several steps of independent f32 rounding: 1e-5 per step
        assert_parity(out, ref, prm["world_size"], tol=1e-5 * max(1, steps) * 2, what=f"session {seed} iteration {it} {prm}")
        ref = out  # continue from the GPU state so that rounding does not accumulate across iterations
    e.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_sharding_emulated(seed):
    """Random world sizes / kernels / parameters: the ranks' partial forces must sum to the full force and the
    ranks' integrate ranges must tile the slots (all ranks run in turn on one device; none waits on another)."""
    rng = np.random.default_rng(7000 + seed)
    prm, _, ts = _case(300 + seed)
    T, W = prm["id_count"], prm["world_size"]
    n = int(rng.choice([900, 5000, 12000]))
    parts = p3.generate_particles(W, n, seed=seed, id_count=T)
    world = int(rng.choice([2, 3, 5, 8]))
    kernel = int(rng.choice([_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS]))
    ref = O.update(prm, ts, parts, mode=O.IDEAL, want_force=True)
    P = p3.Engine.make_params(**prm)
    e = p3.Engine(0)
    e.set_option(_abi.OPT_FORCE_KERNEL, kernel)
    e.set_shard(0, world)
    e.upload(parts, T)
    total = np.zeros((n, 3))
    for r in range(world):
        e.set_shard(r, world)
        e.shard_force(P)
        e.sync()
        total += e.download_forces()
    f = ref["force"].astype(np.float64)
    frms = max(np.sqrt((f ** 2).sum(1).mean()), 1e-30)
    err = np.linalg.norm(total - f, axis=1) / np.maximum(np.linalg.norm(f, axis=1), frms)
    assert err.max() < 1e-5, f"world {world} kernel {kernel} {prm}"
    e.close()


@pytest.mark.parametrize("n", [3, 31, 32, 33, 127, 128, 129, 191, 192, 193, 255, 256, 257, 511, 512, 513, 4095, 4096, 4097, 65535, 65536, 65537])
def test_sizes_around_block_boundaries(n, default_params):
    """Particle counts straddling every granularity in the engine: warp (32), block (128/256), the AUTO
    threshold (192), the 256-block switch (65,536)."""
    W = max(6.0, round(float(n) ** (1 / 3), 1))
    prm = dict(default_params, world_size=W)
    parts = p3.generate_particles(W, n, seed=n)
    ideal = O.update(prm, 1 / 60, parts, mode=O.IDEAL)["out"]
    e = p3.Engine(0)
    P = p3.Engine.make_params(**prm)
    kernels = [_abi.FORCE_AUTO, _abi.FORCE_PAIR, _abi.FORCE_CELLS] + ([_abi.FORCE_REFERENCE_ORDER] if n <= 4097 else [])
    for kernel in kernels:
        e.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        assert_parity(e.update(P, 1 / 60, parts), ideal, W, what=f"n={n} kernel={kernel}")
    e.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_multi_device(seed, monkeypatch):
    """The same random cases through the multi-device handle (p3d_create_multi): 2, 3 or 4 members (sharing cuda:0 on a
    one-GPU box, spread over the visible GPUs otherwise), every force kernel, single steps and short resident runs.
    Covers the split upload + peer all-gather, every kernel's sharding, the fused integrate and the split download
    on random parameters (walls, gravity, r < 1, m > 1, particles outside the box, T up to 8)."""
    import torch

    if seed % 2:  # shard the cell-list / exact kernels too (by default they stay on one device below 8M particles)
        monkeypatch.setenv("P3D_MULTI_CELLS_MIN", "0")
    prm, parts, ts = _case(200 + seed)
    W = prm["world_size"]
    rng = np.random.default_rng(7000 + seed)
    members = int(rng.integers(2, 5))
    ngpu = torch.cuda.device_count()
    devs = [int(k % ngpu) for k in range(members)]
    steps = int(rng.choice([1, 1, 3]))
    ref = parts
    for _ in range(steps):
        ref = O.update(prm, ts, ref, mode=O.IDEAL)["out"]
    e = p3.Engine(devs)
    P = p3.Engine.make_params(**prm)
    for kernel in (_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS, _abi.FORCE_AUTO):
        e.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        if steps == 1:
            out = e.update(P, ts, parts)
        else:
            e.upload(parts, prm["id_count"])
            e.step(P, ts, steps)
            out = e.download()
        # (errors compound over a resident run: the one-step tolerance per step)
        assert_parity(out, ref, W, tol=1e-5 * steps, what=f"multi seed {seed} devices {devs} kernel {kernel} steps {steps} {prm}")
    e.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_resident_cell_runs_with_reslotting(seed):
    """Random parameters on a resident cell-list run long enough to be re-slotted twice (n >= 32,768, 40 steps):
    bitwise equal to the same run stepped through fresh uploads."""
    rng = np.random.default_rng(9000 + seed)
    prm, _, ts = _case(300 + seed)
    n = int(rng.choice([32768, 40000, 70001]))
    W = float(rng.uniform(30.0, 50.0))
    prm["world_size"] = W
    prm["particle_effect_radius"] = float(rng.uniform(0.5, 3.0))
    T = prm["id_count"]
    parts = p3.generate_plummer(W, n, W / 5, seed=seed, id_count=T) if seed % 2 else p3.generate_particles(W, n, seed=seed, id_count=T)
    parts["vx"] = rng.normal(0, 1, n).astype(np.float32)
    P = p3.Engine.make_params(**prm)
    a = p3.Engine(0)
    a.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
    a.upload(parts, T)
    a.step(P, ts, 40)
    b = p3.Engine(0)
    b.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
    cur = parts
    for _ in range(40):
        cur = b.update(P, ts, cur)
    assert a.download().tobytes() == cur.tobytes(), f"seed {seed} {prm}"
    a.close()
    b.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_session_multi_device_and_resident_cells(seed):
    """The device-resident session of test_random_device_resident_session (parameters, box size, walls, kernel and
    step counts change between p3d_step calls; shrinking the box leaves particles outside) on (a) a multi-device
    handle and (b) a single engine at a size where the cell list re-slots its state and the all-pairs path hands
    out-of-box steps to the gated cell pipeline."""
    import torch

    rng = np.random.default_rng(11000 + seed)
    prm, _, _ = _case(400 + seed)
    T = prm["id_count"]
    ngpu = torch.cuda.device_count()
    for mode in ("multi", "large"):
        n = 2500 if mode == "multi" else 36000
        W = prm["world_size"] = float(rng.uniform(14.0, 30.0)) if mode == "multi" else float(rng.uniform(33.0, 45.0))
        prm["particle_effect_radius"] = float(min(prm["particle_effect_radius"], W / 2))
        parts = p3.generate_particles(W, n, seed=seed, id_count=T)
        parts["vy"] = rng.normal(0, 0.5, n).astype(np.float32)
        if mode == "large":  # a sparse handful outside the box from the start: the gated out-of-box paths run
            parts["px"][:: n // 50] += np.float32(W)
        e = p3.Engine([k % ngpu for k in range(3)]) if mode == "multi" else p3.Engine(0)
        first = _abi.FORCE_PAIR if (mode == "multi" or seed % 2) else _abi.FORCE_CELLS
        kernels = ([_abi.FORCE_REFERENCE_ORDER, _abi.FORCE_PAIR, _abi.FORCE_CELLS] if first == _abi.FORCE_PAIR
                   else [_abi.FORCE_CELLS, _abi.FORCE_CELLS, _abi.FORCE_AUTO])  # identity layout: no pair kernel
        if mode == "large" and first == _abi.FORCE_PAIR:
            kernels = [_abi.FORCE_PAIR, _abi.FORCE_CELLS]  # (the exact kernel is O(27 N^2): minutes at this size)
        e.set_option(_abi.OPT_FORCE_KERNEL, first)
        e.upload(parts, T)
        ref = parts
        ts = float(np.float32(1 / 60))
        for it in range(5):
            if rng.integers(0, 2):
                prm["walls"] = not prm["walls"]
            if rng.integers(0, 2):
                # (at 36,000 particles the box only grows: shrinking it with walls on clamps thousands of particles
                # onto the faces, and the violent pile-up that follows amplifies rounding beyond any per-step bound)
                lo = 0.8 if mode == "multi" else 1.0
                prm["world_size"] = float(max(2 * prm["particle_effect_radius"], W * rng.uniform(lo, 1.3)))
            prm["min_pull_ratio"] = float(rng.uniform(0, 1))
            prm["attraction_matrix"] = [float(x) for x in rng.uniform(-1, 1, T * T)]
            e.set_option(_abi.OPT_FORCE_KERNEL, int(rng.choice(kernels)))
            steps = int(rng.choice([1, 2, 7]))
            e.step(p3.Engine.make_params(**prm), ts, steps)
            for _ in range(steps):
                ref = O.update(prm, ts, ref, mode=O.IDEAL)["out"]
            out = e.download()
            assert_parity(out, ref, prm["world_size"], tol=1e-5 * max(1, steps) * 2,
                          what=f"{mode} session {seed} iteration {it} {prm}")
            ref = out
        e.close()
