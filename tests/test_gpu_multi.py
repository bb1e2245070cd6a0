"""The multi-device handle (p3d_create_multi): ONE engine handle drives several devices of the node from the calling
thread, behind the unchanged p3d_update / p3d_upload / p3d_step / p3d_download calls (the reference's
only entry is Particles::update, src/lib.rs:130).  Listing a device more than once makes its members share it, so the
whole path — split upload + peer all-gather, sharded force pass, cross-device event barriers, the fused
reduce-scatter + integrate + all-gather kernel, split download — runs on a one-GPU box too; with more GPUs visible the
same tests run across them."""
import os
import subprocess

import numpy as np
import pytest

import particle_3d as p3
from particle_3d import _abi
from oracle import oracle as O
from helpers import assert_parity, parity_errors

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TS = float(np.float32(1.0 / 60.0))


@pytest.fixture(autouse=True, params=["cells_sharded", "cells_on_one_device"])
def cell_list_placement(request, monkeypatch):
    """Below 8M particles a handle keeps a cell-list / exact-kernel upload on its first device alone (sharding such a
    step costs more host time than the step takes); P3D_MULTI_CELLS_MIN = 0 shards it regardless, which is how the
    sharded cell-list and exact kernels stay covered here.  The all-pairs kernel is sharded either way."""
    if request.param == "cells_sharded":
        monkeypatch.setenv("P3D_MULTI_CELLS_MIN", "0")
    else:
        monkeypatch.delenv("P3D_MULTI_CELLS_MIN", raising=False)
    return request.param


def _device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("kernel", [_abi.FORCE_PAIR, _abi.FORCE_CELLS, _abi.FORCE_REFERENCE_ORDER, _abi.FORCE_AUTO])
def test_multi_device_update_matches_the_oracle(default_params, kernel):
    W, n = 30.0, 27000
    prm = dict(default_params, world_size=W)
    parts = p3.generate_particles(W, n, seed=42)
    ref = O.update(prm, TS, parts, mode=O.IDEAL)["out"]
    P = p3.Engine.make_params(**prm)
    for devs in _device_lists():
        eng = p3.Engine(devs)
        eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        out = eng.update(P, TS, parts)
        assert_parity(out, ref, W, what=f"devices {devs} kernel {kernel}")
        f = eng.download_forces()
        fr = O.update(prm, TS, parts, mode=O.IDEAL, want_force=True)["force"]
        assert np.abs(f - fr).max() < 2e-5 * max(1.0, np.abs(fr).max())
        eng.close()


@pytest.mark.parametrize("kernel", [_abi.FORCE_PAIR, _abi.FORCE_CELLS])
def test_multi_device_steps_match_the_oracle_and_every_block_size(default_params, kernel):
    W, n, steps = 25.4, 16384, 3
    prm = dict(default_params, world_size=W)
    parts = p3.generate_particles(W, n, seed=7)
    ref = parts
    for _ in range(steps):
        ref = O.update(prm, TS, ref, mode=O.IDEAL)["out"]
    P = p3.Engine.make_params(**prm)
    for devs in _device_lists():
        for block in (128, 256):
            eng = p3.Engine(devs)
            eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
            eng.set_option(_abi.OPT_BLOCK_SIZE, block)
            eng.upload(parts, 5)
            eng.step(P, TS, steps)
            out = eng.download()
            dv, dp = parity_errors(out, ref, W)
            assert dv.max() < 5e-5 and dp.max() < 5e-5, (devs, block)
            assert np.array_equal(out["id"], parts["id"])
            d = eng.diagnostics()
            assert d["count"] == n
            eng.close()


def test_multi_device_awkward_sizes_walls_gravity_and_empty(default_params):
    """n not divisible by the device count or the block size, fewer particles than devices, walls + gravity."""
    prm = dict(default_params, world_size=12.0, walls=True, acceleration=(0.0, -1.0, 0.5))
    P = p3.Engine.make_params(**prm)
    for devs in _device_lists():
        for kernel in (_abi.FORCE_PAIR, _abi.FORCE_AUTO):
            eng = p3.Engine(devs)
            eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
            for n in (1, 2, 5, 129, 1001, 4097):
                parts = p3.generate_particles(12.0, n, seed=n)
                parts["vx"] = 0.3
                out = eng.update(P, TS, parts)
                ref = O.update(prm, TS, parts, mode=O.IDEAL)["out"]
                assert_parity(out, ref, 12.0, what=f"devices {devs} kernel {kernel} n={n}")
            assert eng.update(P, TS, np.zeros(0, _abi.PARTICLE)).shape == (0,)
            eng.close()


def test_multi_device_faithful_mode(default_params):
    """K5 through the handle: GPU == the faithful oracle for ALL particles, quirk included."""
    prm = dict(default_params)
    parts = p3.generate_particles(10.0, 1000, seed=42)
    ref = O.update(prm, TS, parts, mode=O.FAITHFUL)["out"]
    P = p3.Engine.make_params(**prm)
    for devs in _device_lists()[:2]:
        eng = p3.Engine(devs)
        eng.set_option(_abi.OPT_FAITHFUL, 1)
        out = eng.update(P, TS, parts)
        assert_parity(out, ref, 10.0, what=f"faithful, devices {devs}")
        eng.close()


def test_multi_device_errors_and_refused_calls(default_params):
    eng = p3.Engine([0, 0])
    prm = dict(default_params)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(10.0, 500, seed=1)
    with pytest.raises(AssertionError):
        eng.update(p3.Engine.make_params(**dict(prm, world_size=3.9)), TS, parts)
    bad = parts.copy()
    bad["id"][77] = 5
    for kernel in (_abi.FORCE_PAIR, _abi.FORCE_CELLS):
        eng.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        with pytest.raises(IndexError):
            eng.update(P, TS, bad)
    for call in (lambda: eng.shard_range(), lambda: eng.set_shard(0, 2), lambda: eng.shard_force(P),
                 lambda: eng.device_buffer(_abi.BUF_POS), lambda: eng.set_stream(None), lambda: eng.ipc_export(),
                 lambda: eng.set_option(_abi.OPT_TIMING, 1)):
        with pytest.raises(p3.P3DError):
            call()
    out = eng.update(P, TS, parts)  # still usable after the refused calls
    assert_parity(out, O.update(prm, TS, parts, mode=O.IDEAL)["out"], 10.0)
    eng.close()


def test_particles_mirror_honours_p3d_devices(default_params, monkeypatch):
    """The host mirror takes the device list from P3D_DEVICES (as the Rust shim and host/particle_3d.hpp do)."""
    monkeypatch.setenv("P3D_DEVICES", "0,0")
    sim = p3.default_scene(n=3000, seed=42)
    assert sim.engine.devices == [0, 0]
    before = sim.active_particles.copy()
    out = sim.update(TS)
    ref = O.update(default_params, TS, before, mode=O.IDEAL)["out"]
    assert_parity(out, ref, 10.0)
    assert np.array_equal(sim.past_particles, before)


def test_c_consumer_of_the_multi_device_handle(tmp_path):
    """tests/c/multi_update.c: p3d_create_multi + p3d_update from C99, checked against the oracle inside the C program."""
    import torch

    pkg = os.path.join(ROOT, "3d-particle-simulation-_b200")
    ora = os.path.join(ROOT, "oracle")
    exe = tmp_path / "multi_update"
    r = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", ora, "-o", str(exe),
                        os.path.join(ROOT, "tests", "c", "multi_update.c"), "-L", pkg, "-lp3d", "-L", ora, "-lp3d_oracle",
                        f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{ora}", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = torch.cuda.device_count()
    for devs in ["0,0"] + ([",".join(str(d) for d in range(min(n, 8)))] if n >= 2 else []):
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, P3D_DEVICES=devs), timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "multi_update ok" in r.stdout


# ---------------------------------------------------------------- sharded host traffic and shard bookkeeping (one GPU)
def test_upload_part_commit_and_download_part_on_one_engine(default_params):
    """p3d_upload_part / p3d_upload_commit / p3d_download_part with world = 1: the parts of the caller's array may
    arrive in any order and any split; the result equals a plain p3d_upload + p3d_step + p3d_download."""
    W, n = 25.4, 16384
    prm = dict(default_params, world_size=W)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(W, n, seed=3)
    for kernel in (_abi.FORCE_PAIR, _abi.FORCE_CELLS):
        a = p3.Engine(0)
        a.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        a.upload(parts, 5)
        a.step(P, TS, 2)
        want = a.download()
        b = p3.Engine(0)
        b.set_option(_abi.OPT_FORCE_KERNEL, kernel)
        for c0, c1 in ((9000, n), (0, 100), (100, 9000)):  # three parts, out of order
            b.upload_part(parts[c0:c1], c0, n, 5)
        ptr, cap = b.device_buffer(_abi.BUF_AOS)
        assert ptr and cap == n
        b.upload_commit(n)
        b.step(P, TS, 2)
        got = np.empty(n, dtype=_abi.PARTICLE)
        for c0, c1 in ((5000, n), (0, 5000)):
            b.download_part_into(got[c0:c1], c0)
        if kernel == _abi.FORCE_CELLS:
            assert got.tobytes() == want.tobytes()
        else:  # float atomics: summation order differs run to run
            dv, dp = parity_errors(got, want, W)
            assert dv.max() < 1e-4 and dp.max() < 1e-4
        with pytest.raises(p3.P3DError):
            b.download_part_into(np.empty(10, dtype=_abi.PARTICLE), n - 5)  # part beyond the resident particles
        with pytest.raises(p3.P3DError):
            b.upload_commit(n)  # nothing staged any more
        with pytest.raises(p3.P3DError):
            b.upload_part(parts[:10], n - 5, n, 5)  # part beyond n
        b.upload_part(parts[:10], 0, n, 5)
        with pytest.raises(p3.P3DError):
            b.upload_part(parts[:10], 10, n + 1, 5)  # another n while an upload is being staged
        b.upload(parts, 5)  # a plain upload discards the staged one ...
        with pytest.raises(p3.P3DError):
            b.upload_commit(n)  # ... so there is nothing to commit
        a.close()
        b.close()


def test_set_shard_after_upload_drops_the_layout(default_params):
    """the slot layout is padded to B * world at upload time; changing `world` afterwards must not leave
    unequal shards behind - the resident state is dropped and the next step asks for an upload."""
    prm = dict(default_params)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(10.0, 1000, seed=5)
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)
    eng.upload(parts, 5)
    eng.set_shard(0, 1)  # same world: the state stays
    eng.step(P, TS, 1)
    eng.set_shard(1, 3)  # another world: dropped
    with pytest.raises(p3.P3DError):
        eng.shard_force(P)
    eng.upload(parts, 5)
    s0, s1 = eng.shard_range()
    ptr, n_slots = eng.device_buffer(_abi.BUF_POS)
    assert n_slots % 3 == 0 and (s1 - s0) * 3 == n_slots and s0 == s1 - s0
    eng.shard_force(P)
    eng.set_shard(0, 1)
    eng.upload(np.zeros(0, _abi.PARTICLE), 5)  # empty upload: id_count is still checked by the step
    with pytest.raises(p3.P3DError):
        eng.step(p3.Engine.make_params(**dict(prm, id_count=3, attraction_matrix=[0.0] * 9)), TS, 1)
    eng.step(P, TS, 1)
    eng.close()


def test_slot_of_reports_the_layout(default_params):
    parts = p3.generate_particles(10.0, 3000, seed=2)
    eng = p3.Engine(0)
    eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
    eng.upload(parts, 5)
    assert np.array_equal(eng.slot_of(), np.arange(3000))
    eng.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_PAIR)
    eng.upload(parts, 5)
    slot = eng.slot_of()
    assert len(np.unique(slot)) == 3000
    order = np.argsort(slot, kind="stable")
    assert np.all(np.diff(parts["id"][order].astype(np.int64)) >= 0)  # slots are grouped by type ...
    for t in range(5):
        idx = np.flatnonzero(parts["id"] == t)
        assert np.all(np.diff(slot[idx].astype(np.int64)) > 0)       # ... in caller order inside a type
    eng.close()


def test_long_resident_cell_run_on_the_handle_is_bitwise_the_single_device_run(default_params):
    """600 resident cell-list steps: the multi-device handle (three members), a single resident engine (re-slotted 18
    times) and a single engine fed through p3d_update every step end in the SAME bits.  (Cell-list forces are complete
    on the member that computes them; the fused kernel only adds the other members' zeros.)"""
    n, W, steps = 60000, 39.1, 600
    prm = dict(default_params, world_size=W)
    P = p3.Engine.make_params(**prm)
    parts = p3.generate_particles(W, n, seed=5)
    outs = []
    for devices in ([0, 0, 0], 0):
        e = p3.Engine(devices)
        e.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
        e.upload(parts, 5)
        e.step(P, TS, steps)
        outs.append(e.download())
        e.close()
    e = p3.Engine(0)
    e.set_option(_abi.OPT_FORCE_KERNEL, _abi.FORCE_CELLS)
    cur = parts
    for _ in range(steps):
        cur = e.update(P, TS, cur)
    e.close()
    assert outs[0].tobytes() == outs[1].tobytes() == cur.tobytes()
    assert np.isfinite(outs[0]["vx"]).all()
