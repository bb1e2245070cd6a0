"""Known-answer tests for the CPU oracle (oracle/p3d_oracle.c).

The reference has no tests: every expected value below is derived by hand from the
source text of /root/reference/src/lib.rs or from the SipHash paper.
"""
import struct

import numpy as np
import pytest

from oracle import oracle as O

TS = float(np.float32(1.0 / 60.0))


def test_siphash24_paper_vectors():
    # SipHash paper, Appendix A: key 00..0f, message 00..0e -> a129ca6149be45e5; empty -> 726fdb47dd0e0e31
    key = bytes(range(16))
    k0, k1 = int.from_bytes(key[:8], "little"), int.from_bytes(key[8:], "little")
    assert O.siphash(2, 4, k0, k1, bytes(range(15))) == 0xA129CA6149BE45E5
    assert O.siphash(2, 4, k0, k1, b"") == 0x726FDB47DD0E0E31


def test_siphash13_default_hasher_anchor():
    # Rust: DefaultHasher::new().finish() == SipHash-1-3(k=0,0)("")
    assert O.siphash(1, 3, 0, 0, b"") == 0xD1FBA762150C532C


@pytest.mark.parametrize("cell", [(0, 0, 0), (1, -2, 3), (-1, -1, -1), (2 ** 40, -(2 ** 41), 7), (2 ** 63 - 1, -(2 ** 63), 0)])
def test_hash_cell_is_siphash13_of_three_le_words(cell):
    # lib.rs:46-52: x.hash(); y.hash(); z.hash(); -> 24-byte message of little-endian isize
    msg = struct.pack("<qqq", *cell)
    assert O.hash_cell(*cell) == O.siphash(1, 3, 0, 0, msg)


def test_cell_coord_truncates_toward_zero():
    # lib.rs:37-43 with r = 2: `as isize` truncates, so cell 0 spans (-2, 2)
    assert O.cell_coord(2.0, (1.9, -1.9, -2.0)) == (0, 0, -1)
    assert O.cell_coord(2.0, (5.0, -5.0, 3.99)) == (2, -2, 1)


def test_cell_coord_saturates_and_nan_is_zero():
    # Rust float->int casts saturate; NaN -> 0.  r = 0 (allowed by the UI, main.rs:308) gives +-inf / NaN.
    assert O.cell_coord(2.0, (float("nan"), 1e30, -1e30)) == (0, 2 ** 63 - 1, -(2 ** 63))
    assert O.cell_coord(0.0, (1.0, -1.0, 0.0)) == (2 ** 63 - 1, -(2 ** 63), 0)


@pytest.mark.parametrize(
    "d,a,expect",
    [
        (0.15, 7.0, -0.5),      # d < m: d/m - 1, independent of a
        (0.3, 1.0, 0.0),        # d == m exactly -> neither branch -> 0
        (0.475, 1.0, 0.5),      # rising edge of the triangle
        (0.65, 1.0, 1.0),       # peak at (1+m)/2
        (0.65, -1.5, -1.5),
        (0.825, 1.0, 0.5),
        (1.0, 1.0, 0.0),        # d >= 1 -> 0
        (1.5, 1.0, 0.0),        # inside r = 2 but past the force range
    ],
)
def test_calculate_force(d, a, expect):
    assert O.calculate_force(0.3, d, a) == pytest.approx(expect, abs=2e-7)


def test_calculate_force_degenerate_thresholds():
    assert O.calculate_force(0.0, 0.25, 2.0) == pytest.approx(2.0 * (1 - abs(0.5 - 1.0) / 1.0))  # m = 0
    assert O.calculate_force(1.0, 0.5, 2.0) == pytest.approx(-0.5)  # m = 1: only the repulsion branch
    assert O.calculate_force(1.0, 1.0, 2.0) == 0.0


def test_walls_and_wrap():
    p = np.zeros((), O.PARTICLE)
    p["px"], p["vx"] = 5.2, 1.0
    q = O.handle_wall_collision(10.0, True, p)   # lib.rs:74-78
    assert q["px"] == 5.0 and q["vx"] == 0.0
    p["vx"] = -1.0
    q = O.handle_wall_collision(10.0, True, p)
    assert q["px"] == 5.0 and q["vx"] == -1.0
    q = O.handle_wall_collision(10.0, False, p)  # lib.rs:80-81
    assert q["px"] == pytest.approx(-4.8, abs=1e-6)
    p["px"] = 15.5
    q = O.handle_wall_collision(10.0, False, p)  # single wrap only (`else if`)
    assert q["px"] == pytest.approx(5.5)
    p["px"], p["py"], p["vy"] = 0.0, -5.5, -2.0
    q = O.handle_wall_collision(10.0, True, p)   # lib.rs:102-105
    assert q["py"] == -5.0 and q["vy"] == 0.0


def _two(default_params, id1, x1=0.65, **over):
    p = np.zeros(2, O.PARTICLE)
    p[1]["px"], p[1]["id"] = x1, id1
    return O.update(dict(default_params, **over), TS, p, want_force=True)


def test_two_body_attraction(default_params):
    # A[0][1] = A[1][0] = 1 at the triangle peak: F = +-1; v = 1*1*2/60 then drag (1 - 0.97/60)
    r = _two(default_params, 1)
    assert r["force"][0, 0] == pytest.approx(1.0, abs=2e-7) and r["force"][1, 0] == pytest.approx(-1.0, abs=2e-7)
    v = (2.0 / 60.0) * (1 - 0.97 / 60.0)
    assert r["out"][0]["vx"] == pytest.approx(v, rel=1e-6)
    assert r["out"][0]["px"] == pytest.approx(v / 60.0, rel=1e-6)
    assert r["out"][1]["px"] == pytest.approx(0.65 - v / 60.0, rel=1e-6)


def test_two_body_asymmetry(default_params):
    # A[0][4] = -1 pushes p0 away, A[4][0] = +1 pulls p1 toward p0: both move in -x (Newton III broken)
    r = _two(default_params, 4)
    assert r["force"][0, 0] == pytest.approx(-1.0, abs=2e-7)
    assert r["force"][1, 0] == pytest.approx(-1.0, abs=2e-7)


def test_periodic_image_applies_with_and_without_walls(default_params):
    for walls in (False, True):  # lib.rs:177-192 has no `walls` test
        p = np.zeros(2, O.PARTICLE)
        p[0]["px"], p[1]["px"] = 4.9, -4.9
        r = O.update(dict(default_params, walls=walls), TS, p, want_force=True)
        assert r["force"][0, 0] == pytest.approx(-1.0 / 3.0, rel=1e-5)
        assert r["force"][1, 0] == pytest.approx(+1.0 / 3.0, rel=1e-5)


def test_drag_clamp_stops_particle(default_params):
    p = np.zeros(1, O.PARTICLE)
    p[0]["vx"] = 3.0
    out = O.update(dict(default_params, coefficient=1.0), 2.0, p)["out"]  # c*ts = 2 > 1 (lib.rs:253-255)
    assert out[0]["vx"] == 0.0 and out[0]["px"] == 0.0


def test_gravity_and_kick_order(default_params):
    p = np.zeros(1, O.PARTICLE)
    out = O.update(dict(default_params, acceleration=(0.0, -9.8, 0.0)), TS, p)["out"]
    ts = np.float32(TS)
    v = np.float32(-9.8) * ts
    v = v - (v * np.float32(0.97)) * ts
    assert out[0]["vy"] == np.float32(v)
    assert out[0]["py"] == np.float32(v * ts)


def test_world_too_small_asserts(default_params):
    with pytest.raises(AssertionError):
        O.update(dict(default_params, world_size=3.9), TS, np.zeros(1, O.PARTICLE))  # lib.rs:132


def test_bad_id_raises(default_params):
    p = np.zeros(2, O.PARTICLE)
    p[1]["id"] = 5
    with pytest.raises(IndexError):
        O.update(default_params, TS, p)


def test_empty_and_single(default_params):
    assert O.update(default_params, TS, np.zeros(0, O.PARTICLE))["out"].shape == (0,)
    p = np.zeros(1, O.PARTICLE)
    p[0]["px"] = 1.0
    out = O.update(default_params, TS, p)["out"]
    assert out[0]["px"] == 1.0 and out[0]["vx"] == 0.0  # self pair is skipped by d2 > 0 (lib.rs:216)


def test_coincident_particles_do_not_interact(default_params):
    p = np.zeros(2, O.PARTICLE)
    p["px"] = 1.25
    r = O.update(default_params, TS, p, want_force=True)
    assert not r["force"].any()


def test_hash_cell_against_cpythons_own_siphash13():
    """An implementation we did not write: CPython >= 3.11 hashes `bytes` with SipHash-1-3, and PYTHONHASHSEED=0
    zeroes its key — the very function Rust's `DefaultHasher::new()` computes (src/lib.rs:46-52 feeds it the three
    cell coordinates as 8 little-endian bytes each).  2,000 random cells, saturated and negative ones included."""
    import json
    import os
    import subprocess
    import sys

    rng = np.random.default_rng(12)
    cells = [tuple(int(v) for v in rng.integers(-(2 ** 63), 2 ** 63 - 1, 3, dtype=np.int64)) for _ in range(1000)]
    cells += [tuple(int(v) for v in rng.integers(-60, 60, 3)) for _ in range(1000)]
    cells += [(2 ** 63 - 1, -(2 ** 63), 0), (0, 0, 0), (-1, -1, -1)]
    code = ("import sys, json\n"
            "assert sys.hash_info.algorithm == 'siphash13' and sys.hash_info.hash_bits == 64, sys.hash_info\n"
            "cells = json.load(sys.stdin)\n"
            "print(json.dumps([hash(b''.join(int(c).to_bytes(8, 'little', signed=True) for c in cell)) & (2**64 - 1)"
            " for cell in cells]))\n")
    env = dict(os.environ, PYTHONHASHSEED="0")
    r = subprocess.run([sys.executable, "-c", code], input=json.dumps(cells), capture_output=True, text=True, env=env)
    if r.returncode != 0 and "hash_info" in r.stderr:
        pytest.skip("this interpreter does not hash bytes with 64-bit SipHash-1-3")
    assert r.returncode == 0, r.stderr
    theirs = json.loads(r.stdout)
    ours = [O.hash_cell(*cell) for cell in cells]
    # CPython maps a hash of -1 to -2 (an error sentinel): such a value cannot be told apart, skip it if it ever occurs
    assert all(a == b for a, b in zip(ours, theirs) if b != 2 ** 64 - 2)
    assert sum(b != 2 ** 64 - 2 for b in theirs) >= len(cells) - 1
