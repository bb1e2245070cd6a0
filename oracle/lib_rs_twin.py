"""Second, independent restatement of /root/reference/src/lib.rs — TEST INFRASTRUCTURE ONLY.

Purpose: the reference ships no tests or golden vectors and cannot be compiled here (no rustc), so the C
oracle (oracle/p3d_oracle.c) cannot be pinned against the reference itself.  This file narrows that gap: it
was written separately from the C oracle, straight from the Rust text, with different machinery (numpy
float32 scalars, Python integers, a byte-message SipHash, a literal `fetch_sub` fill) and follows lib.rs
statement by statement, quirks included (bucket double visits, `as isize` truncation, single wrap).  The
test suite requires the two restatements to agree BIT FOR BIT (tests/test_oracle_twin.py).  Agreement of two
readings is still not the reference's own output: parity stays "unpinned".

Pure-Python loops: small particle counts only (a step costs ~729 hashes per particle).
Execution order where rayon leaves it open: sequential (particles in index order, the 27 image offsets with
x outermost as the nested flat_map yields them, one fold accumulator, `reduce` adding it to zero).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
_M64 = (1 << 64) - 1
_ISIZE_MAX, _ISIZE_MIN = (1 << 63) - 1, -(1 << 63)


# ---- std::collections::hash_map::DefaultHasher: SipHash-1-3, keys (0, 0) -------------------------------
def _rotl(x: int, b: int) -> int:
    return ((x << b) | (x >> (64 - b))) & _M64


def _round(v):
    v0, v1, v2, v3 = v
    v0 = (v0 + v1) & _M64; v1 = _rotl(v1, 13); v1 ^= v0; v0 = _rotl(v0, 32)
    v2 = (v2 + v3) & _M64; v3 = _rotl(v3, 16); v3 ^= v2
    v0 = (v0 + v3) & _M64; v3 = _rotl(v3, 21); v3 ^= v0
    v2 = (v2 + v1) & _M64; v1 = _rotl(v1, 17); v1 ^= v2; v2 = _rotl(v2, 32)
    return [v0, v1, v2, v3]


def siphash(msg: bytes, c: int = 1, d: int = 3, k0: int = 0, k1: int = 0) -> int:
    v = [k0 ^ 0x736F6D6570736575, k1 ^ 0x646F72616E646F6D, k0 ^ 0x6C7967656E657261, k1 ^ 0x7465646279746573]
    tail = len(msg) % 8
    body, rest = msg[: len(msg) - tail], msg[len(msg) - tail:]
    words = [int.from_bytes(body[i: i + 8], "little") for i in range(0, len(body), 8)]
    words.append(int.from_bytes(rest, "little") | ((len(msg) & 0xFF) << 56))
    for m in words:
        v[3] ^= m
        for _ in range(c):
            v = _round(v)
        v[0] ^= m
    v[2] ^= 0xFF
    for _ in range(d):
        v = _round(v)
    return v[0] ^ v[1] ^ v[2] ^ v[3]


def hash_cell(cell) -> int:
    """lib.rs:46-52 — `isize::hash` writes the value's 8 native (little) endian bytes."""
    msg = b"".join(int(c).to_bytes(8, "little", signed=True) for c in cell)
    return siphash(msg)


# ---- lib.rs:37-43 ------------------------------------------------------------------------------------
def _as_isize(x) -> int:
    """Rust `f32 as isize`: truncate toward zero, saturate, NaN -> 0."""
    x = float(x)
    if x != x:
        return 0
    if x >= 9.3e18:
        return _ISIZE_MAX
    if x <= -9.3e18:
        return _ISIZE_MIN
    return int(x)


def _wrap_isize(x: int) -> int:
    """isize addition in a release build wraps."""
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


def cell_coord(radius, v):
    return tuple(_as_isize(f32(c) / f32(radius)) for c in v)


# ---- lib.rs:55-67 ------------------------------------------------------------------------------------
def calculate_force(min_pull_ratio, distance, attraction):
    m, d, a = f32(min_pull_ratio), f32(distance), f32(attraction)
    if d < m:
        return d / m - f32(1.0)
    if m < d and d < f32(1.0):
        return a * (f32(1.0) - abs(f32(2.0) * d - f32(1.0) - m) / (f32(1.0) - m))
    return f32(0.0)


# ---- lib.rs:70-127 -----------------------------------------------------------------------------------
def _rs_min(a, b):
    return b if a != a else (a if b != b else (a if a < b else b))


def _rs_max(a, b):
    return b if a != a else (a if b != b else (a if a > b else b))


def handle_wall_collision(world_size, walls, pos, vel):
    world = f32(world_size)
    half = world * f32(0.5)
    for k in range(3):
        if pos[k] > half:
            if walls:
                pos[k] = half
                vel[k] = _rs_min(vel[k], f32(0.0))
            else:
                pos[k] = pos[k] - world
        elif pos[k] < -half:
            if walls:
                pos[k] = -half
                vel[k] = _rs_max(vel[k], f32(0.0))
            else:
                pos[k] = pos[k] + world


def _mag2(v):
    return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]


# ---- lib.rs:130-272 ----------------------------------------------------------------------------------
def update(prm: dict, ts, particles: np.ndarray):
    """One `Particles::update(ts)`.  `prm` has the keys of oracle.params(); returns (new_active, total_forces)."""
    W, r = f32(prm["world_size"]), f32(prm["particle_effect_radius"])
    k, c, m = f32(prm["interaction_force"]), f32(prm["coefficient"]), f32(prm["min_pull_ratio"])
    g = [f32(x) for x in prm.get("acceleration", (0.0, 0.0, 0.0))]
    T = int(prm["id_count"])
    A = np.asarray(prm["attraction_matrix"], dtype=np.float32).ravel()
    walls = bool(prm.get("walls", False))
    ts = f32(ts)
    assert W >= f32(2.0) * r  # :132
    n = len(particles)
    out = particles.copy()
    forces = np.zeros((n, 3), np.float32)
    if n == 0:
        return out, forces
    pos = [[f32(p["px"]), f32(p["py"]), f32(p["pz"])] for p in particles]
    vel = [[f32(p["vx"]), f32(p["vy"]), f32(p["vz"])] for p in particles]
    ids = [int(p["id"]) for p in particles]

    # :135-164 — count, running totals, fetch_sub fill
    table = [0] * (n + 1)
    bucket = [hash_cell(cell_coord(r, p)) % n for p in pos]
    for b in bucket:
        table[b] += 1
    for i in range(1, n + 1):
        table[i] += table[i - 1]
    indices = [0] * n
    for i, b in enumerate(bucket):
        old = table[b]
        table[b] = old - 1
        indices[old - 1] = i

    r2 = r * r
    with np.errstate(all="ignore"):
        for i in range(n):
            p, pid = pos[i], ids[i]
            acc = [f32(0.0)] * 3
            for xo in (-1, 0, 1):
                for yo in (-1, 0, 1):
                    for zo in (-1, 0, 1):
                        offset = [f32(xo) * W, f32(yo) * W, f32(zo) * W]      # :190-191
                        shifted = [p[0] + offset[0], p[1] + offset[1], p[2] + offset[2]]
                        cell = cell_coord(r, shifted)                          # :192
                        for xc in (-1, 0, 1):
                            for yc in (-1, 0, 1):
                                for zc in (-1, 0, 1):
                                    nb = (_wrap_isize(cell[0] + xc), _wrap_isize(cell[1] + yc),
                                          _wrap_isize(cell[2] + zc))
                                    b = hash_cell(nb) % n                      # :202
                                    for s in range(table[b], table[b + 1]):    # :203-205
                                        j = indices[s]
                                        q = pos[j]
                                        rel = [q[0] - shifted[0], q[1] - shifted[1], q[2] - shifted[2]]
                                        d2 = _mag2(rel)
                                        if d2 > f32(0.0) and d2 < r2:          # :216-220
                                            d = np.sqrt(d2)
                                            f = calculate_force(m, d, A[pid * T + ids[j]])
                                            acc = [acc[0] + rel[0] / d * f, acc[1] + rel[1] / d * f,
                                                   acc[2] + rel[2] / d * f]    # :231
            total = [f32(0.0) + a for a in acc]                                # :240-243 reduce(zero, +)
            forces[i] = total
            v = vel[i]
            v = [v[a] + total[a] * k * r * ts for a in range(3)]               # :246-247
            v = [v[a] + g[a] * ts for a in range(3)]                           # :249
            dv = [v[a] * c * ts for a in range(3)]                             # :252
            if _mag2(dv) > _mag2(v):                                           # :253-259
                v = [f32(0.0)] * 3
            else:
                v = [v[a] - dv[a] for a in range(3)]
            x = [p[a] + v[a] * ts for a in range(3)]                           # :262
            handle_wall_collision(W, walls, x, v)                              # :264
            out[i]["px"], out[i]["py"], out[i]["pz"] = x
            out[i]["vx"], out[i]["vy"], out[i]["vz"] = v
    return out, forces
