"""The algebra the GPU force kernels rest on, checked on the CPU against the oracle's `calculate_force`
(src/lib.rs:55-67): with inv = 1/d,

    f(d, a) / d  =  min(1/m - inv, 0)  +  a * max(0, min(c2 - c2*m*inv, c2*inv - c2)),      c2 = 2/(1-m) (0 for m >= 1)

(csrc/p3d_kernels_pair.cuh `pair_group`, `cell_pair`, `k_force_bxb`), its MPOS variant for 0 < m < 1 (c2*m folded into
the matrix entry, the rising edge shares u = 1/m - inv with the repulsion term), and the reach / cutoff rules of
`canonicalise` (csrc/p3d_engine.cu): the law is zero beyond max(1, m), RCUT masks d >= r when r is smaller.
Evaluated in float32 like the kernels (numpy, no FMA — the identity, not the rounding, is under test)."""
import numpy as np
import pytest

from oracle import oracle as O

f32 = np.float32


def _consts(m):
    m = f32(m)
    with np.errstate(divide="ignore"):
        inv_m = f32(1.0) / m if m > 0 else f32(np.inf)
    c2 = f32(2.0) / (f32(1.0) - m) if m < 1 else f32(0.0)
    return m, inv_m, c2


def kernel_form(d, a, m, r, mpos):
    """f(d, a) the way the kernels evaluate it (times d, to compare with the reference's f)."""
    m, im, c2 = _consts(m)
    d = d.astype(np.float32)
    with np.errstate(all="ignore"):
        inv = (f32(1.0) / np.sqrt(d * d)).astype(np.float32)
        u = (im - inv).astype(np.float32)                      # 1/m - 1/d
        if mpos:
            assert 0 < m < 1
            p2 = (inv * im - im).astype(np.float32)            # (1/d - 1) / m
            ti = np.maximum(np.minimum(u, p2), f32(0.0))
            a_eff = f32(a) * (c2 * m)                          # c2*m folded into the matrix entry
        else:
            p1 = (inv * (-c2 * m) + c2).astype(np.float32)     # c2 * (1 - m/d)
            p2 = (inv * c2 - c2).astype(np.float32)            # c2 * (1/d - 1)
            ti = np.maximum(np.minimum(p1, p2), f32(0.0))
            a_eff = f32(a)
        rs = np.minimum(u, f32(0.0))
        law_range = max(1.0, float(m))
        if r < law_range:                                      # RCUT: src/lib.rs:216-220 cuts inside the law's range
            cut = ~(d * d < f32(r) * f32(r))
            ti = np.where(cut, f32(0.0), ti)
            rs = np.where(cut, f32(0.0), rs)
        s = (a_eff * ti + rs).astype(np.float32)
        return (s * d).astype(np.float64)


def reference(d, a, m, r):
    out = np.array([O.calculate_force(float(m), float(x), float(a)) for x in d])
    return np.where(d.astype(np.float32) ** 2 < f32(r) * f32(r), out, 0.0)  # the caller's cutoff, src/lib.rs:216-220


@pytest.mark.parametrize("m", [0.0, 1e-3, 0.3, 0.5, 0.97, 1.0, 1.5, 3.0])
@pytest.mark.parametrize("a", [-1.5, -1.0, 0.0, 0.5, 1.5])
@pytest.mark.parametrize("r", [0.4, 0.999, 1.0, 2.0, 5.0])
def test_branch_free_law_equals_calculate_force(m, a, r):
    rng = np.random.default_rng(int(1000 * m + 10 * r + 7))
    edges = np.array([m, 1.0, r, 0.5 * (1 + m), 1e-4, 1e-2], dtype=np.float32)
    edges = edges[(edges > 0) & (edges < r)]
    near = np.concatenate([edges * f32(1 - 3e-7), edges, edges * f32(1 + 3e-7)])
    d = np.concatenate([rng.uniform(1e-4, r, 4000).astype(np.float32), near]).astype(np.float32)
    d = d[(d > 0) & (d < f32(r))]
    ref = reference(d, a, m, r)
    for mpos in ([False, True] if 0 < m < 1 else [False]):
        got = kernel_form(d, a, m, r, mpos)
        # the law is continuous on (0, r) (both branches vanish at d = m); its one jump, the cutoff at d = r, is
        # outside the sample.  The bound scales with the slope c2 = 2/(1-m) of the triangle's edges.
        tol = 4e-6 * max(1.0, abs(a)) * max(1.0, 2.0 / max(1e-6, abs(1.0 - m)) if m < 1 else 1.0)
        assert np.max(np.abs(got - ref)) <= tol, (m, a, r, mpos, float(np.max(np.abs(got - ref))))


@pytest.mark.parametrize("m,r", [(0.3, 2.0), (0.3, 0.7), (1.5, 2.5), (1.5, 1.2), (0.0, 1.0), (1.0, 3.0)])
def test_law_is_zero_beyond_reach(m, r):
    """reach = min(r, max(1, m)): what the partition's interior test, the cell size and k_force_bxb's range test use."""
    reach = min(r, max(1.0, m))
    d = np.linspace(reach * (1 + 1e-6), max(r, reach) * 1.5 + 1.0, 2000).astype(np.float32)
    for a in (-1.5, 0.7):
        assert not reference(d, a, m, r).any()
        assert not kernel_form(d, a, m, r, mpos=False).any()
