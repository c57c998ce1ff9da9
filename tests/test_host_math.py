"""The product's device arithmetic (csrc/mpm_math.cuh), compiled for the host and run in the
reference's sequential order, must equal the pinned oracle BIT FOR BIT -- 2D shipped scene,
all three materials, FLIP blend, and the 3D lift.  No GPU needed; the GPU tests then only have to
show that the kernels move the same numbers (and differ by atomic summation order alone)."""
import ctypes

import numpy as np
import pytest

from oracle.cpu import DEFAULT_MATERIALS, make_params
from mpm_flip98a_b200 import scenes
from tests.hostcheck import HostCheck


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


@pytest.fixture(scope="module")
def hc():
    return HostCheck()


def run_both(oracle, hc, dim, n_grid, dt, vol_p, alpha, p0, steps):
    P = make_params(dim=dim, n_grid=n_grid, vol_p=vol_p, alpha=alpha)
    a = p0.copy()
    ga, ta = oracle.advance(P, dt, a, steps, want_grid=True, want_post_p2g=True)
    H = hc.params(n_grid, 1.0, vol_p, (0.0, -200.0, 0.0), 0.05, 0.6, 20.0, alpha, DEFAULT_MATERIALS)
    b = p0.copy()
    gb, tb = hc.advance(H, dim, n_grid, dt, b, steps)
    return (a, ga, ta), (b, gb, tb)


def test_shipped_scene_bitwise(oracle, hc, shipped):
    (a, ga, ta), (b, gb, tb) = run_both(oracle, hc, 2, 80, 1e-4, 1.0, 0.0, shipped["step0"], 300)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(ga), bits(gb)) and np.array_equal(bits(ta), bits(tb))
    # and straight against the reference's own golden state
    (a, _, _), (b, _, _) = run_both(oracle, hc, 2, 80, 1e-4, 1.0, 0.0, shipped["step100"], 1)
    assert np.array_equal(bits(b), bits(shipped["step101"]))


@pytest.mark.parametrize("alpha", [0.0, 0.95])
def test_three_materials_bitwise(oracle, hc, alpha):
    p0 = scenes.commented_three_blocks()
    (a, ga, ta), (b, gb, tb) = run_both(oracle, hc, 2, 80, 1e-4, 1.0, alpha, p0, 400)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(ga), bits(gb)) and np.array_equal(bits(ta), bits(tb))
    assert np.isfinite(a).all()


@pytest.mark.parametrize("alpha", [0.0, 0.95])
def test_3d_lift_bitwise(oracle, hc, alpha):
    n = 24
    dt, vol = scenes.scaled_constants(n)
    p0 = scenes.collapse_3d(n, per_side=2, y_top=0.4, xz=(0.2, 0.8))
    (a, ga, ta), (b, gb, tb) = run_both(oracle, hc, 3, n, dt, vol, alpha, p0, 60)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(ga), bits(gb)) and np.array_equal(bits(ta), bits(tb))
    assert np.isfinite(a).all()
    assert np.abs(a[:, 3:6]).max() > 0.01  # it actually moved


def test_decompositions_bitwise(hc, decomp2, oracle):
    L = hc.lib
    for m, pol, svd in zip(decomp2["m"], decomp2["polar"], decomp2["svd"]):
        m = np.ascontiguousarray(m)
        R, S, U, sg, V = (np.zeros(4, np.float32) for _ in range(5))
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L.hostcheck_polar2(vp(m), vp(R), vp(S))
        L.hostcheck_svd2(vp(m), vp(U), vp(sg), vp(V))
        assert np.array_equal(bits(np.concatenate([R, S])), bits(pol))
        assert np.array_equal(bits(np.concatenate([U, sg, V])), bits(svd))
    rs = np.random.RandomState(3)
    for m in (np.eye(3).reshape(1, 9) + 0.05 * rs.randn(200, 9)).astype(np.float32):
        U, V, s = np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(3, np.float32)
        L.hostcheck_svd3(vp(m), vp(U), vp(s), vp(V))
        Uo, so, Vo = oracle.svd3(m)
        assert np.array_equal(bits(np.concatenate([U, s, V])), bits(np.concatenate([Uo, so, Vo])))


@pytest.mark.parametrize("alpha", [0.0, 0.95])
def test_packed2d_exact_parts_bitwise(oracle, hc, shipped, alpha):
    """mpm_math2.cuh (what the default 2D substep kernel calls): stencil2 / affine2 / g2p_finish2 keep the
    reference association on float2 pairs -> with the exact gather the whole substep is BITWISE the oracle's."""
    for p0, steps in ((shipped["step0"], 300), (scenes.commented_three_blocks(), 400)):
        P = make_params(alpha=alpha)
        a = p0.copy()
        ga, ta = oracle.advance(P, 1e-4, a, steps, want_grid=True, want_post_p2g=True)
        H = hc.params(80, 1.0, 1.0, (0.0, -200.0, 0.0), 0.05, 0.6, 20.0, alpha, DEFAULT_MATERIALS)
        b = p0.copy()
        gb, tb = hc.advance_packed2(H, 80, 1e-4, b, steps, exact_gather=True)
        assert np.array_equal(bits(a), bits(b))
        assert np.array_equal(bits(ga), bits(gb)) and np.array_equal(bits(ta), bits(tb))


@pytest.mark.parametrize("alpha", [0.0, 0.95])
def test_packed2d_fast_gather_one_warm_substep(oracle, hc, shipped, alpha):
    """The separable FMA gather of the kernel (gather2_row) against the oracle on warm states: the 1e-5 bar
    of north_star with an order of magnitude to spare (this is arithmetic only -- no atomics on the host)."""
    from tests.util import fields, rel_l2
    warm = shipped["step1000"].copy()
    three = scenes.commented_three_blocks()
    oracle.advance(make_params(alpha=alpha), 1e-4, three, 1000)
    for p0 in (warm, three):
        P = make_params(alpha=alpha)
        a = p0.copy()
        oracle.advance(P, 1e-4, a, 1)
        H = hc.params(80, 1.0, 1.0, (0.0, -200.0, 0.0), 0.05, 0.6, 20.0, alpha, DEFAULT_MATERIALS)
        b = p0.copy()
        hc.advance_packed2(H, 80, 1e-4, b, 1, exact_gather=False)
        fa, fb = fields(a, 2), fields(b, 2)
        for k in fa:
            assert rel_l2(fb[k], fa[k]) <= 2e-6, (k, rel_l2(fb[k], fa[k]))


def test_polar3_newton_matches_svd_rotation(hc):
    """The fast 3D rotation factor (Newton polar iteration, binned 3D P2G) against U V^T of the Jacobi SVD that the
    oracle defines: snow-like (clamped), jelly-like (large stretch and shear) and near-singular / inverted inputs
    (which must take the SVD fallback and therefore agree exactly)."""
    rs = np.random.RandomState(11)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def rot(m):
        a, b = np.zeros(9, np.float32), np.zeros(9, np.float32)
        hc.lib.hostcheck_rotation3(vp(np.ascontiguousarray(m, np.float32)), vp(a), vp(b))
        return a, b

    def random_rotation():
        q, r = np.linalg.qr(rs.randn(3, 3))
        q = q * np.sign(np.diag(r))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        return q

    worst = 0.0
    for k in range(3000):
        U, V = random_rotation(), random_rotation()
        if k % 3 == 0:
            sig = rs.uniform(0.975, 1.0075, 3)       # snow after the plastic clamp
        elif k % 3 == 1:
            sig = rs.uniform(0.4, 1.8, 3)            # jelly under load
        else:
            sig = np.exp(rs.uniform(-2.5, 1.0, 3))   # extreme but proper
        F = (U * sig) @ V.T
        a, b = rot(F.T.reshape(-1))                  # column-major storage
        worst = max(worst, float(np.abs(a - b).max()))
        R = b.reshape(3, 3).T.astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() < 2e-6 and abs(np.linalg.det(R) - 1) < 2e-6
    assert worst < 1.5e-6, worst
    for F in (np.diag([1.0, 1.0, -0.5]), np.diag([1.0, 0.0, 1.0]), np.zeros((3, 3)), 1e-4 * np.ones((3, 3))):
        a, b = rot(F.T.reshape(-1))
        assert np.array_equal(bits(a), bits(b))      # fallback path == the SVD rotation itself


def test_plastic_project3_matches_svd_clamp(hc, oracle):
    """The 3D snow projection (polar + symmetric Jacobi, oracle/mpm_oracle.cpp plastic_project3 == csrc/mpm_math.cuh):
    host build BITWISE the oracle; against a float64 SVD clamp <= 2e-6; the returned R is the rotation factor of F';
    untouched when every singular value is already inside the window; inverted F takes the SVD fallback."""
    rs = np.random.RandomState(5)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hc.lib.hostcheck_plastic_project3.restype = ctypes.c_float
    lo, hi = np.float32(1 - 2.5e-2), np.float32(1 + 7.5e-3)

    def random_rotation():
        q, r = np.linalg.qr(rs.randn(3, 3))
        q = q * np.sign(np.diag(r))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        return q

    worst, untouched = 0.0, 0
    for k in range(3000):
        U, V = random_rotation(), random_rotation()
        if k % 3 == 0:
            sig = rs.uniform(0.96, 1.02, 3)                  # around the window: some clamped, some not
        elif k % 3 == 1:
            sig = rs.uniform(0.98, 1.005, 3)                 # inside: nothing to clamp
        else:
            sig = np.exp(rs.uniform(-1.0, 0.7, 3))           # far outside
        F = ((U * sig) @ V.T).astype(np.float32)
        m = np.ascontiguousarray(F.T.reshape(-1))            # column-major storage
        mo, ro = oracle.plastic_project3(lo, hi, m)
        mh, R = m.copy(), np.zeros(9, np.float32)
        rh = hc.lib.hostcheck_plastic_project3(ctypes.c_float(lo), ctypes.c_float(hi), vp(mh), vp(R))
        assert np.array_equal(bits(mo), bits(mh)) and bits(np.float32(ro)) == bits(np.float32(rh))
        u, s, vt = np.linalg.svd(F.astype(np.float64))
        want = (u * np.clip(s, lo, hi)) @ vt
        got = mh.reshape(3, 3).T.astype(np.float64)
        worst = max(worst, float(np.abs(got - want).max()))
        assert abs(rh - np.prod(s) / np.prod(np.clip(s, lo, hi))) <= 3e-6 * max(1.0, rh)
        Rm = R.reshape(3, 3).T.astype(np.float64)
        assert np.abs(Rm - u @ vt).max() < 3e-6
        if np.array_equal(bits(mh), bits(m)):
            untouched += 1
            assert rh == 1.0 and s.min() >= lo - 1e-6 and s.max() <= hi + 1e-6
    assert worst < 2e-6, worst
    assert untouched > 300, untouched                        # the early-out is exercised
    F = np.diag([1.0, 1.0, -0.5]).astype(np.float32)         # inverted: SVD fallback, bitwise the oracle's
    m = np.ascontiguousarray(F.T.reshape(-1))
    mo, ro = oracle.plastic_project3(lo, hi, m)
    mh, R = m.copy(), np.zeros(9, np.float32)
    hc.lib.hostcheck_plastic_project3(ctypes.c_float(lo), ctypes.c_float(hi), vp(mh), vp(R))
    assert np.array_equal(bits(mo), bits(mh)) and np.isfinite(mh).all()
