"""Pins the CPU oracle (oracle/mpm_oracle.cpp) to the UNMODIFIED reference.

Golden vectors under tests/golden/ were produced by the reference's own advance()
(/root/reference/cpp_validation/mls-mpm88-explained.cpp:49-180), seeding (:191-196) and
polar_decomp/svd (taichi.h:8375-8420) through oracle/ref_harness.cpp (see oracle/make_golden.py).
The reference has no tests of its own, so these are the known answers.
All comparisons are BITWISE (int32 views) -- the restatement keeps the reference's operation order.
"""
import numpy as np
import pytest

from oracle.cpu import Reference, make_params


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def test_constants_match_reference(oracle, shipped):
    dt, mu0, la0 = shipped["constants"]
    mu, la = oracle.lame(1e2, 0.499)  # mls-mpm88-explained.cpp:19-21,25-26
    assert np.float32(mu) == mu0 and np.float32(la) == la0
    assert np.float32(1e-4) == dt
    # SURVEY section 4 golden table: lambda_0 differs from its fp64 value because 1-2nu cancels in fp32
    assert float(np.float32(la)).hex() == "0x1.0412920000000p+14"
    assert float(np.float32(mu)).hex() == "0x1.0ad8340000000p+5"


def test_seeding_matches_reference(oracle, shipped):
    # shipped seeding: 3000 particles, centre (0.13,0.13)?? -> see :191-196; c = 0x2986CC
    p = oracle.seed_block2(3000, 0.13, 0.13, 0.08, 0x2986CC)
    assert np.array_equal(bits(p), bits(shipped["step0"]))
    rng = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_rng.npz"))
    # first RNG pair of taichi.h:6511 -> particle 0 = (r*2-1)*0.08+0.13
    r = rng["first_pair"].astype(np.float32)
    x0 = (r * np.float32(2.0) - np.float32(1.0)) * np.float32(0.08) + np.float32(0.13)
    assert np.array_equal(bits(x0), bits(shipped["step0"][0, 0:2]))


@pytest.mark.parametrize("upto", [1, 100, 101, 1000])
def test_restatement_bitwise_vs_reference_states(oracle, shipped, upto):
    P = make_params()  # shipped material = last table entry, selected by the colour value in c
    p = shipped["step0"].copy()
    grid, _ = oracle.advance(P, 1e-4, p, upto, want_grid=True)
    assert np.array_equal(bits(p), bits(shipped["step%d" % upto]))
    if upto == 101:
        assert np.array_equal(bits(grid), bits(shipped["grid101"]))


def test_restatement_bulk_2500(oracle, shipped):
    P = make_params()
    p = shipped["step0"].copy()
    oracle.advance(P, 1e-4, p, 2500)
    x, v = p[:, 0:2].astype(np.float64), p[:, 2:4].astype(np.float64)
    got = np.concatenate([x.mean(0), v.sum(0), [0.5 * (v ** 2).sum(), p[:, 12].astype(np.float64).mean()]])
    want = shipped["bulk"][list(shipped["bulk_steps"]).index(2500)]
    assert np.array_equal(got, want)
    assert np.isfinite(p).all()


def test_decompositions_bitwise(oracle, decomp2):
    for m, pol, svd in zip(decomp2["m"], decomp2["polar"], decomp2["svd"]):
        R, S = oracle.polar2(m)
        assert np.array_equal(bits(np.concatenate([R, S])), bits(pol))
        U, s, V = oracle.svd2(m)
        assert np.array_equal(bits(np.concatenate([U, s, V])), bits(svd))


def test_decomposition_properties(oracle, decomp2):
    # property spec from the reference's dead test_simple_decompositions (taichi.h:8422-8447), tol 3e-5
    ms = decomp2["m"][:512]
    for m in ms:
        M = m.reshape(2, 2).T  # column-major storage
        R, S = (a.reshape(2, 2).T for a in oracle.polar2(m))
        tol = 3e-5 * max(1.0, np.abs(M).max())
        assert np.abs(R @ S - M).max() < tol
        assert np.abs(R @ R.T - np.eye(2)).max() < 3e-5
        assert abs(S[0, 1] - S[1, 0]) < tol
        U, sg, V = (a.reshape(2, 2).T for a in oracle.svd2(m))
        assert np.abs(U @ sg @ V.T - M).max() < tol


def test_svd3_properties(oracle):
    rs = np.random.RandomState(7)
    ms = np.concatenate([rs.uniform(-1, 1, (200, 9)), np.eye(3).reshape(1, 9) + 1e-2 * rs.randn(200, 9)])
    for m in ms.astype(np.float32):
        M = m.reshape(3, 3).T
        U, s, V = oracle.svd3(m)
        U, V = U.reshape(3, 3).T, V.reshape(3, 3).T
        assert np.abs(U @ np.diag(s) @ V.T - M).max() < 3e-5 * max(1, np.abs(M).max())
        assert np.abs(U @ U.T - np.eye(3)).max() < 3e-5 and np.abs(V @ V.T - np.eye(3)).max() < 3e-5
        assert abs(np.linalg.det(U) - 1) < 1e-4 and abs(np.linalg.det(V) - 1) < 1e-4
        assert s[0] >= s[1] >= abs(s[2]) - 1e-6


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_one_warm_substep(oracle, shipped):
    # the unmodified reference library itself, driven live from a warm state
    R = Reference()
    R.set(shipped["step100"].copy())
    R.advance(1)
    assert np.array_equal(bits(R.get()), bits(shipped["step101"]))
    p = shipped["step100"].copy()
    oracle.advance(make_params(), 1e-4, p, 1)
    assert np.array_equal(bits(p), bits(R.get()))


def test_alpha0_is_reference_path(oracle, shipped):
    # alpha == 0 must be the reference's pure APIC; a tiny alpha must differ (blend is live)
    p0 = shipped["step100"].copy()
    oracle.advance(make_params(alpha=0.0), 1e-4, p0, 1)
    assert np.array_equal(bits(p0), bits(shipped["step101"]))
    p1 = shipped["step100"].copy()
    oracle.advance(make_params(alpha=0.95), 1e-4, p1, 1)
    assert not np.array_equal(bits(p1[:, 2:4]), bits(p0[:, 2:4]))
    assert np.array_equal(bits(p1[:, 0:2]), bits(p0[:, 0:2]))  # advection uses the grid velocity


def test_binning_oracle(oracle, shipped):
    x = shipped["step100"][:, 0:2]
    cell, key, order, start = oracle.bin(2, 80, 8, x)
    base = (x * np.float32(80.0) - np.float32(0.5)).astype(np.int32)  # trunc, :55
    assert np.array_equal(cell, np.clip(base, 0, 78))
    nb = (80 - 1 + 7) // 8
    assert np.array_equal(key, (cell[:, 0] // 8) * nb + cell[:, 1] // 8)
    assert np.array_equal(order, np.argsort(key, kind="stable").astype(np.int32))
    assert start[-1] == 3000 and np.all(np.diff(start) >= 0)


def test_oracle_extensions_are_frozen(oracle):
    # fluid / jelly, FLIP alpha and the 3D lift are NOT in the reference ("parity unpinned"); this fixture pins them
    # to themselves so that a change of their arithmetic is always deliberate (oracle/make_extension_golden.py)
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_extension_golden import cases
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_extensions.npz"))
    for name, (P, dt, p0, steps) in cases().items():
        p = p0.copy()
        oracle.advance(P, dt, p, steps)
        assert np.array_equal(bits(p), bits(gold[name])), name
    got = np.stack([np.concatenate(oracle.svd3(m)) for m in gold["svd3_in"]])
    assert np.array_equal(bits(got), bits(gold["svd3_out"]))
    got = np.stack([oracle.rotation3(m) for m in gold["svd3_in"]])
    assert np.array_equal(bits(got), bits(gold["rotation3_out"]))
    lo, hi = np.float32(1 - 2.5e-2), np.float32(1 + 7.5e-3)
    got = np.stack([np.concatenate([f, [r]]).astype(np.float32)
                    for f, r in (oracle.plastic_project3(lo, hi, m) for m in gold["project3_in"])])
    assert np.array_equal(bits(got), bits(gold["project3_out"]))


def test_rotation3_properties(oracle):
    """The oracle's 3D rotation factor (Newton polar iteration; no counterpart in the reference) against the spec of
    the reference's own -- dead -- decomposition test (taichi.h:8431-8445: m = R S, R R^T = I, det R = 1, S = S^T,
    tolerance 3e-5) and against U V^T of the oracle's Jacobi SVD."""
    rs = np.random.RandomState(5)
    for k in range(2000):
        m = (np.eye(3) + (0.05 if k % 2 else 0.4) * rs.randn(3, 3)).astype(np.float32)
        if np.linalg.det(m.astype(np.float64)) < 0.05:
            continue
        R = oracle.rotation3(m.T.reshape(-1)).reshape(3, 3).T.astype(np.float64)   # column-major storage
        S = R.T @ m.astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() < 3e-5 and abs(np.linalg.det(R) - 1) < 3e-5
        assert np.abs(S - S.T).max() < 3e-5 * max(1.0, np.abs(S).max())
        U, s, V = oracle.svd3(m.T.reshape(-1))
        Rs = U.reshape(3, 3).T.astype(np.float64) @ V.reshape(3, 3).T.astype(np.float64).T
        assert np.abs(R - Rs).max() < 3e-6
