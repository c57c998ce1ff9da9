"""Freeze the oracle's EXTENSIONS (materials fluid/jelly, FLIP alpha, 3D lift) into tests/golden/oracle_extensions.npz.

TEST INFRASTRUCTURE ONLY.  These parts have no counterpart in the reference ("parity
unpinned"), so the fixture does not pin them to the reference -- it pins them to THEMSELVES: any later edit of
oracle/mpm_oracle.cpp that changes their arithmetic (as the 3D Jacobi convergence rule did in round 1, and the
Newton-polar rotation of the 3D stress and the polar + symmetric-Jacobi snow projection in round 2: 120 substeps of the 3D cases
moved by 2e-6 relative in x and 2e-4 in v, the growth of ~1e-7 per-substep differences in a collapsing pile) has to
regenerate this file on purpose.  Run:  make -C oracle && python oracle/make_extension_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpm_flip98a_b200 import scenes  # noqa: E402  (scene generators only; no engine involved)
from oracle.cpu import Oracle, make_params  # noqa: E402


def cases():
    p2 = scenes.commented_three_blocks()
    n3 = 16
    dt3, vol3 = scenes.scaled_constants(n3)
    p3 = scenes.collapse_3d(n3, per_side=2, y_top=0.4, xz=(0.2, 0.8))
    return {
        "three_materials_apic": (make_params(alpha=0.0), 1e-4, p2, 800),
        "three_materials_flip": (make_params(alpha=0.95), 1e-4, p2, 800),
        "lift3d_apic": (make_params(dim=3, n_grid=n3, vol_p=vol3), dt3, p3, 120),
        "lift3d_flip": (make_params(dim=3, n_grid=n3, vol_p=vol3, alpha=0.95), dt3, p3, 120),
    }


def run():
    O = Oracle()
    out = {}
    for name, (P, dt, p0, steps) in cases().items():
        p = p0.copy()
        O.advance(P, dt, p, steps)
        assert np.isfinite(p).all(), name
        out[name] = p
    rs = np.random.RandomState(21)
    ms = (np.eye(3).reshape(1, 9) + 0.2 * rs.randn(64, 9)).astype(np.float32)
    out["svd3_in"] = ms
    out["svd3_out"] = np.stack([np.concatenate(O.svd3(m)) for m in ms])
    out["rotation3_out"] = np.stack([O.rotation3(m) for m in ms])
    lo, hi = np.float32(1 - 2.5e-2), np.float32(1 + 7.5e-3)
    ms2 = (np.eye(3).reshape(1, 9) + 0.02 * rs.randn(64, 9)).astype(np.float32)  # around the clamp window
    out["project3_in"] = ms2
    out["project3_out"] = np.stack([np.concatenate([f, [r]]).astype(np.float32)
                                    for f, r in (O.plastic_project3(lo, hi, m) for m in ms2)])
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "oracle_extensions.npz")
    np.savez_compressed(path, **run())
    print(path, os.path.getsize(path))
