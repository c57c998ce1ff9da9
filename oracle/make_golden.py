"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libmpmref.so).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    make -C oracle && python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so these
fixtures -- produced by its own advance() (mls-mpm88-explained.cpp:49-180), its own seeding
(:191-196) and its own polar_decomp/svd (taichi.h:8375-8420) -- are what pins the oracle and,
through it, the CUDA engine.  /root/reference is not readable on the GPU box; the fixtures are.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cpu import Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def bulk(p):
    x, v = p[:, 0:2].astype(np.float64), p[:, 2:4].astype(np.float64)
    return dict(com=x.mean(0), mom=v.sum(0), ke=0.5 * (v ** 2).sum(), jp=p[:, 12].astype(np.float64).mean())


def main():
    os.makedirs(OUT, exist_ok=True)
    R = Reference()
    rng_first = np.array([R.lib.ref_rand(), R.lib.ref_rand()], np.float32)  # taichi.h:6511, first two draws
    # NOTE: those two draws advanced the process-wide RNG; a fresh process is needed for the
    # shipped seeding, so the states below come from a second interpreter (see __main__).
    np.savez_compressed(os.path.join(OUT, "ref_rng.npz"), first_pair=rng_first)


def states():
    R = Reference()
    R.lib.ref_clear()
    R.lib.ref_seed_shipped()
    out = {"step0": R.get()}
    bulks = {}
    done = 0
    for target in (1, 100, 101, 1000, 2500):
        R.advance(target - done)
        done = target
        p = R.get()
        b = bulk(p)
        bulks[target] = np.concatenate([b["com"], b["mom"], [b["ke"], b["jp"]]])
        if target in (1, 100, 101, 1000):
            out["step%d" % target] = p
        if target == 101:
            out["grid101"] = R.grid()  # grid after the update of substep 101: (vx, vy, 1|0)
    out["bulk_steps"] = np.array(sorted(bulks), np.int32)
    out["bulk"] = np.stack([bulks[k] for k in sorted(bulks)])  # com(2) mom(2) ke jp, float64
    out["constants"] = np.array([R.lib.ref_dt(), R.lib.ref_mu0(), R.lib.ref_lambda0()], np.float32)
    np.savez_compressed(os.path.join(OUT, "shipped_scene.npz"), **out)

    # decompositions on random and on near-degenerate 2x2 matrices (column-major 4-vectors)
    rs = np.random.RandomState(1234)
    ms = np.concatenate([
        rs.uniform(-1, 1, (512, 4)),
        np.eye(2).reshape(1, 4) + 1e-3 * rs.randn(256, 4),          # near identity (live F values)
        np.eye(2).reshape(1, 4) + 1e-7 * rs.randn(128, 4),          # exercises |S01| < 1e-6
        np.array([[1, 0, 0, 1], [2, 0, 0, 0.5], [0.5, 0, 0, 2], [0, -1, 1, 0], [1, 1e-7, 1e-7, 1]]),
    ]).astype(np.float32)
    pol = np.zeros((len(ms), 8), np.float32)
    svd = np.zeros((len(ms), 12), np.float32)
    for i, m in enumerate(ms):
        Rm, S = R.polar2(m)
        U, sg, V = R.svd2(m)
        pol[i] = np.concatenate([Rm, S])
        svd[i] = np.concatenate([U, sg, V])
    np.savez_compressed(os.path.join(OUT, "decomp2.npz"), m=ms, polar=pol, svd=svd)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "states":
        states()
    else:
        main()
        import subprocess
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "states"])
        for f in sorted(os.listdir(OUT)):
            print(f, os.path.getsize(os.path.join(OUT, f)))
