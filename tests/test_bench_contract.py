"""bench.py contract pieces that can be checked without a GPU: the reference arm prints exactly one JSON line
with the required keys, and the GPU arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--workload", "c2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-substeps/sec" and d["higher_is_better"] is True
    assert d["value"] > 1e5 and d["steps"] == 2 and d["warmup"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["single_thread_value"] > 1e5  # the reference as shipped has no threads: reported beside the threaded rate


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)


def test_rebalanced_cuts_equalise_measured_cost():
    """bench.rebalanced_cuts (multi-GPU load balance): cuts stay ordered, bin-aligned and inside the grid, and move
    towards the cheaper ranks; equal costs leave equal-width filled shares."""
    sys.path.insert(0, ROOT)
    import bench
    from mpm_flip98a_b200 import parallel
    n, edge, world = 8192, 8, 4
    slabs = parallel.partition_filled(n, world, edge)
    same = bench.rebalanced_cuts(slabs, [6.0] * world, n, edge, (0.05, 0.95))
    assert same == slabs
    new = bench.rebalanced_cuts(slabs, [5.0, 6.0, 6.0, 7.0], n, edge, (0.05, 0.95))
    assert new[0][0] == 0 and new[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(new[:-1], new[1:])) and all(hi - lo >= edge and lo % edge == 0 for lo, hi in new)
    assert new[0][1] > slabs[0][1] and new[-1][0] > slabs[-1][0]  # the cheap first rank grows, the costly last one shrinks
    # the predicted cost of the new slabs (piecewise-constant density model) is flat to within one bin column
    lo_f, hi_f = 0.05 * n, 0.95 * n
    dens = []
    for (lo, hi), t in zip(slabs, [5.0, 6.0, 6.0, 7.0]):
        a, b = max(lo, lo_f), min(hi, hi_f)
        dens.append((a, b, t / (b - a)))
    def cost(lo, hi):
        return sum(d * max(0.0, min(hi, b) - max(lo, a)) for a, b, d in dens)
    pred = [cost(lo, hi) for lo, hi in new]
    assert max(pred) - min(pred) <= 2 * edge * max(d for _, _, d in dens)
    # 3D: two slabs of a 324^3 grid, bin edge 4
    s3 = parallel.partition_filled(324, 2, 4)
    n3 = bench.rebalanced_cuts(s3, [3.6, 4.75], 324, 4, (0.05, 0.95))
    assert n3[0][1] > s3[0][1] and n3[0][1] % 4 == 0
