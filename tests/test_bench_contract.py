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
    # `config` names the workload of the command line -- the same object the GPU arm prints -- and the bounded sample the
    # CPU arm actually ran is described beside it
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("c2") and d["config"]["particles"] == 979264
    assert "n_grid=512" in d["sample"] and d["sample"] == cb["sample"]
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


def _fake_profile(substeps, g2p, p2g, migrate, fused):
    ph = {"clear": (0.0, 0), "p2g": p2g, "grid": (0.3 * substeps, substeps), "g2p": g2p, "bin": (1.9, 6),
          "halo": (0.01 * substeps, 2 * substeps), "migrate": migrate}
    ph.update(substeps=substeps, fallback_particles=0, rebin_interval=24, fused_substeps=substeps if fused else 0)
    return ph


def test_roofline_of_the_json_line_is_recomputable():
    """bench.make_line: `roofline.achieved` = algorithmic bytes per particle-substep x this rank's particles / the dominant
    kernel's mean time.  Single GPU: the fused kernel's span; overlapped x-slab runs (2D and 3D): the interior span PLUS
    the boundary launches of the same kernels, which the engine times under "migrate"."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    K, n = 20, 244558728
    args = argparse.Namespace(steps=K, warmup=5, workload="c4", naive=False, warm_substeps=2000)
    clocks = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 6}
    # single GPU, fused 2D kernel
    prof = _fake_profile(K, g2p=(5.9 * K, K), p2g=(0.0, 0), migrate=(0.0, 0), fused=True)
    d = bench.make_line(args, 1, n, n, 14, 2, 8192, 0.0, 1e-6, "c4", 6.3 * K, n / 6.3e-3, 4e9, prof, clocks,
                        scaling="weak", extra_config={})
    r = d["roofline"]
    assert r["kernel"] == "g2p2g" and r["algorithmic_bytes_per_particle"] == 140
    assert abs(r["kernel_ms"] - 5.9) < 1e-9 and abs(r["achieved"] - 140 * n / 5.9e-3 / 1e9) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and d["gpu_launches"] == K + K + 6 + 2 * K
    assert d["ms_per_step"] == pytest.approx(6.3) and d["config"]["name"] == "c4" and d["parity"]
    json.dumps(d)
    # 2 GPUs, overlapped 2D schedule: boundary bins under "migrate"
    prof = _fake_profile(K, g2p=(5.8 * K, K), p2g=(0.0, 0), migrate=(0.1 * K, 5 * K), fused=True)
    d = bench.make_line(args, 2, 2 * n, n, 14, 2, 11584, 0.0, 1e-6, "c4", 6.4 * K, 2 * n / 6.4e-3, 7e9, prof, clocks,
                        scaling="weak", extra_config={"overlap": True, "overlap_3d": False})
    assert d["roofline"]["kernel"] == "g2p2g" and d["roofline"]["kernel_ms"] == pytest.approx(5.9)
    assert d["engine"]["overlap"] is True and d["n_gpus"] == 2 and "overlap" not in d["config"]
    # 2 GPUs, overlapped two-kernel 3D schedule: interior G2P + P2G on the side stream are one "g2p" span
    args5 = argparse.Namespace(steps=K, warmup=5, workload="c5", naive=False, warm_substeps=400)
    n5 = 32163200
    prof = _fake_profile(K, g2p=(3.0 * K, 2 * K), p2g=(0.0, 0), migrate=(0.3 * K, 8 * K), fused=False)
    d = bench.make_line(args5, 2, 2 * n5, n5, 26, 3, 324, 0.0, 3e-5, "c5", 3.6 * K, 2 * n5 / 3.6e-3, 3e9, prof, clocks,
                        scaling="weak", extra_config={"overlap": True, "overlap_3d": True})
    r = d["roofline"]
    assert r["kernel"] == "g2p+p2g" and r["algorithmic_bytes_per_particle"] == 260 and r["kernel_ms"] == pytest.approx(3.3)
    # single GPU 3D: two kernels, the slower one is reported with its own share of the bytes
    prof = _fake_profile(K, g2p=(1.48 * K, K), p2g=(1.73 * K, K), migrate=(0.0, 0), fused=False)
    d = bench.make_line(args5, 1, n5, n5, 26, 3, 256, 0.0, 3e-5, "c5", 3.35 * K, n5 / 3.35e-3, 2e9, prof, clocks,
                        scaling="weak", extra_config={})
    r = d["roofline"]
    assert r["kernel"] == "p2g" and r["algorithmic_bytes_per_particle"] == 104 and r["kernel_ms"] == pytest.approx(1.73)
    assert r["whole_substep"]["algorithmic_bytes_per_particle"] == 260


def test_both_arms_print_the_same_config():
    """bench.workload_config (what `--impl reference` prints as `config`) against the committed GPU-arm lines of this round
    (profiles/r02_bench_*.json: N = 1, 2, 8; weak and strong; 2D and 3D) and against the scene generators at small sizes."""
    sys.path.insert(0, ROOT)
    import glob
    import bench
    from mpm_flip98a_b200 import scenes
    seen = 0
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_c*.json"))):
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c = d["config"]
        w = bench.workload_config(c["name"], d["n_gpus"], d["scaling"])
        for k in ("workload", "name", "dim", "n_grid", "particles", "alpha", "dt"):
            assert w[k] == c[k], (os.path.basename(f), k, w[k], c[k])
        seen += 1
    assert seen >= 9
    assert bench.workload_particles("c2", 256) == len(scenes.three_blocks_2d(256, per_side=4))
    assert bench.workload_particles("c3", 320) == len(scenes.dam_break_2d(320, per_side=3, width=0.47))
    assert bench.workload_particles("c4", 328) == len(scenes.slab_fill_2d(328, per_side=3))
    assert bench.workload_particles("c5", 44) == len(scenes.collapse_3d(44, per_side=2))
    assert bench.workload_particles("c5", 44) == len(scenes.collapse_3d(44, per_side=2, columns=(0, 20))) + \
        len(scenes.collapse_3d(44, per_side=2, columns=(20, 44)))
    # the GPU arm's make_line and the CPU arm's workload_config go through the same config_dict
    assert bench.workload_config("c4", 8)["l2"].startswith("state (13.7 GB per GPU)")
    assert bench.workload_config("c1") == bench.config_dict("c1", 80, 3000, 1e-4, 1)
