#!/usr/bin/env python
"""bench.py -- particle-substeps/s of the MLS-MPM substep (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c5] [--impl reference]

A "step" is one substep (P2G -> grid update -> G2P, the reference's advance(dt),
cpp_validation/mls-mpm88-explained.cpp:49-180) over every particle of the workload.
  value      whole-job particle-substeps/s with the state resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the C-ABI with HOST buffers: mpm_upload_particles (pinned H2D) ->
             mpm_substep(frame) -> mpm_read_particles (D2H), all inside the timed region; `frame` is
             the reference's substeps per rendered frame (frame_dt/dt = 10, :11-12,:217)
  roofline   dominant kernel: algorithmic bytes per launch (2D 140 B / 148 B with FLIP,
             3D 260 B per particle-substep, attributed per kernel below) / its mean CUDA-event time
  cpu_baseline  the oracle's restatement of advance() timed on this box's host on a bounded sample of the same
             scene: on all host threads (the threaded oracle is bitwise its serial self) and, beside it, on one
             thread -- the reference is single-threaded as shipped
`--impl reference` times that CPU path alone and prints the same line with "impl": "reference".
`config` holds only what NAMES the workload (workload, name, dim, n_grid, particles, alpha, dt, l2) and is the same object
in both arms of one command line; what is specific to an arm sits beside it: `engine` (kernel path, warm-up, and for
N > 1 the slab decomposition, calibration, exchange and per-rank phases) / `sample` (the CPU arm's bounded sample).
The timed state is a MOVING scene: c4 starts as a cellular flow (scenes.swirl_velocity, peak speed 3 = 0.024 cells
per substep, the range of the reference's own scene) and every workload is warmed on the GPU (2000 substeps for
c4) before anything is timed; the storage re-sorts that fall due run inside the timed steps (their interval is
capped at K for the timed region so that at least one does), and `resort` / `motion` in the output describe them.
Synthetic data (mpm_flip98a_b200/scenes.py); nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec"
SWIRL = 3.0  # peak speed of c4's initial cellular flow (scenes.swirl_velocity)
WARM = {"c1": 1000, "c2": 3000, "c3": 2000, "c4": 2000, "c5": 1000}  # untimed substeps that put the scene in motion
FRAME = 10  # substeps per e2e call == the reference's int(frame_dt/dt), mls-mpm88-explained.cpp:217

# algorithmic bytes per particle-substep, per kernel: P2G reads the full record;
# G2P reads x,F,Jp,mat (+v with FLIP) and writes x,v,C,F,Jp
ALGO = {2: dict(p2g=56, g2p=32 + 52, flip_extra=8), 3: dict(p2g=104, g2p=56 + 100, flip_extra=12)}

WORKLOADS = {
    # name: (description, dim, n_grid, alpha, scene builder kwargs)
    "c1": ("2D 80^2 grid, 3000 particles: the reference program as shipped (BASELINE configs[0])", 2, 80, 0.0),
    "c2": ("2D 512^2 three-material scene, ~1M particles (BASELINE configs[1])", 2, 512, 0.0),
    "c3": ("2D 2048^2 dam-break fluid, ~16M particles, FLIP alpha=0.95 (BASELINE configs[2])", 2, 2048, 0.95),
    "c4": ("2D 8192^2 pool, ~245M particles, three material bands (BASELINE configs[3])", 2, 8192, 0.0),
    "c5": ("3D 256^3 three-material collapse, ~32M particles (BASELINE configs[4])", 3, 256, 0.0),
}


# the GPU tests that cover each benchmarked configuration at full size (pytest -m gpu)
PARITY_TESTS = {
    "c1": "tests/test_gpu_parity.py::test_shipped_scene_one_warm_substep_vs_reference_golden (golden = the unmodified "
          "reference) and ::test_1000_substeps_chaotic_scenes_within_reference_noise",
    "c2": "tests/test_gpu_parity.py::test_config2_one_warm_substep_full_size",
    "c3": "tests/test_gpu_fullsize.py::test_config3_full_size (+ _apic)",
    "c4": "tests/test_gpu_fullsize.py::test_config4_full_size: same scene, same warm-up, 1 substep <= 1e-5 (C: reference "
          "reorder noise), 10-substep bulk <= 1e-3; fluid/jelly bands and FLIP are north_star extensions whose oracle is "
          "unpinned by the reference (SURVEY 8a M1-M2)",
    "c5": "tests/test_gpu_fullsize.py::test_config5_full_size (3D lift: oracle unpinned by the reference, SURVEY 8a M3)",
}


def build_scene(name, n_grid=None):
    from mpm_flip98a_b200 import scenes
    _, dim, n, alpha = WORKLOADS[name]
    n = n_grid or n
    if name == "c1":  # the reference's own seeding (xorshift128, :191-196), from the committed golden fixture
        p = np.load(os.path.join(ROOT, "tests", "golden", "shipped_scene.npz"))["step0"].copy()
    elif name == "c2":
        p = scenes.three_blocks_2d(n, per_side=4)
    elif name == "c3":
        p = scenes.dam_break_2d(n, per_side=3, width=0.47)
    elif name == "c4":
        p = scenes.slab_fill_2d(n, per_side=3, swirl=SWIRL)
    else:
        p = scenes.collapse_3d(n, per_side=2)
    dt, vol = scenes.scaled_constants(n, dim)
    return p, dim, n, alpha, dt, vol


def scaled_n_grid(name, world, scaling):
    """Grid size of a run on `world` GPUs.  weak scaling: particles AND grid nodes per GPU stay what one GPU has at N = 1
    (2D: n_grid = n0*sqrt(N); 3D: n_grid = n0*cbrt(N)), the domain stays the unit box; strong scaling: the N = 1 problem."""
    _, dim, n0, _ = WORKLOADS[name]
    if world == 1 or scaling != "weak":
        return n0
    edge = 8 if dim == 2 else 4
    f = world ** (0.5 if dim == 2 else 1.0 / 3.0)
    align = edge * world if dim == 2 else edge
    return int(round(n0 * f / align)) * align


def workload_particles(name, n_grid):
    """Particles of workload `name` at grid size n_grid, without generating them (the scenes are cell-aligned jittered
    lattices: cells of the filled boxes x particles per cell; checked against the generators in tests/)."""
    from mpm_flip98a_b200 import scenes
    if name == "c1":
        return 3000
    if name == "c2":
        return sum(scenes.box_count((cx - 0.14, cy - 0.14), (cx + 0.14, cy + 0.14), n_grid, 4)
                   for cx, cy in ((0.55, 0.20), (0.45, 0.49), (0.55, 0.78)))
    if name == "c3":
        return scenes.box_count((0.05, 0.05), (0.05 + 0.47, 0.05 + 0.90), n_grid, 3)
    if name == "c4":
        return scenes.slab_fill_2d_count(n_grid)
    return scenes.box_count((0.05, 0.05, 0.05), (0.95, 0.35, 0.95), n_grid, 2)


def config_dict(name, n_grid, particles, dt, world):
    """`config` of the JSON line: the keys that NAME the workload, identical in both arms (ours / --impl reference) of the
    same command line.  What is specific to an arm (kernel path, slab decomposition, the CPU arm's bounded sample) sits
    in its own top-level object."""
    descr, dim, _, alpha = WORKLOADS[name]
    per_gpu = particles / float(world) * (14 if dim == 2 else 26) * 4
    return {"workload": descr, "name": name, "dim": dim, "n_grid": int(n_grid), "particles": int(particles),
            "alpha": alpha, "dt": float(dt),
            "l2": "state (%.1f GB per GPU) larger than L2; no flush" % (per_gpu / 1e9)
            if per_gpu > 2.5e8 else "state fits L2 (small workload)"}


def workload_config(name, world=1, scaling="weak"):
    """config_dict of a command line, computed without touching a GPU (the reference arm uses this)."""
    from mpm_flip98a_b200 import scenes
    dim = WORKLOADS[name][1]
    n_grid = scaled_n_grid(name, world, scaling)
    dt = 1e-4 if name == "c1" else scenes.scaled_constants(n_grid, dim)[0]  # c1: the shipped constant (:11)
    return config_dict(name, n_grid, workload_particles(name, n_grid), dt, world)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a moment before its first row: do not open the timed region before it is sampling"""
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.03] or [r for t, r in self.rows if t0 - 0.1 <= t <= t1 + 0.1] \
            or [r for _, r in self.rows]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(rows)}
        try:
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(rows[0][1])
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in rows)]
            out["power_w_max"] = max(float(r[2]) for r in rows)
        except Exception:
            pass
        return out


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample(name, max_seconds=20.0, threads=1):
    """Times the oracle (CPU restatement of advance()) on a bounded sample of the workload: the same scene
    generator at a reduced grid, warmed by a few substeps.  threads > 1 = the oracle's threaded loops (bitwise
    identical results).  -> (value, description, state)."""
    from oracle.cpu import Oracle, build, make_params
    build()
    O = Oracle()
    _, dim, n_full, alpha = WORKLOADS[name]
    n_small = {2: min(n_full, 1024), 3: min(n_full, 96)}[dim]
    p, dim, n, alpha, dt, vol = build_scene(name, n_small)
    if name == "c1":
        vol, dt = 1.0, 1e-4  # the shipped constants (:11, :18)
    P = make_params(dim=dim, n_grid=n, vol_p=vol, alpha=alpha)
    O.advance(P, dt, p, 20, threads=host_threads())  # a warm, moving state (not part of the measurement)
    t0 = time.perf_counter()
    steps = 0
    while True:
        O.advance(P, dt, p, 1, threads=threads)
        steps += 1
        el = time.perf_counter() - t0
        if el > max_seconds or steps >= 50 or (steps >= 3 and el > max_seconds / 2):
            break
    val = len(p) * steps / el
    desc = "%s scene at n_grid=%d (%d particles), %d substeps, %.1f s, %d thread%s" % (
        name, n, len(p), steps, el, threads, "" if threads == 1 else "s")
    return val, desc, (O, P, dt, p)


def unmodified_reference_c1(steps=300):
    """The UNMODIFIED reference translation unit (oracle/_ref, built from /root/reference in the build container)
    on the only scene it can run -- its own: 80^2 grid, 3000 particles (BASELINE configs[0]).  Information only."""
    try:
        from oracle.cpu import Reference
        if not Reference.available():
            return None
        # the reference's logger greets on stdout when its library is loaded (taichi.h:16302): keep stdout for
        # the one JSON line by pointing fd 1 at stderr while the library loads
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            R = Reference()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        g = np.load(os.path.join(ROOT, "tests", "golden", "shipped_scene.npz"))
        R.set(g["step100"].copy())
        R.advance(20)
        t0 = time.perf_counter()
        R.advance(steps)
        el = time.perf_counter() - t0
        return {"value": 3000 * steps / el, "unit": "particle-substeps/s", "cores": 1, "kind": "reference",
                "sample": "shipped scene (80^2 grid, 3000 particles), %d substeps from the golden step-100 state" % steps}
    except Exception as e:  # never let the information leg break the bench line
        return {"unavailable": str(e)[:200]}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for the path (oracle port -- the reference
    translation unit itself only compiles for its fixed 80^2 scene) on ALL host threads this process may use;
    the single-thread rate (the reference as shipped has no threads) is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = host_threads()
    val1, desc1, (O, P, dt, p) = cpu_sample(args.workload, max_seconds=5.0, threads=1)
    for _ in range(args.warmup):
        O.advance(P, dt, p, 1, threads=T)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.advance(P, dt, p, 1, threads=T)
    el = time.perf_counter() - t0
    value = len(p) * args.steps / el
    sample = "one substep over a bounded sample per step: %s scene at n_grid=%d, %d particles" % (
        args.workload, P.n_grid, len(p))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "particle-substeps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus, args.scaling),  # == the GPU arm's of this command line
            "sample": sample,
            "cpu_baseline": {"value": value, "unit": "particle-substeps/s", "cores": T, "kind": "port",
                             "sample": sample, "single_thread_value": val1, "single_thread_sample": desc1},
            "e2e": {"value": value, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "host_cores_total": os.cpu_count(),
            "unmodified_reference_c1": unmodified_reference_c1()}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--warm-substeps", type=int, default=None,
                    help="untimed substeps that put the scene in motion before anything is timed "
                         "(default per workload: %s)" % WARM)
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--naive", action="store_true", help="one-thread-per-particle kernels (MPM_FLAG_NAIVE)")
    ap.add_argument("--no-fuse", action="store_true", help="separate P2G and G2P kernels (MPM_FLAG_NO_FUSE)")
    ap.add_argument("--fuse-3d", action="store_true", help="3D: the fused G2P->P2G kernel (MPM_FLAG_FUSE_3D, opt-in)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: plain slab schedule (default: interior bins on a side stream while the slab boundary "
                         "is exchanged, MPM_FLAG_OVERLAP)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak (per-GPU work of N=1, default) or strong (the N=1 problem cut into N slabs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-rebalance", action="store_true",
                    help="N > 1: keep the equal-particle-count slabs (default: move the cuts to equalise measured cost)")
    ap.add_argument("--rebin-every", type=int, default=0, help="storage re-sort interval (0 = engine default)")
    args = ap.parse_args()
    if args.warm_substeps is None:
        args.warm_substeps = WARM[args.workload]
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import mpm_flip98a_b200 as mpm
    from mpm_flip98a_b200.engine import FLAG_NAIVE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
    if world > 1:
        return run_slabs(args, rank, world, local)

    descr, dim, n_grid, alpha = WORKLOADS[args.workload]
    p_np, dim, n_grid, alpha, dt, vol = build_scene(args.workload)
    if args.workload == "c1":
        vol, dt = 1.0, 1e-4  # the shipped constants (:11, :18); material = the shipped one (colour slot -> last entry)
    n = len(p_np)
    words = p_np.shape[1]
    host = torch.empty((n, words), dtype=torch.float32, pin_memory=True)  # pinned: e2e copies are async DMA
    host.numpy()[:] = p_np
    del p_np
    host_out = torch.empty_like(host, pin_memory=True)

    stream = torch.cuda.Stream()
    flags = FLAG_NAIVE if args.naive else ((16 if args.no_fuse else 0) | (128 if args.fuse_3d else 0))
    with torch.cuda.stream(stream):
        eng = mpm.Engine(dim=dim, n_grid=n_grid, capacity=n, dt=dt, vol_p=vol, alpha=alpha, device=local,
                         flags=flags, stream=stream.cuda_stream, rebin_every=args.rebin_every)
        eng.lib.mpm_upload_particles(eng.h, host.data_ptr(), n, 0)
        eng.substep(args.warm_substeps)
        eng.synchronize()
        status = eng.poll_status()
        if status != 0:
            raise SystemExit("engine status %d after warm-up: %s" % (status, eng.lib.mpm_last_error(eng.h)))

        # ---- value: K substeps on HBM-resident state, CUDA events on the launching stream -------
        # long window first (side information): what a substep costs over a few hundred substeps with every
        # storage re-sort at its own adaptive interval
        eng.profile_enable(True)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_long = min(320, max(64, args.warm_substeps // 4))
        l0.record(stream)
        eng.substep(n_long)
        l1.record(stream)
        torch.cuda.synchronize()
        long_prof = eng.profile()
        long_ms = l0.elapsed_time(l1)
        interval = int(long_prof["rebin_interval"])
        # timed region: the re-sort interval is capped at K so that at least one re-sort runs INSIDE the K timed
        # steps whatever their phase (never fewer than the adaptive schedule would run: conservative)
        timed_interval = max(1, min(interval, args.steps)) if interval > 0 else interval
        if args.rebin_every == 0 and not args.naive:
            eng.set_rebin_every(timed_interval)
        for _ in range(args.warmup):
            eng.substep(1)
        eng.synchronize()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        sampler.wait_first()
        eng.substep(3)  # the GPU is already under load when the region opens (clock samples are of a busy GPU)
        eng.synchronize()
        eng.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            eng.substep(1)
        e1.record(stream)
        torch.cuda.synchronize()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        prof = eng.profile()
        eng.profile_enable(False)
        clocks = sampler.stop(t_wall0, t_wall1)
        if eng.poll_status() != 0:
            raise SystemExit("engine flagged an error during the timed region")
        if args.rebin_every == 0 and not args.naive:
            eng.set_rebin_every(0)
        value = n * args.steps / (ms * 1e-3)
        prof["long_window"] = {"substeps": n_long, "ms_per_step": long_ms / n_long,
                               "value": n * n_long / (long_ms * 1e-3), "adaptive_interval": interval,
                               "bin_ms_per_step": long_prof["bin"][0] / n_long,
                               "fallback_fraction": long_prof["fallback_particles"] / float(n * n_long)}
        prof["timed_interval"] = timed_interval
        # what the timed state looks like: read it back once (outside any timed region) and sample it
        eng.lib.mpm_read_particles(eng.h, host_out.data_ptr(), n, 0)
        prof["motion"] = motion_stats(host_out.numpy()[::257], dim, n_grid, dt)

        # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----------------
        eng.lib.mpm_upload_particles(eng.h, host.data_ptr(), n, 0)  # warm the path once
        eng.substep(FRAME)
        eng.lib.mpm_read_particles(eng.h, host_out.data_ptr(), n, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_calls):
            rc = eng.lib.mpm_upload_particles(eng.h, host.data_ptr(), n, 0)
            assert rc == 0
            eng.substep(FRAME)
            rc = eng.lib.mpm_read_particles(eng.h, host_out.data_ptr(), n, 0)
            assert rc == 0
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_value = n * FRAME * args.e2e_calls / e2e_s
        assert torch.isfinite(host_out[:: max(1, n // 100000)]).all()
        eng.close()

    line = make_line(args, world, n, n, words, dim, n_grid, alpha, dt, descr, ms, value, e2e_value, prof, clocks,
                     scaling="weak", extra_config={})
    if not args.no_cpu:
        T = host_threads()
        val, desc, _ = cpu_sample(args.workload, max_seconds=12.0, threads=T)
        val1, desc1, _ = cpu_sample(args.workload, max_seconds=8.0, threads=1)
        line["cpu_baseline"] = {"value": val, "unit": "particle-substeps/s", "cores": T, "kind": "port",
                                "sample": desc, "host_cores_total": os.cpu_count(),
                                "single_thread_value": val1, "single_thread_sample": desc1,
                                "note": "the reference is single-threaded as shipped; the threaded oracle is bitwise "
                                        "its serial self",
                                "unmodified_reference_c1": unmodified_reference_c1()}
    emit(line)


def motion_stats(sample, dim, n_grid, dt):
    """How dynamic the timed state is (a strided sample of the particles read back after the timed region)."""
    d, dd = dim, dim * dim
    v = sample[:, d:2 * d].astype(np.float64)
    speed = np.sqrt((v ** 2).sum(1))
    out = {"rms_speed": float(np.sqrt((speed ** 2).mean())), "max_speed": float(speed.max()),
           "cells_per_substep_rms": float(np.sqrt((speed ** 2).mean()) * dt * n_grid),
           "cells_per_substep_max": float(speed.max() * dt * n_grid),
           "rms_velocity_gradient": float(np.sqrt((sample[:, 2 * d + dd:2 * d + 2 * dd].astype(np.float64) ** 2).mean())),
           "sampled_particles": int(len(sample))}
    mat = sample[:, -1].view(np.int32)
    out["material_fractions"] = {str(k): float((mat == k).mean()) for k in np.unique(mat)[:4]}
    if dim == 2:
        # snow: which branch of the reference's 2x2 svd the plastic projection takes (taichi.h:8393): the cheap
        # one needs |S(0,1)| < 1e-6 of the polar factor of (I + dt C) F
        sn = sample[mat == 2]
        if len(sn):
            F = sn[:, 4:8].astype(np.float64).reshape(-1, 2, 2).transpose(0, 2, 1)   # column-major records
            C = sn[:, 8:12].astype(np.float64).reshape(-1, 2, 2).transpose(0, 2, 1)
            Fn = (np.eye(2) + dt * C) @ F
            x, y = Fn[:, 0, 0] + Fn[:, 1, 1], Fn[:, 1, 0] - Fn[:, 0, 1]
            sc = 1.0 / np.sqrt(x * x + y * y)
            c, s_ = x * sc, y * sc
            s01 = c * Fn[:, 0, 1] + s_ * Fn[:, 1, 1]
            out["snow_svd_general_branch_fraction"] = float((np.abs(s01) >= 1e-6).mean())
            out["snow_Jp_range"] = [float(sn[:, 12].min()), float(sn[:, 12].max())]
    return out


def make_line(args, world, n_total, n_local, words, dim, n_grid, alpha, dt, descr, ms, value, e2e_value, prof, clocks,
              scaling, extra_config):
    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)"
    a = ALGO[dim]
    algo = {"p2g": a["p2g"], "g2p": a["g2p"] + (a["flip_extra"] if alpha != 0 else 0)}
    phases = {k: prof[k] for k in ("clear", "p2g", "grid", "g2p", "bin", "halo", "migrate")}
    if prof.get("fused_substeps", 0) > 0:
        # G2P of a substep and P2G of the next run as ONE kernel (timed under "g2p"): it owns the whole
        # algorithmic traffic of a substep; stand-alone P2G launches (first substep after an upload) are added
        dom = "g2p2g"
        algo[dom] = algo["p2g"] + algo["g2p"]
        dom_ms = (phases["g2p"][0] + phases["p2g"][0]) / max(1, prof["substeps"])
    elif extra_config.get("overlap_3d"):
        # overlapped 3D slab schedule: G2P and the next P2G of the interior run back to back on the side stream and are
        # timed as ONE span (under "g2p"): the pair owns the whole algorithmic traffic of a substep
        dom = "g2p+p2g"
        algo[dom] = algo["p2g"] + algo["g2p"]
        dom_ms = (phases["g2p"][0] + phases["p2g"][0]) / max(1, prof["substeps"])
    else:
        dom = None
    if dom is not None and extra_config.get("overlap"):
        # overlapped slab schedule: the same kernels' launches over the bins next to the cuts are timed under "migrate"
        # (with the immigrant unpack); they process part of this rank's particles, so their time belongs to the kernel
        dom_ms += phases["migrate"][0] / max(1, prof["substeps"])
    if dom is None:
        dom = max(("p2g", "g2p"), key=lambda k: phases[k][0])
        dom_ms = phases[dom][0] / max(1, prof["substeps"])
    achieved = algo[dom] * n_local / (dom_ms * 1e-3) / 1e9  # per launch == this rank's particles
    whole = (algo["p2g"] + algo["g2p"]) * value / world / 1e9  # per GPU
    traffic = None
    try:  # measured DRAM bytes per launch of that kernel (ncu --set full, see profiles/), same particle count
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(args.workload, {})
        if abs(t.get("particles", -1) - n_local) <= 0.01 * n_local:
            traffic = t.get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                "algorithmic_bytes_per_particle": algo[dom], "kernel_ms": dom_ms,
                "whole_substep": {"algorithmic_bytes_per_particle": algo["p2g"] + algo["g2p"],
                                  "achieved": whole, "frac": whole / peak, "frac_of_nominal_8TBs": whole / 8000.0},
                "phase_ms_per_substep": {k: v[0] / max(1, prof["substeps"]) for k, v in phases.items()}}
    launches = int(sum(v[1] for v in phases.values()))

    cfg = config_dict(args.workload, n_grid, n_total, dt, world)
    engine = {"path": "naive" if args.naive else ("binned, fused G2P->P2G" if prof.get("fused_substeps", 0) else "binned"),
              "warm_substeps": args.warm_substeps}
    engine.update(extra_config)  # N > 1: the slab decomposition, calibration, exchange, per-rank phases
    return {"metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "engine": engine,
            "e2e": {"value": e2e_value, "unit": "particle-substeps/s", "h2d_bytes_per_step": n_total * words * 4,
                    "d2h_bytes_per_step": n_total * words * 4, "substeps_per_step": FRAME,
                    "call": "mpm_upload_particles + %d substeps + mpm_read_particles, pinned host buffers" % FRAME},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "fallback_particles": prof.get("fallback_particles", 0),
            "resort": resort_info(prof, ms, args.steps, n_total),
            "motion": prof.get("motion"),
            "parity": PARITY_TESTS.get(args.workload)}


def resort_info(prof, ms, steps, n_total):
    """The storage re-sort is periodic (adaptive interval).  Re-sorts that fall due run INSIDE the timed steps and
    are part of `value`; this object says how many did, what they cost, and what a long window measures."""
    out = {"interval_substeps_adaptive": prof.get("long_window", {}).get("adaptive_interval", prof.get("rebin_interval", 0)),
           "interval_substeps_timed_region": prof.get("timed_interval", prof.get("rebin_interval", 0)),
           "resorts_inside_timed_region": int(round(prof["bin"][1] / 6.0)) if prof["bin"][1] else 0,
           "bin_phase_ms_inside_timed_region": prof["bin"][0],
           "how": "count+rank pass (12 B/particle), scan, then the substep kernel itself writes the new order "
                  "(RESORT variant); no separate reorder pass on the 2D default path"}
    if "long_window" in prof:
        out["long_window"] = prof["long_window"]
    return out


def slab_plan(args, world):
    """grid size, slab cuts and per-slab scene generator of a multi-GPU run.
    weak scaling: particles AND grid nodes per GPU stay what one GPU has at N = 1 (2D: n_grid = 8192*sqrt(N);
    3D: n_grid = 256*cbrt(N)), the domain stays the unit box; strong scaling: the N = 1 problem cut in N."""
    from mpm_flip98a_b200 import parallel, scenes
    name = args.workload
    descr, dim, n0, alpha = WORKLOADS[name]
    edge = 8 if dim == 2 else 4
    n_grid = scaled_n_grid(name, world, args.scaling)
    x_range = (0.05, 0.95) if name in ("c4", "c5") else (0.05, 0.52)
    slabs = parallel.partition_filled(n_grid, world, edge, x_range=x_range)  # equal particle counts

    def generate(lo, hi, out=None):
        if name == "c4":
            return scenes.slab_fill_2d(n_grid, columns=(lo, hi), out=out, swirl=SWIRL)
        if name == "c5":
            return scenes.collapse_3d(n_grid, per_side=2, columns=(lo, hi))
        rec = scenes.dam_break_2d(n_grid, per_side=3, width=0.47)  # c3: small enough to generate whole
        return rec[(rec[:, 0] >= np.float32(lo / n_grid)) & (rec[:, 0] < np.float32(hi / n_grid))]
    return descr, dim, n_grid, alpha, slabs, generate


def rebalanced_cuts(slabs, costs, n_grid, edge, x_range):
    """New slab cuts that equalise the MEASURED per-rank cost, assuming each rank's cost is spread evenly over the filled
    columns it owns (piecewise-constant cost density along x)."""
    world = len(slabs)
    lo_f, hi_f = x_range[0] * n_grid, x_range[1] * n_grid
    segs = []
    for (lo, hi), t in zip(slabs, costs):
        a, b = max(lo, lo_f), min(hi, hi_f)
        segs.append((a, b, t / max(b - a, 1e-9)))
    total = float(sum(costs))
    cuts = [0]
    for r in range(1, world):
        target, acc, x = total * r / world, 0.0, segs[-1][1]
        for a, b, d in segs:
            c = d * (b - a)
            if acc + c >= target:
                x = a + (target - acc) / d
                break
            acc += c
        cut = int(round(x / edge)) * edge
        cuts.append(min(max(cut, cuts[-1] + edge), n_grid - edge * (world - r)))
    cuts.append(n_grid)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def run_slabs(args, rank, world, local):
    """N > 1: one x-slab per rank (mpm_flip98a_b200/parallel.py): ONE fixed-size NCCL P2P message per neighbour and
    substep (ghost-column sums + emigrants), no host synchronisation; by default the interior bins run on a side
    stream while the boundary is exchanged (MPM_FLAG_OVERLAP).  The slabs start as equal particle counts; a short
    calibration run (the full warm-up) measures each rank's device time per substep and, when the slowest is more than 1.5 %
    above the mean, the cuts are
    moved to equalise the measured cost (the material bands of c4 / c5 lie along x and cost differently) and the
    ranks regenerate their particles -- all before anything is timed."""
    import torch
    import torch.distributed as dist
    import mpm_flip98a_b200 as mpm
    from mpm_flip98a_b200 import parallel, scenes
    from mpm_flip98a_b200.engine import FLAG_NAIVE, FLAG_OVERLAP
    if args.workload == "c2":
        raise SystemExit("multi-GPU bench: workloads c4 (default), c5, c3")
    descr, dim, n_grid, alpha, slabs, generate = slab_plan(args, world)
    words = 14 if dim == 2 else 26
    edge = 8 if dim == 2 else 4
    x_range = (0.05, 0.95) if args.workload in ("c4", "c5") else (0.05, 0.52)
    dt, vol = scenes.scaled_constants(n_grid, dim)
    dev = "cuda:%d" % local
    overlap = not args.no_overlap and not args.naive and (dim == 3 or not args.no_fuse)
    # with the overlapped schedule the engine's interior launch runs on a lowest-priority side stream; the main stream
    # (boundary bins, exchange helpers) and NCCL's own stream (TORCH_NCCL_HIGH_PRIORITY, set in __main__) outrank it
    stream = torch.cuda.Stream(priority=-1) if overlap else torch.cuda.Stream()
    flags = FLAG_NAIVE if args.naive else ((FLAG_OVERLAP if overlap else 0) | (16 if args.no_fuse else 0) |
                                           (128 if args.fuse_3d else 0))
    PH = ("clear", "p2g", "grid", "g2p", "bin", "halo", "migrate")

    def setup(slabs):
        """this rank's particles for `slabs` (generated straight into a pinned upload buffer and compacted in place,
        chunk by chunk: no second copy), an engine handle, the exchange object"""
        lo, hi = slabs[rank]
        if args.workload == "c4":
            n_max = scenes.slab_fill_2d_count(n_grid, columns=(lo, hi + 1))
            host = torch.empty((n_max, words), dtype=torch.float32, pin_memory=True)
            rec = generate(lo, hi + 1, out=host.numpy())
        else:
            rec = generate(lo, hi + 1)
            host = torch.empty((len(rec), words), dtype=torch.float32, pin_memory=True)
            host.numpy()[:] = rec
            rec = host.numpy()[:len(rec)]
        n_local = 0
        for c0 in range(0, len(rec), 1 << 22):
            blk = rec[c0:c0 + (1 << 22)]
            b = parallel.base_column(blk[:, 0], n_grid)
            kept = blk[(b >= lo) & (b < hi)]
            host.numpy()[n_local:n_local + len(kept)] = kept
            n_local += len(kept)
        del rec
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = n_local
        dist.all_reduce(counts)
        n_total = int(counts.sum())
        first_id = int(counts[:rank].sum())
        assert n_total < 2 ** 31, "int32 particle ids"
        ids = torch.arange(first_id, first_id + n_local, dtype=torch.int32).pin_memory()
        cap = int(n_local * 1.1) + 65536
        eng = mpm.Engine(dim=dim, n_grid=n_grid, capacity=cap, dt=dt, vol_p=vol, alpha=alpha, device=local,
                         flags=flags, stream=stream.cuda_stream, rebin_every=args.rebin_every, slab=(lo, hi))
        up = lambda: eng._check(eng.lib.mpm_upload_particles_ids(eng.h, host.data_ptr(), ids.data_ptr(), n_local, 0))
        up()
        r = parallel.SlabRank(eng, rank, world, dev)
        ex = parallel.DistExchange(r, shared_stream=True)  # engine and NCCL ops are ordered on `stream`
        return host, ids, n_local, n_total, cap, eng, up, r, ex

    with torch.cuda.stream(stream):
        balance = {"passes": 0, "cost_ms_per_rank": None, "imbalance_max_over_mean": None}
        for attempt in range(2):  # at most ONE rebalancing pass (a pass regenerates and re-warms the scene: ~1.5 min at c4)
            host, ids, n_local, n_total, cap, eng, up, r, ex = setup(slabs)
            # calibration: each rank's device time per substep in the warm, moving state that will be timed
            parallel.step_dist(r, ex, args.warm_substeps, settle=False)
            eng.profile_enable(True)
            parallel.step_dist(r, ex, 24, settle=False)
            prof = eng.profile()
            eng.profile_enable(False)
            mine = sum(prof[k][0] for k in PH) / 24.0
            costs = torch.zeros(world, dtype=torch.float64, device=dev)
            costs[rank] = mine
            dist.all_reduce(costs)
            costs = [float(c) for c in costs]
            imb = max(costs) / (sum(costs) / world)
            balance.update(cost_ms_per_rank=[round(c, 4) for c in costs], imbalance_max_over_mean=round(imb, 4))
            if imb <= 1.015 or attempt == 1 or args.no_rebalance:
                break
            new = rebalanced_cuts(slabs, costs, n_grid, edge, x_range)
            if new == slabs:
                break
            eng.close()
            del host, ids, eng, up, r, ex
            slabs = new
            balance["passes"] += 1
        lo, hi = slabs[rank]
        ids_out = torch.empty(cap, dtype=torch.int32).pin_memory()
        host_out = torch.empty((cap, words), dtype=torch.float32, pin_memory=True)  # e2e read-back (storage order)
        if eng.poll_status() != 0:
            raise SystemExit("rank %d: engine status after warm-up: %s" % (rank, eng.lib.mpm_last_error(eng.h)))
        # cap the re-sort interval at K for the timed region (at least one re-sort inside, see the module docstring)
        interval = int(eng.profile()["rebin_interval"])
        it = torch.tensor([interval], dtype=torch.int64, device=dev)
        dist.all_reduce(it, op=dist.ReduceOp.MIN)
        timed_interval = max(1, min(int(it), args.steps))
        if args.rebin_every == 0 and not args.naive:
            eng.set_rebin_every(timed_interval)
        parallel.step_dist(r, ex, args.warmup, settle=False)
        sampler = ClockSampler(local)
        sampler.start()
        sampler.wait_first()
        dist.barrier()
        torch.cuda.synchronize()
        eng.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        t_wall0 = time.time()
        e0.record(stream)
        parallel.step_dist(r, ex, args.steps, settle=False)
        eng.synchronize()  # the last interior launch runs on the engine's side stream: join it before closing the region
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t_wall1 = time.time()
        ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)  # device time, max over ranks
        ms = float(ms_t)
        prof = eng.profile()
        eng.profile_enable(False)
        clocks = sampler.stop(t_wall0, t_wall1)
        if args.rebin_every == 0 and not args.naive:
            eng.set_rebin_every(0)
        eng.slab_settle()
        if eng.poll_status() != 0:
            raise SystemExit("rank %d: engine flagged an error during the timed region: %s"
                             % (rank, eng.lib.mpm_last_error(eng.h)))
        live = torch.tensor([eng.count], dtype=torch.int64, device=dev)
        dist.all_reduce(live)
        assert int(live) == n_total, "particles lost in migration: %d != %d" % (int(live), n_total)
        value = n_total * args.steps / (ms * 1e-3)
        prof["timed_interval"] = timed_interval
        # device time of this rank's phases vs the wall of the step: what the exchange costs on top of the compute
        phase_sum = sum(prof[k][0] for k in PH) / args.steps
        prof["exchange_bubble_ms_per_step"] = ms / args.steps - phase_sum
        ph_all = torch.zeros(world, dtype=torch.float64, device=dev)
        ph_all[rank] = phase_sum
        dist.all_reduce(ph_all)
        n_all = torch.zeros(world, dtype=torch.int64, device=dev)
        n_all[rank] = n_local
        dist.all_reduce(n_all)

        # ---- e2e: every rank uploads its host buffer, FRAME substeps, reads its particles back -------
        def e2e_call():
            up()
            parallel.step_dist(r, ex, FRAME)
            got = eng.lib.mpm_read_particles_ids(eng.h, host_out.data_ptr(), ids_out.data_ptr(), ids_out.numel(), 0)
            assert got >= 0, eng.lib.mpm_last_error(eng.h)
        e2e_call()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_calls):
            e2e_call()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_value = n_total * FRAME * args.e2e_calls / float(e2e_t)
        d = eng.slab()
        msg_bytes = int(d.bytes)
        eng.close()
    if rank == 0:
        line = make_line(args, world, n_total, n_local, words, dim, n_grid, alpha, dt, descr, ms, value, e2e_value, prof,
                         clocks, scaling=args.scaling,
                         extra_config={"overlap": bool(overlap),
                                       "overlap_3d": bool(overlap and dim == 3 and not args.fuse_3d),
                                       "decomposition": "x-slabs of %d..%d columns per GPU: cut for equal particle counts, "
                                                        "then moved to equalise the device time per substep each rank "
                                                        "measured in a calibration run (%d rebalancing pass%s)"
                                                        % (min(b - a for a, b in slabs), max(b - a for a, b in slabs),
                                                           balance["passes"], "" if balance["passes"] == 1 else "es"),
                                       "slabs": slabs, "particles_per_rank": [int(x) for x in n_all],
                                       "calibration": balance,
                                       "phase_ms_per_substep_per_rank": [round(float(x), 4) for x in ph_all],
                                       "scaling_rule": ("weak: n_grid = %d (N=1: %d), same fill fractions -> particles and "
                                                        "nodes per GPU as at N=1" % (n_grid, WORKLOADS[args.workload][2]))
                                       if args.scaling == "weak" else "strong: the N=1 problem cut into N slabs",
                                       "exchange": "one fixed-size NCCL P2P message per neighbour and substep (%d bytes: "
                                                   "2 ghost node columns + emigrant count + records), no host "
                                                   "synchronisation%s" % (msg_bytes, ", overlapped with the interior bins"
                                                                          if overlap else ""),
                                       "exchange_bubble_ms_per_step": prof["exchange_bubble_ms_per_step"]})
        emit(line)
    dist.barrier()
    dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else this process or its libraries print (NCCL greets
    with its version on stdout, the reference library's logger too) has been pointed at stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and "--no-overlap" not in sys.argv:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run (C libraries included)
    main()
