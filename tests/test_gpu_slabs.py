"""x-slab decomposition on the GPU: N engine handles (slabs) in one process on cuda:0 with
device-to-device exchange (parallel.LocalExchange) must reproduce the oracle / the single-handle
engine: one warm substep within 1e-5, and over many substeps with real migration no particle is lost,
duplicated or mislabelled."""
import numpy as np
import pytest

import mpm_flip98a_b200 as mpm
from mpm_flip98a_b200 import parallel, scenes
from mpm_flip98a_b200.engine import FLAG_FUSE_3D, FLAG_NAIVE, FLAG_NO_FUSE
from oracle.cpu import make_params
from tests.util import bits, fields, rel_l2

pytestmark = pytest.mark.gpu


def run_slabs(p, dim, n_grid, world, steps, dt, vol_p, alpha=0.0, flags=0, rebin_every=0, shared_stream=False):
    ranks, ex, slabs = parallel.make_local_cluster(mpm.Engine, p, dim, n_grid, world, dt=dt, vol_p=vol_p,
                                                   alpha=alpha, flags=flags, rebin_every=rebin_every,
                                                   shared_stream=shared_stream)
    parallel.step_local(ranks, ex, steps)
    out = parallel.collect_local(ranks, len(p), p.shape[1])
    status = [r.e.poll_status() for r in ranks]
    counts = [r.e.count for r in ranks]
    for r in ranks:
        r.e.close()
    return out, status, counts, slabs


@pytest.mark.parametrize("flags", [FLAG_NAIVE, 0, FLAG_NO_FUSE])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_one_warm_substep_matches_oracle(oracle, shipped, world, flags):
    p = shipped["step1000"]  # spread over x in [0.05, 0.97]: every slab owns particles
    want = p.copy()
    oracle.advance(make_params(), 1e-4, want, 1)
    got, status, counts, slabs = run_slabs(p, 2, 80, world, 1, 1e-4, 1.0, flags=flags)
    assert status == [0] * world and sum(counts) == len(p)
    fw, fg = fields(want, 2), fields(got, 2)
    for k in fw:
        assert rel_l2(fg[k], fw[k]) <= 1e-5, (k, rel_l2(fg[k], fw[k]))
    assert np.array_equal(bits(got[:, -1]), bits(p[:, -1]))


@pytest.mark.parametrize("flags", [FLAG_NAIVE, 0, FLAG_NO_FUSE, FLAG_FUSE_3D])
def test_3d_slabs_one_warm_substep(oracle, flags):
    n = 32
    dt, vol = scenes.scaled_constants(n)
    p = scenes.collapse_3d(n, per_side=2, y_top=0.4, xz=(0.15, 0.85))
    P = make_params(dim=3, n_grid=n, vol_p=vol)
    oracle.advance(P, dt, p, 80)
    want = p.copy()
    oracle.advance(P, dt, want, 1)
    got, status, counts, _ = run_slabs(p, 3, n, 2, 1, dt, vol, flags=flags)
    assert status == [0, 0] and sum(counts) == len(p)
    fw, fg = fields(want, 3), fields(got, 3)
    for k in fw:
        assert rel_l2(fg[k], fw[k]) <= 1e-5, (k, rel_l2(fg[k], fw[k]))


@pytest.mark.parametrize("flags", [FLAG_NAIVE, 0, FLAG_NO_FUSE])
def test_migration_conserves_particles_and_tracks_the_single_handle_run(oracle, flags):
    # jelly block thrown sideways across three slab cuts: well-conditioned, so slabs vs one handle
    # vs the CPU oracle must agree closely even after hundreds of substeps
    p = scenes.jelly_drop()
    p[:, 2] = 6.0  # vx: crosses ~0.3 of the domain in 500 substeps
    steps = 500
    want = p.copy()
    oracle.advance(make_params(), 1e-4, want, steps)
    got, status, counts, slabs = run_slabs(p, 2, 80, 4, steps, 1e-4, 1.0, flags=flags, rebin_every=7)
    assert status == [0] * 4 and sum(counts) == len(p)
    own0 = parallel.owner_of(p[:, 0], 80, slabs)
    own1 = parallel.owner_of(got[:, 0], 80, slabs)
    assert (own0 != own1).sum() > 1000, "the scene must migrate particles"
    assert np.array_equal(bits(got[:, -1]), bits(p[:, -1]))
    assert np.isfinite(got).all()
    b0, b1 = scenes.bulk(want, 2), scenes.bulk(got, 2)
    assert np.abs(b0["com"] - b1["com"]).max() <= 1e-3 * np.abs(b0["com"]).max()
    assert abs(b0["ke"] - b1["ke"]) <= 1e-3 * b0["ke"]
    assert rel_l2(got[:, 0:2], want[:, 0:2]) <= 1e-3


@pytest.mark.parametrize("world", [2, 3])
def test_overlapped_schedule_wide_slabs(oracle, world):
    # MPM_FLAG_OVERLAP: interior bins run on a side stream while the boundary bins' emigrants and shared columns
    # are exchanged.  Needs slabs wider than 4 bin columns to have an interior: 512^2 grid, ~1M particles.
    from mpm_flip98a_b200.engine import FLAG_OVERLAP
    n = 512
    dt, vol = scenes.scaled_constants(n)
    p = scenes.three_blocks_2d(n, per_side=4)
    p[:, 2] = np.where(p[:, 1] > 0.4, 3.0, -3.0).astype(np.float32)  # shear: particles cross the cuts both ways
    with mpm.Engine(dim=2, n_grid=n, capacity=len(p), dt=dt, vol_p=vol) as e:  # warm state from one handle
        e.upload(p)
        e.substep(400)
        warm = e.read()
    P = make_params(dim=2, n_grid=n, vol_p=vol)
    want = warm.copy()
    oracle.advance(P, dt, want, 1)
    got, status, counts, slabs = run_slabs(warm, 2, n, world, 1, dt, vol, flags=FLAG_OVERLAP)
    assert status == [0] * world and sum(counts) == len(p)
    # the same on ONE shared stream without any host synchronisation: the interior kernels (side streams) overlap
    # the staging, the copies and the consumption of the messages
    got2, status2, counts2, _ = run_slabs(warm, 2, n, world, 3, dt, vol, flags=FLAG_OVERLAP, shared_stream=True)
    got3, status3, counts3, _ = run_slabs(warm, 2, n, world, 3, dt, vol, flags=0)
    assert status2 == [0] * world and sum(counts2) == len(p) and status3 == [0] * world
    for k, v in fields(got3, 2).items():
        assert rel_l2(fields(got2, 2)[k], v) <= 4e-5, (k, rel_l2(fields(got2, 2)[k], v))
    fw, fg = fields(want, 2), fields(got, 2)
    for k in fw:
        assert rel_l2(fg[k], fw[k]) <= 2e-5, (k, rel_l2(fg[k], fw[k]))  # C at 512^2: the reference's own reorder noise is 1e-5
    # many substeps with re-sorts and real migration: nothing lost, nothing flagged, same bulk as one handle
    steps = 60
    with mpm.Engine(dim=2, n_grid=n, capacity=len(p), dt=dt, vol_p=vol) as e:
        e.upload(warm)
        e.substep(steps)
        single = e.read()
    got, status, counts, slabs = run_slabs(warm, 2, n, world, steps, dt, vol, flags=FLAG_OVERLAP, rebin_every=16,
                                           shared_stream=True)
    assert status == [0] * world and sum(counts) == len(p)
    assert (parallel.owner_of(warm[:, 0], n, slabs) != parallel.owner_of(got[:, 0], n, slabs)).sum() > 100
    b0, b1 = scenes.bulk(single, 2), scenes.bulk(got, 2)
    assert np.abs(b0["com"] - b1["com"]).max() <= 1e-4 and abs(b0["ke"] - b1["ke"]) <= 1e-3 * b0["ke"]


@pytest.mark.parametrize("n_slabs", [1, 2, 3])
def test_group_handle_drives_slabs_from_the_c_abi(oracle, shipped, n_slabs):
    """mpm_group_*: several slabs behind ONE handle (the calls a C++ main() makes): upload, substep, read in upload
    order -- here all slabs on cuda:0 (a device may appear more than once); against the oracle and the migration
    bookkeeping (every particle accounted for, slab particle counts change)."""
    p = shipped["step1000"]
    want = p.copy()
    oracle.advance(make_params(), 1e-4, want, 1)
    with mpm.Group([0] * n_slabs, dim=2, n_grid=80, capacity=len(p)) as g:
        g.upload(p)
        before = g.slabs()
        assert sum(s[3] for s in before) == len(p) and before[0][1] == 0 and before[-1][2] == 80
        g.substep(1)
        got = g.read()
        assert g.poll_status() == 0
        fw, fg = fields(want, 2), fields(got, 2)
        for k in fw:
            assert rel_l2(fg[k], fw[k]) <= 1e-5, (k, rel_l2(fg[k], fw[k]))
        g.substep(150)
        later = g.read()
        assert g.poll_status() == 0
        after = g.slabs()
    assert sum(s[3] for s in after) == len(p)
    assert np.isfinite(later).all() and np.array_equal(bits(later[:, -1]), bits(p[:, -1]))
    if n_slabs > 1:
        assert [s[3] for s in after] != [s[3] for s in before], "the scene must migrate particles between the slabs"
    ref = p.copy()
    oracle.advance(make_params(), 1e-4, ref, 151)
    b0, b1 = scenes.bulk(ref, 2), scenes.bulk(later, 2)
    assert np.abs(b0["com"] - b1["com"]).max() <= 2e-3 * np.abs(b0["com"]).max()  # chaotic scene: reorder noise


@pytest.mark.parametrize("flags", [0, FLAG_NO_FUSE])
def test_run_continues_unsettled_and_changes_dt(oracle, flags):
    """The calling pattern of bench.py / a long run: step_local(..., settle=False) several times in a row (the second
    mpm_slab_begin finds staged-and-exchanged messages and must simply carry on), then a change of dt mid-run
    (mpm_slab_begin takes in the outstanding messages and redoes the P2G).  Against one handle doing the same."""
    p = scenes.jelly_drop()
    p[:, 2] = 6.0
    with mpm.Engine(dim=2, n_grid=80, capacity=len(p), flags=flags) as e:
        e.upload(p)
        e.substep(60, dt=1e-4)
        e.substep(40, dt=5e-5)
        single = e.read()
    ranks, ex, slabs = parallel.make_local_cluster(mpm.Engine, p, 2, 80, 3, flags=flags, rebin_every=9)
    parallel.step_local(ranks, ex, 25, dt=1e-4, settle=False)
    parallel.step_local(ranks, ex, 35, dt=1e-4, settle=False)
    parallel.step_local(ranks, ex, 40, dt=5e-5)
    got = parallel.collect_local(ranks, len(p), p.shape[1])
    assert [r.e.poll_status() for r in ranks] == [0, 0, 0] and sum(r.e.count for r in ranks) == len(p)
    for r in ranks:
        r.e.close()
    assert rel_l2(got[:, 0:2], single[:, 0:2]) <= 1e-4 and rel_l2(got[:, 2:4], single[:, 2:4]) <= 1e-3
    b0, b1 = scenes.bulk(single, 2), scenes.bulk(got, 2)
    assert abs(b0["ke"] - b1["ke"]) <= 1e-3 * b0["ke"]


@pytest.mark.parametrize("fuse", [False, True])
def test_overlapped_schedule_3d(oracle, fuse):
    """MPM_FLAG_OVERLAP in 3D, two slabs of 8 bin columns each (64^3 grid): the default two-kernel path (boundary G2P + P2G
    on the main stream, the interior's on the side stream) and the fused kernel (MPM_FLAG_FUSE_3D: its work list is
    split); one warm substep against the oracle, then 40 substeps with re-sorts against the plain schedule."""
    from mpm_flip98a_b200.engine import FLAG_OVERLAP
    if fuse:
        FLAG_OVERLAP |= FLAG_FUSE_3D
    n = 64
    dt, vol = scenes.scaled_constants(n, 3)
    p = scenes.collapse_3d(n, per_side=2, y_top=0.4, xz=(0.1, 0.9))
    p[:, 3] = np.where(p[:, 1] > 0.2, 2.0, -2.0).astype(np.float32)  # shear along x: particles cross the cut both ways
    with mpm.Engine(dim=3, n_grid=n, capacity=len(p), dt=dt, vol_p=vol) as e:
        e.upload(p)
        e.substep(60)
        warm = e.read()
    P = make_params(dim=3, n_grid=n, vol_p=vol)
    want = warm.copy()
    oracle.advance(P, dt, want, 1)
    got, status, counts, slabs = run_slabs(warm, 3, n, 2, 1, dt, vol, flags=FLAG_OVERLAP, shared_stream=True)
    assert status == [0, 0] and sum(counts) == len(p)
    fw, fg = fields(want, 3), fields(got, 3)
    for k in fw:
        assert rel_l2(fg[k], fw[k]) <= 1e-5, (k, rel_l2(fg[k], fw[k]))
    a, sa, ca, _ = run_slabs(warm, 3, n, 2, 40, dt, vol, flags=FLAG_OVERLAP, rebin_every=8, shared_stream=True)
    b, sb_, cb, _ = run_slabs(warm, 3, n, 2, 40, dt, vol, flags=0, rebin_every=8)  # the default two-kernel 3D path
    assert sa == [0, 0] and sb_ == [0, 0] and sum(ca) == len(p) and sum(cb) == len(p)
    assert (parallel.owner_of(warm[:, 0], n, slabs) != parallel.owner_of(a[:, 0], n, slabs)).sum() > 100
    assert rel_l2(a[:, 0:3], b[:, 0:3]) <= 1e-5 and rel_l2(a[:, 3:6], b[:, 3:6]) <= 1e-3


def test_overlapped_schedule_3d_times_each_launch_once():
    """Per-phase accounting of the overlapped two-kernel 3D schedule (mpm_profile, what bench.py's roofline reads): per
    substep the interior's G2P + P2G are the two launches under "g2p" (side stream), the boundary's five launches and the
    immigrant unpack sit under "migrate", and no enclosing span counts the boundary work a second time."""
    from mpm_flip98a_b200.engine import FLAG_OVERLAP
    n = 64
    dt, vol = scenes.scaled_constants(n, 3)
    p = scenes.collapse_3d(n, per_side=2, y_top=0.4, xz=(0.1, 0.9))
    ranks, ex, slabs = parallel.make_local_cluster(mpm.Engine, p, 3, n, 2, dt=dt, vol_p=vol, flags=FLAG_OVERLAP,
                                                   rebin_every=64, shared_stream=True)
    parallel.step_local(ranks, ex, 3, settle=False)
    for r in ranks:
        r.e.profile_enable(True)
    parallel.step_local(ranks, ex, 6, settle=False)
    profs = [r.e.profile() for r in ranks]
    status = [r.e.poll_status() for r in ranks]
    for r in ranks:
        r.e.close()
    assert status == [0, 0]
    for pr in profs:
        assert pr["substeps"] == 6
        assert pr["g2p"][1] == 2 * 6, pr      # interior G2P + interior P2G, nothing else
        assert pr["p2g"][1] == 0, pr          # the next P2G was launched by the overlapped schedule
        assert pr["migrate"][1] == (5 + 3) * 6, pr
        assert pr["g2p"][0] > 0 and pr["migrate"][0] > 0


def test_overlap_guard_flags_a_too_long_resort_interval():
    """The interior launch of the overlapped schedule skips the migration code; what licenses that -- no particle can
    have travelled from an interior bin to the slab cut since the last re-sort -- is checked on the device every substep
    from the measured displacements.  A fixed, far too long re-sort interval with fast particles must raise MPM_E_CFL."""
    from mpm_flip98a_b200.engine import FLAG_OVERLAP
    n = 512
    dt, vol = scenes.scaled_constants(n)
    p = scenes.three_blocks_2d(n, per_side=2)
    p[:, 2] = 30.0  # 0.24 cells per substep: the guard trips once (substeps since the re-sort + 2) x 0.24 > 8
    ranks, ex, slabs = parallel.make_local_cluster(mpm.Engine, p, 2, n, 2, dt=dt, vol_p=vol, flags=FLAG_OVERLAP,
                                                   rebin_every=400, shared_stream=True)
    parallel.step_local(ranks, ex, 45, settle=False)
    status = [r.e.poll_status() for r in ranks]
    for r in ranks:
        r.e.close()
    assert -5 in status, status  # MPM_E_CFL
    # ... and the adaptive interval keeps the same run clean
    ranks, ex, slabs = parallel.make_local_cluster(mpm.Engine, p, 2, n, 2, dt=dt, vol_p=vol, flags=FLAG_OVERLAP,
                                                   shared_stream=True)
    parallel.step_local(ranks, ex, 45)
    status = [r.e.poll_status() for r in ranks]
    n_live = sum(r.e.count for r in ranks)
    for r in ranks:
        r.e.close()
    assert status == [0, 0] and n_live == len(p), (status, n_live)
