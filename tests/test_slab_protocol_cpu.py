"""N > 1 host logic on CPU (no GPU): slab partition / ownership, and the exchange protocol of
mpm_flip98a_b200.parallel run for real over torch.distributed `gloo`, world_size 2, with a numpy
stand-in for the engine (tests/fake_slab_engine.py).  Checked against the same stand-in on one slab:
identical ghost-summed grids in every substep (although a migrating particle's P2G share is split between sender
and receiver), no particle lost or duplicated, ids preserved."""
import os
import subprocess
import sys

import numpy as np

from mpm_flip98a_b200 import parallel, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_and_ownership():
    for n, w in ((80, 2), (80, 4), (512, 8), (8192, 8), (256, 8)):
        align = 8
        slabs = parallel.partition(n, w, align)
        assert slabs[0][0] == 0 and slabs[-1][1] == n and len(slabs) == w
        for (a, b), (c, d) in zip(slabs[:-1], slabs[1:]):
            assert b == c and a % align == 0 and b - a >= align
    for n, w in ((8192, 8), (23168, 8), (512, 4)):
        slabs = parallel.partition_filled(n, w)
        assert slabs[0][0] == 0 and slabs[-1][1] == n and all(a % 8 == 0 for a, _ in slabs)
        filled = [min(b, 0.95 * n) - max(a, 0.05 * n) for a, b in slabs]
        assert max(filled) - min(filled) <= 16  # equal shares of the filled range, up to alignment
    p = scenes.commented_three_blocks()
    slabs = parallel.partition(80, 4)
    own = parallel.owner_of(p[:, 0], 80, slabs)
    b = parallel.base_column(p[:, 0], 80)
    for r, (lo, hi) in enumerate(slabs):
        assert ((b[own == r] >= lo) & (b[own == r] < hi)).all()
    parts = parallel.scatter_particles(p, 80, slabs)
    assert sum(len(i) for _, i in parts) == len(p)
    back = parallel.gather_particles(parts, len(p), 14)
    assert np.array_equal(back, p)


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mpm_flip98a_b200 import parallel
from tests.fake_slab_engine import FakeEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n = 80
p = np.load(sys.argv[2])
slabs = parallel.partition(n, world)
rec, ids = parallel.scatter_particles(p, n, slabs)[rank]
e = FakeEngine(n, slabs[rank])
e.upload_ids(rec, ids)
r = parallel.SlabRank(e, rank, world, "cpu")
ex = parallel.DistExchange(r)
parallel.step_dist(r, ex, 10, settle=False)
parallel.step_dist(r, ex, 15, settle=False)  # continues an UNSETTLED run (messages staged and exchanged): no begin
parallel.step_dist(r, ex, 15)                # ... and settles at the end
grids = e.complete_grids
rec, ids = e.read_ids()
np.savez(sys.argv[3] + ".%d.npz" % rank, rec=rec, ids=ids, grid=np.stack(grids), lo=slabs[rank][0])
dist.destroy_process_group()
'''


def test_two_rank_gloo_exchange(tmp_path):
    from tests.fake_slab_engine import FakeEngine
    p = scenes.commented_three_blocks()
    p[:, 0] = np.random.RandomState(0).uniform(0.3, 0.7, len(p)).astype(np.float32)  # straddle the cut at 40
    np.save(tmp_path / "p.npy", p)
    (tmp_path / "worker.py").write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(tmp_path / "worker.py"), ROOT, str(tmp_path / "p.npy"),
                               str(tmp_path / "out")], env=dict(env, RANK=str(r))) for r in range(2)]
    for pr in procs:
        assert pr.wait(timeout=120) == 0
    outs = [np.load(str(tmp_path / "out") + ".%d.npz" % r) for r in range(2)]
    # single-slab run of the same stand-in
    e = FakeEngine(80, (0, 80))
    e.upload_ids(p, np.arange(len(p), dtype=np.int32))
    assert not e.slab_begin()
    for s_ in range(40):
        e.slab_step()
    e.slab_settle()
    grids = e.complete_grids
    grids = np.stack(grids)
    moved = 0
    for o in outs:
        lo, g = int(o["lo"]), o["grid"]
        assert np.array_equal(g, grids[:, lo:lo + g.shape[1]])  # ghost-summed columns == global grid
        moved += int((parallel.owner_of(p[o["ids"], 0], 80, parallel.partition(80, 2)) != (0 if lo == 0 else 1)).sum())
    got = parallel.gather_particles([(o["rec"], o["ids"]) for o in outs], len(p), 14)
    assert np.array_equal(got, e.p[np.argsort(e.ids)])
    assert moved > 50, "the test scene must actually migrate particles (moved %d)" % moved
