"""x-slab domain decomposition over several B200s.

The grid is cut into slabs of whole base-cell columns; rank r owns columns [lo_r, hi_r) and every
particle whose base cell (cpp_validation/mls-mpm88-explained.cpp:55) lies in them.  One substep:

    P2G (owned particles)  ->  ghost-column SUM with both x-neighbours (2 node columns each way,
    contiguous memory because x is the major grid index, :47)  ->  grid update (shared columns are
    computed redundantly, bit-identically)  ->  G2P  ->  particle MIGRATION to the neighbours.

The engine (libmpm.so, one handle per GPU) only exposes device buffers and the phase calls; the bytes
are moved here: `DistExchange` = torch.distributed P2P (NCCL over NVLink / NVSwitch, or gloo on CPU
tensors in the tests), `LocalExchange` = device-to-device copies between handles that live in one
process (single-GPU emulation of N slabs, used by the GPU parity tests).  No collective is needed:
the pattern is nearest-neighbour only.
"""
import numpy as np


def partition(n_grid, world, align=8):
    """Equal-width slabs of base-cell columns, boundaries aligned to the bin edge."""
    if world == 1:
        return [(0, n_grid)]
    units = n_grid // align
    assert units >= world, "grid too small for %d slabs" % world
    cuts = [align * ((units * r) // world) for r in range(world)] + [n_grid]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def partition_filled(n_grid, world, align=8, x_range=(0.05, 0.95)):
    """Slabs that split the FILLED x-range evenly (equal particle counts for a uniform fill); the first
    and last slab additionally own the empty margins.  Same conventions as partition()."""
    if world == 1:
        return [(0, n_grid)]
    lo, hi = x_range[0] * n_grid, x_range[1] * n_grid
    cuts = [0] + [int(round((lo + (hi - lo) * r / world) / align)) * align for r in range(1, world)] + [n_grid]
    assert all(b - a >= align for a, b in zip(cuts[:-1], cuts[1:])), "grid too small for %d slabs" % world
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def base_column(x, n_grid):
    """Base cell x-index exactly as the engine computes it (fp32 multiply, subtract, truncate)."""
    inv_dx = np.float32(1.0) / (np.float32(1.0) / np.float32(n_grid))
    b = (x.astype(np.float32) * inv_dx - np.float32(0.5)).astype(np.int32)
    return np.clip(b, 0, n_grid - 2)


def owner_of(x, n_grid, slabs):
    b = base_column(x, n_grid)
    his = np.array([hi for _, hi in slabs])
    return np.searchsorted(his, b, side="right").astype(np.int32)


def scatter_particles(p, n_grid, slabs):
    """Split a global particle set into per-slab (records, global ids)."""
    own = owner_of(p[:, 0], n_grid, slabs)
    out = []
    for r in range(len(slabs)):
        idx = np.nonzero(own == r)[0].astype(np.int32)
        out.append((np.ascontiguousarray(p[idx]), idx))
    return out


def gather_particles(parts, n_total, words):
    """Inverse of scatter_particles after a run: records back in global id order."""
    out = np.zeros((n_total, words), np.float32)
    seen = np.zeros(n_total, np.int32)
    for rec, ids in parts:
        out[ids] = rec
        seen[ids] += 1
    assert (seen == 1).all(), "particles lost or duplicated by migration: %s" % np.bincount(seen)
    return out


class _DevView:
    """A device pointer as a CUDA-array-interface object torch can wrap without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def dev_tensor(ptr, nbytes, device):
    """uint8 torch tensor aliasing `nbytes` at `ptr` (CUDA device memory, or host memory for "cpu")."""
    import torch
    if not ptr or nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if device == "cpu":
        import ctypes
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * int(nbytes)).from_address(int(ptr))))
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


class SlabRank:
    """One engine handle + its place in the slab chain."""

    def __init__(self, engine, rank, world, device):
        self.e, self.rank, self.world, self.device = engine, rank, world, device
        self.has_lo, self.has_hi = rank > 0, rank < world - 1
        h = engine.halo()
        nb = h.bytes
        self.halo_send_lo = dev_tensor(h.send_lo, nb, device)
        self.halo_send_hi = dev_tensor(h.send_hi, nb, device)
        self.halo_recv_lo = dev_tensor(h.recv_lo, nb, device)
        self.halo_recv_hi = dev_tensor(h.recv_hi, nb, device)
        m = engine.migration()
        self.rec_bytes = m.record_bytes
        cap = m.recv_capacity * m.record_bytes
        self.mig_send_lo = dev_tensor(m.send_lo, cap, device)
        self.mig_send_hi = dev_tensor(m.send_hi, cap, device)
        self.mig_recv_lo = dev_tensor(m.recv_lo, cap, device)
        self.mig_recv_hi = dev_tensor(m.recv_hi, cap, device)


class LocalExchange:
    """All slabs live in this process (any mix of devices): plain device-to-device copies."""

    def __init__(self, ranks):
        self.ranks = ranks

    def halo(self):
        for a, b in zip(self.ranks[:-1], self.ranks[1:]):  # a = lower slab, b = upper slab
            a.e.synchronize()
            b.e.synchronize()
            b.halo_recv_lo.copy_(a.halo_send_hi)
            a.halo_recv_hi.copy_(b.halo_send_lo)
        self._sync()

    def migrate(self):
        descs = [r.e.migration() for r in self.ranks]  # synchronises each handle
        n_in = [[0, 0] for _ in self.ranks]
        for k, (a, b) in enumerate(zip(self.ranks[:-1], self.ranks[1:])):
            up, down = descs[k].n_send_hi, descs[k + 1].n_send_lo
            if up:
                b.mig_recv_lo[:up * a.rec_bytes].copy_(a.mig_send_hi[:up * a.rec_bytes])
            if down:
                a.mig_recv_hi[:down * a.rec_bytes].copy_(b.mig_send_lo[:down * a.rec_bytes])
            n_in[k + 1][0] = up
            n_in[k][1] = down
        self._sync()
        return n_in

    def _sync(self):
        import torch
        for d in {r.device for r in self.ranks}:
            torch.cuda.synchronize(d)


class DistExchange:
    """One slab per process: torch.distributed point-to-point with the two x-neighbours."""

    def __init__(self, rank_obj, group=None, shared_stream=False):
        """shared_stream: the engine launches on torch's CURRENT stream (mpm_config.stream), so the
        P2P ops are stream-ordered against its kernels and no host synchronisation is needed around them."""
        import torch.distributed as dist
        self.r, self.dist, self.group, self.shared = rank_obj, dist, group, shared_stream

    def _run(self, ops):
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def halo(self):
        r, d = self.r, self.dist
        if not self.shared:
            r.e.synchronize()  # P2G wrote the send columns on the engine's own stream
        ops = []
        if r.has_hi:
            ops += [d.P2POp(d.isend, r.halo_send_hi, r.rank + 1, self.group),
                    d.P2POp(d.irecv, r.halo_recv_hi, r.rank + 1, self.group)]
        if r.has_lo:
            ops += [d.P2POp(d.isend, r.halo_send_lo, r.rank - 1, self.group),
                    d.P2POp(d.irecv, r.halo_recv_lo, r.rank - 1, self.group)]
        self._run(ops)
        self._sync()

    def migrate(self):
        import torch
        r, d = self.r, self.dist
        m = r.e.migration()  # synchronises; counts of the emigrants packed by G2P
        dev = r.device
        cnt_out = torch.tensor([m.n_send_lo, m.n_send_hi], dtype=torch.int64, device=dev)
        cnt_in = torch.zeros(2, dtype=torch.int64, device=dev)
        ops = []
        if r.has_lo:
            ops += [d.P2POp(d.isend, cnt_out[0:1], r.rank - 1, self.group),
                    d.P2POp(d.irecv, cnt_in[0:1], r.rank - 1, self.group)]
        if r.has_hi:
            ops += [d.P2POp(d.isend, cnt_out[1:2], r.rank + 1, self.group),
                    d.P2POp(d.irecv, cnt_in[1:2], r.rank + 1, self.group)]
        self._run(ops)
        n_lo, n_hi = (int(v) for v in cnt_in.tolist())
        ops = []
        rb = r.rec_bytes
        if r.has_lo:
            if m.n_send_lo:
                ops.append(d.P2POp(d.isend, r.mig_send_lo[:m.n_send_lo * rb], r.rank - 1, self.group))
            if n_lo:
                ops.append(d.P2POp(d.irecv, r.mig_recv_lo[:n_lo * rb], r.rank - 1, self.group))
        if r.has_hi:
            if m.n_send_hi:
                ops.append(d.P2POp(d.isend, r.mig_send_hi[:m.n_send_hi * rb], r.rank + 1, self.group))
            if n_hi:
                ops.append(d.P2POp(d.irecv, r.mig_recv_hi[:n_hi * rb], r.rank + 1, self.group))
        self._run(ops)
        self._sync()
        return n_lo, n_hi

    def _sync(self):
        import torch
        if not self.shared and self.r.device != "cpu" and torch.cuda.is_available():
            torch.cuda.synchronize()


def step_local(ranks, exchange, n_steps=1, dt=0.0):
    """n_steps substeps of all slabs held by this process (LocalExchange)."""
    for _ in range(n_steps):
        for r in ranks:
            r.e.step_p2g(dt)
        exchange.halo()
        for r in ranks:
            r.e.step_halo_add(r.has_lo, r.has_hi)
            r.e.step_grid_g2p(dt)
        n_in = exchange.migrate()
        for r, (n_lo, n_hi) in zip(ranks, n_in):
            r.e.step_immigrate(n_lo, n_hi)


def step_dist(rank_obj, exchange, n_steps=1, dt=0.0):
    """n_steps substeps of this process's slab (DistExchange)."""
    r = rank_obj
    for _ in range(n_steps):
        r.e.step_p2g(dt)
        exchange.halo()
        r.e.step_halo_add(r.has_lo, r.has_hi)
        r.e.step_grid_g2p(dt)
        n_lo, n_hi = exchange.migrate()
        r.e.step_immigrate(n_lo, n_hi)


def make_local_cluster(engine_cls, p, dim, n_grid, world, devices=None, margin=1.5, **engine_kw):
    """N slab handles in this process + their particles uploaded; -> (ranks, exchange, slabs)."""
    slabs = partition(n_grid, world, align=engine_kw.get("bin_edge") or (8 if dim == 2 else 4))
    parts = scatter_particles(p, n_grid, slabs)
    ranks = []
    for r, ((lo, hi), (rec, ids)) in enumerate(zip(slabs, parts)):
        dev = devices[r] if devices else 0
        cap = int(max(len(rec) * margin, len(rec) + 8192))
        e = engine_cls(dim=dim, n_grid=n_grid, capacity=cap, slab=(lo, hi), device=dev, **engine_kw)
        e.upload_ids(rec, ids)
        ranks.append(SlabRank(e, r, world, "cuda:%d" % dev))
    return ranks, LocalExchange(ranks), slabs


def collect_local(ranks, n_total, words):
    return gather_particles([r.e.read_ids() for r in ranks], n_total, words)
