"""x-slab domain decomposition over several B200s.

The grid is cut into slabs of whole base-cell columns; rank r owns columns [lo_r, hi_r) and every
particle whose base cell (cpp_validation/mls-mpm88-explained.cpp:55) lies in them.  Per substep ONE
fixed-size message goes to each x-neighbour (include/mpm.h, mpm_slab_*): the partial sums of the two node
columns the slabs share (contiguous memory because x is the major grid index, :47), the emigrant count and
the emigrant records.  The receiving handle adds the sums (shared columns are then computed redundantly and
bit-identically on both sides), appends the immigrants and adds their share of the next P2G.  Counts and
extents stay on the device, so nothing here ever waits for the GPU:

    [slab_begin -> exchange]   then per substep:   slab_step -> exchange     ...   slab_settle

The engine (libmpm.so, one handle per GPU) only exposes the message buffers; the bytes are moved here:
`DistExchange` = torch.distributed P2P (NCCL over NVLink / NVSwitch, or gloo on CPU tensors in the tests),
one grouped send/recv per substep; `LocalExchange` = device-to-device copies between handles that live in
one process (single-GPU emulation of N slabs, used by the GPU parity tests).  No collective is needed: the
pattern is nearest-neighbour only.
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
        d = engine.slab()
        self.bytes = d.bytes
        self.send_lo = dev_tensor(d.send_lo, d.bytes, device)
        self.send_hi = dev_tensor(d.send_hi, d.bytes, device)
        self.recv_lo = dev_tensor(d.recv_lo, d.bytes, device)
        self.recv_hi = dev_tensor(d.recv_hi, d.bytes, device)


class LocalExchange:
    """All slabs live in this process (any mix of devices): plain device-to-device copies.
    `stream`: a torch stream that every handle ALSO launches on (mpm_config.stream): the copies are then ordered on
    it and nothing synchronises -- with MPM_FLAG_OVERLAP the interior kernels (side streams) really run concurrently
    with the staging, the copies and the next substep's message consumption."""

    def __init__(self, ranks, stream=None):
        self.ranks, self.stream = ranks, stream

    def _copy(self):
        for a, b in zip(self.ranks[:-1], self.ranks[1:]):  # a = lower slab, b = upper slab
            b.recv_lo.copy_(a.send_hi, non_blocking=True)
            a.recv_hi.copy_(b.send_lo, non_blocking=True)

    def exchange(self):
        import torch
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                self._copy()
            return
        for r in self.ranks:
            r.e.synchronize()  # the messages were staged on each handle's own stream
        self._copy()
        for d in {r.device for r in self.ranks}:
            if d != "cpu":
                torch.cuda.synchronize(d)


class DistExchange:
    """One slab per process: torch.distributed point-to-point with the two x-neighbours, one grouped
    send/recv per substep, fixed sizes, no host synchronisation when the engine shares torch's stream."""

    def __init__(self, rank_obj, group=None, shared_stream=False):
        """shared_stream: the engine launches on torch's CURRENT stream (mpm_config.stream), so the
        P2P ops are stream-ordered against its kernels and no host synchronisation is needed around them."""
        import torch.distributed as dist
        self.r, self.dist, self.group, self.shared = rank_obj, dist, group, shared_stream

    def exchange(self):
        r, d = self.r, self.dist
        if not self.shared:
            r.e.synchronize()  # the messages were staged on the engine's own stream
        ops = []
        if r.has_hi:
            ops += [d.P2POp(d.isend, r.send_hi, r.rank + 1, self.group),
                    d.P2POp(d.irecv, r.recv_hi, r.rank + 1, self.group)]
        if r.has_lo:
            ops += [d.P2POp(d.isend, r.send_lo, r.rank - 1, self.group),
                    d.P2POp(d.irecv, r.recv_lo, r.rank - 1, self.group)]
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()  # NCCL: orders the current stream behind the transfer, does not block the host
        if not self.shared and r.device != "cpu":
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()


def step_local(ranks, exchange, n_steps=1, dt=0.0, settle=True):
    """n_steps substeps of all slabs held by this process (LocalExchange)."""
    staged = [r.e.slab_begin(dt) for r in ranks]
    assert all(staged) or not any(staged), "slab handles out of step"
    if any(staged):
        exchange.exchange()
    for _ in range(n_steps):
        for r in ranks:
            r.e.slab_step(dt)
        exchange.exchange()
    if settle:
        for r in ranks:
            r.e.slab_settle()


def step_dist(rank_obj, exchange, n_steps=1, dt=0.0, settle=True):
    """n_steps substeps of this process's slab (DistExchange)."""
    r = rank_obj
    if r.e.slab_begin(dt):
        exchange.exchange()
    for _ in range(n_steps):
        r.e.slab_step(dt)
        exchange.exchange()
    if settle:
        r.e.slab_settle()


def make_local_cluster(engine_cls, p, dim, n_grid, world, devices=None, margin=1.5, shared_stream=False, **engine_kw):
    """N slab handles in this process + their particles uploaded; -> (ranks, exchange, slabs).
    shared_stream (single device): all handles launch on one torch stream and the exchange never synchronises."""
    slabs = partition(n_grid, world, align=engine_kw.get("bin_edge") or (8 if dim == 2 else 4))
    parts = scatter_particles(p, n_grid, slabs)
    ranks = []
    stream = None
    if shared_stream:
        import torch
        assert not devices or len(set(devices)) == 1
        stream = torch.cuda.Stream(device=devices[0] if devices else 0)
        engine_kw = dict(engine_kw, stream=stream.cuda_stream)
    for r, ((lo, hi), (rec, ids)) in enumerate(zip(slabs, parts)):
        dev = devices[r] if devices else 0
        cap = int(max(len(rec) * margin, len(rec) + 8192))
        e = engine_cls(dim=dim, n_grid=n_grid, capacity=cap, slab=(lo, hi), device=dev, **engine_kw)
        e.upload_ids(rec, ids)
        ranks.append(SlabRank(e, r, world, "cuda:%d" % dev))
    return ranks, LocalExchange(ranks, stream), slabs


def collect_local(ranks, n_total, words):
    return gather_particles([r.e.read_ids() for r in ranks], n_total, words)
