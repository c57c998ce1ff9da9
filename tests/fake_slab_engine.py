"""A numpy stand-in for one engine handle (TEST ONLY) with the same x-slab protocol calls and message layout as
mpm_flip98a_b200.Engine (include/mpm.h: mpm_slab_describe / begin / step / settle), so the exchange protocol of
parallel.py (partition, ownership, ONE fixed-size message per neighbour and substep carrying ghost-column sums +
emigrant count + records, the split P2G share of migrating particles, id bookkeeping) can run on CPU tensors over
gloo.  Its "physics" is deliberately trivial and exactly reproducible: P2G deposits integer-valued weights on the
3x3 stencil, G2P moves every particle by a fixed integer-derived step."""
import numpy as np

from mpm_flip98a_b200.engine import SlabDesc
from mpm_flip98a_b200.parallel import base_column

REC = 16  # floats per migration record (2D): 14 AoS words + id + pad
FRESH, STAGED, SETTLED = 0, 1, 2


class FakeEngine:
    def __init__(self, n_grid, slab, cap=4096):
        self.n, (self.lo, self.hi) = n_grid, slab
        self.n1 = n_grid + 1
        xhi = min(self.hi, n_grid - 1)
        self.ncol = xhi - self.lo + 2
        self.has_lo, self.has_hi = self.lo > 0, self.hi < n_grid
        self.grid = np.zeros((self.ncol, self.n1, 4), np.float32)
        self.K = cap
        self.halo_floats = 2 * self.n1 * 4
        self.msg_floats = self.halo_floats + 4 + cap * REC
        self.send = [np.zeros(self.msg_floats, np.float32) for _ in range(2)]
        self.recv = [np.zeros(self.msg_floats, np.float32) for _ in range(2)]
        self.p = np.zeros((0, 14), np.float32)
        self.ids = np.zeros(0, np.int32)
        self.state = FRESH
        self.complete_grids = []  # the whole (ghost-summed) P2G grid of every substep, as the grid update sees it

    # ---- message views ----
    def _halo(self, m):
        return m[:self.halo_floats].reshape(2, self.n1, 4)

    def _count(self, m):
        return m[self.halo_floats:self.halo_floats + 4].view(np.int32)

    def _recs(self, m):
        return m[self.halo_floats + 4:].reshape(self.K, REC)

    def upload_ids(self, p, ids):
        self.p, self.ids = p.copy(), ids.copy()
        self.state = FRESH

    def read_ids(self):
        return self.p.copy(), self.ids.copy()

    def synchronize(self):
        pass

    def slab(self):
        d = SlabDesc()
        d.send_lo, d.send_hi = self.send[0].ctypes.data, self.send[1].ctypes.data
        d.recv_lo, d.recv_hi = self.recv[0].ctypes.data, self.recv[1].ctypes.data
        d.bytes, d.halo_bytes = self.msg_floats * 4, self.halo_floats * 4
        d.record_bytes, d.record_capacity = REC * 4, self.K
        d.has_lo, d.has_hi = int(self.has_lo), int(self.has_hi)
        return d

    # ---- the stand-in physics ----
    def _scatter(self, grid, x, col_lo, col_hi):
        """P2G of positions x into `grid`, restricted to the global node columns [col_lo, col_hi)"""
        bx, by = base_column(x[:, 0], self.n), base_column(x[:, 1], self.n)
        for a in range(3):
            col = bx + a
            ok = (col >= col_lo) & (col < col_hi)
            for b in range(3):
                np.add.at(grid, (col[ok] - self.lo, by[ok] + b, 2), float((a + 1) * (b + 1)))

    def _stage(self, emigrants):
        for k, have in ((0, self.has_lo), (1, self.has_hi)):
            if not have:
                continue
            rec, ids = emigrants[k]
            assert len(rec) <= self.K
            # the emigrants' share of the next P2G on the node columns this handle holds
            self._scatter(self.grid, rec, self.lo, self.lo + self.ncol)
        for k, have in ((0, self.has_lo), (1, self.has_hi)):
            if not have:
                continue
            rec, ids = emigrants[k]
            self._halo(self.send[k])[:] = self.grid[:2] if k == 0 else self.grid[self.ncol - 2:]
            self._count(self.send[k])[0] = len(rec)
            self._recs(self.send[k])[:len(rec), :14] = rec
            self._recs(self.send[k])[:len(rec), 14] = ids.view(np.float32)
        self.state = STAGED

    def _consume(self):
        top = self.lo + self.ncol
        for k, have in ((0, self.has_lo), (1, self.has_hi)):
            if not have:
                continue
            if k == 0:
                self.grid[:2] += self._halo(self.recv[0])
            else:
                self.grid[self.ncol - 2:] += self._halo(self.recv[1])
        for k, have in ((0, self.has_lo), (1, self.has_hi)):
            if not have:
                continue
            cnt = int(self._count(self.recv[k])[0])
            rec = self._recs(self.recv[k])[:cnt]
            self.p = np.concatenate([self.p, rec[:, :14]])
            self.ids = np.concatenate([self.ids, rec[:, 14].view(np.int32)])
            # arrivals from below were scattered into the shared columns lo, lo+1 by their sender; from above hi, hi+1
            lo, hi = (self.lo + 2, top) if k == 0 else (self.lo, top - 2)
            self._scatter(self.grid, rec[:, :14], lo, hi)

    def slab_begin(self, dt=0.0):
        if self.state != FRESH:  # the run continues (the stand-in has no dt): nothing to recompute or exchange
            return False
        self.grid[:] = 0
        bx = base_column(self.p[:, 0], self.n)
        assert ((bx >= self.lo) & (bx < self.hi)).all(), "particle outside its slab"
        self._scatter(self.grid, self.p, self.lo, self.lo + self.ncol)
        empty = (np.zeros((0, 14), np.float32), np.zeros(0, np.int32))
        self._stage([empty, empty])
        if not (self.has_lo or self.has_hi):
            self.state = SETTLED
            return False
        return True

    def slab_step(self, dt=0.0):
        assert self.state != FRESH
        if self.state == STAGED:
            self._consume()
        self.complete_grids.append(self.grid.copy())
        # "G2P": a shear flow that crosses slab boundaries both ways
        self.p[:, 0] += np.where(self.p[:, 1] > 0.5, 1.0, -1.0).astype(np.float32) * np.float32(0.37 / self.n)
        self.p[:, 0] = np.clip(self.p[:, 0], 0.02, 0.98)
        bx = base_column(self.p[:, 0], self.n)
        side = np.where(bx < self.lo, 0, np.where(bx >= self.hi, 1, -1))
        emigrants = [(self.p[side == k].copy(), self.ids[side == k].copy()) for k in (0, 1)]
        keep = side < 0
        self.p, self.ids = self.p[keep], self.ids[keep]
        # next P2G of the residents
        self.grid[:] = 0
        self._scatter(self.grid, self.p, self.lo, self.lo + self.ncol)
        self._stage(emigrants)
        if not (self.has_lo or self.has_hi):
            self.state = SETTLED

    def slab_settle(self):
        if self.state == STAGED:
            self._consume()
            self.state = SETTLED
