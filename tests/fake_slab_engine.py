"""A numpy stand-in for one engine handle (TEST ONLY) with the same phase calls and buffer
descriptors as mpm_flip98a_b200.Engine, so the x-slab exchange protocol (parallel.py: partition,
ownership, ghost-column sums, migration with counts + payload, id bookkeeping) can run on CPU
tensors over gloo.  Its "physics" is deliberately trivial and exactly reproducible: P2G deposits
integer-valued weights on the 3x3 stencil, G2P moves every particle by a fixed integer-derived step."""
import ctypes

import numpy as np

from mpm_flip98a_b200.engine import HaloDesc, MigrationDesc
from mpm_flip98a_b200.parallel import base_column

REC = 16  # floats per migration record (2D): 14 AoS words + id + pad


class FakeEngine:
    def __init__(self, n_grid, slab, cap=4096):
        self.n, (self.lo, self.hi) = n_grid, slab
        self.n1 = n_grid + 1
        xhi = min(self.hi, n_grid - 1)
        self.ncol = xhi - self.lo + 2
        self.grid = np.zeros((self.ncol, self.n1, 4), np.float32)
        self.recv_lo = np.zeros((2, self.n1, 4), np.float32)
        self.recv_hi = np.zeros((2, self.n1, 4), np.float32)
        self.cap = cap
        self.send = [np.zeros((cap, REC), np.float32), np.zeros((cap, REC), np.float32)]
        self.mrecv = [np.zeros((cap, REC), np.float32), np.zeros((cap, REC), np.float32)]
        self.nsend = [0, 0]
        self.p = np.zeros((0, 14), np.float32)
        self.ids = np.zeros(0, np.int32)

    def upload_ids(self, p, ids):
        self.p, self.ids = p.copy(), ids.copy()

    def read_ids(self):
        return self.p.copy(), self.ids.copy()

    def synchronize(self):
        pass

    def halo(self):
        d = HaloDesc()
        d.send_lo = self.grid[:2].ctypes.data
        d.send_hi = self.grid[self.ncol - 2:].ctypes.data
        d.recv_lo = self.recv_lo.ctypes.data
        d.recv_hi = self.recv_hi.ctypes.data
        d.bytes = self.recv_lo.nbytes
        return d

    def migration(self):
        d = MigrationDesc()
        d.send_lo, d.send_hi = self.send[0].ctypes.data, self.send[1].ctypes.data
        d.recv_lo, d.recv_hi = self.mrecv[0].ctypes.data, self.mrecv[1].ctypes.data
        d.n_send_lo, d.n_send_hi = self.nsend
        d.recv_capacity, d.record_bytes = self.cap, REC * 4
        return d

    def step_p2g(self, dt=0.0):
        self.grid[:] = 0
        bx = base_column(self.p[:, 0], self.n)
        by = base_column(self.p[:, 1], self.n)
        assert ((bx >= self.lo) & (bx < self.hi)).all(), "particle outside its slab"
        for a in range(3):
            for b in range(3):
                np.add.at(self.grid, (bx - self.lo + a, by + b, 2), float((a + 1) * (b + 1)))

    def step_halo_add(self, have_lo, have_hi):
        if have_lo:
            self.grid[:2] += self.recv_lo
        if have_hi:
            self.grid[self.ncol - 2:] += self.recv_hi

    def step_grid_g2p(self, dt=0.0):
        # a shear flow that crosses slab boundaries both ways
        self.p[:, 0] += np.where(self.p[:, 1] > 0.5, 1.0, -1.0).astype(np.float32) * np.float32(0.37 / self.n)
        self.p[:, 0] = np.clip(self.p[:, 0], 0.02, 0.98)
        bx = base_column(self.p[:, 0], self.n)
        side = np.where(bx < self.lo, 0, np.where(bx >= self.hi, 1, -1))
        for k in (0, 1):
            sel = np.nonzero(side == k)[0]
            self.nsend[k] = len(sel)
            self.send[k][:len(sel), :14] = self.p[sel]
            self.send[k][:len(sel), 14] = self.ids[sel].view(np.float32)
        keep = side < 0
        self.p, self.ids = self.p[keep], self.ids[keep]

    def step_immigrate(self, n_lo, n_hi):
        for k, n in ((0, n_lo), (1, n_hi)):
            if n:
                self.p = np.concatenate([self.p, self.mrecv[k][:n, :14]])
                self.ids = np.concatenate([self.ids, self.mrecv[k][:n, 14].view(np.int32)])
