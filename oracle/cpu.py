"""ctypes front-end of the CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``    : oracle/build/liboracle.so, the parametrised restatement (mpm_oracle.cpp) of
                  /root/reference/cpp_validation/mls-mpm88-explained.cpp:49-180.
* ``Reference`` : oracle/_ref/libmpmref.so, the UNMODIFIED reference translation unit driven
                  through ref_harness.cpp (shipped scene only: its constants are compile-time).

Nothing here is imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmpmref.so")

KIND_FLUID, KIND_JELLY, KIND_SNOW = 0, 1, 2


def build(verbose=False):
    """Compile the checkers (g++ only).  The reference library is rebuilt only when the
    reference tree is present; on the GPU box the prebuilt oracle/_ref/ travels as is."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-2000:], out.stderr[-2000:])
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


class _Material(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("E", ctypes.c_float), ("nu", ctypes.c_float),
                ("hardening", ctypes.c_float), ("sig_lo", ctypes.c_float), ("sig_hi", ctypes.c_float)]


class OracleParams(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int), ("n_grid", ctypes.c_int), ("mass_p", ctypes.c_float),
                ("vol_p", ctypes.c_float), ("gravity", ctypes.c_float * 3), ("boundary", ctypes.c_float),
                ("jp_min", ctypes.c_float), ("jp_max", ctypes.c_float), ("alpha", ctypes.c_float),
                ("n_materials", ctypes.c_int), ("mat", _Material * 4)]


# (kind, E, nu, hardening, sig_lo, sig_hi).  Entries 0-2: upstream mls-mpm88 constants
# (E=1e4, nu=0.2, snow hardening 10, jelly factor 0.3); entry 3: the scene as shipped
# (mls-mpm88-explained.cpp:19-21: E=1e2, nu=0.499, hardening=1).
DEFAULT_MATERIALS = [
    (KIND_FLUID, 1e4, 0.2, 0.0, 1.0 - 2.5e-2, 1.0 + 7.5e-3),
    (KIND_JELLY, 1e4, 0.2, 0.3, 1.0 - 2.5e-2, 1.0 + 7.5e-3),
    (KIND_SNOW, 1e4, 0.2, 10.0, 1.0 - 2.5e-2, 1.0 + 7.5e-3),
    (KIND_SNOW, 1e2, 0.499, 1.0, 1.0 - 2.5e-2, 1.0 + 7.5e-3),
]


def make_params(dim=2, n_grid=80, mass_p=1.0, vol_p=1.0, gravity=(0.0, -200.0, 0.0), boundary=0.05,
                jp_min=0.6, jp_max=20.0, alpha=0.0, materials=None):
    p = OracleParams()
    p.dim, p.n_grid = dim, n_grid
    p.mass_p, p.vol_p = mass_p, vol_p
    for k in range(3):
        p.gravity[k] = gravity[k]
    p.boundary, p.jp_min, p.jp_max, p.alpha = boundary, jp_min, jp_max, alpha
    mats = DEFAULT_MATERIALS if materials is None else materials
    p.n_materials = len(mats)
    for i, m in enumerate(mats):
        # clamp limits are formed in fp32 like the reference's `1.0f - 2.5e-2f` (:169)
        kind, E, nu, h, lo, hi = m
        p.mat[i].kind, p.mat[i].E, p.mat[i].nu, p.mat[i].hardening = kind, E, nu, h
        p.mat[i].sig_lo = float(np.float32(lo))
        p.mat[i].sig_hi = float(np.float32(hi))
    return p


def record_words(dim):
    return 14 if dim == 2 else 26


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = L = ctypes.CDLL(ORACLE_SO)
        L.oracle_advance.restype = ctypes.c_int
        L.oracle_advance.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_advance_mt.restype = ctypes.c_int
        L.oracle_advance_mt.argtypes = L.oracle_advance.argtypes + [ctypes.c_int]
        L.oracle_bin.restype = ctypes.c_int
        L.oracle_bin.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]
        L.oracle_seed_block2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_ulonglong]
        L.oracle_lame.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        assert L.oracle_params_bytes() == ctypes.sizeof(OracleParams)

    def advance(self, params, dt, particles, n_steps=1, want_grid=False, want_post_p2g=False, threads=1):
        """In-place n_steps substeps on an (n, 14|26) float32 AoS array (last word = int32 id).
        threads > 1 runs the same loops on several host threads with BITWISE identical results
        (every grid node still receives its contributions in particle-index order)."""
        assert particles.dtype == np.float32 and particles.flags.c_contiguous
        assert particles.shape[1] == record_words(params.dim)
        n1 = params.n_grid + 1
        shape = (n1,) * params.dim + (params.dim + 1,)
        grid = np.zeros(shape, np.float32) if want_grid else None
        tap = np.zeros(shape, np.float32) if want_post_p2g else None
        rc = self.lib.oracle_advance_mt(ctypes.byref(params), ctypes.c_float(dt), particles.ctypes.data,
                                        particles.shape[0], n_steps,
                                        grid.ctypes.data if grid is not None else None,
                                        tap.ctypes.data if tap is not None else None, int(threads))
        if rc != 0:
            raise RuntimeError("oracle_advance rc=%d" % rc)
        return grid, tap

    def seed_block2(self, n, cx, cy, half, c, skip=0):
        out = np.zeros((n, 14), np.float32)
        self.lib.oracle_seed_block2(out.ctypes.data, n, cx, cy, half, c, skip)
        return out

    def bin(self, dim, n_grid, bin_edge, x, order_in=None):
        """Stable binning oracle -> (cell[n,dim], key[n], order[n], bin_start[n_bins+1])."""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        nb = (n_grid - 1 + bin_edge - 1) // bin_edge
        n_bins = nb ** dim
        cell = np.zeros((n, dim), np.int32)
        key = np.zeros(n, np.int32)
        order = np.zeros(n, np.int32)
        start = np.zeros(n_bins + 1, np.int32)
        oin = None
        if order_in is not None:
            order_in = np.ascontiguousarray(order_in, np.int32)
            oin = order_in.ctypes.data
        got = self.lib.oracle_bin(dim, n_grid, bin_edge, x.ctypes.data, n, oin, cell.ctypes.data,
                                  key.ctypes.data, order.ctypes.data, start.ctypes.data)
        assert got == n_bins
        return cell, key, order, start

    def lame(self, E, nu):
        mu, la = ctypes.c_float(), ctypes.c_float()
        self.lib.oracle_lame(E, nu, ctypes.byref(mu), ctypes.byref(la))
        return mu.value, la.value

    def polar2(self, m):
        m = np.ascontiguousarray(m, np.float32)
        R, S = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self.lib.oracle_polar2(m.ctypes.data_as(ctypes.c_void_p), R.ctypes.data_as(ctypes.c_void_p),
                               S.ctypes.data_as(ctypes.c_void_p))
        return R, S

    def svd2(self, m):
        m = np.ascontiguousarray(m, np.float32)
        U, s, V = (np.zeros(4, np.float32) for _ in range(3))
        self.lib.oracle_svd2(m.ctypes.data_as(ctypes.c_void_p), U.ctypes.data_as(ctypes.c_void_p),
                             s.ctypes.data_as(ctypes.c_void_p), V.ctypes.data_as(ctypes.c_void_p))
        return U, s, V

    def rotation3(self, m):
        """rotation factor of the 3D stress (Newton polar iteration, SVD fallback): 9 floats, column-major"""
        m = np.ascontiguousarray(m, np.float32)
        R = np.zeros(9, np.float32)
        self.lib.oracle_rotation3(m.ctypes.data_as(ctypes.c_void_p), R.ctypes.data_as(ctypes.c_void_p))
        return R

    def plastic_project3(self, lo, hi, m):
        """-> (F', det F / det F') of the 3D snow projection (oracle/mpm_oracle.cpp plastic_project3)"""
        m = np.ascontiguousarray(m, np.float32).copy()
        self.lib.oracle_plastic_project3.restype = ctypes.c_float
        r = self.lib.oracle_plastic_project3(ctypes.c_float(lo), ctypes.c_float(hi), m.ctypes.data_as(ctypes.c_void_p))
        return m, float(r)

    def svd3(self, m):
        m = np.ascontiguousarray(m, np.float32)
        U, V = np.zeros(9, np.float32), np.zeros(9, np.float32)
        s = np.zeros(3, np.float32)
        self.lib.oracle_svd3(m.ctypes.data_as(ctypes.c_void_p), U.ctypes.data_as(ctypes.c_void_p),
                             s.ctypes.data_as(ctypes.c_void_p), V.ctypes.data_as(ctypes.c_void_p))
        return U, s, V


class Reference:
    """The unmodified reference program's advance() (shipped constants only)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (run `make -C oracle` where /root/reference exists)")
        self.lib = L = ctypes.CDLL(REF_SO)
        L.ref_dt.restype = ctypes.c_float
        L.ref_mu0.restype = ctypes.c_float
        L.ref_lambda0.restype = ctypes.c_float
        L.ref_rand.restype = ctypes.c_float
        L.ref_set.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ref_get.argtypes = [ctypes.c_void_p]
        L.ref_get_grid.argtypes = [ctypes.c_void_p]
        L.ref_advance_dt.argtypes = [ctypes.c_float, ctypes.c_int]
        for f in (L.ref_polar2, L.ref_svd2):
            f.argtypes = None
        assert L.ref_particle_bytes() == 56

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def set(self, particles):
        assert particles.dtype == np.float32 and particles.shape[1] == 14 and particles.flags.c_contiguous
        self.lib.ref_set(particles.ctypes.data, particles.shape[0])

    def get(self):
        out = np.zeros((self.lib.ref_count(), 14), np.float32)
        self.lib.ref_get(out.ctypes.data)
        return out

    def grid(self):
        n1 = self.lib.ref_num_grid() + 1
        g = np.zeros((n1, n1, 3), np.float32)
        self.lib.ref_get_grid(g.ctypes.data)
        return g

    def advance(self, n_steps=1):
        self.lib.ref_advance(n_steps)

    def polar2(self, m):
        m = np.ascontiguousarray(m, np.float32)
        R, S = np.zeros(4, np.float32), np.zeros(4, np.float32)
        self.lib.ref_polar2(ctypes.c_void_p(m.ctypes.data), ctypes.c_void_p(R.ctypes.data),
                            ctypes.c_void_p(S.ctypes.data))
        return R, S

    def svd2(self, m):
        m = np.ascontiguousarray(m, np.float32)
        U, s, V = (np.zeros(4, np.float32) for _ in range(3))
        self.lib.ref_svd2(ctypes.c_void_p(m.ctypes.data), ctypes.c_void_p(U.ctypes.data),
                          ctypes.c_void_p(s.ctypes.data), ctypes.c_void_p(V.ctypes.data))
        return U, s, V
