"""ctypes binding of include/mpm.h.

Mirrors the call surface a user of the reference has: build the scene (``add_object``,
cpp_validation/mls-mpm88-explained.cpp:191-196) -> ``upload``; the ``for ... advance(dt)`` loop
(:214-215) -> ``substep``; reading ``particles`` (:220-222) -> ``read``.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPM_LIBRARY") or os.path.join(HERE, "libmpm.so")  # override: A/B builds only

MPM_ABI_VERSION = 1
KIND_FLUID, KIND_JELLY, KIND_SNOW = 0, 1, 2
FLAG_CAPTURE_POST_P2G = 1
FLAG_NAIVE = 2
FLAG_STRICT = 4
FLAG_G2P_TILE = 8
FLAG_NO_FUSE = 16
FLAG_OVERLAP = 32
FLAG_DETERMINISTIC = 64
FLAG_FUSE_3D = 128

_ERRORS = {-1: "MPM_E_INVALID", -2: "MPM_E_CUDA", -3: "MPM_E_CAPACITY", -4: "MPM_E_DOMAIN", -5: "MPM_E_CFL",
           -6: "MPM_E_STATE"}


class MpmError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s (%d): %s" % (_ERRORS.get(code, "MPM_E_?"), code, text))
        self.code = code


class Material(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("E", ctypes.c_float), ("nu", ctypes.c_float),
                ("hardening", ctypes.c_float), ("sig_lo", ctypes.c_float), ("sig_hi", ctypes.c_float)]


class Config(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int), ("dim", ctypes.c_int), ("n_grid", ctypes.c_int),
                ("dt", ctypes.c_float), ("mass_p", ctypes.c_float), ("vol_p", ctypes.c_float),
                ("gravity", ctypes.c_float * 3), ("boundary", ctypes.c_float), ("jp_min", ctypes.c_float),
                ("jp_max", ctypes.c_float), ("alpha", ctypes.c_float), ("n_materials", ctypes.c_int),
                ("materials", Material * 4), ("capacity", ctypes.c_longlong), ("device", ctypes.c_int),
                ("flags", ctypes.c_int), ("slab_lo", ctypes.c_int), ("slab_hi", ctypes.c_int),
                ("stream", ctypes.c_void_p), ("bin_edge", ctypes.c_int), ("rebin_every", ctypes.c_int),
                ("mig_records", ctypes.c_int), ("reserved", ctypes.c_int * 5)]


class SlabDesc(ctypes.Structure):
    _fields_ = [("send_lo", ctypes.c_void_p), ("send_hi", ctypes.c_void_p), ("recv_lo", ctypes.c_void_p),
                ("recv_hi", ctypes.c_void_p), ("bytes", ctypes.c_longlong), ("halo_bytes", ctypes.c_longlong),
                ("record_bytes", ctypes.c_int), ("record_capacity", ctypes.c_int), ("has_lo", ctypes.c_int),
                ("has_hi", ctypes.c_int)]


class Profile(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * 8), ("launches", ctypes.c_longlong * 8), ("substeps", ctypes.c_longlong),
                ("fallback_particles", ctypes.c_longlong), ("rebin_interval", ctypes.c_longlong),
                ("fused_substeps", ctypes.c_longlong)]


PHASES = ("clear", "p2g", "grid", "g2p", "bin", "halo", "migrate")

# every symbol include/mpm.h declares: name -> (restype, argtypes)
_H = ctypes.c_void_p
SYMBOLS = {
    "mpm_default_config": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.c_int]),
    "mpm_config_bytes": (ctypes.c_int, []),
    "mpm_create": (_H, [ctypes.POINTER(Config)]),
    "mpm_destroy": (None, [_H]),
    "mpm_last_error": (ctypes.c_char_p, [_H]),
    "mpm_upload_particles": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]),
    "mpm_substep": (ctypes.c_int, [_H, ctypes.c_float, ctypes.c_int]),
    "mpm_read_particles": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]),
    "mpm_read_grid": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_void_p]),
    "mpm_upload_particles_ids": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]),
    "mpm_read_particles_ids": (ctypes.c_longlong, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]),
    "mpm_storage_extent": (ctypes.c_longlong, [_H]),
    "mpm_particle_count": (ctypes.c_longlong, [_H]),
    "mpm_resort": (ctypes.c_int, [_H]),
    "mpm_set_rebin_every": (ctypes.c_int, [_H, ctypes.c_int]),
    "mpm_synchronize": (ctypes.c_int, [_H]),
    "mpm_poll_status": (ctypes.c_int, [_H]),
    "mpm_profile_enable": (ctypes.c_int, [_H, ctypes.c_int]),
    "mpm_profile_read": (ctypes.c_int, [_H, ctypes.POINTER(Profile)]),
    "mpm_bin_particles": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mpm_slab_describe": (ctypes.c_int, [_H, ctypes.POINTER(SlabDesc)]),
    "mpm_slab_begin": (ctypes.c_int, [_H, ctypes.c_float]),
    "mpm_slab_step": (ctypes.c_int, [_H, ctypes.c_float]),
    "mpm_slab_settle": (ctypes.c_int, [_H]),
    "mpm_group_create": (_H, [ctypes.POINTER(Config), ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "mpm_group_destroy": (None, [_H]),
    "mpm_group_last_error": (ctypes.c_char_p, [_H]),
    "mpm_group_upload_particles": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_longlong]),
    "mpm_group_substep": (ctypes.c_int, [_H, ctypes.c_float, ctypes.c_int]),
    "mpm_group_synchronize": (ctypes.c_int, [_H]),
    "mpm_group_read_particles": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_longlong]),
    "mpm_group_poll_status": (ctypes.c_int, [_H]),
    "mpm_group_slab": (ctypes.c_int, [_H, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]),
}

_lib = None


def load_library(path=None):
    """dlopen libmpm.so and type every entry point.  Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "or `make -C mpm_flip98a_b200/csrc`" % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name, None)
        if f is None:
            if os.environ.get("MPM_LIBRARY"):  # A/B runs against an older build
                continue
            raise ImportError("%s does not export %s" % (p, name))
        f.restype = res
        f.argtypes = args
    if lib.mpm_config_bytes() != ctypes.sizeof(Config):
        raise ImportError("mpm_config layout mismatch: library %d bytes, binding %d" %
                          (lib.mpm_config_bytes(), ctypes.sizeof(Config)))
    if path is None:
        _lib = lib
    return lib


def default_config(dim=2):
    lib = load_library()
    c = Config()
    rc = lib.mpm_default_config(ctypes.byref(c), dim)
    if rc != 0:
        raise MpmError(rc, "mpm_default_config")
    return c


class Engine:
    """One handle == one GPU == one (slab of a) simulation."""

    def __init__(self, dim=2, n_grid=80, capacity=1 << 20, dt=None, vol_p=None, alpha=0.0, materials=None,
                 device=0, flags=0, slab=None, bin_edge=0, rebin_every=0, gravity=None, stream=None, **overrides):
        self.lib = load_library()
        c = default_config(dim)
        c.n_grid = n_grid
        c.capacity = capacity
        if dt is not None:
            c.dt = dt
        if vol_p is not None:
            c.vol_p = vol_p
        c.alpha = alpha
        c.device = device
        c.flags = flags
        c.slab_lo, c.slab_hi = slab if slab is not None else (0, n_grid)
        c.bin_edge = bin_edge
        c.rebin_every = rebin_every
        if gravity is not None:
            for k in range(3):
                c.gravity[k] = gravity[k]
        if stream is not None:
            c.stream = stream
        if materials is not None:
            c.n_materials = len(materials)
            for i, m in enumerate(materials):
                kind, E, nu, h, lo, hi = m
                c.materials[i] = Material(kind, E, nu, h, float(np.float32(lo)), float(np.float32(hi)))
        for k, v in overrides.items():
            setattr(c, k, v)
        self.cfg = c
        self.dim = dim
        self.words = 2 * dim + 2 * dim * dim + 2
        self.h = self.lib.mpm_create(ctypes.byref(c))
        if not self.h:
            raise MpmError(-1, self.lib.mpm_last_error(None).decode())

    def _check(self, rc):
        if rc < 0:
            raise MpmError(rc, self.lib.mpm_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpm_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def count(self):
        return self.lib.mpm_particle_count(self.h)

    def upload(self, particles):
        p = np.ascontiguousarray(particles, np.float32)
        assert p.ndim == 2 and p.shape[1] == self.words, "records must be (n, %d) float32" % self.words
        self._check(self.lib.mpm_upload_particles(self.h, p.ctypes.data, p.shape[0], 0))

    def upload_ids(self, particles, ids):
        p = np.ascontiguousarray(particles, np.float32)
        ids = np.ascontiguousarray(ids, np.int32)
        assert p.ndim == 2 and p.shape[1] == self.words and len(ids) == len(p)
        self._check(self.lib.mpm_upload_particles_ids(self.h, p.ctypes.data, ids.ctypes.data, p.shape[0], 0))

    def read_ids(self):
        """(records, ids) of the live particles of this handle, storage order."""
        ext = self.lib.mpm_storage_extent(self.h)
        out = np.empty((ext, self.words), np.float32)
        ids = np.empty(ext, np.int32)
        got = self._check(self.lib.mpm_read_particles_ids(self.h, out.ctypes.data, ids.ctypes.data, ext, 0))
        keep = ids[:got] >= 0
        return out[:got][keep], ids[:got][keep]

    def slab(self):
        d = SlabDesc()
        self._check(self.lib.mpm_slab_describe(self.h, ctypes.byref(d)))
        return d

    def slab_begin(self, dt=0.0):
        """P2G of the resident particles + staged messages; True when the caller must exchange them."""
        return self._check(self.lib.mpm_slab_begin(self.h, dt)) == 1

    def slab_step(self, dt=0.0):
        self._check(self.lib.mpm_slab_step(self.h, dt))

    def slab_settle(self):
        self._check(self.lib.mpm_slab_settle(self.h))

    def upload_device(self, dev_ptr, n):
        self._check(self.lib.mpm_upload_particles(self.h, dev_ptr, n, 1))

    def substep(self, n_steps=1, dt=0.0):
        self._check(self.lib.mpm_substep(self.h, dt, n_steps))

    def resort(self):
        self._check(self.lib.mpm_resort(self.h))

    def set_rebin_every(self, every):
        if hasattr(self.lib, "mpm_set_rebin_every"):  # absent from older builds (MPM_LIBRARY A/B runs)
            self._check(self.lib.mpm_set_rebin_every(self.h, int(every)))

    def synchronize(self):
        self._check(self.lib.mpm_synchronize(self.h))

    def poll_status(self):
        return self.lib.mpm_poll_status(self.h)

    def read(self, n=None):
        n = self.count if n is None else n
        out = np.empty((n, self.words), np.float32)
        self._check(self.lib.mpm_read_particles(self.h, out.ctypes.data, n, 0))
        return out

    def read_device(self, dev_ptr, n):
        self._check(self.lib.mpm_read_particles(self.h, dev_ptr, n, 1))

    def profile_enable(self, on=True):
        self._check(self.lib.mpm_profile_enable(self.h, 1 if on else 0))

    def profile(self):
        """{phase: (device_ms, kernel_launches)} accumulated since profile_enable, plus 'substeps'."""
        pr = Profile()
        self._check(self.lib.mpm_profile_read(self.h, ctypes.byref(pr)))
        out = {name: (pr.ms[i], pr.launches[i]) for i, name in enumerate(PHASES)}
        out["substeps"] = pr.substeps
        out["fallback_particles"] = pr.fallback_particles
        out["rebin_interval"] = pr.rebin_interval
        out["fused_substeps"] = pr.fused_substeps
        return out

    def grid_shape(self):
        n1 = self.cfg.n_grid + 1
        xhi = min(self.cfg.slab_hi, self.cfg.n_grid - 1)
        ncol = xhi - self.cfg.slab_lo + 2
        return (ncol,) + (n1,) * (self.dim - 1) + (self.dim + 1,)

    def read_grid(self, stage=0):
        out = np.empty(self.grid_shape(), np.float32)
        self._check(self.lib.mpm_read_grid(self.h, stage, out.ctypes.data))
        return out

    def bin_particles(self, want_cell=True):
        n = self.count
        cell = np.empty((n, self.dim), np.int32) if want_cell else None
        key = np.empty(n, np.int32)
        order = np.empty(n, np.int32)
        # n_bins is only known from the call; size the start array generously from the config
        edge = self.cfg.bin_edge or (8 if self.dim == 2 else 4)
        nb = (self.cfg.n_grid - 1 + edge - 1) // edge
        start = np.empty(nb ** self.dim + 1, np.int32)
        n_bins = self._check(self.lib.mpm_bin_particles(self.h, cell.ctypes.data if want_cell else None,
                                                        key.ctypes.data, order.ctypes.data, start.ctypes.data))
        return cell, key, order, start[:n_bins + 1]


class Group:
    """Several GPUs behind one handle (include/mpm.h, mpm_group_*): `devices` lists one CUDA ordinal per x-slab."""

    def __init__(self, devices, dim=2, n_grid=80, capacity=1 << 20, dt=None, vol_p=None, alpha=0.0, flags=0,
                 rebin_every=0):
        self.lib = load_library()
        c = default_config(dim)
        c.n_grid, c.capacity, c.alpha, c.flags, c.rebin_every = n_grid, capacity, alpha, flags, rebin_every
        if dt is not None:
            c.dt = dt
        if vol_p is not None:
            c.vol_p = vol_p
        self.cfg, self.dim, self.words = c, dim, 2 * dim + 2 * dim * dim + 2
        arr = (ctypes.c_int * len(devices))(*devices)
        self.n_slabs = len(devices)
        self.g = self.lib.mpm_group_create(ctypes.byref(c), arr, len(devices))
        if not self.g:
            raise MpmError(-1, "mpm_group_create")
        self.n = 0

    def _check(self, rc):
        if rc < 0:
            raise MpmError(rc, self.lib.mpm_group_last_error(self.g).decode())
        return rc

    def close(self):
        if getattr(self, "g", None):
            self.lib.mpm_group_destroy(self.g)
            self.g = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, particles):
        p = np.ascontiguousarray(particles, np.float32)
        assert p.ndim == 2 and p.shape[1] == self.words
        self._check(self.lib.mpm_group_upload_particles(self.g, p.ctypes.data, p.shape[0]))
        self.n = p.shape[0]

    def substep(self, n_steps=1, dt=0.0):
        self._check(self.lib.mpm_group_substep(self.g, dt, n_steps))

    def read(self):
        out = np.empty((self.n, self.words), np.float32)
        self._check(self.lib.mpm_group_read_particles(self.g, out.ctypes.data, self.n))
        return out

    def poll_status(self):
        return self.lib.mpm_group_poll_status(self.g)

    def synchronize(self):
        self._check(self.lib.mpm_group_synchronize(self.g))

    def slabs(self):
        out = []
        for k in range(self.n_slabs):
            d, lo, hi, cnt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
            self._check(self.lib.mpm_group_slab(self.g, k, ctypes.byref(d), ctypes.byref(lo), ctypes.byref(hi),
                                                ctypes.byref(cnt)))
            out.append((d.value, lo.value, hi.value, cnt.value))
        return out
