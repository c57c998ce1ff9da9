"""ctypes loader of tests/host_check.cpp (TEST ONLY): the product's mpm_math.cuh run on the CPU."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "build", "libhostcheck.so")
SRC = os.path.join(HERE, "host_check.cpp")
HDR = os.path.join(HERE, "..", "mpm_flip98a_b200", "csrc", "mpm_math.cuh")
HDR2 = os.path.join(HERE, "..", "mpm_flip98a_b200", "csrc", "mpm_math2.cuh")


def build():
    if os.path.exists(SO) and os.path.getmtime(SO) > max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2)):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-Wno-unknown-pragmas", SRC, "-o", SO])


class HostCheck:
    def __init__(self):
        build()
        self.lib = ctypes.CDLL(SO)

    def params(self, n_grid, mass_p, vol_p, gravity, boundary, jp_min, jp_max, alpha, materials):
        buf = ctypes.create_string_buffer(self.lib.hostcheck_params_bytes())
        g = (ctypes.c_float * 3)(*gravity)
        m = np.array([[float(x) for x in row] for row in materials], np.float32)
        self.lib.hostcheck_make_params(buf, n_grid, ctypes.c_float(mass_p), ctypes.c_float(vol_p), g,
                                       ctypes.c_float(boundary), ctypes.c_float(jp_min), ctypes.c_float(jp_max),
                                       ctypes.c_float(alpha), len(materials), m.ctypes.data_as(ctypes.c_void_p))
        return buf

    def advance(self, P, dim, n_grid, dt, particles, n_steps=1):
        n1 = n_grid + 1
        grid = np.zeros((n1,) * dim + (dim + 1,), np.float32)
        tap = np.zeros_like(grid)
        self.lib.hostcheck_advance(P, dim, ctypes.c_float(dt), particles.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_longlong(particles.shape[0]), n_steps,
                                   grid.ctypes.data_as(ctypes.c_void_p), tap.ctypes.data_as(ctypes.c_void_p))
        return grid, tap

    def advance_packed2(self, P, n_grid, dt, particles, n_steps=1, exact_gather=True):
        """The packed 2D path of the default substep kernel (mpm_math2.cuh), sequential order."""
        n1 = n_grid + 1
        grid = np.zeros((n1, n1, 3), np.float32)
        tap = np.zeros_like(grid)
        self.lib.hostcheck_advance_packed2(P, ctypes.c_float(dt), particles.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_longlong(particles.shape[0]), n_steps,
                                           grid.ctypes.data_as(ctypes.c_void_p), tap.ctypes.data_as(ctypes.c_void_p),
                                           1 if exact_gather else 0)
        return grid, tap
