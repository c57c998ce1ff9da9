"""The five names the reference's Python driver imports from its withheld solver module
(`exec.py:5`: ``createFilePaths, progressBar, initialization, post_process, subStep``), backed by
libmpm.so.  A maintainer changes one line of exec.py,

    from mpm_flip98a_b200.exec_shim import createFilePaths, progressBar, initialization, post_process, subStep

and the loop of exec.py:20-29 runs unchanged (``subStep()`` takes no arguments, exec.py:24).  The physics is
this repository's MLS-MPM substep -- fluid material, FLIP blend = ``flipBlendParameter`` (config.py:29) --
NOT the thesis' stabilised solver, whose source is not in the reference repository (README.md:23-25).
The dam-break geometry follows config.py:30-37: a column of width : height = 0.057 : 0.114 in a box of
0.4375 m, 65 x 130 particles, mapped onto the engine's unit square.  `taichi` is not needed.
"""
import os
import sys

import numpy as np

from . import scenes
from .engine import Engine

_state = {"engine": None, "settings": None, "frame": 0}


class Settings:
    """The fields of config.py's NumericalSettings that the driver loop touches (exec.py:18-29)."""

    def __init__(self, n_grid=104, time_step=None, flip_blend=0.0, frame_substeps=100):
        self.numGrids = n_grid + 1          # config.py:38 (105 nodes)
        self.numCells = n_grid              # config.py:39
        self.numParticlesX, self.numParticlesY = 65, 130   # config.py:30-31
        self.numParticles = self.numParticlesX * self.numParticlesY
        self.domainLength = 0.4375          # config.py:33
        self.fluidWidth, self.fluidHeight = 0.057, 0.114   # config.py:34-35
        self.flipBlendParameter = flip_blend  # config.py:29
        dt, vol = scenes.scaled_constants(n_grid)
        self.timeStep = time_step or dt
        self.volume = vol
        self.frameRate = self.timeStep * frame_substeps    # exec.py:21: substeps per frame = frameRate // timeStep
        self.totalTime = 0.0
        self.simulationTime = 3.0


def initialization(settings=None, device=0):
    """exec.py:12 -- builds the dam-break column and uploads it; returns the settings object."""
    s = settings or Settings()
    w = s.fluidWidth / s.domainLength
    hgt = s.fluidHeight / s.domainLength
    xs = 0.05 + (np.arange(s.numParticlesX) + 0.5) * (w / s.numParticlesX)
    ys = 0.05 + (np.arange(s.numParticlesY) + 0.5) * (hgt / s.numParticlesY)
    x = np.stack(np.meshgrid(xs, ys, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    p = scenes.make_records(x, scenes.FLUID, 2)
    if _state["engine"] is not None:
        _state["engine"].close()
    e = Engine(dim=2, n_grid=s.numCells, capacity=len(p), dt=s.timeStep, vol_p=s.volume, alpha=s.flipBlendParameter,
               device=device)
    e.upload(p)
    _state.update(engine=e, settings=s, frame=0)
    return s


def subStep():
    """exec.py:24 -- one substep of the engine's configured dt."""
    _state["engine"].substep(1)


def createFilePaths(numerical=None):
    """exec.py:16 -- output directories, named like the reference's ignored ones (.gitignore:3-4)."""
    s = numerical or _state["settings"]
    tag = "dt%.0e_pointwise" % (s.timeStep if s is not None else 1e-6)
    filepath, vtkpath = "mov_" + tag, "vtk_" + tag
    os.makedirs(filepath, exist_ok=True)
    os.makedirs(vtkpath, exist_ok=True)
    return filepath, vtkpath


def progressBar(t, total, width=40, stream=sys.stdout):
    """exec.py:28."""
    f = min(1.0, max(0.0, t / total if total else 1.0))
    stream.write("\r[%s%s] %5.1f%%" % ("#" * int(f * width), "." * (width - int(f * width)), 100 * f))
    stream.flush()


def write_frame(path, x, res=512, background=0x112F41, colour=0x068587):
    """One frame of the "mov" directory: the particles splatted as single pixels into a res x res binary PPM
    (background = the colour exec.py:14 gives its ti.GUI; y up like the GUI).  Headless stand-in for gui.show(file)."""
    img = np.empty((res, res, 3), np.uint8)
    img[:] = [(background >> 16) & 255, (background >> 8) & 255, background & 255]
    ix = np.clip((x[:, 0] * res).astype(np.int64), 0, res - 1)
    iy = np.clip(((1.0 - x[:, 1]) * res).astype(np.int64), 0, res - 1)
    img[iy, ix] = [(colour >> 16) & 255, (colour >> 8) & 255, colour & 255]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (res, res))
        f.write(img.tobytes())
    return path


def post_process(numParticles, gui, vtkpath, filepath, num_substeps, count):
    """exec.py:29 -- reads the particles back and writes one legacy-VTK point cloud (x, v, J = det F) into `vtkpath`
    and one frame of the movie (PPM point splat) into `filepath`.  `gui` is accepted for signature compatibility
    and ignored (headless: taichi's GUI is not available here)."""
    p = _state["engine"].read()[:numParticles]
    n = len(p)
    frame = _state["frame"]
    _state["frame"] += 1
    if filepath:
        write_frame(os.path.join(filepath, "frame_%06d.ppm" % frame), p[:, 0:2])
    name = os.path.join(vtkpath, "particles_%06d.vtk" % frame)
    J = p[:, 4] * p[:, 7] - p[:, 5] * p[:, 6]
    with open(name, "w") as f:
        f.write("# vtk DataFile Version 3.0\nMLS-MPM substep %d\nASCII\nDATASET POLYDATA\nPOINTS %d float\n" % (count, n))
        np.savetxt(f, np.column_stack([p[:, 0], p[:, 1], np.zeros(n)]), fmt="%.7g")
        f.write("POINT_DATA %d\nVECTORS velocity float\n" % n)
        np.savetxt(f, np.column_stack([p[:, 2], p[:, 3], np.zeros(n)]), fmt="%.7g")
        f.write("SCALARS detF float 1\nLOOKUP_TABLE default\n")
        np.savetxt(f, J, fmt="%.7g")
    return name
