"""SURVEY 8f "next" rows: the reference's two callers on top of the C-ABI.
 * examples/mls_mpm88_driver.cpp == main() of cpp_validation/mls-mpm88-explained.cpp:203-227 (headless)
 * mpm_flip98a_b200/exec_shim.py == the five names exec.py:5 imports from the withheld solver module"""
import os
import subprocess

import numpy as np
import pytest

from tests.util import fields, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "examples", "build", "mls_mpm88_driver")


def build_driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    assert os.path.exists(DRIVER)


def test_cpp_driver_builds_against_the_header():
    build_driver()  # plain g++ against include/mpm.h + libmpm.so: the ABI is usable from C++ without CUDA headers


def test_exec_shim_exports_the_reference_names():
    from mpm_flip98a_b200 import exec_shim
    for name in ("createFilePaths", "progressBar", "initialization", "post_process", "subStep"):  # exec.py:5
        assert callable(getattr(exec_shim, name))


@pytest.mark.gpu
def test_cpp_driver_reproduces_the_reference_scene(tmp_path, shipped):
    build_driver()
    dump = tmp_path / "p.bin"
    frames = tmp_path / "frames"
    frames.mkdir()
    out = subprocess.run([DRIVER, "--steps", "100", "--dump", str(dump), "--frames", str(frames)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    p = np.fromfile(dump, np.float32).reshape(-1, 14)
    # its own seeding is the reference's (bit-exact initial state => same particles as golden step0) ...
    want = shipped["step100"]
    fw, fg = fields(want, 2), fields(p, 2)
    # ... and 100 substeps later it is still on the reference's trajectory (this scene amplifies the 1e-7
    # summation-order noise of every substep: measured 1.7e-5 on x here, 1.7e-3 after 1000 substeps on the CPU)
    assert rel_l2(fg["x"], fw["x"]) < 1e-4 and rel_l2(fg["v"], fw["v"]) < 2e-2
    assert np.abs(p[:, 0:2].mean(0) - want[:, 0:2].mean(0)).max() < 1e-5
    assert len(list(frames.glob("*.ppm"))) == 10  # every int(frame_dt/dt) = 10 substeps, :217
    assert (frames / "00000.ppm").read_bytes()[:2] == b"P6"


@pytest.mark.gpu
def test_cpp_driver_initial_state_is_bit_exact(tmp_path, shipped):
    build_driver()
    dump = tmp_path / "p0.bin"
    out = subprocess.run([DRIVER, "--steps", "0", "--dump", str(dump)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    p = np.fromfile(dump, np.float32).reshape(-1, 14)
    assert np.array_equal(p.view(np.int32), shipped["step0"].view(np.int32))  # upload -> read is lossless too


@pytest.mark.gpu
def test_exec_shim_runs_the_reference_loop(tmp_path, monkeypatch):
    from mpm_flip98a_b200 import exec_shim as fc
    monkeypatch.chdir(tmp_path)
    numerical = fc.initialization(fc.Settings(frame_substeps=50))
    filepath, vtkpath = fc.createFilePaths(numerical)
    count = 0
    for _ in range(2):  # the loop of exec.py:20-29, two frames
        num_substeps = int(numerical.frameRate // numerical.timeStep)
        for s in range(num_substeps):
            fc.subStep()
            count += 1
            numerical.totalTime += numerical.timeStep
        fc.progressBar(numerical.totalTime, numerical.simulationTime, stream=open(os.devnull, "w"))
        name = fc.post_process(numerical.numParticles, None, vtkpath, filepath, num_substeps, count)
    assert count in (98, 100)  # float floor of frameRate // timeStep, as in the reference loop
    text = open(name).read()
    assert "POINTS 8450 float" in text and "VECTORS velocity" in text  # 65 x 130 particles, config.py:30-32
    p = fc._state["engine"].read()
    assert np.isfinite(p).all() and p[:, 3].mean() < 0  # the column falls
    frames = sorted(os.listdir(filepath))  # the "mov" directory of exec.py:16,29 (.gitignore:3): one PPM per frame
    assert frames == ["frame_000000.ppm", "frame_000001.ppm"]
    raw = open(os.path.join(filepath, frames[-1]), "rb").read()
    assert raw.startswith(b"P6\n512 512\n255\n") and len(raw) == 15 + 512 * 512 * 3


def test_checkpoint_format_roundtrip_cpu(tmp_path):
    # format only (no GPU): header + the reference's 56-byte records, readable with a plain fread
    from mpm_flip98a_b200 import checkpoint, scenes

    class FakeEngine:  # the two attributes save() uses
        dim = 2

        def __init__(self, p):
            from mpm_flip98a_b200.engine import default_config
            self.cfg, self._p = default_config(2), p

        def read(self):
            return self._p
    p = scenes.commented_three_blocks()
    checkpoint.save(FakeEngine(p), tmp_path / "c.mpm", step=7)
    h, q = checkpoint.load(tmp_path / "c.mpm")
    assert h["step"] == 7 and h["n_particles"] == 3000 and h["record_bytes"] == 56 and len(h["materials"]) == 4
    assert np.array_equal(q.view(np.int32), p.view(np.int32))
    raw = open(tmp_path / "c.mpm", "rb").read()
    assert raw[-56 * 3000:] == p.tobytes()  # records are the file's tail, verbatim


@pytest.mark.gpu
def test_checkpoint_restart_continues_the_run(tmp_path):
    import mpm_flip98a_b200 as mpm
    from mpm_flip98a_b200 import checkpoint, scenes
    p = scenes.jelly_drop()
    with mpm.Engine(capacity=len(p)) as e:
        e.upload(p)
        e.substep(300)
        checkpoint.save(e, tmp_path / "c.mpm", step=300)
        e.substep(300)
        straight = e.read()
    e2, h = checkpoint.restore(mpm.Engine, tmp_path / "c.mpm")
    assert h["step"] == 300
    e2.substep(300)
    resumed = e2.read()
    e2.close()
    for k, (a, b) in {k: (fields(resumed, 2)[k], fields(straight, 2)[k]) for k in ("x", "v", "F")}.items():
        assert rel_l2(a, b) < 1e-4, k  # identical up to atomic summation order over 300 substeps


@pytest.mark.gpu
def test_cpp_driver_on_several_slabs(tmp_path, shipped):
    """--devices: the same main() loop through mpm_group_* (several x-slabs behind one handle; here all on cuda:0)"""
    build_driver()
    dump = tmp_path / "p.bin"
    out = subprocess.run([DRIVER, "--steps", "100", "--dump", str(dump), "--devices", "0,0,0"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "slabs 3:" in out.stdout
    p = np.fromfile(dump, np.float32).reshape(-1, 14)
    want = shipped["step100"]
    fw, fg = fields(want, 2), fields(p, 2)
    assert rel_l2(fg["x"], fw["x"]) < 1e-4 and rel_l2(fg["v"], fw["v"]) < 2e-2
    assert np.abs(p[:, 0:2].mean(0) - want[:, 0:2].mean(0)).max() < 1e-5


def test_exec_shim_frame_writer_cpu(tmp_path):
    """the movie frame of exec.py's "mov" directory: a binary PPM point splat, y up (no GPU involved)"""
    from mpm_flip98a_b200 import exec_shim as fc
    x = np.array([[0.25, 0.75], [0.999, 0.001]], np.float32)
    path = fc.write_frame(str(tmp_path / "f.ppm"), x, res=64)
    raw = open(path, "rb").read()
    hdr = b"P6\n64 64\n255\n"
    assert raw.startswith(hdr) and len(raw) == len(hdr) + 64 * 64 * 3
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(64, 64, 3)
    assert tuple(img[16, 16]) == (0x06, 0x85, 0x87) and tuple(img[63, 63]) == (0x06, 0x85, 0x87)  # row = (1 - y) * res
    assert tuple(img[0, 0]) == (0x11, 0x2F, 0x41)  # background of exec.py:14
