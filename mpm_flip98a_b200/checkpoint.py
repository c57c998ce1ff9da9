"""Checkpoint / restart.  The on-disk particle format is the reference's own
56-byte record (cpp_validation/mls-mpm88-explained.cpp:28-42; 104 bytes in 3D) in upload order, behind a
small self-describing header -- so a checkpoint is also a fixture any user of the reference can read with
``std::vector<Particle>`` + fread.  The reference itself has no checkpointing (its long runs,
config.py:24-26, are 3 M substeps)."""
import json
import struct

import numpy as np

MAGIC = b"MPMCKPT1"


def save(engine, path, step=0, extra=None):
    """Synchronises, reads every particle back (upload order) and writes header + records."""
    p = engine.read()
    c = engine.cfg
    header = {"dim": engine.dim, "n_grid": c.n_grid, "dt": c.dt, "mass_p": c.mass_p, "vol_p": c.vol_p,
              "alpha": c.alpha, "gravity": list(c.gravity), "boundary": c.boundary, "n_particles": len(p),
              "record_bytes": 4 * p.shape[1], "step": step, "extra": extra or {},
              "materials": [[m.kind, m.E, m.nu, m.hardening, m.sig_lo, m.sig_hi]
                            for m in list(c.materials)[:c.n_materials]]}
    blob = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<q", len(blob)) + blob)
        f.write(np.ascontiguousarray(p, np.float32).tobytes())
    return header


def load(path):
    """-> (header dict, (n, 14|26) float32 records)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s is not an MPM checkpoint" % path)
        (n,) = struct.unpack("<q", f.read(8))
        header = json.loads(f.read(n).decode())
        words = header["record_bytes"] // 4
        p = np.frombuffer(f.read(), np.float32).reshape(-1, words).copy()
    if len(p) != header["n_particles"]:
        raise ValueError("truncated checkpoint: %d of %d records" % (len(p), header["n_particles"]))
    return header, p


def restore(engine_cls, path, **overrides):
    """New engine with the checkpoint's configuration and particles."""
    h, p = load(path)
    kw = dict(dim=h["dim"], n_grid=h["n_grid"], capacity=max(len(p), 1), dt=h["dt"], vol_p=h["vol_p"], alpha=h["alpha"],
              gravity=h["gravity"], materials=[tuple(m) for m in h["materials"]], mass_p=h["mass_p"],
              boundary=h["boundary"])
    kw.update(overrides)
    e = engine_cls(**kw)
    e.upload(p)
    return e, h
