"""Seeded synthetic scenes in the reference's AoS record layout.

Records are float32 rows: 2D ``x[2] v[2] F[4] C[4] Jp c`` (14 words, 56 B -- the reference's
``Particle``, cpp_validation/mls-mpm88-explained.cpp:28-42) and the 3D lift ``x[3] v[3] F[9] C[9]
Jp c`` (26 words); the last word holds the int32 material id.  All scenes start at rest with
F = I, C = 0, Jp = 1 like the reference constructor (:35-41).

Scaling rule for n_grid != 80 (the reference's mass_p = vol_p = 1 are not physical, :17-18): keep
vol_p*inv_dx^2 and dt*inv_dx fixed, i.e. vol_p = (80/n)^2 and dt = 1e-4*80/n (dt capped at 1e-4 for n < 80,
where the elastic wave speed, not advection, limits the step).
"""
import numpy as np

FLUID, JELLY, SNOW, SHIPPED = 0, 1, 2, 3


def record_words(dim):
    return 2 * dim + 2 * dim * dim + 2


def scaled_constants(n_grid, dim=2):
    """(dt, vol_p) that keep the shipped scene's effective stiffness and CFL at another resolution."""
    s = 80.0 / n_grid
    return 1e-4 * min(1.0, s), s ** 2


def make_records(x, mat, dim):
    n = x.shape[0]
    w = record_words(dim)
    p = np.zeros((n, w), np.float32)
    p[:, 0:dim] = x
    eye = np.eye(dim, dtype=np.float32).reshape(-1)
    p[:, 2 * dim:2 * dim + dim * dim] = eye
    p[:, 2 * dim + 2 * dim * dim] = 1.0
    p[:, w - 1] = np.asarray(mat, np.int32).view(np.float32) if np.ndim(mat) else np.full(n, mat, np.int32).view(np.float32)
    return p


def _jittered_box(lo, hi, n_grid, per_side, rng, dim):
    """Jittered lattice: per_side^dim particles in every cell of the box [lo, hi) (cell-aligned)."""
    dx = 1.0 / n_grid
    c_lo = [int(np.ceil(l * n_grid - 1e-9)) for l in lo]
    c_hi = [int(np.floor(h * n_grid + 1e-9)) for h in hi]
    axes = [np.arange(c_lo[k] * per_side, c_hi[k] * per_side, dtype=np.float64) for k in range(dim)]
    mesh = np.meshgrid(*axes, indexing="ij")
    sub = np.stack([m.reshape(-1) for m in mesh], 1)
    jit = rng.uniform(0.1, 0.9, sub.shape)
    return ((sub + jit) * (dx / per_side)).astype(np.float32)


def box_count(lo, hi, n_grid, per_side):
    """Number of particles _jittered_box generates for the box [lo, hi): its cells x per_side^dim."""
    n = 1
    for l, h in zip(lo, hi):
        n *= max(0, int(np.floor(h * n_grid + 1e-9)) - int(np.ceil(l * n_grid - 1e-9))) * per_side
    return n


def three_blocks_2d(n_grid=512, per_side=4, seed=1, side=0.28):
    """BASELINE config 2: three squares (fluid / jelly / snow), per_side^2 particles per cell."""
    rng = np.random.RandomState(seed)
    xs, ms = [], []
    for (cx, cy), m in zip(((0.55, 0.20), (0.45, 0.49), (0.55, 0.78)), (FLUID, JELLY, SNOW)):
        x = _jittered_box((cx - side / 2, cy - side / 2), (cx + side / 2, cy + side / 2), n_grid, per_side, rng, 2)
        xs.append(x)
        ms.append(np.full(len(x), m, np.int32))
    return make_records(np.concatenate(xs), np.concatenate(ms), 2)


def dam_break_2d(n_grid=2048, per_side=3, seed=2, width=0.45, height=0.90, mat=FLUID):
    """BASELINE config 3: a fluid column x in [0.05, 0.05+width], y in [0.05, 0.05+height]."""
    rng = np.random.RandomState(seed)
    x = _jittered_box((0.05, 0.05), (0.05 + width, 0.05 + height), n_grid, per_side, rng, 2)
    return make_records(x, mat, 2)


def swirl_velocity(x, x_range=(0.05, 0.95), y_range=(0.05, 0.50), speed=3.0, kx=8, ky=4):
    """Divergence-free cellular flow v = curl(psi), psi = A sin(kx pi X) sin(ky pi Y) over the pool box: kx x ky
    counter-rotating vortices, zero normal velocity at the side walls, floor and the initial free surface.
    `speed` is the peak of each component; 3 puts the per-substep motion (speed*dt/dx = 0.024 cells with
    scaled_constants) in the range of the reference's own scene (rms speed 5.7 at substep 1000 of the shipped
    run, :214, i.e. 0.046 cells per substep)."""
    X = (x[:, 0].astype(np.float64) - x_range[0]) / (x_range[1] - x_range[0])
    Y = (x[:, 1].astype(np.float64) - y_range[0]) / (y_range[1] - y_range[0])
    v = np.empty((len(x), 2), np.float32)
    v[:, 0] = speed * np.sin(kx * np.pi * X) * np.cos(ky * np.pi * Y)
    v[:, 1] = -speed * np.cos(kx * np.pi * X) * np.sin(ky * np.pi * Y)
    return v


def slab_fill_2d(n_grid=8192, per_side=3, seed=3, x_range=(0.05, 0.95), y_range=(0.05, 0.50), bands=True,
                 columns=None, out=None, strip=256, swirl=0.0):
    """BASELINE config 4: a wide pool, three material bands along x.  `swirl` > 0 seeds the cellular flow of
    swirl_velocity() (peak speed = swirl) so the scene is in motion from the first substep.

    `columns=(c_lo, c_hi)` restricts generation to the cell columns of one x-slab (each column strip has
    its own seeded stream, so any partition yields the same global scene); `out` (an (n,14) float32
    array, e.g. pinned memory) is filled strip by strip so no second full-size temporary exists.
    Returns the records (a view of `out` when given)."""
    c0 = int(np.ceil(x_range[0] * n_grid - 1e-9))
    c1 = int(np.floor(x_range[1] * n_grid + 1e-9))
    if columns is not None:
        c0, c1 = max(c0, columns[0]), min(c1, columns[1])
    r0 = int(np.ceil(y_range[0] * n_grid - 1e-9))
    r1 = int(np.floor(y_range[1] * n_grid + 1e-9))
    n = max(0, c1 - c0) * (r1 - r0) * per_side * per_side
    if out is None:
        out = np.empty((n, 14), np.float32)
    assert out.shape[0] >= n and out.shape[1] == 14
    pos = 0
    for s0 in range((c0 // strip) * strip, c1, strip):
        a, b = max(s0, c0), min(s0 + strip, c1)
        if a >= b:
            continue
        rng = np.random.RandomState((seed * 1000003 + s0) % (2 ** 31))
        full = _jittered_box((s0 / n_grid, y_range[0]), ((s0 + strip) / n_grid, y_range[1]), n_grid, per_side, rng, 2)
        keep = (full[:, 0] >= np.float32(a / n_grid)) & (full[:, 0] < np.float32(b / n_grid)) if (a, b) != (s0, s0 + strip) \
            else slice(None)
        x = full[keep]
        if bands:
            t = (x[:, 0] - x_range[0]) / (x_range[1] - x_range[0])
            mat = np.clip((t * 3).astype(np.int32), 0, 2)
        else:
            mat = FLUID
        out[pos:pos + len(x)] = make_records(x, mat, 2)
        if swirl:
            out[pos:pos + len(x), 2:4] = swirl_velocity(x, x_range, y_range, swirl)
        pos += len(x)
    return out[:pos]


def slab_fill_2d_count(n_grid, per_side=3, x_range=(0.05, 0.95), y_range=(0.05, 0.50), columns=None):
    """Upper bound of the number of particles slab_fill_2d generates (for sizing buffers)."""
    c0 = int(np.ceil(x_range[0] * n_grid - 1e-9))
    c1 = int(np.floor(x_range[1] * n_grid + 1e-9))
    if columns is not None:
        c0, c1 = max(c0, columns[0]), min(c1, columns[1])
    r0 = int(np.ceil(y_range[0] * n_grid - 1e-9))
    r1 = int(np.floor(y_range[1] * n_grid + 1e-9))
    return max(0, c1 - c0) * (r1 - r0) * per_side * per_side


def collapse_3d(n_grid=256, per_side=2, seed=4, y_top=0.35, xz=(0.05, 0.95), columns=None, strip=16):
    """BASELINE config 5: a 3D slab x,z in [0.05,0.95], y in [0.05,y_top], three bands along x.

    `columns=(c_lo, c_hi)` generates only the cell columns (x) of one x-slab, strip by strip with one seeded
    stream per strip of `strip` columns, so that any partition yields the same global scene (the default,
    columns=None, draws the whole box from a single stream)."""
    def records(x):
        t = (x[:, 0] - xz[0]) / (xz[1] - xz[0])
        return make_records(x, np.clip((t * 3).astype(np.int32), 0, 2), 3)

    if columns is None:
        rng = np.random.RandomState(seed)
        return records(_jittered_box((xz[0], 0.05, xz[0]), (xz[1], y_top, xz[1]), n_grid, per_side, rng, 3))
    c0 = max(int(np.ceil(xz[0] * n_grid - 1e-9)), columns[0])
    c1 = min(int(np.floor(xz[1] * n_grid + 1e-9)), columns[1])
    parts = []
    for s0 in range((c0 // strip) * strip, c1, strip):
        a, b = max(s0, c0), min(s0 + strip, c1)
        if a >= b:
            continue
        rng = np.random.RandomState((seed * 1000003 + s0) % (2 ** 31))
        full = _jittered_box((max(s0 / n_grid, xz[0]), 0.05, xz[0]), (min((s0 + strip) / n_grid, xz[1]), y_top, xz[1]),
                             n_grid, per_side, rng, 3)
        keep = (full[:, 0] >= np.float32(a / n_grid)) & (full[:, 0] < np.float32(b / n_grid))
        parts.append(records(full[keep]))
    return np.concatenate(parts) if parts else np.zeros((0, 26), np.float32)


def commented_three_blocks(seed=5):
    """BASELINE config 1 wording: the seeding left commented in the reference (:182-188) --
    3 x 1000 uniformly random particles, half-width 0.08, fluid / jelly / snow."""
    rng = np.random.RandomState(seed)
    xs, ms = [], []
    for (cx, cy), m in zip(((0.55, 0.45), (0.45, 0.65), (0.55, 0.85)), (FLUID, JELLY, SNOW)):
        x = (rng.uniform(-1, 1, (1000, 2)) * 0.08 + np.array([cx, cy])).astype(np.float32)
        xs.append(x)
        ms.append(np.full(1000, m, np.int32))
    return make_records(np.concatenate(xs), np.concatenate(ms), 2)


def jelly_drop(seed=11, n=3000):
    """A single elastic block dropped onto the floor (n_grid 80): well-conditioned over 1000 substeps
    (two runs of the reference differing only in summation order agree to ~1e-6 in bulk)."""
    rng = np.random.RandomState(seed)
    x = (rng.uniform(-1, 1, (n, 2)) * 0.1 + np.array([0.5, 0.3])).astype(np.float32)
    return make_records(x, JELLY, 2)


def fluid_pool(seed=12, n_grid=80, per_side=3):
    """A fluid pool settling under gravity (well-conditioned, like jelly_drop)."""
    rng = np.random.RandomState(seed)
    return make_records(_jittered_box((0.1, 0.06), (0.9, 0.3), n_grid, per_side, rng, 2), FLUID, 2)


def bulk(p, dim):
    """Bulk diagnostics of north_star's 1000-substep criterion: centre of mass, momentum, KE."""
    x = p[:, 0:dim].astype(np.float64)
    v = p[:, dim:2 * dim].astype(np.float64)
    return dict(com=x.mean(0), mom=v.sum(0), ke=0.5 * (v ** 2).sum())
