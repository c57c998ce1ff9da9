import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def fields(p, dim):
    """Split AoS records into the named state fields (reference Particle, :28-42)."""
    d, dd = dim, dim * dim
    return dict(x=p[:, 0:d], v=p[:, d:2 * d], F=p[:, 2 * d:2 * d + dd], C=p[:, 2 * d + dd:2 * d + 2 * dd],
                Jp=p[:, 2 * d + 2 * dd])


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)
