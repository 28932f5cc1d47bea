"""Phantoms used by the reference's benchmarks (``tomosipo/phantom.py``)."""
import numpy as np


def hollow_box(vd):
    """Fill a volume dataset with a box (outer 20 % margin) that is hollow (inner 40 % margin)."""
    shape = np.array(vd.data.shape)
    outer = tuple(slice(a, n - a) for a, n in zip(shape * 20 // 100, shape))
    inner = tuple(slice(a, n - a) for a, n in zip(shape * 40 // 100, shape))
    vd.data[:] = 0.0
    vd.data[outer] = 1.0
    vd.data[inner] = 0.0
    return vd
