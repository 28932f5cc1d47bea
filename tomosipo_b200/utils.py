"""Small helpers: NumPy print options for reprs and 1-D interval slicing.

API mirror of the reference's ``tomosipo/utils.py``.  ``slice_interval`` is what
geometry indexing (``vg[:1]``, ``pg.to_vec()[:, :1, :]`` of
``notebooks/learned_pd.py:55``) is built on.
"""
from numbers import Integral

import numpy as np


def print_options():
    """Context manager pinning the NumPy print options used by geometry reprs."""
    return np.printoptions(
        edgeitems=3, threshold=1000, floatmode="maxprec", precision=8, suppress=False, linewidth=71,
        nanstr="nan", infstr="inf", sign="-", formatter=None, legacy=False,
    )


def up_slice(key):
    """Turn an integer index into the length-one slice selecting it."""
    if isinstance(key, Integral):
        return slice(key, None) if key == -1 else slice(key, key + 1)
    return key


def slice_interval(left, right, length, key):
    """Slice the interval ``[left, right]`` that is divided into ``length`` cells.

    Returns ``(new_left, new_right, new_length, new_cell_size)``.  A step > 1
    bins cells: the new cell size is ``step`` times the old one and the new
    interval is centred on the selected cells (so it may stick out of the
    original interval), matching detector / voxel binning.

    >>> slice_interval(0, 4, 4, slice(0, 4, 2))
    (-0.5, 3.5, 2, 2.0)
    >>> slice_interval(0, 4, 4, slice(1, 4, 2))
    (0.5, 4.5, 2, 2.0)
    """
    cell = 1 if length == 0 else (right - left) / length
    start, stop, step = up_slice(key).indices(length)
    count = max(0, -(-(stop - start) // step))
    last = max(start, start + (count - 1) * step + 1)  # one past the last selected cell
    grow = 0.5 * cell * (step - 1)
    return (left + start * cell - grow, left + last * cell + grow, count, cell * step)
