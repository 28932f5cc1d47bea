"""Time-dependent vector algebra on ``(num_steps, 3)`` arrays.

API mirror of the reference's ``tomosipo/vector_calc.py`` (names, shapes,
broadcast rules, exception types), NumPy-2 clean.  Host-only fp64 math: it
feeds the geometry conversion (``Operator.py:11-60``) and never touches the
GPU path.
"""
from contextlib import contextmanager

import numpy as np

_EPS = 1e-8


# ------------------------------------------------------------- construction --
def to_vec(x):
    """(N, 3) array; homogeneous (N, 4) input is de-homogenised."""
    x = np.asarray(x, dtype=np.float64)
    original = x.shape
    x = np.atleast_2d(x)
    if x.ndim == 2 and x.shape[1] == 3:
        return x
    if x.ndim == 2 and x.shape[1] == 4:
        w = x[:, 3:4].copy()
        w[np.abs(w) < _EPS] = 1.0  # direction vectors carry w == 0
        return x[:, :3] / w
    raise ValueError(f"Shape {original} cannot be converted to vector. ")


def to_scalar(x):
    """(N, 1) array from a scalar or 1-D array."""
    x = np.asarray(x, dtype=np.float64)
    original = x.shape
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, None]
    if x.ndim == 2 and x.shape[1] == 1:
        return x
    raise ValueError(f"Shape {original} cannot be converted to scalar. ")


def to_homogeneous(x, s):
    x, _ = _broadcastv(x, x)
    if x.ndim == 2 and x.shape[1] == 4:
        return x
    if x.ndim == 2 and x.shape[1] == 3:
        return np.concatenate([x, np.full((x.shape[0], 1), float(s))], axis=1)
    raise ValueError(
        "Could not convert array to homogeneous coordinates. "
        f"Expected shape (3,) or (n_rows, 3) but got {x.shape}"
    )


def to_homogeneous_vec(x):
    return to_homogeneous(x, 0)


def to_homogeneous_point(x):
    return to_homogeneous(x, 1)


# -------------------------------------------------------------- broadcasting --
def broadcast_lengths(len_a, len_b):
    if len_a == 1:
        return len_b
    if len_b == 1 or len_a == len_b:
        return len_a
    raise ValueError("Operands could not be broadcast together.")


def _broadcastv(x, y):
    x, y = np.asarray(x), np.asarray(y)
    sx, sy = x.shape, y.shape
    x, y = np.atleast_2d(x), np.atleast_2d(y)
    if x.ndim == 2 and y.ndim == 2:
        if x.shape[0] == 1:
            x = np.broadcast_to(x, (y.shape[0], x.shape[1]))
        elif y.shape[0] == 1:
            y = np.broadcast_to(y, (x.shape[0], y.shape[1]))
    if x.ndim != 2 or y.ndim != 2 or x.shape != y.shape:
        raise ValueError(f"Arguments of shape {sx} and {sy} could not be broadcast together.")
    return x, y


def _atleast_3d_front(m):
    m = np.asarray(m)
    while m.ndim < 3:
        m = m[None]
    return m


def _broadcastmv(M, x):
    M, x = np.asarray(M), np.asarray(x)
    sm, sx = M.shape, x.shape
    M, x = _atleast_3d_front(M), np.atleast_2d(x)
    if M.ndim == 3 and x.ndim == 2:
        if x.shape[0] == 1:
            x = np.broadcast_to(x, (M.shape[0], x.shape[1]))
        if M.shape[0] == 1:
            M = np.broadcast_to(M, (x.shape[0],) + M.shape[1:])
    if M.ndim != 3 or x.ndim != 2 or M.shape[2] != x.shape[1] or M.shape[0] != x.shape[0]:
        raise ValueError(f"Arguments of shape {sm} and {sx} could not be broadcast together.")
    return M, x


def _broadcastmm(M1, M2):
    M1, M2 = np.asarray(M1), np.asarray(M2)
    s1, s2 = M1.shape, M2.shape
    M1, M2 = _atleast_3d_front(M1), _atleast_3d_front(M2)
    if M1.ndim == 3 and M2.ndim == 3:
        if M1.shape[0] == 1:
            M1 = np.broadcast_to(M1, (M2.shape[0],) + M1.shape[1:])
        elif M2.shape[0] == 1:
            M2 = np.broadcast_to(M2, (M1.shape[0],) + M2.shape[1:])
    if M1.ndim != 3 or M2.ndim != 3 or M1.shape[2] != M2.shape[1] or M1.shape[0] != M2.shape[0]:
        raise ValueError(f"Arguments of shape {s1} and {s2} could not be broadcast together.")
    return M1, M2


# ---------------------------------------------------------------- operations --
def cross_product(x, y):
    x, y = _broadcastv(x, y)
    return np.stack(
        [
            x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1],
            x[:, 2] * y[:, 0] - x[:, 0] * y[:, 2],
            x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0],
        ],
        axis=-1,
    )


def dot(x, y):
    x, y = _broadcastv(x, y)
    return np.sum(x * y, axis=1)


def squared_norm(x):
    return dot(x, x)


def norm(x):
    return np.sqrt(squared_norm(x))


def intersect(v_origin, v_direction, plane_origin, plane_normal):
    """Intersection of the lines ``o + t d`` with planes; NaN when parallel."""
    o, d = to_vec(v_origin), to_vec(v_direction)
    p0, n = to_vec(plane_origin), to_vec(plane_normal)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (np.sum(n * (p0 - o), axis=1) / np.sum(n * d, axis=1))[:, None]
        t[np.isinf(t)] = np.nan
        return o + t * d


def orthogonal_basis_from_axis(axis):
    """Left-handed orthonormal basis whose first vector is ``axis``/|axis|.

    Closed form of the reference (``tomosipo/vector_calc.py:187-262``); the
    degenerate case (axis parallel to the third coordinate) uses a fixed
    completion.
    """
    a = to_homogeneous_vec(axis)
    w0, w1, w2 = a[:, 0], a[:, 1], a[:, 2]
    n2 = w0 * w0 + w1 * w1 + w2 * w2
    r2 = w0 * w0 + w1 * w1
    degenerate = (np.abs(w0) < _EPS) & (np.abs(w1) < _EPS)
    with np.errstate(divide="ignore", invalid="ignore"):
        n, r = np.sqrt(n2), np.sqrt(r2)
        e0 = np.stack([w0 / n, w1 / n, w2 / n], axis=1)
        e1 = np.stack([w0 * w2 / (r * n), w1 * w2 / (r * n), -r / n], axis=1)
        e2 = np.stack([-w1 / r, w0 / r, np.zeros_like(w0)], axis=1)
        sgn = w2 / np.abs(w2)
    z, o = np.zeros_like(w0), np.ones_like(w0)
    d0 = np.stack([z, z, sgn], axis=1)
    d1 = np.stack([-o, z, z], axis=1)
    d2 = np.stack([z, -o, z], axis=1)
    pick = degenerate[:, None]
    return (
        np.where(pick, np.nan_to_num(d0), np.nan_to_num(e0)),
        np.where(pick, d1, np.nan_to_num(e1)),
        np.where(pick, d2, np.nan_to_num(e2)),
    )


def matrix_transform(M, x):
    """Row-wise ``M[i] @ x[i]`` for stacked 4x4 matrices and homogeneous vectors."""
    M, x = _broadcastmv(M, x)
    return np.einsum("nij,nj->ni", M, x)


def matrix_matrix_transform(M1, M2):
    M1, M2 = _broadcastmm(M1, M2)
    return np.matmul(M1, M2)


def invert_transformation_matrix(M):
    M, _ = _broadcastmm(M, M)
    try:
        return np.linalg.inv(M)
    except np.linalg.LinAlgError:
        raise ValueError(f"Inverting matrix failed, {M}")


# ------------------------------------------------------------------ utilities --
@contextmanager
def ignore_divide_by_zero():
    with np.errstate(divide="ignore", invalid="ignore"):
        yield


def check_same_shapes(*args):
    shapes = [x.shape for x in args]
    if min(shapes) != max(shapes):
        raise ValueError(f"Not all arguments are the same shape. Got: {shapes}")
