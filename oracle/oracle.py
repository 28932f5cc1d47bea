"""ctypes front-end of the CPU oracle (``oracle/tsp_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The
product package ``tomosipo_b200`` never imports this module.

The geometry helpers at the bottom restate ``astra.geom_2vec`` for the two
circular geometries (formulas: SURVEY.md section 3.5, pinned by
``doc/topics/geometries.rst:366-410`` of the reference) so that the oracle
can be driven without the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtsp_oracle.so")

CONE_VEC = 0
PARALLEL_VEC = 1


class Geometry(ctypes.Structure):
    """Mirror of ``oracle_geometry`` (and of ``tsp_geometry``)."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("nz", ctypes.c_int32),
        ("win_min", ctypes.c_double * 3),
        ("win_max", ctypes.c_double * 3),
        ("det_rows", ctypes.c_int32),
        ("det_cols", ctypes.c_int32),
        ("n_angles", ctypes.c_int32),
        ("vectors", ctypes.POINTER(ctypes.c_double)),
        ("voxel_supersampling", ctypes.c_int32),
        ("detector_supersampling", ctypes.c_int32),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        gp = ctypes.POINTER(Geometry)
        for sfx, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
            p = ctypes.POINTER(ct)
            fp = getattr(_lib, f"oracle_fp_{sfx}")
            fp.argtypes = [gp, p, p, ctypes.c_int]
            fp.restype = ctypes.c_int
            bp = getattr(_lib, f"oracle_bp_{sfx}")
            bp.argtypes = [gp, p, p, ctypes.c_int]
            bp.restype = ctypes.c_int
        _lib.oracle_marching_axes_f64.argtypes = [gp, ctypes.POINTER(ctypes.c_int32)]
        i32p, f32p, f64p = (ctypes.POINTER(t) for t in (ctypes.c_int32, ctypes.c_float, ctypes.c_double))
        _lib.oracle_fp_angles_mixed.argtypes = [gp, f32p, i32p, ctypes.c_int, i32p, ctypes.c_int, f64p]
        _lib.oracle_bp_window_mixed.argtypes = [gp, f32p, i32p, i32p, f64p]
        _lib.oracle_bp_map.argtypes = [gp, ctypes.c_int, f64p, f64p]
        for sfx in ("f64", "f32"):
            getattr(_lib, f"oracle_set_weight_bits_{sfx}").argtypes = [ctypes.c_int]
    return _lib


class OracleProjector:
    """Holds one geometry; ``fp``/``bp`` run the C restatement.

    Parameters mirror what ``astra.create_projector('cuda3d', pg, vg, opts)``
    receives from tomosipo (``tomosipo/astra.py:90-98``): the volume window
    (x, y, z order), the detector shape and the ASTRA 12-column vectors.
    """

    def __init__(self, kind, vol_shape_zyx, win_min_xyz, win_max_xyz, det_shape_vu, vectors,
                 voxel_supersampling=1, detector_supersampling=1):
        self.vectors = np.ascontiguousarray(vectors, dtype=np.float64).reshape(-1, 12)
        nz, ny, nx = (int(s) for s in vol_shape_zyx)
        g = Geometry()
        g.kind = kind
        g.nx, g.ny, g.nz = nx, ny, nz
        for i in range(3):
            g.win_min[i] = float(win_min_xyz[i])
            g.win_max[i] = float(win_max_xyz[i])
        g.det_rows, g.det_cols = int(det_shape_vu[0]), int(det_shape_vu[1])
        g.n_angles = self.vectors.shape[0]
        g.vectors = self.vectors.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        g.voxel_supersampling = int(voxel_supersampling)
        g.detector_supersampling = int(detector_supersampling)
        self.g = g
        self.vol_shape = (nz, ny, nx)
        self.proj_shape = (g.det_rows, g.n_angles, g.det_cols)

    def _run(self, name, dtype, vol, proj, additive):
        ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
        sfx = "f64" if dtype == np.float64 else "f32"
        fn = getattr(lib(), f"oracle_{name}_{sfx}")
        p = ctypes.POINTER(ct)
        rc = fn(ctypes.byref(self.g), vol.ctypes.data_as(p), proj.ctypes.data_as(p), int(additive))
        if rc != 0:
            raise RuntimeError(f"oracle_{name}_{sfx} failed with {rc}")

    def fp(self, vol, out=None, additive=False, dtype=np.float64):
        vol = np.ascontiguousarray(vol, dtype=dtype)
        assert vol.shape == self.vol_shape, (vol.shape, self.vol_shape)
        if out is None:
            out = np.zeros(self.proj_shape, dtype=dtype)
        assert out.dtype == dtype and out.flags.c_contiguous and out.shape == self.proj_shape
        self._run("fp", dtype, vol, out, additive)
        return out

    def bp(self, proj, out=None, additive=False, dtype=np.float64):
        proj = np.ascontiguousarray(proj, dtype=dtype)
        assert proj.shape == self.proj_shape, (proj.shape, self.proj_shape)
        if out is None:
            out = np.zeros(self.vol_shape, dtype=dtype)
        assert out.dtype == dtype and out.flags.c_contiguous and out.shape == self.vol_shape
        self._run("bp", dtype, out, proj, additive)
        return out

    # ---- spot checks at benchmark sizes: float32 storage (what the GPU saw), fp64 arithmetic
    def fp_angles(self, vol_f32, angle_list, row_list=None):
        """``(A vol)[row_list, angle_list, :]`` (all rows by default) in fp64 from a float32 volume."""
        vol = np.ascontiguousarray(vol_f32, dtype=np.float32)
        assert vol.shape == self.vol_shape
        idx = np.ascontiguousarray(angle_list, dtype=np.int32)
        rows = np.ascontiguousarray(np.arange(self.g.det_rows) if row_list is None else row_list, dtype=np.int32)
        assert idx.min() >= 0 and idx.max() < self.g.n_angles and rows.min() >= 0 and rows.max() < self.g.det_rows
        out = np.zeros((len(rows), len(idx), self.g.det_cols))
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = lib().oracle_fp_angles_mixed(ctypes.byref(self.g), vol.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                          idx.ctypes.data_as(i32p), len(idx), rows.ctypes.data_as(i32p), len(rows),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert rc == 0
        return out

    def bp_window(self, proj_f32, z, y, x):
        """``(A^T proj)[z0:z1, y0:y1, x0:x1]`` in fp64 from a float32 projection stack; z, y, x are (lo, hi)."""
        proj = np.ascontiguousarray(proj_f32, dtype=np.float32)
        assert proj.shape == self.proj_shape
        lo = np.array([x[0], y[0], z[0]], dtype=np.int32)
        hi = np.array([x[1], y[1], z[1]], dtype=np.int32)
        assert (lo >= 0).all() and (hi > lo).all() and (hi <= np.array(self.vol_shape[::-1])).all()
        out = np.zeros((z[1] - z[0], y[1] - y[0], x[1] - x[0]))
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = lib().oracle_bp_window_mixed(ctypes.byref(self.g), proj.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                          lo.ctypes.data_as(i32p), hi.ctypes.data_as(i32p),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert rc == 0
        return out

    def bp_map(self, angle, xyz):
        """(U, V, weight) of a world-frame point (x, y, z) on the detector of ``angle``: the backprojector's map."""
        p = np.ascontiguousarray(xyz, dtype=np.float64)
        out = np.zeros(3)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib().oracle_bp_map(ctypes.byref(self.g), int(angle), p.ctypes.data_as(f64p), out.ctypes.data_as(f64p))
        return out

    def marching_axes(self):
        axes = np.zeros(self.g.n_angles, dtype=np.int32)
        lib().oracle_marching_axes_f64(ctypes.byref(self.g), axes.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        return axes


class astra_texture_weights:
    """Context manager: interpolation weights rounded to ``bits`` fractional bits (8 = CUDA's texture unit,
    which is what ASTRA interpolates with); see ``oracle_set_weight_bits`` in tsp_oracle.c."""

    def __init__(self, bits=8):
        self.bits = bits

    def __enter__(self):
        for sfx in ("f64", "f32"):
            getattr(lib(), f"oracle_set_weight_bits_{sfx}")(self.bits)

    def __exit__(self, *exc):
        for sfx in ("f64", "f32"):
            getattr(lib(), f"oracle_set_weight_bits_{sfx}")(0)


# --------------------------------------------------------------------------
# Circular geometries -> ASTRA vectors (restated astra.geom_2vec, x,y,z order)
# --------------------------------------------------------------------------

def parallel_vectors(angles, det_spacing_x, det_spacing_y):
    """``parallel3d`` -> ``parallel3d_vec`` rows [ray | centre | u | v]."""
    t = np.asarray(angles, dtype=np.float64)
    out = np.zeros((len(t), 12))
    out[:, 0] = np.sin(t)
    out[:, 1] = -np.cos(t)
    out[:, 6] = np.cos(t) * det_spacing_x
    out[:, 7] = np.sin(t) * det_spacing_x
    out[:, 11] = det_spacing_y
    return out


def cone_vectors(angles, det_spacing_x, det_spacing_y, src_orig, orig_det):
    """``cone`` -> ``cone_vec`` rows [src | centre | u | v]."""
    t = np.asarray(angles, dtype=np.float64)
    out = np.zeros((len(t), 12))
    out[:, 0] = np.sin(t) * src_orig
    out[:, 1] = -np.cos(t) * src_orig
    out[:, 3] = -np.sin(t) * orig_det
    out[:, 4] = np.cos(t) * orig_det
    out[:, 6] = np.cos(t) * det_spacing_x
    out[:, 7] = np.sin(t) * det_spacing_x
    out[:, 11] = det_spacing_y
    return out


def hollow_box(shape):
    """The reference's phantom (``tomosipo/phantom.py:4-25``) as an array."""
    shape = np.array((shape,) * 3 if np.isscalar(shape) else shape)
    x = np.zeros(tuple(shape), dtype=np.float32)
    a, b = shape * 20 // 100, shape * 40 // 100
    x[tuple(slice(i, n - i) for i, n in zip(a, shape))] = 1.0
    x[tuple(slice(i, n - i) for i, n in zip(b, shape))] = 0.0
    return x
