"""Pure-Python restatement of the three ASTRA dict helpers tomosipo calls.

The reference calls ``astra.create_vol_geom`` (``geometry/volume.py:301``),
``astra.create_proj_geom('cone', ...)`` (``geometry/cone.py:257-265``) and
``astra.geom_2vec`` (``geometry/cone.py:252``, ``geometry/parallel.py:195``).
They are pure dict builders inside the un-vendored ASTRA toolbox
(astra-toolbox >= 2.0); their semantics are pinned by the reference's own
documentation (``doc/topics/geometries.rst:366-410``) and tests
(``tests/geometry/test_cone_vec.py:143-201``) and restated here so that the
geometry layer needs no ASTRA install.  All vectors are in ASTRA's (x, y, z)
order.
"""
import numpy as np


def create_vol_geom(rows_y, cols_x, slices_z, min_x, max_x, min_y, max_y, min_z, max_z):
    """``astra.create_vol_geom(Y, X, Z, minx, maxx, miny, maxy, minz, maxz)``."""
    return {
        "GridRowCount": int(rows_y),
        "GridColCount": int(cols_x),
        "GridSliceCount": int(slices_z),
        "option": {
            "WindowMinX": float(min_x),
            "WindowMaxX": float(max_x),
            "WindowMinY": float(min_y),
            "WindowMaxY": float(max_y),
            "WindowMinZ": float(min_z),
            "WindowMaxZ": float(max_z),
        },
    }


def create_proj_geom_cone(spacing_x, spacing_y, det_rows, det_cols, angles, src_origin, origin_det):
    """``astra.create_proj_geom('cone', ...)``."""
    return {
        "type": "cone",
        "DetectorSpacingX": float(spacing_x),
        "DetectorSpacingY": float(spacing_y),
        "DetectorRowCount": int(det_rows),
        "DetectorColCount": int(det_cols),
        "ProjectionAngles": np.asarray(angles, dtype=np.float64),
        "DistanceOriginSource": float(src_origin),
        "DistanceOriginDetector": float(origin_det),
    }


def geom_2vec(pg):
    """``astra.geom_2vec`` for the two circular 3D geometries.

    parallel3d:  ray = ( sin t, -cos t, 0),  centre = 0,
                 u = (cos t, sin t, 0) * sx,  v = (0, 0, sy)
    cone:        src = ( sin t, -cos t, 0) * SOD,  centre = (-sin t, cos t, 0) * ODD,
                 u, v as above.
    """
    kind = pg["type"]
    t = np.asarray(pg["ProjectionAngles"], dtype=np.float64)
    sx, sy = float(pg["DetectorSpacingX"]), float(pg["DetectorSpacingY"])
    vectors = np.zeros((len(t), 12))
    sin, cos = np.sin(t), np.cos(t)
    if kind == "cone":
        sod, odd = float(pg["DistanceOriginSource"]), float(pg["DistanceOriginDetector"])
        vectors[:, 0], vectors[:, 1] = sin * sod, -cos * sod
        vectors[:, 3], vectors[:, 4] = -sin * odd, cos * odd
        out_kind = "cone_vec"
    elif kind == "parallel3d":
        vectors[:, 0], vectors[:, 1] = sin, -cos
        out_kind = "parallel3d_vec"
    else:
        raise ValueError(f"geom_2vec: unsupported geometry type {kind!r}")
    vectors[:, 6], vectors[:, 7] = cos * sx, sin * sx
    vectors[:, 11] = sy
    return {
        "type": out_kind,
        "DetectorRowCount": pg["DetectorRowCount"],
        "DetectorColCount": pg["DetectorColCount"],
        "Vectors": vectors,
    }
