#!/usr/bin/env python
"""Generate golden geometry vectors by running the *reference* tomosipo code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference cannot be imported as is: it needs the ASTRA toolbox (absent,
un-installable offline) and uses ``np.array(..., copy=False)``, which NumPy 2
rejects.  This script therefore
  * installs a stub ``astra`` module whose three dict helpers
    (``create_vol_geom``, ``create_proj_geom``, ``geom_2vec``) are restated from
    the ASTRA documentation -- everything else in the conversion path
    (``to_astra`` of the vector geometries, transforms, slicing,
    ``to_astra_compatible_operator_geometry``, ``project_point``) is the
    reference's own, unmodified code;
  * wraps ``numpy.array`` so that ``copy=False`` means "copy if needed" (its
    NumPy 1 meaning).
The outputs are small fp64 arrays, stored in ``geometry_golden.npz``.
The ``geom_2vec`` restatement itself is pinned separately against the numeric
dump in the reference documentation (tests/test_geometry_kat.py).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    _array = np.array

    def array(obj, *a, **kw):
        if kw.get("copy", True) is False:
            kw["copy"] = None
        return _array(obj, *a, **kw)

    np.array = array

    astra = types.ModuleType("astra")
    astra.__version__ = "2.1.0"

    def create_vol_geom(y, x, z, minx, maxx, miny, maxy, minz, maxz):
        return {"GridRowCount": y, "GridColCount": x, "GridSliceCount": z,
                "option": {"WindowMinX": minx, "WindowMaxX": maxx, "WindowMinY": miny, "WindowMaxY": maxy,
                           "WindowMinZ": minz, "WindowMaxZ": maxz}}

    def create_proj_geom(kind, sx, sy, rows, cols, angles, sod, odd):
        assert kind == "cone"
        return {"type": "cone", "DetectorSpacingX": sx, "DetectorSpacingY": sy, "DetectorRowCount": rows,
                "DetectorColCount": cols, "ProjectionAngles": angles, "DistanceOriginSource": sod,
                "DistanceOriginDetector": odd}

    def geom_2vec(pg):
        t = np.asarray(pg["ProjectionAngles"], dtype=np.float64)
        v = np.zeros((len(t), 12))
        if pg["type"] == "cone":
            v[:, 0] = np.sin(t) * pg["DistanceOriginSource"]
            v[:, 1] = -np.cos(t) * pg["DistanceOriginSource"]
            v[:, 3] = -np.sin(t) * pg["DistanceOriginDetector"]
            v[:, 4] = np.cos(t) * pg["DistanceOriginDetector"]
            kind = "cone_vec"
        else:
            v[:, 0] = np.sin(t)
            v[:, 1] = -np.cos(t)
            kind = "parallel3d_vec"
        v[:, 6] = np.cos(t) * pg["DetectorSpacingX"]
        v[:, 7] = np.sin(t) * pg["DetectorSpacingX"]
        v[:, 11] = pg["DetectorSpacingY"]
        return {"type": kind, "DetectorRowCount": pg["DetectorRowCount"], "DetectorColCount": pg["DetectorColCount"],
                "Vectors": v}

    astra.create_vol_geom = create_vol_geom
    astra.create_proj_geom = create_proj_geom
    astra.geom_2vec = geom_2vec
    astra.create_projector = lambda *a, **k: 0
    astra.use_cuda = lambda: False
    exp = types.ModuleType("astra.experimental")
    exp.accumulate_FDK = exp.do_composite = exp.direct_FPBP3D = lambda *a, **k: None
    astra.experimental = exp
    d3 = types.ModuleType("astra.data3d")
    d3.link = lambda *a, **k: 0
    d3.delete = lambda *a, **k: None
    d3.GPULink = lambda *a: a
    astra.data3d = d3
    sys.modules["astra"] = astra
    sys.modules["astra.experimental"] = exp
    sys.modules["astra.data3d"] = d3


def cases(ts):
    """name -> (volume geometry, projection geometry); shared with tests/test_geometry_golden.py."""
    rng = np.random.default_rng(1234)
    T = (ts.rotate(pos=(0.1, -0.2, 0.3), axis=(1.0, 0.5, -0.2), angles=0.7)
         * ts.translate((0.3, -0.4, 0.5)) * ts.scale((1.0, 1.5, 0.75)))
    angles = rng.uniform(0, 2 * np.pi, size=7)
    out = {
        "readme_cone": (ts.volume(shape=128), ts.cone(size=np.sqrt(2), cone_angle=1 / 2, angles=100, shape=(128, 192))),
        "cfg3_cone": (ts.volume(shape=32, size=1),
                      ts.cone(angles=24, shape=(32, 48), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)),
        "parallel": (ts.volume(shape=(10, 12, 14), pos=(0.5, -1, 2), size=(1, 2, 3)),
                     ts.parallel(angles=angles, shape=(9, 11), size=(2.0, 3.5))),
        "slab": (ts.volume(shape=16, size=1)[:1],
                 ts.parallel(angles=12, shape=(16, 24), size=(1, 1.5)).to_vec()[:, :1, :]),
        "cone_vec_T": (ts.volume(shape=(8, 9, 10), size=(2, 3, 4)),
                       T * ts.cone(angles=angles, shape=(6, 7), size=(3, 4), src_orig_dist=5, src_det_dist=8).to_vec()),
        "par_vec_T": (ts.volume(shape=(8, 9, 10), size=(2, 3, 4)),
                      T * ts.parallel(angles=angles, shape=(6, 7), size=(3, 4)).to_vec()),
        "vol_vec_cone": (T * ts.volume(shape=(8, 9, 10), size=(2, 3, 4)).to_vec(),
                         ts.cone(angles=5, shape=(6, 7), size=(3, 4), src_orig_dist=5, src_det_dist=8)),
        "vol_vec_par": (T * ts.volume(shape=(8, 9, 10), pos=(1, 2, 3), size=(2, 3, 4)).to_vec(),
                        ts.parallel(angles=5, shape=(6, 7), size=(3, 4))),
        "binned": (ts.volume(shape=(8, 8, 12), size=(2, 2, 3))[::2, 1:7:3, 2:],
                   ts.cone(angles=6, shape=(8, 12), size=(2, 3), src_orig_dist=3, src_det_dist=5).to_vec()[1:5, ::2, 1:9:4]),
    }
    return out


def describe(ts, vg, pg):
    """Everything the backend receives for one operator, as plain arrays."""
    from_op = ts.Operator.to_astra_compatible_operator_geometry if hasattr(ts.Operator, "to_astra_compatible_operator_geometry") else None
    cvg, cpg = from_op(vg, pg)
    avg, apg = cvg.to_astra(), cpg.to_astra()
    if apg["type"] in ("cone", "parallel3d"):
        apg = cpg.to_vec().to_astra()
    o = avg["option"]
    window = np.array([o["WindowMinX"], o["WindowMaxX"], o["WindowMinY"], o["WindowMaxY"], o["WindowMinZ"], o["WindowMaxZ"]])
    shape = np.array([avg["GridSliceCount"], avg["GridRowCount"], avg["GridColCount"]])
    det = np.array([apg["DetectorRowCount"], apg["DetectorColCount"]])
    pts = np.array([[0.1, 0.2, 0.3], [-0.5, 0.25, 0.0]])
    pp = np.stack([cpg.project_point(p) for p in pts])
    return {"window": window, "shape": shape, "det": det, "vectors": np.asarray(apg["Vectors"], dtype=np.float64),
            "kind": np.array([0 if apg["type"] == "cone_vec" else 1]), "project_point": pp}


def project_point_cases(ts):
    """Geometries with orthogonal detector axes (the reference's project_point divides by |u|^2 and |v|^2
    separately, which is the detector coordinate only then): name -> projection geometry."""
    rng = np.random.default_rng(4321)
    R = (ts.rotate(pos=(0.1, -0.2, 0.3), axis=(1.0, 0.5, -0.2), angles=0.7) * ts.translate((0.3, -0.4, 0.5)))
    angles = rng.uniform(0, 2 * np.pi, size=9)
    return {
        # tests/geometry/test_cone_vec.py:143-173 of the reference (its own known answers live on these two)
        "ref_test_coarse": ts.cone(angles=1, shape=(10, 40), size=(30, 80), src_orig_dist=10, src_det_dist=10).to_vec(),
        "ref_test_fine": ts.cone(angles=1, shape=(100, 400), size=(30, 80), src_orig_dist=10, src_det_dist=10).to_vec(),
        "cfg3_cone": ts.cone(angles=24, shape=(32, 48), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec(),
        "cone_vec_R": R * ts.cone(angles=angles, shape=(6, 7), size=(3, 4), src_orig_dist=5, src_det_dist=8).to_vec(),
        "parallel": ts.parallel(angles=angles, shape=(9, 11), size=(2.0, 3.5)).to_vec(),
        "par_vec_R": R * ts.parallel(angles=angles, shape=(6, 7), size=(3, 4)).to_vec(),
    }


def describe_project_point(ts, pg, seed):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-0.6, 0.6, size=(16, 3))               # (z, y, x), the reference's point order
    pp = np.stack([pg.project_point(p) for p in pts])        # (16, n_angles, 2) = (v, u) in pixels from the detector centre
    apg = pg.to_astra()
    return {"points_zyx": pts, "project_point": pp, "vectors": np.asarray(apg["Vectors"], dtype=np.float64),
            "det": np.array([apg["DetectorRowCount"], apg["DetectorColCount"]]),
            "kind": np.array([0 if apg["type"] == "cone_vec" else 1])}


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import tomosipo as ts

    out = {}
    for name, (vg, pg) in cases(ts).items():
        for k, v in describe(ts, vg, pg).items():
            out[f"{name}/{k}"] = v
    # transform algebra samples
    T = ts.rotate(pos=(0.1, -0.2, 0.3), axis=(1.0, 0.5, -0.2), angles=[0.0, 0.7, 2.1])
    out["transform/rotate"] = T.matrix
    out["transform/reflect"] = ts.reflect(pos=(1, 2, 3), axis=(0.3, -1, 0.2)).matrix
    out["transform/scale"] = ts.scale((1, 2, 3), pos=(1, 0, -1), alpha=[1.0, 0.5]).matrix
    out["transform/perspective"] = ts.from_perspective(pos=(1, 2, 3), w=(0, 1, 0), v=(0, 0, 2), u=(3, 0, 0)).matrix
    np.savez_compressed(os.path.join(HERE, "geometry_golden.npz"), **out)
    print(f"wrote {len(out)} arrays")
    # voxel -> detector map (the (U, V) level of the backprojector) from the reference's project_point
    out = {}
    for i, (name, pg) in enumerate(project_point_cases(ts).items()):
        for k, v in describe_project_point(ts, pg, 100 + i).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "project_point_golden.npz"), **out)
    print(f"wrote {len(out)} project_point arrays")


if __name__ == "__main__":
    main()
