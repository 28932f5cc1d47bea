"""The C-ABI library: loads, exports every symbol of include/tsproj.h, validates
arguments, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from tomosipo_b200 import _backend as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tsproj.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsp_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(B.LIB_PATH)
    names = declared_symbols()
    assert set(names) == set(B.EXPORTED_SYMBOLS), (names, B.EXPORTED_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert B.lib().tsp_version() == 100


def test_struct_layout_matches_oracle_struct():
    # one ctypes Structure serves both libraries
    assert [f[0] for f in B.tsp_geometry._fields_] == [f[0] for f in O.Geometry._fields_]
    assert ctypes.sizeof(B.tsp_geometry) == ctypes.sizeof(O.Geometry)


def _vec(n=5):
    return O.cone_vectors(np.linspace(0, 2 * np.pi, n, endpoint=False), 0.1, 0.1, 5.0, 3.0)


def test_create_validates_geometry():
    win = [(-1, 1)] * 3
    with pytest.raises(ValueError, match="shape"):
        B.Projector(0, (0, 4, 4), win, (4, 4), _vec())
    with pytest.raises(ValueError, match="window"):
        B.Projector(0, (4, 4, 4), [(1, 1)] * 3, (4, 4), _vec())
    with pytest.raises(ValueError, match="kind"):
        B.Projector(7, (4, 4, 4), win, (4, 4), _vec())
    with pytest.raises(ValueError, match="supersampling"):
        B.Projector(0, (4, 4, 4), win, (4, 4), _vec(), voxel_supersampling=0)
    bad = _vec(); bad[1, 3] = np.nan
    with pytest.raises(ValueError, match="non-finite"):
        B.Projector(0, (4, 4, 4), win, (4, 4), bad)
    with pytest.raises(ValueError, match="12"):
        B.Projector(0, (4, 4, 4), win, (4, 4), np.zeros((3, 11)))


def test_projector_info_and_marching_axes_match_oracle():
    rng = np.random.default_rng(3)
    vec = rng.normal(size=(40, 12))
    win = [(-1.0, 1.5), (-2.0, 1.0), (-0.5, 0.75)]
    P = B.Projector(0, (9, 11, 13), win, (7, 5), vec)
    Q = O.OracleProjector(0, (9, 11, 13), [w[0] for w in win], [w[1] for w in win], (7, 5), vec)
    np.testing.assert_array_equal(P.marching_axes(), Q.marching_axes())
    info = P.info()
    assert info.n_angles == 40
    assert info.n_march_x + info.n_march_y + info.n_march_z == 40
    np.testing.assert_allclose(list(info.voxel_size), [2.5 / 13, 3.0 / 11, 1.25 / 9])


def test_project_argument_checks_and_no_cpu_fallback():
    P = B.Projector(0, (4, 4, 4), [(-1, 1)] * 3, (4, 4), _vec())
    x = np.zeros(P.vol_shape, np.float32)
    y = np.zeros(P.proj_shape, np.float32)
    with pytest.raises(ValueError, match="direction"):
        P.project(5, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST)
    with pytest.raises(ValueError, match="NULL"):
        P.project(B.FP, False, 0, y.ctypes.data, B.MEM_HOST)
    with pytest.raises(ValueError, match="batch"):
        P.project(B.FP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, batch=0)
    if not B.cuda_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            P.project(B.FP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tomosipo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "tsproj.cu" and "oracle" not in text, (dirpath, f)
