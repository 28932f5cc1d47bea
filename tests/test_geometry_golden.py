"""Host-side geometry conversion against golden vectors produced by the
reference's own code (tests/golden/make_golden.py, run against /root/reference
in the build container)."""
import importlib.util
import os

import numpy as np
import pytest

import tomosipo_b200 as ts

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)

GOLD = np.load(os.path.join(HERE, "golden", "geometry_golden.npz"))
CASES = make_golden.cases(ts)


@pytest.mark.parametrize("name", sorted(CASES))
def test_operator_geometry_matches_reference(name):
    vg, pg = CASES[name]
    got = make_golden.describe(ts, vg, pg)
    for key, val in got.items():
        ref = GOLD[f"{name}/{key}"]
        assert ref.shape == np.asarray(val).shape, key
        np.testing.assert_allclose(val, ref, rtol=0, atol=1e-12, err_msg=f"{name}/{key}")


def test_transforms_match_reference():
    T = ts.rotate(pos=(0.1, -0.2, 0.3), axis=(1.0, 0.5, -0.2), angles=[0.0, 0.7, 2.1])
    np.testing.assert_allclose(T.matrix, GOLD["transform/rotate"], atol=1e-13)
    np.testing.assert_allclose(ts.reflect(pos=(1, 2, 3), axis=(0.3, -1, 0.2)).matrix, GOLD["transform/reflect"], atol=1e-13)
    np.testing.assert_allclose(ts.scale((1, 2, 3), pos=(1, 0, -1), alpha=[1.0, 0.5]).matrix, GOLD["transform/scale"], atol=1e-13)
    np.testing.assert_allclose(
        ts.from_perspective(pos=(1, 2, 3), w=(0, 1, 0), v=(0, 0, 2), u=(3, 0, 0)).matrix,
        GOLD["transform/perspective"], atol=1e-13)
