"""Known answers the reference's documentation and tests hold for the
geometry -> vector conversion (crumbs 1 and 2)."""
import numpy as np

import tomosipo_b200 as ts


def test_parallel_vec_dump_from_docs():
    # doc/topics/geometries.rst:366-410: ts.parallel(angles=3, shape=(10, 15), size=(1, 1.5)).to_vec()
    pg = ts.parallel(angles=3, shape=(10, 15), size=(1, 1.5)).to_vec()
    np.testing.assert_allclose(pg.ray_dir, [[0, -1, 0], [0, -0.5, 0.8660254], [0, 0.5, 0.8660254]], atol=1e-7)
    np.testing.assert_allclose(pg.det_pos, np.zeros((3, 3)), atol=1e-12)
    np.testing.assert_allclose(pg.det_v, [[0.1, 0, 0]] * 3, atol=1e-12)
    np.testing.assert_allclose(pg.det_u, [[0, 0, 0.1], [0, 0.08660254, 0.05], [0, 0.08660254, -0.05]], atol=1e-8)
    np.testing.assert_allclose(
        pg.det_normal, [[0, 0.01, 0], [0, 0.005, -0.00866025], [0, -0.005, -0.00866025]], atol=1e-8)
    np.testing.assert_allclose(
        pg.lower_left_corner, [[-0.5, 0, -0.75], [-0.5, -0.64951905, -0.375], [-0.5, -0.64951905, 0.375]], atol=1e-8)
    np.testing.assert_allclose(
        pg.corners,
        [[[-0.5, 0, -0.75], [0.5, 0, -0.75], [-0.5, 0, 0.75], [0.5, 0, 0.75]],
         [[-0.5, -0.64951905, -0.375], [0.5, -0.64951905, -0.375], [-0.5, 0.64951905, 0.375], [0.5, 0.64951905, 0.375]],
         [[-0.5, -0.64951905, 0.375], [0.5, -0.64951905, 0.375], [-0.5, 0.64951905, -0.375], [0.5, 0.64951905, -0.375]]],
        atol=1e-8)

def test_cone_project_point():
    # tests/geometry/test_cone_vec.py:143-173 of the reference: 3x2-unit pixels, SOD = SDD = 10
    pg = ts.cone(shape=(10, 10), size=(30, 20), angles=1, src_orig_dist=10, src_det_dist=10).to_vec()
    np.testing.assert_allclose(pg.project_point((0, 0, 0)), [[0, 0]], atol=1e-12)
    np.testing.assert_allclose(pg.project_point((3, 0, 0)), [[1, 0]], atol=1e-12)
    np.testing.assert_allclose(pg.project_point((0, 0, 2)), [[0, 1]], atol=1e-12)


def test_cone_source_detector_placement():
    # doc/howto/cone_beam_template.rst:16-35: source on the negative y side, detector on the positive
    pg = ts.cone(angles=1, shape=4, size=4, src_orig_dist=3, src_det_dist=5).to_vec()
    np.testing.assert_allclose(pg.src_pos, [[0, -3, 0]], atol=1e-12)
    np.testing.assert_allclose(pg.det_pos, [[0, 2, 0]], atol=1e-12)


def test_default_arcs():
    # geometry/parallel.py:86-88 and geometry/cone.py:164-166: half / full circle, endpoint excluded
    np.testing.assert_allclose(ts.parallel(angles=4).angles, np.linspace(0, np.pi, 4, endpoint=False))
    np.testing.assert_allclose(ts.cone(angles=4, cone_angle=1).angles, np.linspace(0, 2 * np.pi, 4, endpoint=False))


def test_shapes_from_docs():
    # doc/intro/forward_projection.rst:83-86
    A = ts.operator(ts.volume(shape=32), ts.parallel(angles=32, shape=48))
    assert A.domain_shape == (32, 32, 32) and A.range_shape == (48, 32, 48)


def test_hollow_box_mean():
    # tests/test_phantom.py:11-17
    vd = ts.phantom.hollow_box(ts.data(ts.volume(shape=100)))
    assert abs(vd.data.mean() - 0.208) < 1e-6
