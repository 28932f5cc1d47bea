"""Behaviour of the geometry classes (own condensed restatement of what the
reference pins in tests/geometry/*.py; the reference suite itself was run
against this package during development)."""
import numpy as np
import pytest

import tomosipo_b200 as ts
from tomosipo_b200.geometry import random_cone, random_cone_vec, random_parallel, random_parallel_vec
from tomosipo_b200.geometry import random_transform, random_volume, random_volume_vec


def test_volume_constructors_and_properties():
    vg = ts.volume(shape=(2, 4, 8), pos=(1, 2, 3), size=(1, 2, 4))
    assert vg.shape == (2, 4, 8) and vg.size == (1.0, 2.0, 4.0) and vg.voxel_size == (0.5, 0.5, 0.5)
    assert vg.extent == ((0.5, 1.5), (1.0, 3.0), (1.0, 5.0))
    assert ts.volume(shape=(2, 4, 8), extent=vg.extent) == vg
    assert ts.volume(shape=3).size == (3.0, 3.0, 3.0)
    with pytest.raises(ValueError):
        ts.volume(shape=1, extent=((0, 1),) * 3, pos=0)
    with pytest.raises(TypeError):
        ts.volume(shape=(1.5, 1, 1))
    assert repr(ts.volume(shape=1)) == "ts.volume(\n    shape=(1, 1, 1),\n    pos=(0.0, 0.0, 0.0),\n    size=(1.0, 1.0, 1.0),\n)"
    assert eval(repr(vg), {"ts": ts}) == vg
    assert vg in ts.volume(shape=1, pos=(1, 2, 3), size=10)
    assert vg.translate((1, 0, 0)).untranslate((1, 0, 0)) == vg
    assert vg.scale(2).size == (2.0, 4.0, 8.0) and vg.multiply(2).pos[0, 0] == 2.0
    assert vg.with_voxel_size(1.0).shape == (1, 2, 4)
    assert vg.reshape(4).voxel_size == (0.25, 0.5, 1.0)


def test_volume_slicing_and_binning():
    vg = ts.volume(shape=(4, 4, 8), size=(4, 4, 8))
    assert vg[:1] == ts.volume(shape=(1, 4, 8), pos=(-1.5, 0, 0), size=(1, 4, 8))
    assert vg[:, :, ::2].voxel_size == (1.0, 1.0, 2.0)
    assert vg[1, 2, 3].shape == (1, 1, 1)
    assert vg[-1] == vg[3:]
    assert vg.to_vec()[0, ::2] == vg[::2].to_vec()


def test_astra_round_trips():
    for g in (random_volume(), random_cone(), random_cone_vec(), random_parallel(), random_parallel_vec()):
        assert ts.from_astra(ts.to_astra(g)) == g
    d = ts.volume(shape=(2, 3, 4), extent=((0, 1), (2, 4), (5, 9))).to_astra()
    assert (d["GridSliceCount"], d["GridRowCount"], d["GridColCount"]) == (2, 3, 4)
    assert (d["option"]["WindowMinX"], d["option"]["WindowMaxX"]) == (5.0, 9.0)
    assert (d["option"]["WindowMinZ"], d["option"]["WindowMaxZ"]) == (0.0, 1.0)
    with pytest.raises(TypeError):
        ts.from_astra(ts.volume(shape=1))
    with pytest.raises(TypeError):
        ts.to_astra({})
    pg = ts.cone(angles=3, shape=(4, 6), size=(2, 3), src_orig_dist=5, src_det_dist=8)
    d = pg.to_astra()
    assert d["type"] == "cone" and d["DistanceOriginDetector"] == 3.0 and d["DetectorSpacingX"] == 0.5
    v = pg.to_vec().to_astra()
    assert v["type"] == "cone_vec" and v["Vectors"].shape == (3, 12)
    np.testing.assert_allclose(v["Vectors"][0], [0, -5, 0, 0, 3, 0, 0.5, 0, 0, 0, 0, 0.5], atol=1e-12)


def test_cone_parameters():
    pg = ts.cone(size=np.sqrt(2), cone_angle=1 / 2, angles=100, shape=(128, 192))   # README.md:140
    assert abs(pg.src_orig_dist - 2 * np.sqrt(2)) < 1e-12 and pg.src_det_dist == pg.src_orig_dist
    assert ts.cone(src_orig_dist=3, angles=1).src_det_dist == 3.0
    with pytest.raises(ValueError):
        ts.cone(angles=1)
    with pytest.raises(ValueError):
        ts.cone(angles=1, cone_angle=1, src_orig_dist=2)
    with pytest.raises(TypeError):
        ts.cone(angles=[], cone_angle=1)
    with pytest.raises(TypeError):
        ts.parallel(angles=[])
    assert pg[:10].num_angles == 10 and pg[3].num_angles == 1
    with pytest.raises(ValueError):
        pg[0, 0]
    assert pg.rescale_det(2).det_shape == (64, 96) and pg.reshape(10).det_shape == (10, 10)
    assert eval(repr(ts.cone(angles=3, cone_angle=1)), {"ts": ts}) == ts.cone(angles=3, cone_angle=1)
    assert eval(repr(ts.parallel(angles=3, shape=2)), {"ts": ts, "array": np.array}) == ts.parallel(angles=3, shape=2)


def test_vector_geometry_slicing_transform_and_projection():
    pg = ts.cone(angles=6, shape=(8, 12), size=(2, 3), src_orig_dist=3, src_det_dist=5).to_vec()
    sub = pg[1:5, ::2, 2:10:4]
    assert sub.num_angles == 4 and sub.det_shape == (4, 2)
    np.testing.assert_allclose(sub.det_v, 2 * pg.det_v[1:5])
    np.testing.assert_allclose(sub.det_u, 4 * pg.det_u[1:5])
    np.testing.assert_allclose(sub.src_pos, pg.src_pos[1:5])
    np.testing.assert_allclose(sub.det_sizes, [[2, 2]] * 4)
    row = ts.parallel(angles=4, shape=(16, 24), size=(1, 1.5)).to_vec()[:, :1, :]       # learned_pd.py:55
    np.testing.assert_allclose(row.det_pos[:, 0], -0.5 + 0.5 / 16)
    assert row.det_shape == (1, 24)

    T = random_transform()
    moved = T * pg
    p = np.array([0.3, -0.2, 0.1])
    np.testing.assert_allclose(moved.project_point(T.transform_point(p)[0]), pg.project_point(p), atol=1e-9)
    assert (T.inv * moved) == pg
    assert moved.rescale_det((2, 3)).det_shape == (4, 4)
    np.testing.assert_allclose(moved.reshape((4, 6)).det_sizes, moved.det_sizes)
    with pytest.raises(NotImplementedError):
        pg.ray_dir
    with pytest.raises(NotImplementedError):
        ts.parallel(angles=2).to_vec().src_pos
    with pytest.warns(UserWarning):
        T * ts.parallel(angles=2)
    with pytest.warns(UserWarning):
        ts.rotate(pos=0, axis=(1, 0, 0), angles=1.0) * ts.volume(shape=2)
    assert (ts.translate((1, 2, 3)) * ts.scale(2)) * ts.volume(shape=2) == ts.volume(shape=2, pos=(1, 2, 3), size=4)


def test_transform_algebra():
    T = random_transform()
    assert T * T.inv == ts.geometry.transform.identity()
    R = ts.rotate(pos=(1, 2, 3), axis=(0, 0, 1), angles=[0.0, np.pi / 2])
    assert R.num_steps == 2 and R[0] == ts.geometry.transform.identity()
    # left-handed (z, y, x) frame: rotating e_z about e_x by +90 degrees gives -e_y (reference doctest)
    np.testing.assert_allclose(ts.rotate(pos=0, axis=(0, 0, 1), angles=np.pi / 2).transform_vec((1, 0, 0)),
                               [[0, -1, 0]], atol=1e-12)
    assert ts.reflect(pos=0, axis=(1, 0, 0)) * ts.reflect(pos=0, axis=(1, 0, 0)) == ts.geometry.transform.identity()
    assert ts.scale(2) * ts.scale(0.5) == ts.geometry.transform.identity()
    P = ts.to_perspective(pos=(1, 2, 3), w=(0, 2, 0), v=(0, 0, 1), u=(1, 0, 0))
    np.testing.assert_allclose(P.transform_point((0, 0, 0)), [[1, 2, 3]])
    np.testing.assert_allclose(P.transform_vec((1, 0, 0)), [[0, 1, 0]])
    assert ts.from_perspective(pos=(1, 2, 3), w=(0, 2, 0), v=(0, 0, 1), u=(1, 0, 0)) == P.inv
    with pytest.raises(ValueError):
        ts.translate(np.zeros((2, 3))) * ts.translate(np.zeros((3, 3)))
    with pytest.warns(DeprecationWarning):
        ts.rotate(pos=0, axis=(1, 0, 0), deg=90)
    with pytest.raises(ValueError):
        ts.rotate(pos=0, axis=(1, 0, 0))
    assert ts.concatenate([ts.translate((1, 0, 0)), ts.translate((2, 0, 0))]).num_steps == 2


def test_concatenate_and_volume_vec():
    a, b = ts.parallel(angles=3, shape=4), ts.parallel(angles=2, shape=4)
    assert ts.concatenate([a, b]).num_angles == 5
    c = ts.cone(angles=3, shape=4, cone_angle=1)
    assert ts.concatenate([c, c.to_vec()]).num_angles == 6
    with pytest.raises(TypeError):
        ts.concatenate([a, c])
    with pytest.raises(ValueError):
        ts.concatenate([a, ts.parallel(angles=2, shape=5)])
    with pytest.raises(ValueError):
        ts.concatenate([])
    vv = random_volume_vec()
    assert ts.concatenate([vv, vv]).num_steps == 2 * vv.num_steps
    assert vv.corners.shape == (vv.num_steps, 8, 3)
    np.testing.assert_allclose(vv.corners.mean(axis=1), vv.pos, atol=1e-9)
    assert vv.reshape(2).sizes == pytest.approx(vv.sizes)
    vg = ts.volume_vec(shape=(2, 3, 4), pos=np.zeros((3, 3)), w=[(1, 0, 0), (2, 0, 0), (3, 0, 0)])
    with pytest.raises(ValueError):
        vg.size
    assert vg.sizes.shape == (3, 3)


def test_random_geometries_fuzz_conversion():
    # geometry -> vectors must keep the detector frame consistent for arbitrary geometries
    for _ in range(10):
        pg = random_cone_vec()
        v = pg.to_astra()["Vectors"]
        np.testing.assert_allclose(v[:, 0:3][:, ::-1], pg.src_pos)
        np.testing.assert_allclose(v[:, 9:12][:, ::-1], pg.det_v)
        centre = pg.lower_left_corner + pg.det_shape[0] / 2 * pg.det_v + pg.det_shape[1] / 2 * pg.det_u
        np.testing.assert_allclose(centre, pg.det_pos, atol=1e-9)


def test_fdk_angle_table_of_a_circular_scan():
    """Host part of algorithms.fdk: per-angle constants recovered from the cone-beam vectors
    (pixel pitches, SDD, SOD, principal point) for a circular scan and for a shifted detector."""
    import tomosipo_b200 as ts
    from tomosipo_b200.algorithms import _fdk_angle_table

    vg = ts.volume(shape=(8, 10, 12), size=(0.8, 1.0, 1.2))
    pg = ts.cone(angles=7, shape=(16, 24), size=(3.2, 3.6), src_orig_dist=5, src_det_dist=8)
    t = _fdk_angle_table(ts.operator(vg, pg))
    assert np.allclose(t["pu"], 3.6 / 24) and np.allclose(t["pv"], 3.2 / 16)
    assert np.allclose(t["sdd"], 8) and np.allclose(t["sod"], 5)
    assert np.allclose(t["ppu"], 0, atol=1e-12) and np.allclose(t["ppv"], 0, atol=1e-12)
    assert np.isclose(t["vox"], 0.1 ** 3)
    # detector shifted by 2.5 pixels along u and -1 pixel along v: the principal point moves the other way
    v = pg.to_vec()
    shifted = ts.cone_vec(shape=v.det_shape, src_pos=v.src_pos, det_pos=v.det_pos + 2.5 * v.det_u - 1.0 * v.det_v,
                          det_v=v.det_v, det_u=v.det_u)
    t2 = _fdk_angle_table(ts.operator(vg, shifted))
    assert np.allclose(t2["ppu"], -2.5) and np.allclose(t2["ppv"], 1.0) and np.allclose(t2["sdd"], 8)
    with pytest.raises(TypeError):
        _fdk_angle_table(ts.operator(vg, ts.parallel(angles=5, shape=(16, 24))))
