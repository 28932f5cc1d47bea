"""The CPU oracle against every known answer the reference holds for the
projection path and against closed-form line integrals."""
import numpy as np
import pytest

from oracle import oracle as O


def unit_cube(n, size=1.0):
    return [-size / 2] * 3, [size / 2] * 3


def test_fp_statistics_of_real_astra_output():
    # notebooks/cupy.ipynb cell 4 of the reference: the only real ASTRA FP output it ships.
    # 128^3 hollow box (mean 0.21790314) in a unit cube, parallel beam, 128x128 detector of
    # size sqrt(2), 180 angles over 2 pi  ->  sino mean 0.109146185, max 0.8471311, min 0.0
    n = 128
    x = O.hollow_box(n)
    assert abs(float(x.mean()) - 0.21790314) < 1e-7
    s = np.sqrt(2) / n
    vec = O.parallel_vectors(np.linspace(0, 2 * np.pi, 180, endpoint=False), s, s)
    lo, hi = unit_cube(n)
    y = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), lo, hi, (n, n), vec).fp(x)
    assert abs(y.mean() - 0.109146185) / 0.109146185 < 2e-6
    assert abs(y.max() - 0.8471311) / 0.8471311 < 1e-6
    assert y.min() == 0.0


def test_fp_axis_aligned_chords():
    # ones volume of 16 voxels, parallel beam along y: every interior ray integrates to 16 * voxel
    n = 16
    vec = O.parallel_vectors([0.0], 1.0, 1.0)
    P = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), [-8] * 3, [8] * 3, (n, n), vec)
    y = P.fp(np.ones((n, n, n)))
    np.testing.assert_allclose(y[:, 0, :], 16.0, atol=1e-12)
    # anisotropic voxels (x, y, z) = (0.5, 2, 1): chord along y is 16 * 2
    P = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), [-4, -16, -8], [4, 16, 8], (n, n), O.parallel_vectors([0.0], 0.5, 1.0))
    np.testing.assert_allclose(P.fp(np.ones((n, n, n)))[:, 0, :], 32.0, atol=1e-12)


def test_fp_oblique_box_chord_matches_closed_form():
    # ray through the centre of a cube of side L at angle t (in the x-y plane): chord = L / max(|cos|,|sin|)
    n = 64
    for t in (0.3, 0.9, 1.2, 2.5):
        vec = O.parallel_vectors([t], 1.0 / n, 1.0 / n)
        lo, hi = unit_cube(n)
        P = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), lo, hi, (2, 2), vec)
        y = P.fp(np.ones((n, n, n)))
        expect = 1.0 / max(abs(np.cos(t)), abs(np.sin(t)))
        # the 2x2 detector straddles the centre by half a pixel; Joseph is exact for a constant box
        # up to the linear fade at the faces
        assert abs(y.mean() - expect) < 2.5 / n


def test_marching_axis_ties_prefer_x_then_y():
    t = np.array([np.pi / 4, 3 * np.pi / 4, 0.0, np.pi / 2])
    P = O.OracleProjector(O.PARALLEL_VEC, (8, 8, 8), [-4] * 3, [4] * 3, (8, 8), O.parallel_vectors(t, 1, 1))
    ax = P.marching_axes()
    assert ax[2] == 1 and ax[3] == 0          # ray along -y marches y; along x marches x
    assert ax[0] in (0, 1) and ax[1] in (0, 1)


@pytest.mark.parametrize("kind", ["parallel", "cone"])
def test_bp_is_scaled_adjoint(kind):
    rng = np.random.default_rng(0)
    n = 32
    t = np.linspace(0, 2 * np.pi, 24, endpoint=False)
    if kind == "parallel":
        vec, k = O.parallel_vectors(t, 1.5 / 48, 1.0 / 32), O.PARALLEL_VEC
    else:
        vec, k = O.cone_vectors(t, 2.8125 / 48, 1.875 / 32, 4.0, 2.0), O.CONE_VEC
    lo, hi = unit_cube(n)
    P = O.OracleProjector(k, (n, n, n), lo, hi, (32, 48), vec)
    x, y = rng.random(P.vol_shape), rng.random(P.proj_shape)
    ratio = (P.fp(x) * y).sum() / (x * P.bp(y)).sum()
    assert abs(ratio - 1) < 0.02, ratio


def test_sirt_weights_reconstruct():
    # README.md:150-164 of the reference: R = 1/A(1), C = 1/A^T(1); the iteration must converge
    n = 24
    vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 30, endpoint=False), 3.0 / 36, 2.0 / 24, 6.0, 3.0)
    P = O.OracleProjector(O.CONE_VEC, (n, n, n), [-.5] * 3, [.5] * 3, (24, 36), vec)
    phantom = np.zeros((n, n, n)); phantom[6:14, 8:16, 5:12] = 1.0
    R = 1 / np.maximum(P.fp(np.ones_like(phantom)), 1e-8)
    C = 1 / np.maximum(P.bp(np.ones(P.proj_shape)), 1e-8)
    y = P.fp(phantom)
    x = np.zeros_like(phantom)
    res = []
    for _ in range(30):
        r = y - P.fp(x)
        res.append(np.linalg.norm(r))
        x += C * P.bp(R * r)
    assert res[-1] < 0.15 * res[0]
    assert np.linalg.norm(x - phantom) < 0.5 * np.linalg.norm(phantom)


def test_additive_and_float32_build():
    rng = np.random.default_rng(1)
    n = 12
    vec = O.parallel_vectors(np.linspace(0, np.pi, 7, endpoint=False), 0.1, 0.12)
    P = O.OracleProjector(O.PARALLEL_VEC, (n, n + 1, n + 2), [-.6, -.7, -.5], [.8, .6, .7], (9, 13), vec)
    x, y0 = rng.random(P.vol_shape), rng.random(P.proj_shape)
    np.testing.assert_allclose(P.fp(x, out=y0.copy(), additive=True), P.fp(x) + y0, atol=1e-13)
    np.testing.assert_allclose(P.bp(y0, out=x.copy(), additive=True), P.bp(y0) + x, atol=1e-13)
    y32 = P.fp(x.astype(np.float32), dtype=np.float32)
    assert np.linalg.norm(y32 - P.fp(x)) / np.linalg.norm(P.fp(x)) < 1e-5


def test_supersampling_converges_to_finer_grid():
    # detector supersampling 2 == average of the 2x2 finer detector pixels
    n = 16
    t = np.linspace(0, np.pi, 5, endpoint=False)
    lo, hi = unit_cube(n)
    coarse = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), lo, hi, (8, 10), O.parallel_vectors(t, 0.2, 0.2),
                               detector_supersampling=2)
    fine = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), lo, hi, (16, 20), O.parallel_vectors(t, 0.1, 0.1))
    x = np.random.default_rng(2).random((n, n, n))
    yf = fine.fp(x)
    avg = yf.reshape(8, 2, 5, 10, 2).mean(axis=(1, 4))
    np.testing.assert_allclose(coarse.fp(x), avg, atol=1e-12)


def test_bp_of_ones_closed_forms():
    """Backprojection of a stack of ones (the C = 1 / A.T(1) of README.md:151-154) in closed form.

    Bilinear interpolation of a constant is exact, so a voxel whose shadow stays on the detector gets
    n_angles x (ray-density weight) x (voxel volume):
      parallel:  1 / (pixel area)                       -> n_angles * s^3 / (pu * pv)
      cone, voxel on the rotation axis: (SDD / SOD)^2 / (pixel area) at every angle
                                                        -> n_angles * s^3 * (SDD / SOD)^2 / (pu * pv)
    """
    n, na = 32, 20
    lo, hi = unit_cube(n)
    s = 1.0 / n
    t = np.linspace(0, 2 * np.pi, na, endpoint=False)
    pu, pv = 1.5 / 48, 1.25 / 32
    P = O.OracleProjector(O.PARALLEL_VEC, (n, n, n), lo, hi, (32, 48), O.parallel_vectors(t, pu, pv))
    x = P.bp(np.ones(P.proj_shape))
    inner = x[8:24, 8:24, 8:24]
    assert np.allclose(inner, na * s ** 3 / (pu * pv), rtol=1e-12)
    sod, odd = 4.0, 2.0
    pu, pv = 2.8125 / 48, 1.875 / 32
    Pc = O.OracleProjector(O.CONE_VEC, (n + 1, n + 1, n + 1), [-(n + 1) * s / 2] * 3, [(n + 1) * s / 2] * 3, (32, 48),
                           O.cone_vectors(t, pu, pv, sod, odd))
    xc = Pc.bp(np.ones(Pc.proj_shape))
    c = n // 2                                        # odd grid: voxel c is centred on the rotation axis, z = 0
    assert abs(xc[c, c, c] / (na * s ** 3 * ((sod + odd) / sod) ** 2 / (pu * pv)) - 1) < 1e-12
    # off the mid-plane the weight uses the distance along the central ray only: same value along the axis
    assert abs(xc[c + 5, c, c] / xc[c, c, c] - 1) < 1e-12
