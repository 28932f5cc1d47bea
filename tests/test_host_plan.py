"""Host-array pipeline plan (tsp_projector_host_plan) against the geometry, on the CPU: the sub-problems must
cover everything, and the detector rows / volume slices each one is given must contain every interpolation tap
its voxels / rays can touch (the bounds are computed in tsproj.cu: slab_row_range, block_z_range)."""
import numpy as np
import pytest

from oracle import oracle as O
from tomosipo_b200 import _backend as B


def geometries():
    ang = np.linspace(0, 2 * np.pi, 48, endpoint=False)
    out = []
    # cfg-3-like circular cone, cubic voxels
    out.append(("cone", B.KIND_CONE_VEC, (128, 128, 128), [(-.5, .5)] * 3, (128, 192),
                O.cone_vectors(ang, 2.8125 / 192, 1.875 / 128, 4.0, 2.0)))
    # wide cone angle, anisotropic off-centre volume, more rows than slices
    out.append(("cone_wide", B.KIND_CONE_VEC, (96, 80, 112), [(-.6, .8), (-.5, .5), (-.3, .5)], (160, 144),
                O.cone_vectors(ang, 3.2 / 144, 2.6 / 160, 2.0, 1.5)))
    # parallel beam, tilted: rays leave the horizontal plane
    v = O.parallel_vectors(np.linspace(0, np.pi, 40, endpoint=False), 1.6 / 128, 1.4 / 128)
    tilt = 0.2
    v[:, 2] = tilt                        # ray z component
    out.append(("par_tilted", B.KIND_PARALLEL_VEC, (128, 96, 96), [(-.5, .5)] * 3, (128, 128), v))
    return out


def project(kind, vec, pts, det_shape):
    """Detector index coordinates (u, v) (texel-centre convention: pixel i covers [i, i+1)) of points `pts` [N, 3]
    (x, y, z) for one angle's 12-vector; nan where the ray is parallel to the detector."""
    d, du, dv = vec[3:6], vec[6:9], vec[9:12]
    nrm = np.cross(du, dv)
    if kind == B.KIND_CONE_VEC:
        s = vec[0:3]
        dirs = pts - s
        t = ((d - s) @ nrm) / (dirs @ nrm)
        hit = s + t[:, None] * dirs
    else:
        r = vec[0:3]
        t = ((d - pts) @ nrm) / (r @ nrm)
        hit = pts + t[:, None] * r
    rel = hit - d
    return rel @ du / (du @ du) + det_shape[1] / 2, rel @ dv / (dv @ dv) + det_shape[0] / 2


@pytest.mark.parametrize("geom", geometries(), ids=lambda g: g[0])
def test_host_plan_covers_every_tap(geom):
    name, kind, vol_shape, window, det_shape, vec = geom
    nz, ny, nx = vol_shape
    V, U = det_shape
    P = B.Projector(kind, vol_shape, window, det_shape, vec)
    bp, fp = P.host_plan(B.BP), P.host_plan(B.FP)
    assert bp and fp
    # the chunks partition the volume (BP) / the detector rows (FP)
    assert sorted((z0, z1) for z0, z1, _, _ in bp)[0][0] == 0 and sum(z1 - z0 for z0, z1, _, _ in bp) == nz
    assert sorted((v0, v1) for _, _, v0, v1 in fp)[0][0] == 0 and sum(v1 - v0 for _, _, v0, v1 in fp) == V
    zs = sorted((z0, z1) for z0, z1, _, _ in bp)
    assert all(a[1] == b[0] for a, b in zip(zs, zs[1:]))
    vs = sorted((v0, v1) for _, _, v0, v1 in fp)
    assert all(a[1] == b[0] for a, b in zip(vs, vs[1:]))

    (x0, x1), (y0, y1), (zlo, zhi) = window
    sx, sy, sz = (x1 - x0) / nx, (y1 - y0) / ny, (zhi - zlo) / nz
    rng = np.random.default_rng(0)

    # ---- BP: a voxel centre of slab [z0, z1) samples rows floor(v - 0.5), +1; those inside the detector must be staged
    for z0, z1, v0, v1 in bp:
        ix = rng.integers(0, nx, 400); iy = rng.integers(0, ny, 400)
        iz = np.concatenate([rng.integers(z0, z1, 396), [z0, z0, z1 - 1, z1 - 1]])
        ix[-4:] = [0, nx - 1, 0, nx - 1]; iy[-4:] = [0, ny - 1, ny - 1, 0]
        pts = np.stack([x0 + (ix + .5) * sx, y0 + (iy + .5) * sy, zlo + (iz + .5) * sz], axis=1)
        for a in range(len(vec)):
            _, v = project(kind, vec[a], pts, det_shape)
            r = np.floor(v - 0.5)
            for tap in (r, r + 1):
                inside = (tap >= 0) & (tap < V) & np.isfinite(v)
                assert np.all((tap[inside] >= v0) & (tap[inside] < v1)), (name, "BP", (z0, z1, v0, v1), a)

    # ---- FP: points of the rays of row block [v0, v1) inside the volume's x-y extent interpolate between slices
    #      floor(q), floor(q) + 1 with q the z index coordinate; those inside the volume must be uploaded
    for z0, z1, v0, v1 in fp:
        for a in range(0, len(vec), 3):
            w = vec[a]
            d, du, dv = w[3:6], w[6:9], w[9:12]
            iu = rng.uniform(0, U, 64); iv = np.concatenate([rng.uniform(v0, v1, 60), [v0, v0, v1, v1]])
            pix = d + (iu - U / 2)[:, None] * du + (iv - V / 2)[:, None] * dv
            if kind == B.KIND_CONE_VEC:
                org = np.broadcast_to(w[0:3], pix.shape); dirs = pix - org
            else:
                org = pix; dirs = np.broadcast_to(w[0:3], pix.shape)
            for t in np.linspace(-3, 3, 241):
                p = org + t * dirs
                ok = (p[:, 0] >= x0) & (p[:, 0] <= x1) & (p[:, 1] >= y0) & (p[:, 1] <= y1)
                q = (p[:, 2] - zlo) / sz - 0.5
                for tap in (np.floor(q), np.floor(q) + 1):
                    inside = ok & (tap >= 0) & (tap < nz)
                    assert np.all((tap[inside] >= z0) & (tap[inside] < z1)), (name, "FP", (z0, z1, v0, v1), a)


def test_small_problems_are_not_pipelined():
    vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 8, endpoint=False), 0.05, 0.05, 4.0, 2.0)
    P = B.Projector(B.KIND_CONE_VEC, (32, 32, 32), [(-.5, .5)] * 3, (32, 48), vec)
    assert P.host_plan(B.FP) == [] and P.host_plan(B.BP) == []
    with pytest.raises(ValueError):
        P.host_plan(7)
