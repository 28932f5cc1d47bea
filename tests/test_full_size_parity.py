"""CUDA path vs. the fp64 oracle ON the benchmark configurations themselves (BASELINE.json configs[2] and [3]).

The small-size parity tests cannot reach the code only big problems run: 16 z-tiles x 32-voxel runs, offset arithmetic
on 1.1 GB / 9 GB projection stacks, hull skipping of rays that miss the volume, TMA boxes of 45-degree angles on a
768 / 1536-column detector, the row blocks / z-slabs of the host-array pipeline.  A full fp64 oracle run at these sizes
takes hours, so the oracle forms only part of each output (oracle/tsp_oracle.c: oracle_fp_angles_mixed,
oracle_bp_window_mixed) from the very float32 arrays the GPU saw:
  * FP: whole detector rows (every 8th row plus the two border rows at each end) of a handful of angles - the first,
    the last, both sides of every marching-axis switch (45, 135, 225, 315 degrees), and a few in between;
  * BP: all angles for voxel windows at a corner, at the far corner (last tiles), on the rotation axis and off-axis.
Tolerance: relative L2 <= 1e-5 per angle / per window (north_star: "<= 1e-5 against an fp64 Joseph / voxel-driven
reference"), and max abs error <= 1e-5 of the largest value.
"""
import numpy as np
import pytest
import torch

import tomosipo_b200 as ts

from .test_operator_gpu import oracle_of, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cfg3():
    """BASELINE.json configs[2]: cone_vec 512^3, 720 angles, 512 x 768."""
    n = 512
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=720, shape=(n, 768), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
    return vg, pg


def cfg4():
    """BASELINE.json configs[3]: cone 1024^3, 1440 angles, 1024 x 1536."""
    n = 1024
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=1440, shape=(n, 1536), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
    return vg, pg


def sample_angles(A, count):
    """First, last, both sides of every marching-axis switch, then evenly spread ones up to `count`."""
    axes = A.astra_projector.marching_axes()
    na = len(axes)
    pick = [0, na - 1]
    for a in range(1, na):
        if axes[a] != axes[a - 1]:
            pick += [a - 1, a]
    pick = sorted(set(pick))
    extra = [a for a in np.linspace(0, na - 1, count + 2).astype(int)[1:-1] if a not in pick]
    while len(pick) < count and extra:
        pick.append(int(extra.pop(len(extra) // 2)))
    return sorted(set(pick))[:max(count, 10)]


def windows(shape):
    """Voxel windows (z, y, x) as (lo, hi): corner, far corner (last tiles), centre, off-axis."""
    nz, ny, nx = shape
    return [((0, 12), (0, 8), (0, 32)),
            ((nz - 9, nz), (ny - 8, ny), (nx - 32, nx)),
            ((nz // 2 - 4, nz // 2 + 4), (ny // 2 - 4, ny // 2 + 4), (nx // 2 - 16, nx // 2 + 16)),
            ((nz // 4, nz // 4 + 6), (3 * ny // 4, 3 * ny // 4 + 8), (nx // 8, nx // 8 + 32))]


def sample_rows(det_rows):
    return sorted(set(list(range(0, det_rows, 8)) + [1, det_rows - 2, det_rows - 1]))


def check_fp(want, y_host, angles, rows, what):
    got = y_host[np.ix_(rows, angles, np.arange(y_host.shape[2]))]
    scale = np.abs(want).max()
    for j, a in enumerate(angles):
        e = rel_l2(got[:, j, :], want[:, j, :])
        assert e < TOL, f"{what}: FP angle {a}: rel-L2 {e:.2e}"
        assert np.abs(got[:, j, :] - want[:, j, :]).max() < TOL * scale, f"{what}: FP angle {a}: max error"


def check_bp(P, y_host, xb_host, shape, what):
    for z, y, x in windows(shape):
        want = P.bp_window(y_host, z, y, x)
        got = xb_host[z[0]:z[1], y[0]:y[1], x[0]:x[1]]
        e = rel_l2(got, want)
        assert e < TOL, f"{what}: BP window z{z} y{y} x{x}: rel-L2 {e:.2e}"
        assert np.abs(got - want).max() < TOL * np.abs(want).max(), f"{what}: BP window z{z} y{y} x{x}: max error"


def run_config(vg, pg, n_fp_angles, host_arrays):
    A = ts.operator(vg, pg)
    P = oracle_of(A)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(A.domain_shape, device="cuda", generator=g)
    w = torch.rand(A.range_shape, device="cuda", generator=g)
    angles, rows = sample_angles(A, n_fp_angles), sample_rows(A.range_shape[0])
    x_host, w_host = x.cpu().numpy(), w.cpu().numpy()
    want_fp = P.fp_angles(x_host, angles, rows)
    # device-resident path
    y = A(x)
    assert A.astra_projector.info().fp_uses_tma == 1
    y_host = y.cpu().numpy()
    del y
    check_fp(want_fp, y_host, angles, rows, "device")
    xb = A.T(w)
    assert A.astra_projector.info().bp_uses_tma == 1
    check_bp(P, w_host, xb.cpu().numpy(), A.domain_shape, "device")
    del xb, x, w
    torch.cuda.empty_cache()
    if host_arrays:
        # host-array path (chunked copy / compute pipeline on sub-projectors): same numbers
        y2 = A(x_host)
        assert A.astra_projector.info().host_pipelined == 1
        check_fp(want_fp, y2, angles, rows, "host pipeline")
        np.testing.assert_allclose(y2[:, angles, :], y_host[:, angles, :], rtol=0, atol=1e-5 * np.abs(y_host).max())
        xb2 = A.T(w_host)
        assert A.astra_projector.info().host_pipelined == 1
        check_bp(P, w_host, xb2, A.domain_shape, "host pipeline")


def test_configs2_cone_vec_512_against_oracle():
    vg, pg = cfg3()
    run_config(vg, pg, n_fp_angles=12, host_arrays=True)


def test_configs3_cone_1024_against_oracle():
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~50 GB of device memory")
    vg, pg = cfg4()
    run_config(vg, pg, n_fp_angles=10, host_arrays=False)


def test_configs1_parallel_256_against_oracle():
    """BASELINE.json configs[1]: parallel3d 256^3, 180 angles, 256 x 256 - small enough for every angle."""
    vg = ts.volume(shape=256)
    pg = ts.parallel(angles=180, shape=(256, 256))
    A = ts.operator(vg, pg)
    P = oracle_of(A)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand(A.domain_shape, device="cuda", generator=g)
    w = torch.rand(A.range_shape, device="cuda", generator=g)
    y = A(x).cpu().numpy()
    angles, rows = sorted(set(list(range(0, 180, 9)) + [44, 45, 46, 134, 135, 136, 179])), list(range(256))
    check_fp(P.fp_angles(x.cpu().numpy(), angles, rows), y, angles, rows, "device")
    check_bp(P, w.cpu().numpy(), A.T(w).cpu().numpy(), A.domain_shape, "device")
