"""Every kernel variant of the hot path against the fp64 oracle and against each other.

The library picks its FP / BP kernels from the geometry (TMA-staged vs. plain
loads, rows per thread, z-run length, shared-memory pitch).  Tuning
environment variables (read by libtsproj at projector creation / launch) force
each variant here so that all of them stay parity-green, including the
fallback paths a default run of the large configurations never takes.

Tolerance: relative L2 <= 1e-5 against the fp64 oracle (BASELINE.json
north_star); variants among themselves differ only by fp32 summation order
(<= 2e-6).
"""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from oracle import oracle as O

from .test_gpu_kernels import TOL, _run, cases, make, rel_l2

pytestmark = pytest.mark.gpu


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


FP_VARIANTS = [
    {},                                   # default plan
    {"TSP_FP_NO_TMA": 1},                 # plain-load kernels (fp_cols_kernel / fp_kernel)
    {"TSP_FP_R": 4},
    {"TSP_FP_R": 8},
    {"TSP_FP_ONE_PITCH": 1},
    {"TSP_FP_BOX_SHRINK": 3},             # undersized box: some slices take the "unfit" global path
    {"TSP_FP_BOX_SHRINK": 40},            # box far too small: every slice unfit
    {"TSP_FP_STAGES": 2},
    {"TSP_FP_SEGMENTS": 2},               # marching axis in segments, later ones accumulate (large volumes: L2-sized slabs)
    {"TSP_FP_SEGMENTS": 3},
]
BP_VARIANTS = [{}, {"TSP_BP_NO_TMA": 1}, {"TSP_BP_ZPT": 1}, {"TSP_BP_ZPT": 4}, {"TSP_BP_ZPT": 8}, {"TSP_BP_ZPT": 16},
               {"TSP_BP_ZPT": 32}, {"TSP_BP_ZPT": 64}, {"TSP_BP_ROWS": 2}]


def _ids(v):
    return "default" if not v else "-".join(f"{k[4:].lower()}{val}" for k, val in v.items())


@pytest.mark.parametrize("variant", FP_VARIANTS, ids=_ids)
@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_fp_variant_matches_oracle(case, variant):
    name, kind, vs, win, ds, vec = case
    rng = np.random.default_rng(0)
    x = rng.random(vs).astype(np.float32)
    with env(**variant):
        P, Q = make(kind, vs, win, ds, vec)
        _, y = _run(P, 0, x, np.zeros(Q.proj_shape, np.float32))
    ref = Q.fp(x.astype(np.float64))
    assert rel_l2(y, ref) <= TOL, rel_l2(y, ref)


@pytest.mark.parametrize("variant", BP_VARIANTS, ids=_ids)
@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_bp_variant_matches_oracle(case, variant):
    name, kind, vs, win, ds, vec = case
    rng = np.random.default_rng(0)
    with env(**variant):
        P, Q = make(kind, vs, win, ds, vec)
        y = rng.random(Q.proj_shape).astype(np.float32)
        x, _ = _run(P, 1, np.zeros(vs, np.float32), y)
    ref = Q.bp(y.astype(np.float64))
    assert rel_l2(x, ref) <= TOL, rel_l2(x, ref)


def test_default_plan_uses_tma_kernels():
    """The TMA-staged kernels are the ones that run for TMA-compatible shapes (no silent fallback)."""
    name, kind, vs, win, ds, vec = cases()[0]
    P, Q = make(kind, vs, win, ds, vec)
    x = np.ones(vs, np.float32)
    _, y = _run(P, 0, x, np.zeros(Q.proj_shape, np.float32))
    assert P.info().fp_uses_tma == 1
    _run(P, 1, x, y)
    assert P.info().bp_uses_tma == 1
    with env(TSP_FP_NO_TMA=1, TSP_BP_NO_TMA=1):
        P2, _ = make(kind, vs, win, ds, vec)
        _, y2 = _run(P2, 0, x, np.zeros(Q.proj_shape, np.float32))
        assert P2.info().fp_uses_tma == 0
        _run(P2, 1, x, y2)
        assert P2.info().bp_uses_tma == 0


def _cfg3_like(n, n_angles):
    det = (n, 3 * n // 2)
    vec = O.cone_vectors(np.linspace(0, 2 * np.pi, n_angles, endpoint=False), 2.8125 / det[1], 1.875 / det[0], 4.0, 2.0)
    return O.CONE_VEC, (n, n, n), [(-.5, .5)] * 3, det, vec


def test_variants_agree_at_256():
    """cfg-3 geometry at 256^3 x 90 angles (too slow for the fp64 oracle in a unit test): TMA-staged and
    plain-load kernels must agree to fp32 summation-order noise, in both directions."""
    import torch
    from tomosipo_b200 import _backend as B

    kind, vs, win, ds, vec = _cfg3_like(256, 90)
    torch.manual_seed(0)
    x = torch.rand(vs, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for variant in ({}, {"TSP_FP_NO_TMA": 1, "TSP_BP_NO_TMA": 1}, {"TSP_FP_R": 4, "TSP_BP_ZPT": 16}):
        with env(**variant):
            P = B.Projector(kind, vs, win, ds, vec)
            y = torch.empty(P.proj_shape, device="cuda")
            xb = torch.empty(P.vol_shape, device="cuda")
            P.project(B.FP, False, x.data_ptr(), y.data_ptr(), B.MEM_DEVICE, 0, s)
            P.project(B.BP, False, xb.data_ptr(), y.data_ptr(), B.MEM_DEVICE, 0, s)
            torch.cuda.synchronize()
        outs.append((y.double(), xb.double()))
    for y, xb in outs[1:]:
        assert float((y - outs[0][0]).norm() / outs[0][0].norm()) <= 2e-6
        assert float((xb - outs[0][1]).norm() / outs[0][1].norm()) <= 2e-6
    # linearity at full kernel size: A(2x + x') == 2 A(x) + A(x')
    with env():
        P = B.Projector(kind, vs, win, ds, vec)
        x2 = torch.rand(vs, device="cuda")
        ya, yb, yc = (torch.empty(P.proj_shape, device="cuda") for _ in range(3))
        xs = 2 * x + x2
        for src, dst in ((x, ya), (x2, yb), (xs, yc)):
            P.project(B.FP, False, src.data_ptr(), dst.data_ptr(), B.MEM_DEVICE, 0, s)
        torch.cuda.synchronize()
        lin = (2 * ya + yb).double()
        assert float((yc.double() - lin).norm() / lin.norm()) <= 2e-6


def test_fused_sirt_matches_explicit_loop_and_oracle():
    """tsp_sirt (residual / update fused into the FP / BP stores) == the reference loop
    (notebooks/sirt_benchmark.py:130-136) with separate elementwise passes, and == the fp64 oracle loop."""
    import torch
    import tomosipo_b200 as ts
    from tomosipo_b200.algorithms import sirt

    n = 32
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=24, shape=(n, 48), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
    A = ts.operator(vg, pg)
    phantom = torch.zeros(A.domain_shape, device="cuda")
    phantom[5:12, 5:12, 5:12] = 1.0
    y = A(phantom)
    x_fused = sirt(A, y, 6)

    # explicit loop on the GPU
    eps = ts.epsilon
    R = A(torch.ones(A.domain_shape, device="cuda")); R[R < eps] = float("inf"); R.reciprocal_()
    C = A.T(torch.ones(A.range_shape, device="cuda")); C[C < eps] = float("inf"); C.reciprocal_()
    x = torch.zeros(A.domain_shape, device="cuda")
    for _ in range(6):
        x -= C * A.T(R * (A(x) - y))
    assert float((x_fused - x).norm() / x.norm()) <= 2e-6

    # fp64 oracle loop
    from .test_operator_gpu import oracle_of

    Q = oracle_of(A)
    yq = Q.fp(phantom.cpu().numpy().astype(np.float64))
    with np.errstate(divide="ignore"):
        Rq = np.minimum(1 / Q.fp(np.ones(A.domain_shape)), 1 / eps)
        Cq = np.minimum(1 / Q.bp(np.ones(A.range_shape)), 1 / eps)
    xq = np.zeros(A.domain_shape)
    for _ in range(6):
        xq += Cq * Q.bp(Rq * (yq - Q.fp(xq)))
    assert rel_l2(x_fused.cpu().numpy(), xq) <= 1e-4

    # numpy in, numpy out (explicit loop through the host path)
    x_np = sirt(A, y.cpu().numpy(), 6)
    assert isinstance(x_np, np.ndarray) and rel_l2(x_np, xq) <= 1e-4


def test_project_fused_halves_match_explicit_passes():
    """tsp_project_fused: FP store forms mul * (A x - sub), BP store forms x -= mul * A^T y
    (the two halves of tsp_sirt, used by the sharded SIRT)."""
    import torch
    import tomosipo_b200 as ts
    from tomosipo_b200 import _backend as B

    vg = ts.volume(shape=(40, 36, 44), size=(1.0, 0.9, 1.1))
    pg = ts.cone(angles=19, shape=(40, 60), size=(2.0, 3.0), src_orig_dist=4, src_det_dist=7)
    A = ts.operator(vg, pg)
    P = A.astra_projector
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(A.domain_shape, device="cuda", generator=g)
    y = torch.rand(A.range_shape, device="cuda", generator=g)
    R = torch.rand(A.range_shape, device="cuda", generator=g)
    C = torch.rand(A.domain_shape, device="cuda", generator=g)
    s = torch.cuda.current_stream().cuda_stream
    r = torch.empty_like(y)
    P.project_fused(B.FP, x.data_ptr(), r.data_ptr(), y.data_ptr(), R.data_ptr(), device=0, stream=s)
    torch.testing.assert_close(r, R * (A(x) - y), rtol=1e-5, atol=1e-6)
    x2 = x.clone()
    P.project_fused(B.BP, x2.data_ptr(), r.data_ptr(), None, C.data_ptr(), device=0, stream=s)
    torch.testing.assert_close(x2, x - C * A.T(r), rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        P.project_fused(B.FP, x.data_ptr(), r.data_ptr(), None, R.data_ptr(), device=0, stream=s)
    with pytest.raises(ValueError):
        P.project_fused(B.BP, x.data_ptr(), r.data_ptr(), y.data_ptr(), C.data_ptr(), device=0, stream=s)


def test_segmented_fp_keeps_add_mode_and_the_fused_residual():
    """A forward projection split into segments of the marching axis (TSP_FP_SEGMENTS; automatic for volumes whose
    per-row-tile slab outgrows L2): ADD adds to the caller's data once, the fused residual is formed from the whole sum."""
    import torch
    import tomosipo_b200 as ts
    from tomosipo_b200 import _backend as B

    vg = ts.volume(shape=(40, 72, 68), size=(1.0, 1.8, 1.7))
    pg = ts.cone(angles=23, shape=(40, 96), size=(2.0, 4.8), src_orig_dist=5, src_det_dist=8)
    g = torch.Generator(device="cuda").manual_seed(2)
    s = torch.cuda.current_stream().cuda_stream
    A1 = ts.operator(vg, pg)
    x = torch.rand(A1.domain_shape, device="cuda", generator=g)
    y = torch.rand(A1.range_shape, device="cuda", generator=g)
    R = torch.rand(A1.range_shape, device="cuda", generator=g)
    ref = A1(x)
    with env(TSP_FP_SEGMENTS=3):
        A, Aadd = ts.operator(vg, pg), ts.operator(vg, pg, additive=True)
        n0 = A.astra_projector.info().kernel_launches
        torch.testing.assert_close(A(x), ref, rtol=1e-5, atol=1e-5 * float(ref.max()))
        assert A.astra_projector.info().kernel_launches - n0 >= 3          # one launch per segment (and group)
        acc = y.clone()
        Aadd(x, out=acc)
        torch.testing.assert_close(acc, y + ref, rtol=1e-5, atol=1e-5 * float(ref.max()))
        r = torch.empty_like(y)
        A.astra_projector.project_fused(B.FP, x.data_ptr(), r.data_ptr(), y.data_ptr(), R.data_ptr(), device=0, stream=s)
        torch.testing.assert_close(r, R * (ref - y), rtol=1e-5, atol=1e-5 * float(ref.max()))


@pytest.mark.parametrize("kindname", ["par_slab", "cone_thin"])
def test_thin_batched_matches_oracle_per_item(kindname):
    """cfg-5 shapes: a batch of thin volumes in ONE C-ABI call (tsp_project(batch=B), thin_kernels.cuh)
    == the oracle applied to every batch element (the reference's loop, torch_support.py:49-53)."""
    import torch
    from tomosipo_b200 import _backend as B

    if kindname == "par_slab":   # exact 2-D problem: 1 x 48 x 52 slab, one detector row
        name, kind, vs, win, ds, vec = [c for c in cases() if c[0] == "slab"][0]
    else:                        # 3 slices, 2 detector rows, cone beam, ragged sizes
        kind, vs, win, ds = O.CONE_VEC, (3, 45, 50), [(-1.0, 1.0), (-.9, .9), (-.06, .06)], (2, 71)
        vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 37, endpoint=False), 0.05, 0.07, 6.0, 3.0)
    P, Q = make(kind, vs, win, ds, vec)
    nb = 5
    rng = np.random.default_rng(3)
    x = rng.random((nb,) + tuple(vs)).astype(np.float32)
    y = rng.random((nb,) + tuple(Q.proj_shape)).astype(np.float32)
    s = torch.cuda.current_stream().cuda_stream
    dx, dy = torch.from_numpy(x).cuda(), torch.empty((nb,) + tuple(Q.proj_shape), device="cuda")
    P.project(B.FP, False, dx.data_ptr(), dy.data_ptr(), B.MEM_DEVICE, 0, s, batch=nb)
    dy2, dx2 = torch.from_numpy(y).cuda(), torch.empty((nb,) + tuple(vs), device="cuda")
    P.project(B.BP, False, dx2.data_ptr(), dy2.data_ptr(), B.MEM_DEVICE, 0, s, batch=nb)
    torch.cuda.synchronize()
    for b in range(nb):
        assert rel_l2(dy[b].cpu().numpy(), Q.fp(x[b].astype(np.float64))) <= TOL
        assert rel_l2(dx2[b].cpu().numpy(), Q.bp(y[b].astype(np.float64))) <= TOL
    # additive batch and the generic kernels on the same shapes agree
    with env(TSP_NO_THIN=1):
        P2, _ = make(kind, vs, win, ds, vec)
        dy3 = torch.empty_like(dy)
        P2.project(B.FP, False, dx.data_ptr(), dy3.data_ptr(), B.MEM_DEVICE, 0, s, batch=nb)
        torch.cuda.synchronize()
    assert float((dy3 - dy).norm() / dy.norm()) <= 2e-6
    P.project(B.FP, True, dx.data_ptr(), dy.data_ptr(), B.MEM_DEVICE, 0, s, batch=nb)
    torch.cuda.synchronize()
    assert float((dy - 2 * dy3).norm() / dy3.norm()) <= 4e-6


def test_fdk_reconstructs_phantom():
    """FDK (ts.astra.fdk / algorithms.fdk; reference tomosipo/astra.py:374-406, tests/test_astra.py:66-84):
    a full-circle cone-beam scan of the hollow box is reconstructed to the phantom's scale."""
    import torch
    import tomosipo_b200 as ts
    from tomosipo_b200.algorithms import fdk

    n = 64
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=192, shape=(96, 128), size=(1.8, 2.4), src_orig_dist=6, src_det_dist=9)
    A = ts.operator(vg, pg)
    x = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data.copy()).cuda()
    y = A(x)
    rec = fdk(A, y)
    assert rec.shape == x.shape and rec.is_cuda
    # smooth discretisation error only: global relative L2 and the plateau value inside the box wall
    err = float((rec - x).norm() / x.norm())
    assert err < 0.30, err
    wall = x > 0.5
    interior = torch.zeros_like(wall)
    interior[4:-4, 4:-4, 4:-4] = True
    assert abs(float(rec[wall].mean()) - 1.0) < 0.08
    hole = (x < 0.5) & interior
    assert abs(float(rec[hole].mean())) < 0.08

    # legacy Data interface, numpy arrays (reference tests/test_astra.py::test_fdk)
    vd, pd = ts.data(vg), ts.data(pg, y.cpu().numpy())
    ts.astra.fdk(vd, pd)
    assert rel_l2(vd.data, rec.cpu().numpy().astype(np.float64)) < 1e-5


@pytest.mark.parametrize("kindname", ["cone", "parallel"])
def test_host_array_pipeline_matches_single_shot(kindname):
    """Host arrays (the reference's numpy path, README.md:139-164): the chunked copy/compute pipeline
    (z-slabs + their detector rows for BP, detector row blocks for FP) == one upload / launch / download,
    and both == the fp64 oracle."""
    from tomosipo_b200 import _backend as B

    n = 96
    det = (96, 128)
    if kindname == "cone":
        kind = O.CONE_VEC
        vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 40, endpoint=False), 2.8 / det[1], 2.1 / det[0], 4.0, 2.0)
    else:
        kind = O.PARALLEL_VEC
        vec = O.parallel_vectors(np.linspace(0, np.pi, 40, endpoint=False), 1.6 / det[1], 1.2 / det[0])
    win = [(-.5, .5)] * 3
    rng = np.random.default_rng(5)
    x = rng.random((n, n, n)).astype(np.float32)
    outs = []
    for variant in ({"TSP_HOST_PIPELINE_MIN_MB": 0}, {"TSP_HOST_NO_PIPELINE": 1}):
        with env(**variant):
            P, Q = make(kind, (n, n, n), win, det, vec)
            y = np.zeros(Q.proj_shape, np.float32)
            P.project(B.FP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
            piped_fp = P.info().host_pipelined
            xb = np.zeros((n, n, n), np.float32)
            P.project(B.BP, False, xb.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
            assert P.info().host_pipelined == piped_fp == (1 if "TSP_HOST_PIPELINE_MIN_MB" in variant else 0)
        outs.append((y, xb))
    assert rel_l2(outs[0][0], outs[1][0].astype(np.float64)) <= 2e-6
    assert rel_l2(outs[0][1], outs[1][1].astype(np.float64)) <= 2e-6
    assert rel_l2(outs[0][0], Q.fp(x.astype(np.float64))) <= TOL
    assert rel_l2(outs[0][1], Q.bp(outs[0][0].astype(np.float64))) <= TOL


@pytest.mark.parametrize("kindname", ["cone", "parallel"])
def test_host_array_pipeline_out_of_core_ring(kindname):
    """Bounded device memory (ASTRA's CompositeGeometryManager splits jobs that do not fit,
    reference doc/topics/operator.rst:226-261): with a device-memory budget below the size of the two arrays the
    host-array pipeline runs out of a ring of three chunk-sized buffer pairs; same numbers as the whole-array mode
    and the fp64 oracle.  A budget too small for the ring itself is a MemoryError, not a crash."""
    from tomosipo_b200 import _backend as B

    n, det = 256, (256, 384)
    if kindname == "cone":
        kind = O.CONE_VEC
        vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 60, endpoint=False), 2.8125 / det[1], 1.875 / det[0], 4.0, 2.0)
    else:
        kind = O.PARALLEL_VEC
        vec = O.parallel_vectors(np.linspace(0, np.pi, 60, endpoint=False), 1.6 / det[1], 1.2 / det[0])
    win = [(-.5, .5)] * 3
    x = np.random.default_rng(6).random((n, n, n)).astype(np.float32)
    total_mb = (x.size + det[0] * 60 * det[1]) * 4 / 2 ** 20          # 64 + 22.5 MB
    outs = []
    for cap in (None, 64):                                             # MB; None: whole-array mode
        kw = {"TSP_HOST_PIPELINE_MIN_MB": 0, "TSP_HOST_CHUNKS": 8}
        if cap:
            kw["TSP_HOST_MEM_CAP_MB"] = cap
            assert cap < total_mb
        with env(**kw):
            P, Q = make(kind, (n, n, n), win, det, vec)
            y = np.zeros(Q.proj_shape, np.float32)
            P.project(B.FP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
            assert P.info().host_pipelined == 1 and P.info().host_ring == (1 if cap else 0)
            xb = np.zeros((n, n, n), np.float32)
            P.project(B.BP, False, xb.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)
            assert P.info().host_pipelined == 1 and P.info().host_ring == (1 if cap else 0)
        outs.append((y, xb))
    assert rel_l2(outs[1][0], outs[0][0].astype(np.float64)) <= 2e-6
    assert rel_l2(outs[1][1], outs[0][1].astype(np.float64)) <= 2e-6
    sel = [0, 7, 8, 31, 59]
    assert rel_l2(outs[1][0][:, sel, :], Q.fp_angles(x, sel)) <= TOL
    for z, yy, xx in (((0, 8), (0, 8), (0, 32)), ((124, 132), (100, 108), (64, 96)), ((250, 256), (248, 256), (224, 256))):
        assert rel_l2(outs[1][1][z[0]:z[1], yy[0]:yy[1], xx[0]:xx[1]], Q.bp_window(outs[1][0], z, yy, xx)) <= TOL
    with env(TSP_HOST_PIPELINE_MIN_MB=0, TSP_HOST_CHUNKS=8, TSP_HOST_MEM_CAP_MB=8):
        P, Q = make(kind, (n, n, n), win, det, vec)
        with pytest.raises(MemoryError):
            P.project(B.FP, False, x.ctypes.data, np.zeros(Q.proj_shape, np.float32).ctypes.data, B.MEM_HOST, 0, 0)


def test_host_arrays_divided_over_several_gpus():
    """`ts.astra.set_gpu_index([0, 1, ...])` (the reference's astra.set_gpu_index for ndarray inputs,
    doc/topics/operator.rst:233-245): tsp_project_multi deals the chunks of the host plan out to the GPUs."""
    import torch

    import tomosipo_b200 as ts

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    vg = ts.volume(shape=128, size=1)
    pg = ts.cone(angles=60, shape=(128, 192), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
    A = ts.operator(vg, pg)
    x = np.random.default_rng(8).random(A.domain_shape).astype(np.float32)
    with env(TSP_HOST_PIPELINE_MIN_MB=0):
        y1 = A(x)
        xb1 = A.T(y1)
        try:
            ts.astra.set_gpu_index(list(range(ngpu)))
            y2 = A(x)
            assert A.astra_projector.info().host_devices == ngpu
            xb2 = A.T(y1)
            assert A.astra_projector.info().host_devices == ngpu
        finally:
            ts.astra.set_gpu_index(None)
    assert rel_l2(y2, y1.astype(np.float64)) <= 2e-6
    assert rel_l2(xb2, xb1.astype(np.float64)) <= 2e-6


def test_thin_kernels_keep_non_finite_interior_values_local():
    """Non-finite INTERIOR voxels / pixels affect exactly the rays / voxels the tiled kernels let them affect.
    (Non-finite values in the outermost two rows / columns can additionally reach samples within one element outside
    the array in the default build of the thin kernels: see thin_kernels.cuh, THIN_STRICT_BORDERS.)"""
    import torch
    import tomosipo_b200 as ts

    vg = ts.volume(shape=(1, 64, 64), size=(1 / 64, 1, 1))
    pg = ts.parallel(angles=40, shape=(1, 96), size=(1 / 64, 1.5))
    x = torch.rand((1, 64, 64), device="cuda")
    x[0, 20, 31] = float("inf"); x[0, 40, 17] = float("nan")
    y = torch.rand((1, 40, 96), device="cuda")
    y[0, 3, 30] = float("inf"); y[0, 7, 60] = float("nan")
    A = ts.operator(vg, pg)
    fp_thin, bp_thin = A(x), A.T(y)
    with env(TSP_NO_THIN=1):
        B_ = ts.operator(vg, pg)
        fp_ref, bp_ref = B_(x), B_.T(y)
    for thin, ref in ((fp_thin, fp_ref), (bp_thin, bp_ref)):
        bad_thin, bad_ref = ~torch.isfinite(thin), ~torch.isfinite(ref)
        assert int((bad_thin & ~bad_ref).sum()) == 0          # nothing poisoned that the tiled kernels keep finite
        ok = ~bad_ref & ~bad_thin
        assert float((thin[ok] - ref[ok]).abs().max()) <= 1e-5 * float(ref[ok].abs().max())
