"""FDK (reference tomosipo/astra.py:374-406 -> astra.experimental.accumulate_FDK).

CPU: the fp64 FDK oracle (oracle/fdk_oracle.py: direct spatial convolution, no FFT) against closed forms - the analytic
cone-beam projections of a uniform ball are reconstructed to the ball's density - for the full circle and for a
Parker-weighted short scan; conjugate rays' Parker weights sum to one.
GPU: tomosipo_b200.algorithms.fdk (hand-written weighting / ramp / crop kernels around cuFFT + the library's
backprojector) against that oracle, relative L2 <= 1e-4 (float32 FFT of 2U points), full and short scan."""
import numpy as np
import pytest

from oracle import fdk_oracle as F
from oracle import oracle as O

N, DET, SOD, ODD = 40, (40, 60), 6.0, 3.0
PU, PV = 2.4 / DET[1], 1.8 / DET[0]
LO, HI = [-0.5] * 3, [0.5] * 3
CENTRE, RADIUS = np.array([0.12, -0.1, 0.03]), 0.27


def ball_projections(vec):
    """Exact line integrals of a unit-density ball through the pixel centres of every angle."""
    V, U = DET
    y = np.zeros((V, vec.shape[0], U))
    iu, iv = np.arange(U) + 0.5 - U / 2, np.arange(V) + 0.5 - V / 2
    for a in range(vec.shape[0]):
        s = vec[a, 0:3]
        pix = vec[a, 3:6][None, None, :] + iu[None, :, None] * vec[a, 6:9] + iv[:, None, None] * vec[a, 9:12]
        d = pix - s
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        w = CENTRE - s
        dist2 = (w * w).sum(-1) - ((w * d).sum(-1)) ** 2
        y[:, a, :] = 2 * np.sqrt(np.maximum(RADIUS ** 2 - dist2, 0))
    return y


def regions():
    z, y, x = np.meshgrid(*[(np.arange(N) + 0.5) / N - 0.5] * 3, indexing="ij")
    r2 = (x - CENTRE[0]) ** 2 + (y - CENTRE[1]) ** 2 + (z - CENTRE[2]) ** 2
    return r2 < (RADIUS - 0.08) ** 2, r2 > (RADIUS + 0.1) ** 2


def short_scan_angles(n=72):
    fan = np.arctan(1.2 / (SOD + ODD))
    return np.arange(n) * (np.pi + 2 * fan + 0.25) / n


def test_fdk_oracle_reconstructs_ball_density_full_circle():
    vec = O.cone_vectors(np.linspace(0, 2 * np.pi, 100, endpoint=False), PU, PV, SOD, ODD)
    rec = F.fdk(vec, (N, N, N), LO, HI, DET, ball_projections(vec))
    inside, outside = regions()
    assert abs(rec[inside].mean() - 1.0) < 5e-3               # the scale of FDK: a factor 2 or a pitch slip shows here
    assert rec[inside].std() < 0.02 and np.abs(rec[outside]).mean() < 0.02


def test_fdk_oracle_short_scan_with_parker_weights():
    vec = O.cone_vectors(short_scan_angles(), PU, PV, SOD, ODD)
    rec = F.fdk(vec, (N, N, N), LO, HI, DET, ball_projections(vec), short_scan=True)
    inside, outside = regions()
    assert abs(rec[inside].mean() - 1.0) < 5e-3
    assert rec[inside].std() < 0.02 and np.abs(rec[outside]).mean() < 0.03   # (the wrong fan-angle sign gives std 0.04)


def test_parker_weights_of_conjugate_rays_sum_to_one():
    beta = short_scan_angles(4000)
    step = beta[1] - beta[0]
    delta = 0.5 * (beta[-1] + step - np.pi)
    gam = np.linspace(-0.9 * delta, 0.9 * delta, 31)
    w = F.parker_weights(beta, gam)
    for j, g in enumerate(gam):
        for i in range(0, len(beta), 371):
            b2 = beta[i] + np.pi + 2 * g                       # the same line, measured from the other side
            if b2 > beta[-1]:
                continue
            i2 = b2 / step
            lo = int(np.floor(i2))
            if lo + 1 >= len(beta):
                continue
            w2 = (1 - (i2 - lo)) * w[lo, len(gam) - 1 - j] + (i2 - lo) * w[lo + 1, len(gam) - 1 - j]
            assert abs(w[i, j] + w2 - 1) < 2e-3, (i, j)


@pytest.mark.gpu
@pytest.mark.parametrize("short", [False, True], ids=["full", "short"])
def test_fdk_matches_fp64_oracle(short):
    import torch

    import tomosipo_b200 as ts
    from tomosipo_b200.algorithms import fdk

    from .test_operator_gpu import rel_l2

    angles = short_scan_angles() if short else 100
    vg = ts.volume(shape=N, size=1)
    pg = ts.cone(angles=angles, shape=DET, size=(1.8, 2.4), src_orig_dist=SOD, src_det_dist=SOD + ODD)
    A = ts.operator(vg, pg)
    vec = A.astra_compat_pg.to_vec().to_astra()["Vectors"]
    y = ball_projections(np.asarray(vec)) + 0.05 * np.random.default_rng(1).random(tuple(A.range_shape))
    want = F.fdk(vec, (N, N, N), LO, HI, DET, y, short_scan=short)
    got = fdk(A, torch.from_numpy(y.astype(np.float32)).cuda(), short_scan=short)
    assert got.is_cuda
    assert rel_l2(got.cpu().numpy(), want) < 1e-4
    # numpy in, numpy out; the legacy Data interface accumulates like astra.experimental.accumulate_FDK
    got_np = fdk(A, y.astype(np.float32), short_scan=short)
    assert isinstance(got_np, np.ndarray) and rel_l2(got_np, want) < 1e-4
    if not short:
        vd, pd = ts.data(vg, np.ones(tuple(A.domain_shape), np.float32)), ts.data(pg, y.astype(np.float32))
        ts.astra.fdk(vd, pd)
        assert rel_l2(vd.data - 1.0, want) < 1e-4
