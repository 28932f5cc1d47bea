"""SIRT on top of an operator: the caller on either side of the projection path
in every benchmark of the reference (``README.md:150-164``,
``notebooks/sirt_benchmark.py:116-139``; the ``ts_algorithms.sirt`` call of
``doc/intro/fast_reconstruction.rst``).

For float32 CUDA tensors the iterations run inside the C library
(``tsp_sirt``): the residual ``R * (A x - y)`` is formed in the forward
projector's store and the update ``x -= C * A^T r`` in the backprojector's
store, so an iteration is exactly one FP and one BP pass with no elementwise
kernels in between.  Other inputs take the reference's explicit loop.
"""
import numpy as np

import tomosipo_b200 as ts
from .Operator import Operator


def _weights(A, like, eps):
    import torch

    ones_v = torch.ones(tuple(A.domain_shape), device=like.device)
    ones_p = torch.ones(tuple(A.range_shape), device=like.device)
    R = A(ones_v)
    R[R < eps] = float("inf")
    R.reciprocal_()
    C = A.T(ones_p)
    C[C < eps] = float("inf")
    C.reciprocal_()
    return R, C


def sirt(A, y, num_iterations, x_init=None, eps=None):
    """``num_iterations`` of ``x += C * A.T(R * (y - A(x)))`` with ``R = 1/A(1)``, ``C = 1/A.T(1)``.

    ``y``: torch tensor (CUDA: fused path; CPU: explicit loop through the host
    path of the library) or numpy array (explicit loop).  Returns the same kind.
    """
    eps = ts.epsilon if eps is None else eps
    try:
        import torch
    except ModuleNotFoundError:  # pragma: no cover
        torch = None
    if torch is None or not isinstance(y, torch.Tensor):
        y = np.asarray(y, dtype=np.float32)
        with np.errstate(divide="ignore"):
            R = np.minimum(1 / A(np.ones(A.domain_shape, np.float32)), 1 / eps)
            C = np.minimum(1 / A.T(np.ones(A.range_shape, np.float32)), 1 / eps)
        x = np.zeros(A.domain_shape, np.float32) if x_init is None else np.array(x_init, dtype=np.float32)
        for _ in range(num_iterations):
            x += C * A.T(R * (y - A(x)))
        return x

    y = y.to(torch.float32).contiguous()
    R, C = _weights(A, y, eps)
    x = torch.zeros(tuple(A.domain_shape), device=y.device) if x_init is None else x_init.to(torch.float32).contiguous().clone()
    fused = y.is_cuda and isinstance(A, Operator) and not A.additive
    if fused:
        y_tmp = torch.empty_like(y)
        with torch.cuda.device_of(y):
            stream = torch.cuda.current_stream(y.device).cuda_stream
            A.astra_projector.sirt(x.data_ptr(), y.data_ptr(), R.data_ptr(), C.data_ptr(), y_tmp.data_ptr(),
                                   num_iterations, device=y.device.index, stream=stream)
        return x
    y_tmp = torch.empty_like(y)
    x_tmp = torch.empty_like(x)
    for _ in range(num_iterations):
        A(x, out=y_tmp)
        y_tmp -= y
        y_tmp *= R
        A.T(y_tmp, out=x_tmp)
        x_tmp *= C
        x -= x_tmp
    return x


def _fdk_angle_table(A):
    """Per-angle FDK constants from the operator's cone-beam vectors (fp64, host)."""
    pg = A.astra_compat_pg.to_vec()
    if not ts.geometry.is_cone(pg):
        raise TypeError("FDK needs a cone-beam projection geometry.")
    vg = A.astra_compat_vg
    src, det, u, v = (np.asarray(a, dtype=np.float64) for a in (pg.src_pos, pg.det_pos, pg.det_u, pg.det_v))
    pu, pv = np.linalg.norm(u, axis=1), np.linalg.norm(v, axis=1)
    n = np.cross(u, v)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    sdd = np.abs(np.sum((det - src) * n, axis=1))
    sod = np.abs(np.sum((np.asarray(vg.pos, dtype=np.float64) - src) * n, axis=1))
    foot = src + np.sum((det - src) * n, axis=1, keepdims=True) * n   # principal point on the detector plane
    ppu = np.sum((foot - det) * u, axis=1) / pu ** 2                  # in pixels, relative to the detector centre
    ppv = np.sum((foot - det) * v, axis=1) / pv ** 2
    vox = float(np.prod(np.asarray(vg.voxel_size, dtype=np.float64)))
    return dict(pu=pu, pv=pv, sdd=sdd, sod=sod, ppu=ppu, ppv=ppv, vox=vox)


def fdk(A, y, angle_weights=None):
    """Feldkamp-Davis-Kress reconstruction for a circular cone-beam operator ``A``.

    Replaces ``astra.experimental.accumulate_FDK`` behind ``ts.astra.fdk``
    (reference ``tomosipo/astra.py:374-406``): cosine pre-weighting, Ram-Lak
    ramp filter along det_u (FFT, zero-padded to >= 2U), then the library's
    backprojector, whose cone weight ``SDD^2 / (|u||v| (SOD - depth)^2)`` is the
    FDK distance weight up to the per-angle constant applied here.

    ``angle_weights``: integration weight per angle (radians); default
    ``2 pi / num_angles`` (full, uniformly sampled circle).  Short scans need
    Parker weights supplied by the caller.  ``y``: torch tensor or numpy array
    ``(V, angles, U)``; the result is of the same kind.
    """
    import torch

    is_np = not isinstance(y, torch.Tensor)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float32) if is_np else y)
    dev = yt.device if yt.is_cuda else torch.device("cuda", torch.cuda.current_device())
    yt = yt.to(device=dev, dtype=torch.float32)
    nv, na, nu = yt.shape
    t = _fdk_angle_table(A)
    f64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
    pu, pv, sdd, sod, ppu, ppv = (f64(t[k]) for k in ("pu", "pv", "sdd", "sod", "ppu", "ppv"))
    w_angle = f64(np.full(na, 2 * np.pi / na) if angle_weights is None else np.asarray(angle_weights, dtype=np.float64))

    # 1. cosine weighting  SDD / sqrt(SDD^2 + U^2 + V^2)  (physical detector coordinates about the principal point)
    iu = torch.arange(nu, device=dev, dtype=torch.float64) + 0.5 - nu / 2
    iv = torch.arange(nv, device=dev, dtype=torch.float64) + 0.5 - nv / 2
    U = (iu[None, :] - ppu[:, None]) * pu[:, None]                       # [A, U]
    V = (iv[:, None] - ppv[None, :]) * pv[None, :]                       # [V, A]
    cosw = sdd[None, :, None] / torch.sqrt(sdd[None, :, None] ** 2 + U[None] ** 2 + V[:, :, None] ** 2)
    p1 = yt * cosw.to(torch.float32)
    del cosw

    # 2. ramp filter: q = (1 / tau) * (p1 conv g),  g[0] = 1/4, g[n odd] = -1 / (pi n)^2,  tau = pixel pitch at the isocentre
    nfft = 1 << int(np.ceil(np.log2(2 * nu)))
    k = torch.arange(nfft, device=dev, dtype=torch.float64)
    k = torch.minimum(k, nfft - k)
    g = torch.where(k == 0, torch.full_like(k, 0.25), torch.where(k % 2 == 1, -1.0 / (np.pi * k) ** 2, torch.zeros_like(k)))
    G = torch.fft.rfft(g).real.to(torch.float32)
    q = torch.fft.irfft(torch.fft.rfft(p1, n=nfft, dim=-1) * G, n=nfft, dim=-1)[..., :nu]
    del p1

    # 3. per-angle constant:  (d_beta / 2) * SOD^2 |u||v| / (SDD^2 V_vox) / tau
    tau = pu * sod / sdd
    c = 0.5 * w_angle * sod ** 2 * pu * pv / (sdd ** 2 * t["vox"]) / tau
    q = (q * c.to(torch.float32)[None, :, None]).contiguous()
    rec = A.T(q)
    return rec.cpu().numpy() if is_np else rec.to(y.device)
