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
        # same clamping as the tensor path (notebooks/sirt_benchmark.py:116-128): weight 0 where a row / column of A
        # is empty.  (The README loop clamps to 1 / eps instead, README.md:151-154; both appear in the reference -
        # one convention here, so that numpy and torch inputs give the same reconstruction.)
        R = A(np.ones(A.domain_shape, np.float32))
        C = A.T(np.ones(A.range_shape, np.float32))
        with np.errstate(divide="ignore"):
            R = np.where(R < eps, 0.0, 1.0 / R).astype(np.float32)
            C = np.where(C < eps, 0.0, 1.0 / C).astype(np.float32)
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


def parker_weights(A):
    """Parker (1982) redundancy weights ``w[angle, column]`` for a short circular scan (source angles ascending,
    uniformly spaced, covering at least pi + the fan angle; any over-scan up to the full circle is used).

    Every line through the object is measured once or twice; ``w`` of the two measurements of a line sums to 1 and
    falls smoothly to 0 at both ends of the arc.  In this library's geometry the ray of source angle ``beta`` and
    fan angle ``gamma = atan(U / SDD)`` is measured again at ``(beta + pi - 2 gamma, -gamma)``, i.e. Parker's
    formulas hold with ``gamma -> -gamma``."""
    t = _fdk_angle_table(A)
    pg = A.astra_compat_pg.to_vec()
    src = np.asarray(pg.src_pos, dtype=np.float64)           # (z, y, x)
    beta = np.unwrap(np.arctan2(src[:, 2], -src[:, 1]))
    if len(beta) < 2:
        raise ValueError("A short scan needs at least two angles.")
    steps = np.diff(beta)
    step = float(np.mean(steps))
    if step < 0:
        beta, step, flip = -beta, -step, True
    else:
        flip = False
    if np.abs(steps - np.mean(steps)).max() > 1e-6 * abs(step) + 1e-12:
        raise ValueError("Parker weights need uniformly spaced angles.")
    nu = pg.det_shape[1]
    iu = np.arange(nu) + 0.5 - nu / 2
    b = (beta - beta[0])[:, None]
    rng = b[-1, 0] + step
    delta = 0.5 * (rng - np.pi)
    up = (iu[None, :] - t["ppu"][:, None]) * t["pu"][:, None]
    g = -np.arctan2(up, t["sdd"][:, None])                   # Parker's sign convention
    if flip:
        g = -g
    if delta < np.abs(g).max() - 1e-9:
        raise ValueError(f"The scan covers {rng:.4f} rad: less than pi + the fan angle ({np.pi + 2 * np.abs(g).max():.4f}).")
    with np.errstate(divide="ignore", invalid="ignore"):
        ramp_up = np.sin(0.25 * np.pi * b / (delta - g)) ** 2
        ramp_dn = np.sin(0.25 * np.pi * (np.pi + 2 * delta - b) / (delta + g)) ** 2
    w = np.where(b < 2 * delta - 2 * g, ramp_up, 1.0)
    w = np.where(b > np.pi - 2 * g, ramp_dn, w)
    return np.clip(np.nan_to_num(w, nan=0.0), 0.0, 1.0), abs(step)


def fdk(A, y, angle_weights=None, short_scan=False):
    """Feldkamp-Davis-Kress reconstruction for a circular cone-beam operator ``A``.

    Replaces ``astra.experimental.accumulate_FDK`` behind ``ts.astra.fdk``
    (reference ``tomosipo/astra.py:374-406``).  Cosine weighting, redundancy
    weighting and the FFT's zero padding are one kernel (``tsp_fdk_stage`` 0),
    the Ram-Lak multiply in the frequency domain another (stage 1), the crop of
    the padded rows with the per-angle constant a third (stage 2); the
    transforms themselves are cuFFT (``torch.fft``), the backprojection is the
    library's, whose cone weight ``SDD^2 / (|u||v| (SOD - depth)^2)`` is the FDK
    distance weight up to that per-angle constant.

    ``angle_weights``: integration weight per angle (radians); default
    ``2 pi / num_angles`` (full, uniformly sampled circle) or the angular step
    of a short scan.  ``short_scan``: weight the projections with Parker's
    redundancy weights (:func:`parker_weights`) instead of the full circle's
    1/2.  ``y``: torch tensor or numpy array ``(V, angles, U)``; the result is
    of the same kind.
    """
    import torch

    from . import _backend

    if A.additive:
        raise ValueError("FDK needs a non-additive operator.")
    is_np = not isinstance(y, torch.Tensor)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float32) if is_np else y)
    dev = yt.device if yt.is_cuda else torch.device("cuda", torch.cuda.current_device())
    yt = yt.to(device=dev, dtype=torch.float32).contiguous()
    if tuple(yt.shape) != tuple(A.range_shape):
        raise ValueError(f"Expected projections of shape {tuple(A.range_shape)}. Got {tuple(yt.shape)}")
    nv, na, nu = yt.shape
    _fdk_angle_table(A)                                      # raises TypeError for parallel beams
    red = None
    if short_scan:
        w, step = parker_weights(A)
        red = torch.as_tensor(w, dtype=torch.float32, device=dev).contiguous()
        if angle_weights is None:
            angle_weights = np.full(na, step)
    aw = None if angle_weights is None else np.ascontiguousarray(angle_weights, dtype=np.float64)
    if aw is not None and aw.shape != (na,):
        raise ValueError(f"Expected {na} angle weights. Got {aw.shape}")
    P = A.astra_projector
    nfft = 1 << int(np.ceil(np.log2(2 * nu)))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        padded = torch.empty((nv, na, nfft), dtype=torch.float32, device=dev)
        P.fdk_stage(0, yt.data_ptr(), padded.data_ptr(), nfft, 0, red.data_ptr() if red is not None else 0, None,
                    device=dev.index, stream=stream)
        spec = torch.fft.rfft(padded, dim=-1)
        del padded
        P.fdk_stage(1, 0, spec.data_ptr(), nfft // 2 + 1, nfft, 0, None, device=dev.index, stream=stream)
        filtered = torch.fft.irfft(spec, n=nfft, dim=-1)
        del spec
        q = torch.empty_like(yt)
        P.fdk_stage(2, filtered.data_ptr(), q.data_ptr(), nfft, 0, 0, aw, device=dev.index, stream=stream)
        del filtered
        rec = A.T(q)
    return rec.cpu().numpy() if is_np else rec.to(y.device)
