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
