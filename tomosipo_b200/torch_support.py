"""PyTorch autograd support for tomographic operators.

API mirror of the reference's ``tomosipo/torch_support.py``:
``OperatorFunction``, ``to_autograd``, ``AutogradOperator``,
``autograd_operator`` and the legacy ``forward`` / ``backward`` functions.

Difference below the API: where the reference loops in Python over the leading
(batch, channel, ...) dimensions and issues one projector call per sub-tensor
(``torch_support.py:49-53,70-74``), a dense float32 batch is handed to the C
ABI in one call (``tsp_project(..., batch=B)``); the per-sub-tensor loop remains
as the general path and produces identical results.
"""
import itertools
import math
import warnings

import tomosipo_b200 as ts
from . import _backend
from .Operator import BackprojectionOperator, Operator

try:
    import torch
except ModuleNotFoundError:
    warnings.warn(
        "\n------------------------------------------------------------\n\n"
        "Cannot import torch package. \n"
        "Please make sure to install torch. \n"
        "You can install torch using: \n\n"
        " > conda install pytorch -c pytorch \n"
        "\n------------------------------------------------------------\n\n"
    )
    raise
from torch.autograd import Function


def _project_batched(operator, src, dst, extra_dims):
    """Apply ``operator`` to every sub-tensor ``src[idx]`` -> ``dst[idx]``.

    One C-ABI call when the layout allows it, otherwise the reference's loop.
    """
    if len(extra_dims) == 0:
        operator(src, out=dst)
        return
    base = operator.parent if isinstance(operator, BackprojectionOperator) else operator
    forward = not isinstance(operator, BackprojectionOperator)
    n = math.prod(extra_dims)
    if (isinstance(base, Operator) and not base.additive and n > 0 and src.dtype == torch.float32
            and not src.is_contiguous() and dst.is_contiguous()):
        # e.g. the expanded gradient of a sum: one copy of the batch (with the link's warning, once)
        # instead of one copy + one projector call per sub-tensor
        warnings.warn(
            "The parameter initial_value should be contiguous. "
            "It has been automatically made contiguous. "
            "Use `ts.link(x.contiguous())' to inhibit this warning. "
        )
        src = src.contiguous()
    dense = (
        isinstance(base, Operator)
        and not base.additive
        and n > 0
        and src.dtype == torch.float32
        and dst.dtype == torch.float32
        and src.is_contiguous()
        and dst.is_contiguous()
        and src.device == dst.device
        and tuple(src.shape[len(extra_dims):]) == tuple(operator.domain_shape)
    )
    if not dense:
        for idx in itertools.product(*(range(d) for d in extra_dims)):
            operator(src[idx], out=dst[idx])
        return
    vol, proj = (src, dst) if forward else (dst, src)
    vol, proj = vol.detach(), proj.detach()
    if src.is_cuda:
        with torch.cuda.device_of(src):
            stream = torch.cuda.current_stream(src.device).cuda_stream
            base.astra_projector.project(
                _backend.FP if forward else _backend.BP, False, vol.data_ptr(), proj.data_ptr(),
                _backend.MEM_DEVICE, device=src.device.index, stream=stream, batch=n,
            )
    else:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        base.astra_projector.project(
            _backend.FP if forward else _backend.BP, False, vol.data_ptr(), proj.data_ptr(),
            _backend.MEM_HOST, device=device, stream=0, batch=n,
        )


class OperatorFunction(Function):
    """``torch.autograd.Function`` whose backward pass is the transposed operator."""

    @staticmethod
    def forward(ctx, input, operator, num_extra_dims=0, is_2d=False):
        extra_dims = input.size()[:num_extra_dims]
        if input.requires_grad:
            ctx.operator = operator
            ctx.extra_dims = extra_dims
            ctx.is_2d = is_2d

        expected_ndim = (2 if is_2d else 3) + num_extra_dims
        assert input.ndim == expected_ndim, (
            f"Tomosipo autograd operator expected {expected_ndim} dimensions "
            f"but got {input.ndim}.\n"
            "The interface of to_autograd was changed in Tomosipo 0.6.0 to "
            "by default match standard Tomosipo operators and extra arguments are "
            "provided to match Pytorch NN functions.\n"
            "To add batch and channel dimensions set argument num_extra_dims=2\n"
            "To remove the first operator dimension set argument is_2d=True\n"
        )
        output = input.new_empty(extra_dims + operator.range_shape, dtype=torch.float32)
        if is_2d:
            input = torch.unsqueeze(input, dim=-3)
        _project_batched(operator, input, output, tuple(extra_dims))
        if is_2d:
            output = torch.squeeze(output, dim=-3)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        operator, extra_dims, is_2d = ctx.operator, ctx.extra_dims, ctx.is_2d
        grad_input = grad_output.new_empty(extra_dims + operator.domain_shape, dtype=torch.float32)
        if is_2d:
            grad_output = torch.unsqueeze(grad_output, dim=-3)
        _project_batched(operator.T, grad_output, grad_input, tuple(extra_dims))
        if is_2d:
            grad_input = torch.squeeze(grad_input, dim=-3)
        # no gradient for operator, num_extra_dims, is_2d
        return grad_input, None, None, None


def to_autograd(operator, num_extra_dims=0, is_2d=False):
    """Wrap an operator (or its transpose) as an autograd-enabled function.

    ``num_extra_dims`` leading dimensions (e.g. batch and channel) are mapped
    over; ``is_2d`` drops the operator's first (length-one) dimension.

    >>> A = ts.operator(ts.volume(shape=10), ts.parallel(angles=10, shape=10))
    >>> f = to_autograd(A)
    """

    def f(x):
        return OperatorFunction.apply(x, operator, num_extra_dims, is_2d)

    return f


class _LegacyProjection(Function):
    """Shared body of the legacy ``Forward`` / ``Backward`` functions (``Data``-free here)."""

    @staticmethod
    def forward(ctx, input, vg, pg, projector, is_forward):
        if input.requires_grad:
            ctx.vg, ctx.pg, ctx.projector, ctx.is_forward = vg, pg, projector, is_forward
        A = ts.operator(vg, pg)
        return (A if is_forward else A.T)(input)

    @staticmethod
    def backward(ctx, grad_output):
        A = ts.operator(ctx.vg, ctx.pg)
        return (A.T if ctx.is_forward else A)(grad_output), None, None, None, None


def forward(input, vg, pg, projector=None):
    """Legacy functional forward projection with autograd (``torch_support.py:133-165``)."""
    return _LegacyProjection.apply(input, vg, pg, projector, True)


def backward(input, vg, pg, projector=None):
    """Legacy functional backprojection with autograd (``torch_support.py:168-202``).

    The reference's gradient of this function calls ``ts.backward`` where the
    forward projection is meant; the mathematically correct transpose is used here.
    """
    return _LegacyProjection.apply(input, vg, pg, projector, False)


class AutogradOperator:
    """Operator look-alike with autograd support (``torch_support.py:205-319``).

    Additive operators, numpy arrays and ``Data`` objects are not supported.
    """

    def __init__(self, operator, num_extra_dims=0, is_2d=False):
        if operator.additive:
            raise ValueError("Additive operators are not supported")
        self.operator = operator
        self._fp_op = to_autograd(operator, num_extra_dims, is_2d)
        self._bp_op = to_autograd(operator.T, num_extra_dims, is_2d)
        self._transpose = BackprojectionOperator(self)

    def _fp(self, volume, out=None):
        if out is None:
            return self._fp_op(volume)
        out[...] = self._fp_op(volume)
        return out

    def _bp(self, projection, out=None):
        if out is None:
            return self._bp_op(projection)
        out[...] = self._bp_op(projection)
        return out

    def __call__(self, volume, out=None):
        """Forward-project a ``torch.Tensor`` (optionally into ``out``)."""
        return self._fp(volume, out)

    def transpose(self):
        return self._transpose

    @property
    def T(self):
        return self.transpose()

    @property
    def astra_compat_vg(self):
        return self.operator.astra_compat_vg

    @property
    def astra_compat_pg(self):
        return self.operator.astra_compat_pg

    @property
    def domain(self):
        return self.operator.domain

    @property
    def range(self):
        return self.operator.range

    @property
    def domain_shape(self):
        return self.operator.domain_shape

    @property
    def range_shape(self):
        return self.operator.range_shape


def autograd_operator(volume_geometry, projection_geometry, voxel_supersampling=1, detector_supersampling=1,
                      num_extra_dims=0, is_2d=False):
    """``ts.operator`` + :class:`AutogradOperator` in one call."""
    op = ts.operator(
        volume_geometry=volume_geometry,
        projection_geometry=projection_geometry,
        voxel_supersampling=voxel_supersampling,
        detector_supersampling=detector_supersampling,
    )
    return AutogradOperator(operator=op, num_extra_dims=num_extra_dims, is_2d=is_2d)
