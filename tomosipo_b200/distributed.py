"""Multi-GPU projection: one process per GPU, angle-sharded projections,
z-slab-sharded volumes (BASELINE.json north_star).

The reference has no multi-GPU path for device arrays ("you must distribute the
data over multiple GPUs yourself", ``doc/topics/operator.rst:246-249``); this is
the natural sharding of its operator:

=========  =====================================  ==========================
step       per rank                               exchange (NCCL / NVLink)
=========  =====================================  ==========================
state      x, C: z-slab ``[nz/N, ny, nx]``;       --
           y, R: angle block ``[V, A/N, U]``
``A(x)``   FP of the rank's angle block from the  ``all_gather`` of z-slabs
           replicated volume
``A.T(y)`` BP of the rank's angle block into a    ``reduce_scatter(sum)`` of
           full-size partial volume               partial volumes -> z-slabs
=========  =====================================  ==========================

FP is linear in the volume and independent per angle; BP is a sum over angles,
so the sharded operator equals the single-GPU operator exactly up to the order
of the fp32 sum over angle blocks.

Overlap (``pipeline=True``, the default on NCCL).  The backprojection is cut
into the ``N`` z-slabs of the sharding: slab ``j`` is back-projected by a
sub-operator on ``vg[z_j]`` and handed to a second stream, which reduces it to
its owner ``j`` while slab ``j + 1`` is being computed; only the last slab's
reduce is exposed.  :func:`sirt` extends the same pipeline across the
iteration boundary: the owner applies ``x_j -= C_j * slab`` as soon as its slab
has arrived and broadcasts the new ``x_j`` into every rank's replicated volume,
so the next forward projection needs no all-gather.  The residual
``R * (A x - y)`` is formed in the forward projector's store
(``tsp_project_fused``): an iteration is one FP launch group, ``N`` BP launches
and ``2 N`` slab-sized collectives hidden behind them.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import tomosipo_b200 as ts


def shard_bounds(n, world, rank):
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank``."""
    return rank * n // world, (rank + 1) * n // world


class ShardedOperator:
    """Angle-/slab-sharded view of ``ts.operator(vg, pg)`` over a process group.

    ``make_local(vg, pg_block)`` builds the rank-local operator (default:
    ``ts.operator``); it only has to be callable as ``op(x, out=...)`` /
    ``op.T(y, out=...)`` on the arrays it is given.
    """

    def __init__(self, volume_geometry, projection_geometry, group=None, make_local=None, device=None, pipeline=None):
        if not isinstance(volume_geometry, ts.geometry.VolumeGeometry):
            raise TypeError("ShardedOperator needs an axis-aligned VolumeGeometry (z-slab sharding).")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.volume_geometry = volume_geometry
        self.projection_geometry = projection_geometry
        pg = projection_geometry.to_vec()
        self.angle_lo, self.angle_hi = shard_bounds(pg.num_angles, self.world, self.rank)
        if self.angle_hi <= self.angle_lo:
            raise ValueError(f"rank {self.rank} would own no projection angle ({pg.num_angles} angles, {self.world} ranks)")
        self.local_pg = pg[self.angle_lo:self.angle_hi]
        self.local = (make_local or ts.operator)(volume_geometry, self.local_pg)
        nz, ny, nx = volume_geometry.shape
        self.slab_nz = -(-nz // self.world)  # slabs are padded to equal height for the collectives
        self.z_lo = min(self.rank * self.slab_nz, nz)
        self.z_hi = min(self.z_lo + self.slab_nz, nz)
        self.vol_shape = (nz, ny, nx)
        self.padded_shape = (self.slab_nz * self.world, ny, nx)
        self.slab_shape = (self.slab_nz, ny, nx)          # rows >= z_hi - z_lo are padding (zeros)
        self.proj_shape = (pg.det_shape[0], self.angle_hi - self.angle_lo, pg.det_shape[1])
        self.device = device
        self._full = None
        self._partial = None
        self._transpose = _ShardedTranspose(self)
        # z-slab sub-operators of the pipelined backprojection (built on first use)
        self._make_local = make_local or ts.operator
        if pipeline is None:  # TSP_SHARD_NO_PIPELINE: measurement aid (one reduce_scatter / all_gather per call)
            pipeline = not os.environ.get("TSP_SHARD_NO_PIPELINE")
        self.pipeline = bool(pipeline) and self.world > 1
        self._slab_ops = None
        self._comm_stream = None

    # ------------------------------------------------------------- buffers --
    def _full_volume(self, like):
        if self._full is None or self._full.device != like.device:
            self._full = torch.zeros(self.padded_shape, dtype=torch.float32, device=like.device)
        return self._full

    def _partial_volume(self, like):
        """Second full-size buffer: the partial backprojection of this rank's angle block."""
        if self._partial is None or self._partial.device != like.device:
            self._partial = torch.zeros(self.padded_shape, dtype=torch.float32, device=like.device)
        return self._partial

    def slab_geometry(self):
        """Geometry of this rank's (unpadded) z-slab."""
        return self.volume_geometry[self.z_lo:self.z_hi]

    def slab_bounds(self, j):
        """Unpadded z range ``[lo, hi)`` of rank ``j``'s slab."""
        nz = self.vol_shape[0]
        lo = min(j * self.slab_nz, nz)
        return lo, min(lo + self.slab_nz, nz)

    def slab_operators(self):
        """``[(z_lo, z_hi, operator on vg[z_lo:z_hi] x this rank's angle block)]``, empty slabs left out."""
        if self._slab_ops is None:
            ops = []
            for j in range(self.world):
                lo, hi = self.slab_bounds(j)
                ops.append((j, lo, hi, self._make_local(self.volume_geometry[lo:hi], self.local_pg) if hi > lo else None))
            self._slab_ops = ops
        return self._slab_ops

    def _side_stream(self, like):
        """(compute stream, communication stream) on CUDA, (None, None) on CPU."""
        if not like.is_cuda:
            return None, None
        if self._comm_stream is None or self._comm_stream.device != like.device:
            self._comm_stream = torch.cuda.Stream(device=like.device)
        return torch.cuda.current_stream(like.device), self._comm_stream

    def _bp_slabs(self, y_block, partial, after_slab):
        """Back-project slab by slab into ``partial``; ``after_slab(j, slab_view)`` is issued on the
        communication stream once slab ``j`` is complete (slab ``j + 1`` is computed meanwhile)."""
        compute, comm = self._side_stream(y_block)
        slabs = partial.view(self.world, *self.slab_shape)
        if comm is not None:
            comm.wait_stream(compute)                      # earlier users of `partial` / the replicated volume
        for j, lo, hi, op in self.slab_operators():
            if op is not None:
                op.T(y_block, out=slabs[j][: hi - lo])
            if comm is None:
                after_slab(j, slabs[j])
                continue
            ev = torch.cuda.Event()
            ev.record(compute)
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                after_slab(j, slabs[j])
        if comm is not None:
            compute.wait_stream(comm)

    def scatter_volume(self, full):
        """This rank's padded slab of a replicated ``[nz, ny, nx]`` tensor."""
        slab = torch.zeros(self.slab_shape, dtype=torch.float32, device=full.device)
        slab[: self.z_hi - self.z_lo] = full[self.z_lo:self.z_hi]
        return slab

    def gather_volume(self, slab):
        """Replicated ``[nz, ny, nx]`` tensor from all ranks' slabs."""
        full = self._full_volume(slab)
        self._all_gather(full, slab)
        return full[: self.vol_shape[0]].clone()

    # --------------------------------------------------------- collectives --
    def _nccl(self):
        return dist.is_initialized() and dist.get_backend(self.group) == "nccl"

    def _all_gather(self, full, slab):
        if self.world == 1:
            full.copy_(slab)
        elif self._nccl():
            dist.all_gather_into_tensor(full, slab.contiguous(), group=self.group)
        else:  # gloo (CPU tests)
            parts = list(full.view(self.world, *self.slab_shape).unbind(0))
            dist.all_gather(parts, slab.contiguous(), group=self.group)

    def _reduce_scatter(self, slab, full):
        if self.world == 1:
            slab.copy_(full)
        elif self._nccl():
            dist.reduce_scatter_tensor(slab, full, op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo has no reduce_scatter: all_reduce, then keep the own slab
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            slab.copy_(full.view(self.world, *self.slab_shape)[self.rank])

    # ------------------------------------------------------------ operator --
    def __call__(self, x_slab, out=None):
        """``y_block = A[angle block] (all_gather(x_slab))``."""
        if tuple(x_slab.shape) != self.slab_shape:
            raise ValueError(f"Expected a padded z-slab of shape {self.slab_shape}. Got {tuple(x_slab.shape)}")
        if self.world == 1:
            return self.local(x_slab, out=out)
        full = self._full_volume(x_slab)
        self._all_gather(full, x_slab)
        if out is None:
            out = torch.empty(self.proj_shape, dtype=torch.float32, device=x_slab.device)
        self.local(full[: self.vol_shape[0]], out=out)
        return out

    def _bp(self, y_block, out=None):
        """``x_slab = reduce_scatter(A[angle block]^T y_block)``."""
        if tuple(y_block.shape) != self.proj_shape:
            raise ValueError(f"Expected an angle block of shape {self.proj_shape}. Got {tuple(y_block.shape)}")
        if self.world == 1:
            return self.local.T(y_block, out=out)
        if out is None:
            out = torch.empty(self.slab_shape, dtype=torch.float32, device=y_block.device)
        if self.pipeline:
            partial = self._partial_volume(y_block)       # padding rows are never written: they stay zero
            self._bp_slabs(y_block, partial, lambda j, slab: dist.reduce(slab, dst=self._global(j), group=self.group))
            out.copy_(partial.view(self.world, *self.slab_shape)[self.rank])
            return out
        full = self._full_volume(y_block)
        if self.padded_shape != self.vol_shape:
            full[self.vol_shape[0]:].zero_()
        self.local.T(y_block, out=full[: self.vol_shape[0]])
        self._reduce_scatter(out, full)
        return out

    def _global(self, j):
        """Global rank of group rank ``j`` (collectives with ``dst`` / ``src`` take global ranks)."""
        return j if self.group is None else dist.get_global_rank(self.group, j)

    # ------------------------------------------------------ fused residual --
    def residual(self, x_full, y, R, out):
        """``out = R * (A[angle block] x_full - y)`` for a replicated volume: in the projector's
        store when the rank-local operator is the library's (CUDA), else the explicit three passes."""
        proj = getattr(self.local, "astra_projector", None)
        if (proj is not None and x_full.is_cuda and hasattr(proj, "project_fused") and not self.local.additive
                and all(t.dtype == torch.float32 and t.is_contiguous() for t in (x_full, y, R, out))):
            from . import _backend

            with torch.cuda.device_of(x_full):
                stream = torch.cuda.current_stream(x_full.device).cuda_stream
                proj.project_fused(_backend.FP, x_full.data_ptr(), out.data_ptr(), y.data_ptr(), R.data_ptr(),
                                   device=x_full.device.index, stream=stream)
            return out
        self.local(x_full, out=out)
        out -= y
        out *= R
        return out

    @property
    def T(self):
        return self._transpose

    def transpose(self):
        return self._transpose


class _ShardedTranspose:
    def __init__(self, parent):
        self.parent = parent

    def __call__(self, y_block, out=None):
        return self.parent._bp(y_block, out)

    @property
    def T(self):
        return self.parent

    def transpose(self):
        return self.parent


def sirt_weights(A, device, eps=None):
    """``(R, C)`` = ``(1 / A(1), 1 / A.T(1))`` clamped like ``notebooks/sirt_benchmark.py:116-128``:
    ``R`` for this rank's angle block, ``C`` for its z-slab."""
    eps = ts.epsilon if eps is None else eps
    y_tmp = torch.ones(A.proj_shape, device=device)
    C = A.T(y_tmp)
    C[C < eps] = float("inf")
    C.reciprocal_()
    x_tmp = torch.ones(A.slab_shape, device=device)
    if A.z_hi - A.z_lo < A.slab_nz:
        x_tmp[A.z_hi - A.z_lo:] = 0  # padding rows stay empty
    R = A(x_tmp)
    R[R < eps] = float("inf")
    R.reciprocal_()
    return R, C


def sirt(A, y, num_iterations, x=None, eps=None, weights=None):
    """SIRT with the volume sharded in z (loop of ``notebooks/sirt_benchmark.py:116-139``).

    ``A`` is a :class:`ShardedOperator` (or any operator with the same call
    signature), ``y`` this rank's angle block, ``weights`` an optional
    ``(R, C)`` from :func:`sirt_weights`.  Returns this rank's padded z-slab of
    the reconstruction.
    """
    dev = y.device
    R, C = sirt_weights(A, dev, eps) if weights is None else weights
    y_tmp = torch.empty(A.proj_shape, device=dev)
    x_tmp = torch.empty(A.slab_shape, device=dev)
    x_cur = torch.zeros(A.slab_shape, device=dev) if x is None else x
    if getattr(A, "pipeline", False):
        return _sirt_pipelined(A, y, R, C, x_cur, y_tmp, num_iterations)
    for _ in range(num_iterations):
        A(x_cur, out=y_tmp)
        y_tmp -= y
        y_tmp *= R
        A.T(y_tmp, out=x_tmp)
        x_tmp *= C
        x_cur -= x_tmp
    return x_cur


def _sirt_pipelined(A, y, R, C, x_cur, y_tmp, num_iterations):
    """The loop above with every exchange hidden behind the slab-wise backprojection
    (module docstring, "Overlap").  Invariant at the top of an iteration: ``x_full`` holds
    the current reconstruction on every rank, ``x_cur`` this rank's slab of it."""
    x_full = A._full_volume(y)
    A._all_gather(x_full, x_cur)
    partial = A._partial_volume(y)
    x_slabs = x_full.view(A.world, *A.slab_shape)
    y = y.contiguous()

    def after_slab(j, slab):
        dist.reduce(slab, dst=A._global(j), group=A.group)
        if j == A.rank:
            x_cur.addcmul_(C, slab, value=-1.0)           # x_j -= C_j * (sum over ranks of A_r^T y_tmp)
            x_slabs[j].copy_(x_cur)
        dist.broadcast(x_slabs[j], src=A._global(j), group=A.group)

    for _ in range(num_iterations):
        A.residual(x_full[: A.vol_shape[0]], y, R, y_tmp)
        A._bp_slabs(y_tmp, partial, after_slab)
    return x_cur
