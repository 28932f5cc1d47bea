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
"""
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

    def __init__(self, volume_geometry, projection_geometry, group=None, make_local=None, device=None):
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
        self._transpose = _ShardedTranspose(self)

    # ------------------------------------------------------------- buffers --
    def _full_volume(self, like):
        if self._full is None or self._full.device != like.device:
            self._full = torch.zeros(self.padded_shape, dtype=torch.float32, device=like.device)
        return self._full

    def slab_geometry(self):
        """Geometry of this rank's (unpadded) z-slab."""
        return self.volume_geometry[self.z_lo:self.z_hi]

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
        full = self._full_volume(y_block)
        if self.padded_shape != self.vol_shape:
            full[self.vol_shape[0]:].zero_()
        self.local.T(y_block, out=full[: self.vol_shape[0]])
        if out is None:
            out = torch.empty(self.slab_shape, dtype=torch.float32, device=y_block.device)
        self._reduce_scatter(out, full)
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


def sirt(A, y, num_iterations, x=None, eps=None):
    """SIRT with the volume sharded in z (loop of ``notebooks/sirt_benchmark.py:116-139``).

    ``A`` is a :class:`ShardedOperator` (or any operator with the same call
    signature), ``y`` this rank's angle block.  Returns this rank's padded
    z-slab of the reconstruction.
    """
    eps = ts.epsilon if eps is None else eps
    dev = y.device
    y_tmp = torch.ones(A.proj_shape, device=dev)
    C = A.T(y_tmp)
    C[C < eps] = float("inf")
    C.reciprocal_()
    x_tmp = torch.ones(A.slab_shape, device=dev)
    if A.z_hi - A.z_lo < A.slab_nz:
        x_tmp[A.z_hi - A.z_lo:] = 0  # padding rows stay empty
    R = A(x_tmp)
    R[R < eps] = float("inf")
    R.reciprocal_()
    x_cur = torch.zeros(A.slab_shape, device=dev) if x is None else x
    for _ in range(num_iterations):
        A(x_cur, out=y_tmp)
        y_tmp -= y
        y_tmp *= R
        A.T(y_tmp, out=x_tmp)
        x_tmp *= C
        x_cur -= x_tmp
    return x_cur
