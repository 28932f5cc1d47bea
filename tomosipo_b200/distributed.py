"""Multi-GPU projection: one process per GPU, angle-sharded projections,
z-sharded volumes (BASELINE.json north_star).

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

Overlap (``chunks = K > 1``).  The z axis is cut into ``K`` chunks and every
chunk into ``N`` pieces, one per rank: rank ``r`` owns piece ``r`` of *every*
chunk, and its "slab" tensor is those ``K`` pieces stacked.  A chunk is then a
contiguous z range on which the bandwidth-optimal collectives apply directly:
the backprojection runs chunk by chunk (a sub-operator on ``vg[chunk]``) and
chunk ``c`` is handed to a second stream for its ``reduce_scatter`` while chunk
``c + 1`` is being computed, so only ``1/K`` of the exchange is exposed.
:func:`sirt` extends the pipeline across the iteration boundary: as soon as a
rank's piece of chunk ``c`` has arrived it applies ``x -= C * piece`` and the
chunk is all-gathered into every rank's replicated volume, again behind the
remaining chunks' kernels; the next forward projection starts without an
all-gather.  The residual ``R * (A x - y)`` is formed in the forward
projector's store (``tsp_project_fused``).  ``K = 1`` is the plain scheme of
the table (contiguous slabs, one collective per call).

Row exchange (``bp_exchange="rows"``).  The backprojection can also be sharded
by *volume*: every rank back-projects ALL angles into its own (contiguous)
z-slab and nothing has to be summed across ranks.  A z-slab only sees a band of
detector rows (:func:`slab_row_bounds`), so what is exchanged is that band of
every other rank's angle block - an ``all_to_all`` of contiguous row ranges
(the projection layout is ``[V, angles, U]``) - instead of partial volumes.
The kernel then runs one launch over all angles on the slab, there is no
partial-volume buffer, nothing to reduce and no collective competing with the
kernel for SMs; the exchange itself is exposed, and so is the all-gather of the
updated slabs in :func:`sirt` (the chunked ``"volume"`` scheme hides both).  :func:`default_bp_exchange`
picks the scheme; ``TSP_SHARD_BP=rows|volume`` overrides it.

(Measured on 8 x B200, profiles/r01_bench_n8_*.json: reducing each slab to its
owner and broadcasting it back -- the first pipelining scheme tried -- loses to
the un-overlapped collectives, because a reduce / broadcast of one slab uses
one ring where all_gather / reduce_scatter use every NVLink port at once.)
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import tomosipo_b200 as ts


def shard_bounds(n, world, rank):
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank``."""
    return rank * n // world, (rank + 1) * n // world


def default_chunks(nz, world):
    """Chunks of the z axis for the overlapped exchange: 4 when the pieces stay >= 8 slices thick."""
    if world == 1:
        return 1
    env = os.environ.get("TSP_SHARD_CHUNKS")  # measurement aid: 1 = one collective per call
    if env:
        return max(1, int(env))
    for k in (4, 2):
        if nz >= 8 * k * world:
            return k
    return 1


def slab_row_bounds(volume_geometry, projection_geometry, margin=2):
    """Detector rows ``[lo, hi)`` that a backprojection into ``volume_geometry`` (any angle of
    ``projection_geometry``) can read: the projections of the box's corners bound those of its voxels
    (the detector coordinate is a projective function of the position, the box is convex and in front
    of the source), plus the bilinear taps and ``margin`` rows of slack.  The whole detector when a
    corner does not project (source inside the slab's plane range)."""
    pg = projection_geometry.to_vec()
    V = pg.det_shape[0]
    with np.errstate(all="ignore"):
        v = np.stack([pg.project_point(tuple(c))[:, 0] for c in np.asarray(volume_geometry.corners).reshape(-1, 3)])
    if not np.all(np.isfinite(v)):
        return 0, V
    lo = int(np.floor(v.min() + V / 2.0 - 0.5)) - margin
    hi = int(np.ceil(v.max() + V / 2.0 + 0.5)) + 1 + margin
    lo, hi = max(lo, 0), min(hi, V)
    return (lo, hi) if hi > lo else (0, 0)


def crop_detector_rows(projection_geometry, lo, hi, angles=None):
    """The vector geometry restricted to detector rows ``[lo, hi)`` (same pixels, same positions) and, optionally,
    to the angles of an index array."""
    pg = projection_geometry.to_vec()
    V, U = pg.det_shape
    sel = slice(None) if angles is None else np.asarray(angles, dtype=np.int64)
    det_v, det_u = pg.det_v[sel], pg.det_u[sel]
    det_pos = pg.det_pos[sel] + ((lo + hi) / 2.0 - V / 2.0) * det_v
    if pg.is_cone:
        return ts.cone_vec(shape=(hi - lo, U), src_pos=pg.src_pos[sel], det_pos=det_pos, det_v=det_v, det_u=det_u)
    return ts.parallel_vec(shape=(hi - lo, U), ray_dir=pg.ray_dir[sel], det_pos=det_pos, det_v=det_v, det_u=det_u)


def default_bp_exchange(volume_geometry, projection_geometry, world):
    """``"rows"`` or ``"volume"`` (module docstring) from what was measured on B200s:

    * rows need bands that are not much larger (<= 1.25 x) than the partial volumes a reduce_scatter would move -
      circular scans around z; every slab of a scan around another axis sees the whole detector;
    * the chunked volume scheme loses about a millisecond per call to its sub-launches and to NCCL sharing the SMs,
      whatever the size, and hides its whole exchange in SIRT; the row scheme leaves SIRT's all-gather exposed, a
      cost that grows with the volume.  Rows win while a chunk launch of the volume scheme would be shorter than
      about 12 ms (2.5e10 voxel updates): cfg 3 at any N, not cfg 4 up to N = 8."""
    env = os.environ.get("TSP_SHARD_BP")
    if env in ("rows", "volume"):
        return env
    if world == 1:
        return "volume"
    nz, ny, nx = volume_geometry.shape
    pg = projection_geometry.to_vec()
    if float(nz) * ny * nx * pg.num_angles / (world * default_chunks(nz, world)) > 2.5e10:
        return "volume"
    rows = 0
    for r in range(world):
        lo, hi = shard_bounds(nz, world, r)
        if hi > lo:
            b = slab_row_bounds(volume_geometry[lo:hi], pg)
            rows = max(rows, b[1] - b[0])
    return "rows" if rows * pg.num_angles * pg.det_shape[1] <= 1.25 * nz * ny * nx else "volume"


class ShardedOperator:
    """Angle-/z-sharded view of ``ts.operator(vg, pg)`` over a process group.

    ``make_local(vg, pg_block)`` builds the rank-local operator (default:
    ``ts.operator``); it only has to be callable as ``op(x, out=...)`` /
    ``op.T(y, out=...)`` on the arrays it is given.  ``chunks``: see the module
    docstring (default :func:`default_chunks`).  ``bp_exchange``: ``"volume"`` (reduce_scatter of partial
    volumes), ``"rows"`` (all_to_all of detector row bands, contiguous slabs, ``chunks = 1``) or None
    (:func:`default_bp_exchange`; an explicit ``chunks > 1`` selects ``"volume"``).
    """

    def __init__(self, volume_geometry, projection_geometry, group=None, make_local=None, device=None, chunks=None,
                 bp_exchange=None):
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
        self._make_local = make_local or ts.operator
        self.local = self._make_local(volume_geometry, self.local_pg)
        nz, ny, nx = volume_geometry.shape
        if bp_exchange is None:
            bp_exchange = ("volume" if (chunks is not None and int(chunks) > 1)
                           else default_bp_exchange(volume_geometry, pg, self.world))
        if bp_exchange not in ("rows", "volume"):
            raise ValueError(f"bp_exchange must be 'rows', 'volume' or None. Got {bp_exchange!r}")
        self.bp_exchange = bp_exchange if self.world > 1 else "volume"
        if self.bp_exchange == "rows":
            if chunks is not None and int(chunks) > 1:
                raise ValueError("bp_exchange='rows' shards the volume in contiguous slabs (chunks = 1)")
            chunks = 1
        self.chunks = default_chunks(nz, self.world) if chunks is None else (max(1, int(chunks)) if self.world > 1 else 1)
        self.piece_nz = -(-nz // (self.chunks * self.world))  # pieces are padded to equal height for the collectives
        self.chunk_nz = self.piece_nz * self.world
        self.slab_nz = self.piece_nz * self.chunks
        self.vol_shape = (nz, ny, nx)
        self.padded_shape = (self.chunk_nz * self.chunks, ny, nx)
        self.slab_shape = (self.slab_nz, ny, nx)          # rows beyond a piece's valid height are padding (zeros)
        self.proj_shape = (pg.det_shape[0], self.angle_hi - self.angle_lo, pg.det_shape[1])
        # contiguous slab of the K = 1 scheme (None when the ownership is interleaved)
        self.z_lo, self.z_hi = self.piece_bounds(0, self.rank) if self.chunks == 1 else (None, None)
        self.device = device
        self._full = None
        self._full_t = False      # not decided yet (None: no caller-made transposed copy)
        self._partial = None
        self._chunk_ops = None
        self._comm_stream = None
        self._transpose = _ShardedTranspose(self)
        self.angle_bounds = [shard_bounds(pg.num_angles, self.world, r) for r in range(self.world)]
        self.row_bounds = None    # rows mode: detector rows [lo, hi) every rank's slab reads
        self._row_op = self._row_own = self._row_rest = None
        self._rows = self._row_staging = None
        self._peer = False        # not tried yet (None: unavailable, the NCCL all_to_all is used)
        # rows mode, opt-in variant (library operator): the rank's own angle block is back-projected first (it needs no
        # exchange), the all_to_all runs behind that launch and a second, additive launch takes the other ranks'
        # bands.  Measured at N = 2, cfg 3: 24.08 vs 23.97 ms - the second launch's wave tail costs what the hidden
        # exchange saves - so one launch over all angles is the default.
        self._row_split = (self.bp_exchange == "rows" and make_local is None and self._nccl()
                           and bool(os.environ.get("TSP_SHARD_ROWS_SPLIT")))
        if self.bp_exchange == "rows":
            self.row_bounds = []
            for r in range(self.world):
                lo, hi = self.piece_bounds(0, r)
                self.row_bounds.append(slab_row_bounds(volume_geometry[lo:hi], pg) if hi > lo else (0, 0))
            lo, hi = self.row_bounds[self.rank]
            if hi > lo and self.z_hi > self.z_lo:
                slab_vg = volume_geometry[self.z_lo:self.z_hi]
                if self._row_split:
                    own = np.arange(self.angle_lo, self.angle_hi)
                    rest = np.concatenate([np.arange(0, self.angle_lo), np.arange(self.angle_hi, pg.num_angles)])
                    self._row_own = ts.operator(slab_vg, crop_detector_rows(pg, lo, hi, own))
                    self._row_rest = ts.operator(slab_vg, crop_detector_rows(pg, lo, hi, rest), additive=True)
                else:
                    self._row_op = self._make_local(slab_vg, crop_detector_rows(pg, lo, hi))
        # The exchange only overlaps the kernels if its CTAs are scheduled ahead of the backprojector's
        # queued ones: NCCL must run on high-priority streams.  With the default group that is a
        # construction-time option the caller may not have set, so the overlapped scheme talks over its
        # own communicator (measured at N = 8 without it: zero overlap, profiles/r01_sirt_breakdown_n8.txt).
        if group is None and (self.chunks > 1 or self._row_split) and self._nccl():
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            self.group = dist.new_group(backend="nccl", pg_options=opts)

    # ------------------------------------------------------------- layout --
    def piece_bounds(self, c, r):
        """Unpadded z range ``[lo, hi)`` of rank ``r``'s piece of chunk ``c``."""
        nz = self.vol_shape[0]
        lo = min((c * self.world + r) * self.piece_nz, nz)
        return lo, min(lo + self.piece_nz, nz)

    def chunk_bounds(self, c):
        """Unpadded z range ``[lo, hi)`` of chunk ``c``."""
        nz = self.vol_shape[0]
        lo = min(c * self.chunk_nz, nz)
        return lo, min(lo + self.chunk_nz, nz)

    def slab_pieces(self, rank=None):
        """``[(slab_row, z_lo, z_hi)]``: rows ``slab_row : slab_row + z_hi - z_lo`` of the rank's slab
        tensor hold the volume slices ``z_lo : z_hi``."""
        r = self.rank if rank is None else rank
        return [(c * self.piece_nz,) + self.piece_bounds(c, r) for c in range(self.chunks)]

    def slab_geometry(self):
        """Geometry of this rank's (unpadded) z-slab; only for the contiguous ``chunks == 1`` layout."""
        if self.chunks != 1:
            raise ValueError("the slab is not contiguous when chunks > 1; see slab_pieces()")
        return self.volume_geometry[self.z_lo:self.z_hi]

    def zero_padding_(self, slab):
        """Zero the padding rows of a slab tensor (in place)."""
        for row, lo, hi in self.slab_pieces():
            slab[row + hi - lo: row + self.piece_nz].zero_()
        return slab

    def scatter_volume(self, full):
        """This rank's padded slab of a replicated ``[nz, ny, nx]`` tensor."""
        slab = torch.zeros(self.slab_shape, dtype=torch.float32, device=full.device)
        for row, lo, hi in self.slab_pieces():
            slab[row: row + hi - lo] = full[lo:hi]
        return slab

    def gather_volume(self, slab):
        """Replicated ``[nz, ny, nx]`` tensor from all ranks' slabs."""
        full = self._full_volume(slab)
        slab = slab.contiguous()
        for c in range(self.chunks):
            self._all_gather_chunk(full, slab, c)
        return full[: self.vol_shape[0]].clone()

    # ------------------------------------------------------------- buffers --
    def _full_volume(self, like):
        if self._full is None or self._full.device != like.device:
            self._full = torch.zeros(self.padded_shape, dtype=torch.float32, device=like.device)
        return self._full

    def _partial_volume(self, like):
        """Second full-size buffer: the partial backprojection of this rank's angle block
        (padding rows are never written and stay zero)."""
        if self._partial is None or self._partial.device != like.device:
            self._partial = torch.zeros(self.padded_shape, dtype=torch.float32, device=like.device)
        return self._partial

    def _transposed_volume(self, like):
        """Device buffer for the (x <-> y)-transposed copy of the replicated volume that this rank's forward projector
        reads (angles marching along x), or None (not needed / rank-local operator is not the library's).  With it,
        every z-chunk is transposed as it arrives (behind the all-gather of the next chunk, or behind the next
        chunk's backprojection in SIRT) instead of the whole volume being transposed in front of every FP."""
        if self._full_t is False:
            self._full_t = None
            proj = getattr(self.local, "astra_projector", None)
            if (like.is_cuda and proj is not None and hasattr(proj, "fp_transposed_elems") and self.world > 1
                    and not getattr(self.local, "additive", False) and not os.environ.get("TSP_SHARD_NO_PRETRANSPOSE")):
                n = proj.fp_transposed_elems()
                if n > 0:
                    self._full_t = torch.empty(n, dtype=torch.float32, device=like.device)
        return self._full_t

    def _transpose_chunk(self, full, full_t, c):
        """Transpose the slices of chunk ``c`` of ``full`` into ``full_t`` on the current stream."""
        lo, hi = self.chunk_bounds(c)
        if hi > lo:
            stream = torch.cuda.current_stream(full.device).cuda_stream
            self.local.astra_projector.transpose_slices(full.data_ptr(), full_t.data_ptr(), lo, hi, device=full.device.index,
                                                        stream=stream)

    def _lib_fp_ok(self, full, *arrays):
        """The rank-local operator is the library's and the arrays can go to it by pointer."""
        proj = getattr(self.local, "astra_projector", None)
        return (full.is_cuda and proj is not None and hasattr(proj, "fp_pre_transposed")
                and not getattr(self.local, "additive", False)
                and all(t.dtype == torch.float32 and t.is_contiguous() for t in (full,) + arrays))

    def _fp_local(self, full, full_t, out, y=None, R=None, turn=None):
        """``out = A[block] full`` (or ``R * (A full - y)``) from the replicated volume and its transposed copy
        (None: the library transposes itself if it has to).  ``turn``: band buffer of the peer-memory exchange that the
        projector's store fills on every rank as it goes (``tsp_fp_push``)."""
        proj = self.local.astra_projector
        stream = torch.cuda.current_stream(full.device).cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()
        P = self._peer
        with torch.cuda.device_of(full):
            if turn is not None and P and P["fp_push"]:
                U, a_all = self.proj_shape[2], self.angle_bounds[-1][1]
                peers = [(P["ptrs"][turn][q] + self.angle_lo * U * 4, a, b) for q, (a, b) in enumerate(self.row_bounds) if b > a]
                proj.fp_push(full.data_ptr(), ptr(full_t), out.data_ptr(), ptr(y), ptr(R), peers, a_all * U,
                             device=full.device.index, stream=stream)
                try:
                    torch.autograd.graph.increment_version(out)  # written behind torch's back
                    P["pushed"] = (out, out._version)
                except RuntimeError:                             # no version counter (inference tensor): not tracked
                    P["pushed"] = None
            else:
                proj.fp_pre_transposed(full.data_ptr(), ptr(full_t), out.data_ptr(), ptr(y), ptr(R),
                                       device=full.device.index, stream=stream)
        return out

    def chunk_operators(self):
        """``[(c, z_lo, z_hi, operator on vg[z_lo:z_hi] x this rank's angle block | None if empty)]``."""
        if self._chunk_ops is None:
            ops = []
            for c in range(self.chunks):
                lo, hi = self.chunk_bounds(c)
                if self.chunks == 1:
                    op = self.local
                else:
                    op = self._make_local(self.volume_geometry[lo:hi], self.local_pg) if hi > lo else None
                ops.append((c, lo, hi, op))
            self._chunk_ops = ops
        return self._chunk_ops

    def bp_operators(self):
        """The rank-local operators whose transposes run in ``A.T`` (for launch counts / kernel records)."""
        if self.bp_exchange == "rows":
            return [op for op in (self._row_op, self._row_own, self._row_rest) if op is not None]
        return [op for _, _, _, op in self.chunk_operators() if op is not None]

    def _streams(self, like):
        """(compute stream, communication stream) on CUDA, (None, None) on CPU."""
        if not like.is_cuda:
            return None, None
        if self._comm_stream is None or self._comm_stream.device != like.device:
            self._comm_stream = torch.cuda.Stream(device=like.device, priority=-1)
        return torch.cuda.current_stream(like.device), self._comm_stream

    # --------------------------------------------------------- collectives --
    def _nccl(self):
        return dist.is_initialized() and dist.get_backend(self.group) == "nccl"

    def _chunk_view(self, full, c):
        return full[c * self.chunk_nz: (c + 1) * self.chunk_nz]

    def _piece_view(self, slab, c):
        return slab[c * self.piece_nz: (c + 1) * self.piece_nz]

    def _all_gather_chunk(self, full, slab, c):
        dst, src = self._chunk_view(full, c), self._piece_view(slab, c)
        if self.world == 1:
            dst.copy_(src)
        elif self._nccl():
            dist.all_gather_into_tensor(dst, src, group=self.group)
        else:  # gloo (CPU tests)
            parts = list(dst.view(self.world, self.piece_nz, *self.slab_shape[1:]).unbind(0))
            dist.all_gather(parts, src.contiguous(), group=self.group)

    def _reduce_scatter_chunk(self, piece, full, c):
        """``piece`` (one piece-sized tensor) = this rank's part of the sum over ranks of chunk ``c`` of ``full``."""
        src = self._chunk_view(full, c)
        if self.world == 1:
            piece.copy_(src)
        elif self._nccl():
            dist.reduce_scatter_tensor(piece, src, op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo has no reduce_scatter: all_reduce, then keep the own piece
            dist.all_reduce(src, op=dist.ReduceOp.SUM, group=self.group)
            piece.copy_(src.view(self.world, self.piece_nz, *self.slab_shape[1:])[self.rank])

    # ------------------------------------------------- peer-memory exchange --
    def _peer_setup(self, y_block):
        """Two band buffers per rank (alternating calls), allocated by the library and mapped by every peer through
        CUDA IPC; None when any rank cannot (then every rank uses the NCCL all_to_all)."""
        from . import _backend as B

        dev = y_block.device
        lo, hi = self.row_bounds[self.rank]
        shape = (hi - lo, self.angle_bounds[-1][1], self.proj_shape[2])
        own, mapped, ok = [], [], 1.0
        try:
            own = [B.peer_alloc(max(1, int(np.prod(shape))) * 4, dev.index) for _ in range(2)]
        except Exception:
            ok = 0.0
        handles = [None] * self.world
        dist.all_gather_object(handles, [h for _, h in own] if ok else None, group=self.group)
        ptrs = [[None] * self.world for _ in range(2)]
        if ok and all(h is not None for h in handles):
            try:
                for q in range(self.world):
                    for k in range(2):
                        if q == self.rank:
                            ptrs[k][q] = own[k][0]
                        else:
                            ptrs[k][q] = B.peer_open(handles[q][k], dev.index)
                            mapped.append(ptrs[k][q])
            except Exception:
                ok = 0.0
        else:
            ok = 0.0
        views = []
        if ok:
            views = [torch.as_tensor(B.DeviceBuffer(ptr, shape), device=dev) for ptr, _ in own]
            if any(v.numel() and (v.data_ptr() != ptr or v.device != dev) for v, (ptr, _) in zip(views, own)):
                ok = 0.0                                   # torch copied instead of wrapping the buffer
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if float(flag.item()) < 1.0:                       # some rank failed: nobody uses peer memory
            for ptr in mapped:
                B.peer_close(ptr, dev.index)
            dist.barrier(group=self.group)
            for ptr, _ in own:
                B.peer_free(ptr, dev.index)
            return None
        return {"ptrs": ptrs, "views": views, "turn": 0, "flag": torch.zeros(1, device=dev), "mapped": mapped,
                "own": [ptr for ptr, _ in own], "device": dev.index,
                # the band buffer a forward projection has claimed (on every rank) for the backprojection that follows,
                # and the (tensor, version) whose rows this rank's projector has already stored there
                "pending": None, "pushed": None,
                "fp_push": (not os.environ.get("TSP_SHARD_NO_FP_PUSH") and self.world <= 16      # FP_MAX_PEERS
                            and hasattr(getattr(self.local, "astra_projector", None), "fp_push"))}

    def _take_turn(self, like, pending=False):
        """Next band buffer (they alternate) of the peer-memory exchange, or None when it is not in use.  Every rank
        calls this at the same points of the program, so all ranks agree on the buffer."""
        if self.bp_exchange != "rows" or self.world == 1:
            return None
        if (self._peer is False and like.is_cuda and self._nccl() and not os.environ.get("TSP_SHARD_NO_P2P")
                and not self._row_split):
            self._peer = self._peer_setup(like)
        P = self._peer
        if not P or not like.is_cuda:
            return None
        k = P["turn"]
        P["turn"] = k ^ 1
        P["pending"], P["pushed"] = (k if pending else None), None
        return k

    def _exchange_rows_peer(self, y_block):
        """The band of every rank's angle block, stored by the ranks themselves: one kernel of NVLink stores per rank
        (``tsp_push_rows``) puts the rows each peer's slab reads into that peer's buffer at this rank's angle offset;
        a one-element all_reduce on the same stream tells every rank that all stores have landed.  The buffers
        alternate between calls: a rank may start storing for call k + 1 while a peer still back-projects call k."""
        from . import _backend as B

        P = self._peer
        k, pushed = P["pending"], P["pushed"]
        P["pending"] = P["pushed"] = None
        if k is None:
            k = self._take_turn(y_block)
        if pushed is not None:                             # did this rank's projector store exactly this array's rows?
            t, version = pushed
            try:
                pushed = (t.data_ptr() == y_block.data_ptr() and t.shape == y_block.shape and y_block._version == version)
            except RuntimeError:                           # no version counter (inference tensor)
                pushed = False
        if not pushed:
            U = self.proj_shape[2]
            a_me, a_all = self.angle_hi - self.angle_lo, self.angle_bounds[-1][1]
            base = y_block.data_ptr()
            jobs = [(base + a * a_me * U * 4, P["ptrs"][k][q] + self.angle_lo * U * 4, b - a, a_me * U, a_me * U, a_all * U)
                    for q, (a, b) in enumerate(self.row_bounds) if b > a]
            stream = torch.cuda.current_stream(y_block.device).cuda_stream
            B.push_rows(jobs, device=P["device"], stream=stream, projector=getattr(self.local, "astra_projector", None))
        dist.all_reduce(P["flag"], group=self.group)
        return P["views"][k]

    def __del__(self):
        # not collective: only this process's mappings of the peers' buffers can go; the own buffers must outlive the
        # peers' mappings and stay until close() or process exit
        P = getattr(self, "_peer", None)
        if P and P.get("mapped"):
            try:
                from . import _backend as B

                torch.cuda.synchronize(P["device"])
                for ptr in P["mapped"]:
                    B.peer_close(ptr, P["device"])
                P["mapped"] = []
            except Exception:
                pass

    def close(self):
        """Unmap the peers' band buffers and release the own ones (collective: every rank calls it; without it the
        own buffers live until the process exits)."""
        P, self._peer = self._peer, None
        if P:
            from . import _backend as B

            torch.cuda.synchronize(P["device"])
            for ptr in P["mapped"]:
                B.peer_close(ptr, P["device"])
            dist.barrier(group=self.group)
            for ptr in P["own"]:
                B.peer_free(ptr, P["device"])

    def _exchange_rows(self, y_block, with_own=True):
        """``[rows of this rank's band, angles, U]`` from every rank's angle block: rank ``q`` receives rows
        ``row_bounds[q]`` of each block (contiguous: rows are the outermost axis) and interleaves the blocks by angle
        (all angles, or all but the rank's own block)."""
        if with_own and y_block.is_cuda and self._peer is False:
            self._take_turn(y_block)                       # sets the peer-memory exchange up (the turn itself is unused)
        if with_own and self._peer and y_block.is_cuda:
            if y_block.dtype != torch.float32 or not y_block.is_contiguous():
                y_block = y_block.to(torch.float32).contiguous()
            return self._exchange_rows_peer(y_block)
        lo, hi = self.row_bounds[self.rank]
        U = self.proj_shape[2]
        n_angles = self.angle_bounds[-1][1] - (0 if with_own else self.angle_hi - self.angle_lo)
        if self._rows is None or self._rows.device != y_block.device:
            self._rows = torch.empty((hi - lo, n_angles, U), dtype=torch.float32, device=y_block.device)
            self._row_staging = torch.empty((hi - lo) * n_angles * U, dtype=torch.float32, device=y_block.device)
        # the own block never travels; empty tensors stand in for it in the exchange
        none = y_block[:0]
        send = [none if q == self.rank else y_block[a:b] for q, (a, b) in enumerate(self.row_bounds)]
        recv, off = [], 0
        for q, (a, b) in enumerate(self.angle_bounds):
            n = 0 if q == self.rank else (hi - lo) * (b - a) * U
            recv.append(self._row_staging[off: off + n].view(hi - lo if n else 0, b - a, U))
            off += n
        if self._nccl():
            dist.all_to_all(recv, send, group=self.group)
        else:  # gloo has no all_to_all: pairwise, lower rank sends first
            for q in range(self.world):
                if q == self.rank:
                    continue
                ops = [lambda: dist.send(send[q].contiguous(), q, group=self.group) if send[q].numel() else None,
                       lambda: dist.recv(recv[q], q, group=self.group) if recv[q].numel() else None]
                for op in (ops if self.rank < q else ops[::-1]):
                    op()
        at = 0
        for q, ((a, b), part) in enumerate(zip(self.angle_bounds, recv)):
            if q == self.rank and not with_own:
                continue
            self._rows[:, at: at + b - a].copy_(y_block[lo:hi] if q == self.rank else part)
            at += b - a
        return self._rows

    def _bp_rows(self, y_block, out):
        """``out`` = the rank's z-slab of ``A^T y``: all angles, from the exchanged row band."""
        valid = self.z_hi - self.z_lo
        y_block = y_block.contiguous()
        compute, comm = self._streams(y_block)
        if self._row_split and comm is not None:
            lo, hi = self.row_bounds[self.rank]
            comm.wait_stream(compute)                      # producer of y_block, last reader of the band buffer
            with torch.cuda.stream(comm):
                rest = self._exchange_rows(y_block, with_own=False)
                ev = torch.cuda.Event()
                ev.record(comm)
            y_block.record_stream(comm)
            if self._row_own is None:
                compute.wait_event(ev)
                out.zero_()
                return out
            self._row_own.T(y_block[lo:hi], out=out[:valid])      # own angles: no exchange needed
            compute.wait_event(ev)
            self._row_rest.T(rest, out=out[:valid])               # += the other ranks' angles
        else:
            rows = self._exchange_rows(y_block)
            if self._row_op is None:
                out.zero_()
                return out
            self._row_op.T(rows, out=out[:valid])
        if valid < out.shape[0]:
            out[valid:].zero_()
        return out

    def _bp_chunks(self, y_block, partial, after_chunk):
        """Back-project chunk by chunk into ``partial``; ``after_chunk(c)`` is issued on the communication
        stream once chunk ``c`` is complete (chunk ``c + 1`` is computed meanwhile)."""
        compute, comm = self._streams(y_block)
        if comm is not None:
            comm.wait_stream(compute)                      # earlier users of the buffers the exchange touches
        for c, lo, hi, op in self.chunk_operators():
            if op is not None:
                op.T(y_block, out=self._chunk_view(partial, c)[: hi - lo])
            if comm is None:
                after_chunk(c)
                continue
            ev = torch.cuda.Event()
            ev.record(compute)
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                after_chunk(c)
        if comm is not None:
            compute.wait_stream(comm)

    # ------------------------------------------------------------ operator --
    def __call__(self, x_slab, out=None):
        """``y_block = A[angle block] (all_gather(x_slab))``."""
        if tuple(x_slab.shape) != self.slab_shape:
            raise ValueError(f"Expected a padded z-slab of shape {self.slab_shape}. Got {tuple(x_slab.shape)}")
        if self.world == 1:
            return self.local(x_slab, out=out)
        full = self._full_volume(x_slab)
        x_slab = x_slab.contiguous()
        if out is None:
            out = torch.empty(self.proj_shape, dtype=torch.float32, device=x_slab.device)
        # (A forward projection cut into detector row blocks that start behind the z-chunks they need, so that the
        # rest of the all-gather hides behind their kernels, was built and measured in round 2: the blocks' extra
        # launches, wave tails and per-block transposes cost more than the all-gather they hide - 24.5 vs 22.5 ms at
        # N = 2, 6.56 vs 6.29 ms at N = 8, cfg 3 - and it was dropped.)
        lib_ok = self._lib_fp_ok(full, out)
        full_t = self._transposed_volume(x_slab) if lib_ok else None
        compute, comm = self._streams(x_slab)
        turn = self._take_turn(x_slab, pending=True)       # peer-memory exchange: the band buffer this projection fills
        if full_t is None or comm is None:
            for c in range(self.chunks):
                self._all_gather_chunk(full, x_slab, c)
            if lib_ok:
                self._fp_local(full, None, out, turn=turn)
            else:
                self.local(full[: self.vol_shape[0]], out=out)
            return out
        # x-marching ranks: gather chunk by chunk on the communication stream and transpose every chunk as it lands,
        # behind the all-gather of the next one; the forward projection then reads the finished copy
        comm.wait_stream(compute)                          # earlier readers of the replicated buffers, producer of x_slab
        for c in range(self.chunks):
            with torch.cuda.stream(comm):
                self._all_gather_chunk(full, x_slab, c)
                ev = torch.cuda.Event()
                ev.record(comm)
            compute.wait_event(ev)
            self._transpose_chunk(full, full_t, c)
        x_slab.record_stream(comm)
        self._fp_local(full, full_t, out, turn=turn)
        return out

    def _bp(self, y_block, out=None):
        """``x_slab = reduce_scatter(A[angle block]^T y_block)``, chunk by chunk."""
        if tuple(y_block.shape) != self.proj_shape:
            raise ValueError(f"Expected an angle block of shape {self.proj_shape}. Got {tuple(y_block.shape)}")
        if self.world == 1:
            return self.local.T(y_block, out=out)
        if out is None:
            out = torch.empty(self.slab_shape, dtype=torch.float32, device=y_block.device)
        elif not out.is_contiguous():
            raise ValueError("out must be contiguous")
        if self.bp_exchange == "rows":
            return self._bp_rows(y_block, out)
        partial = self._partial_volume(y_block)
        self._bp_chunks(y_block, partial, lambda c: self._reduce_scatter_chunk(self._piece_view(out, c), partial, c))
        return out

    # ------------------------------------------------------ fused residual --
    def residual(self, x_full, y, R, out, x_full_t=None):
        """``out = R * (A[angle block] x_full - y)`` for a replicated volume: in the projector's
        store when the rank-local operator is the library's (CUDA), else the explicit three passes.
        ``x_full_t``: the caller-maintained transposed copy of ``x_full`` (see :meth:`_transposed_volume`)."""
        turn = self._take_turn(x_full, pending=True) if self.world > 1 else None
        if self.world > 1 and self._lib_fp_ok(x_full, y, R, out):
            return self._fp_local(x_full, x_full_t, out, y, R, turn=turn)
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
    if hasattr(A, "zero_padding_"):
        A.zero_padding_(x_tmp)  # padding rows stay empty
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
    x_cur = torch.zeros(A.slab_shape, device=dev) if x is None else x
    if getattr(A, "world", 1) > 1:
        return _sirt_overlapped(A, y, R, C, x_cur, y_tmp, num_iterations)
    x_tmp = torch.empty(A.slab_shape, device=dev)
    for _ in range(num_iterations):
        A(x_cur, out=y_tmp)
        y_tmp -= y
        y_tmp *= R
        A.T(y_tmp, out=x_tmp)
        x_tmp *= C
        x_cur -= x_tmp
    return x_cur


def _sirt_overlapped(A, y, R, C, x_cur, y_tmp, num_iterations):
    """The loop above on a :class:`ShardedOperator` with the exchange behind the chunk-wise
    backprojection (module docstring, "Overlap").  Invariant at the top of an iteration:
    ``x_full`` holds the current reconstruction on every rank, ``x_cur`` this rank's pieces of it."""
    if not x_cur.is_contiguous():
        raise ValueError("x must be contiguous")
    x_full = A._full_volume(y)
    x_full_t = A._transposed_volume(y)                     # None unless this rank's angles march along x
    if A.bp_exchange == "rows":
        # volume-sharded backprojection: nothing to sum across ranks, the update is local and the new slabs are
        # all-gathered (and transposed) in front of the next fused residual
        x_tmp = torch.empty_like(x_cur)
        y = y.contiguous()
        for _ in range(num_iterations):
            A._all_gather_chunk(x_full, x_cur, 0)
            if x_full_t is not None:
                A._transpose_chunk(x_full, x_full_t, 0)
            A.residual(x_full[: A.vol_shape[0]], y, R, y_tmp, x_full_t)
            A._bp_rows(y_tmp, x_tmp)
            x_cur.addcmul_(C, x_tmp, value=-1.0)
        return x_cur
    for c in range(A.chunks):
        A._all_gather_chunk(x_full, x_cur, c)
        if x_full_t is not None:
            A._transpose_chunk(x_full, x_full_t, c)
    partial = A._partial_volume(y)
    piece = torch.empty((A.piece_nz,) + tuple(A.slab_shape[1:]), dtype=torch.float32, device=y.device)
    y = y.contiguous()

    def after_chunk(c):
        A._reduce_scatter_chunk(piece, partial, c)
        A._piece_view(x_cur, c).addcmul_(A._piece_view(C, c), piece, value=-1.0)   # x -= C * sum_r A_r^T y_tmp
        A._all_gather_chunk(x_full, x_cur, c)
        if x_full_t is not None:                           # on the communication stream, behind chunk c + 1's kernel
            A._transpose_chunk(x_full, x_full_t, c)

    for _ in range(num_iterations):
        A.residual(x_full[: A.vol_shape[0]], y, R, y_tmp, x_full_t)
        A._bp_chunks(y_tmp, partial, after_chunk)
    return x_cur
