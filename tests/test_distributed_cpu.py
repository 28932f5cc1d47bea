"""Host-side logic of the angle-/slab-sharded operator on CPU: two gloo ranks,
the rank-local projector replaced by the oracle (test infrastructure)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tomosipo_b200 as ts
from tomosipo_b200.distributed import ShardedOperator, shard_bounds, sirt


class OracleOperator:
    """``ts.operator`` look-alike backed by the CPU oracle (torch CPU tensors)."""

    def __init__(self, vg, pg):
        from oracle import oracle as O

        avg, apg = vg.to_astra(), pg.to_vec().to_astra()
        o = avg["option"]
        kind = O.CONE_VEC if apg["type"] == "cone_vec" else O.PARALLEL_VEC
        self.Q = O.OracleProjector(
            kind, vg.shape, [o["WindowMinX"], o["WindowMinY"], o["WindowMinZ"]],
            [o["WindowMaxX"], o["WindowMaxY"], o["WindowMaxZ"]], pg.det_shape, apg["Vectors"])
        self.T = self._T(self)

    def __call__(self, x, out=None):
        y = torch.from_numpy(self.Q.fp(x.numpy(), dtype=np.float32))
        if out is None:
            return y
        out.copy_(y)
        return out

    class _T:
        def __init__(self, parent):
            self.parent = parent

        def __call__(self, y, out=None):
            x = torch.from_numpy(self.parent.Q.bp(y.numpy(), dtype=np.float32))
            if out is None:
                return x
            out.copy_(x)
            return out


def geometries():
    vg = ts.volume(shape=(9, 10, 12), size=(0.9, 1.0, 1.2))          # 9 slices: uneven split over 2 ranks
    pg = ts.cone(angles=7, shape=(10, 14), size=(2.0, 2.8), src_orig_dist=4, src_det_dist=7)
    return vg, pg


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        vg, pg = geometries()
        ref = OracleOperator(vg, pg)
        x = torch.rand(vg.shape)
        y_full = ref(x)
        w = torch.rand(y_full.shape)
        full_bp = ref.T(w)
        lo, hi = shard_bounds(pg.num_angles, world, rank)
        recs = []
        # chunks = 1: contiguous slabs, one collective per call; chunks = 2: interleaved pieces, the
        # overlapped exchange (9 slices over 2 x 2 pieces of 3: the last piece is all padding)
        for chunks, mode in ((1, "volume"), (2, "volume"), (1, "rows")):
            S = ShardedOperator(vg, pg, make_local=OracleOperator, chunks=chunks, bp_exchange=mode)
            assert (S.angle_lo, S.angle_hi) == (lo, hi) and S.chunks == chunks and S.bp_exchange == mode
            if mode == "rows":    # every rank back-projects all angles from a band of rows; the bands overlap
                assert len(S.bp_operators()) == 1 and S.row_bounds[0][0] == 0 and S.row_bounds[1][1] == pg.det_shape[0]
                assert S.row_bounds[0][1] > S.row_bounds[1][0] > 0
            if chunks == 1:
                assert (S.z_lo, S.z_hi) == ((0, 5) if rank == 0 else (5, 9)) and S.slab_geometry().shape[0] == S.z_hi - S.z_lo
            else:
                assert S.slab_pieces() == ([(0, 0, 3), (3, 6, 9)] if rank == 0 else [(0, 3, 6), (3, 9, 9)])
                with pytest.raises(ValueError):
                    S.slab_geometry()
            # forward: own angle block of the full projection
            y_blk = S(S.scatter_volume(x))
            torch.testing.assert_close(y_blk, y_full[:, lo:hi, :], rtol=1e-5, atol=1e-6)
            # backward: own pieces of the full backprojection, padding rows stay empty
            slab = S.T(w[:, lo:hi, :].contiguous())
            for row, z0, z1 in S.slab_pieces():
                torch.testing.assert_close(slab[row: row + z1 - z0], full_bp[z0:z1], rtol=1e-4, atol=1e-5)
                assert float(slab[row + z1 - z0: row + S.piece_nz].abs().sum()) == 0.0
            torch.testing.assert_close(S.gather_volume(slab), full_bp, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(S.gather_volume(S.scatter_volume(x)), x)
            # SIRT: sharded == single-process (compared below), both layouts agree
            recs.append(S.gather_volume(sirt(S, y_full[:, lo:hi, :].contiguous(), 4)))
        torch.testing.assert_close(recs[0], recs[1], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(recs[0], recs[2], rtol=1e-4, atol=1e-5)
        with pytest.raises(ValueError):
            ShardedOperator(vg, pg, make_local=OracleOperator, chunks=2, bp_exchange="rows")
        rec = recs[1]
        torch.save(rec, os.path.join(tmpdir, f"rec{rank}.pt"))
        with pytest.raises(ValueError):
            S(torch.zeros(3, 3, 3))
    finally:
        dist.destroy_process_group()


def test_sharded_operator_and_sirt_two_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rec0.pt"), torch.load(tmp_path / "rec1.pt")
    torch.testing.assert_close(r0, r1)
    # single-process reference SIRT with the same loop
    vg, pg = geometries()
    torch.manual_seed(0)
    ref = OracleOperator(vg, pg)
    x = torch.rand(vg.shape)

    class Single:
        proj_shape, slab_shape = tuple(ref(x).shape), vg.shape
        T = ref.T

        def __call__(self, v, out=None):
            return ref(v, out=out)

    rec = sirt(Single(), ref(x), 4)
    torch.testing.assert_close(r0, rec, rtol=1e-4, atol=1e-5)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 720, 1440):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def test_single_process_sharded_operator_is_identity_wrapper():
    vg, pg = geometries()
    S = ShardedOperator(vg, pg, make_local=OracleOperator)
    assert S.world == 1 and S.chunks == 1 and S.slab_shape == vg.shape and S.T.T is S
    x = torch.rand(vg.shape)
    torch.testing.assert_close(S(x), OracleOperator(vg, pg)(x))
    with pytest.raises(TypeError):
        ShardedOperator(vg.to_vec(), pg)


def test_slab_row_bounds_and_cropped_geometry(monkeypatch):
    """The row band of a z-slab holds every bilinear tap of every voxel of the slab (brute force over the voxel
    centres), is tight to a few rows, and the cropped geometry addresses the same pixels."""
    from tomosipo_b200.distributed import crop_detector_rows, default_bp_exchange, slab_row_bounds

    vg = ts.volume(shape=(24, 16, 20), size=(2.4, 1.6, 2.0))
    for pg in (ts.cone(angles=11, shape=(40, 30), size=(6.0, 4.5), src_orig_dist=4, src_det_dist=7),
               ts.parallel(angles=9, shape=(40, 30), size=(4.0, 3.0))):
        V = pg.det_shape[0]
        for z0, z1 in ((0, 6), (6, 12), (9, 15), (18, 24), (0, 24)):
            sub = vg[z0:z1]
            lo, hi = slab_row_bounds(sub, pg)
            zz, yy, xx = np.meshgrid(*[np.arange(n) for n in sub.shape], indexing="ij")
            centres = np.asarray(sub.lower_left_corner).reshape(1, 3) + (np.stack([zz, yy, xx], -1).reshape(-1, 3) + 0.5) * np.asarray(sub.voxel_size)[None]
            v = np.concatenate([pg.to_vec().project_point(tuple(c))[:, 0] for c in centres]) + V / 2.0
            t0, t1 = int(np.floor(v.min() - 0.5)), int(np.floor(v.max() - 0.5)) + 1      # first / last tap row
            assert lo <= max(t0, 0) and min(t1, V - 1) < hi
            assert max(t0, 0) - lo <= 4 and hi - 1 - min(t1, V - 1) <= 4
            if hi - lo < V:
                sub_pg = crop_detector_rows(pg, lo, hi)
                assert tuple(sub_pg.det_shape) == (hi - lo, pg.det_shape[1]) and sub_pg.num_angles == pg.num_angles
                p = centres[len(centres) // 3]
                a, b = pg.to_vec().project_point(tuple(p)), sub_pg.project_point(tuple(p))
                np.testing.assert_allclose(a[:, 0] + V / 2.0 - lo, b[:, 0] + (hi - lo) / 2.0, atol=1e-9)
                np.testing.assert_allclose(a[:, 1], b[:, 1], atol=1e-9)
    # angle subsets (the opt-in two-launch variant back-projects the own block and the others separately)
    pg = ts.cone(angles=11, shape=(40, 30), size=(6.0, 4.5), src_orig_dist=4, src_det_dist=7).to_vec()
    pick = np.array([7, 8, 0, 3])
    sub_pg = crop_detector_rows(pg, 5, 29, pick)
    assert sub_pg.num_angles == 4 and tuple(sub_pg.det_shape) == (24, 30)
    p = (0.3, -0.2, 0.5)
    np.testing.assert_allclose(pg.project_point(p)[pick, 0] + 20.0 - 5, sub_pg.project_point(p)[:, 0] + 12.0, atol=1e-9)
    np.testing.assert_allclose(pg.project_point(p)[pick, 1], sub_pg.project_point(p)[:, 1], atol=1e-9)
    monkeypatch.delenv("TSP_SHARD_BP", raising=False)
    big = ts.volume(shape=(64, 64, 64))
    circ = ts.cone(angles=96, shape=(64, 96), size=(64 * 1.5, 96 * 1.5), src_orig_dist=256, src_det_dist=384)
    assert default_bp_exchange(big, circ, 1) == "volume" and default_bp_exchange(big, circ, 8) == "rows"
    tilted = ts.rotate(pos=0, axis=(0, 1, 0), angles=np.pi / 2) * circ.to_vec()     # scan around y: slabs see all rows
    assert default_bp_exchange(big, tilted, 8) == "volume"
    monkeypatch.setenv("TSP_SHARD_BP", "volume")
    assert default_bp_exchange(big, circ, 8) == "volume"


def test_default_chunks_and_piece_layout(monkeypatch):
    from tomosipo_b200.distributed import default_chunks

    monkeypatch.delenv("TSP_SHARD_CHUNKS", raising=False)
    assert default_chunks(512, 1) == 1 and default_chunks(512, 8) == 4 and default_chunks(1024, 8) == 4
    assert default_chunks(128, 8) == 2 and default_chunks(50, 4) == 1
    monkeypatch.setenv("TSP_SHARD_CHUNKS", "1")
    assert default_chunks(512, 8) == 1
    # every slice is owned exactly once, whatever the padding
    for nz, world, k in ((512, 8, 4), (50, 4, 3), (9, 2, 2), (7, 8, 1), (100, 3, 4)):
        S = ShardedOperator.__new__(ShardedOperator)
        S.world, S.chunks, S.vol_shape = world, k, (nz, 4, 4)
        S.piece_nz = -(-nz // (k * world))
        S.chunk_nz = S.piece_nz * world
        owned = np.zeros(nz, int)
        for r in range(world):
            for row, lo, hi in S.slab_pieces(rank=r):
                assert row % S.piece_nz == 0 and 0 <= hi - lo <= S.piece_nz
                owned[lo:hi] += 1
        assert (owned == 1).all()
        for c in range(k):
            lo, hi = S.chunk_bounds(c)
            assert lo == min(c * S.chunk_nz, nz) and hi == min((c + 1) * S.chunk_nz, nz)
