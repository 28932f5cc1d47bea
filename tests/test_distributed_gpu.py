"""Angle-/slab-sharded operator over NCCL (needs >= 2 GPUs; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    import tomosipo_b200 as ts
    from tomosipo_b200.distributed import ShardedOperator, sirt

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        vg = ts.volume(shape=(50, 48, 56), size=(1.0, 0.96, 1.12))      # 50 slices: padded slabs for world = 4
        pg = ts.cone(angles=37, shape=(40, 64), size=(2.0, 3.2), src_orig_dist=4, src_det_dist=7)
        A = ts.operator(vg, pg)
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand(vg.shape, device="cuda", generator=g)
        w = torch.rand(A.range_shape, device="cuda", generator=g)
        Rw = torch.rand(A.range_shape, device="cuda", generator=g)
        y_full, bp_full = A(x), A.T(w)
        recs = []
        # chunks = 1: contiguous slabs, one collective per call; chunks = 3: interleaved pieces with the
        # exchange overlapped chunk by chunk on a second stream (50 slices: padded pieces); "rows": every rank
        # back-projects all angles into its own slab from an all_to_all of detector row bands
        # (row bands stored into the peers' buffers over NVLink by tsp_push_rows; or exchanged by NCCL all_to_all;
        # or, opt-in, the own angle block first with the all_to_all behind it and the other blocks added by a second launch)
        # default: the forward projector's own store fills the peers' buffers (tsp_fp_push), tsp_push_rows only for
        # arrays that did not come out of the forward projection
        env_keys = ("TSP_SHARD_NO_P2P", "TSP_SHARD_ROWS_SPLIT", "TSP_SHARD_NO_FP_PUSH")
        variants = ((1, "volume", {}), (3, "volume", {}), (1, "rows", {}), (1, "rows", {"TSP_SHARD_NO_FP_PUSH": "1"}),
                    (1, "rows", {"TSP_SHARD_NO_P2P": "1"}), (1, "rows", {"TSP_SHARD_ROWS_SPLIT": "1"}))
        bp_of_fp = A.T(y_full)
        for chunks, mode, env in variants:
            for key in env_keys:
                os.environ.pop(key, None)
            os.environ.update(env)
            S = ShardedOperator(vg, pg, chunks=chunks, bp_exchange=mode)
            assert S.bp_exchange == mode
            assert len(S.bp_operators()) == (chunks if mode == "volume" else 2 if "TSP_SHARD_ROWS_SPLIT" in env else 1)
            blk = slice(S.angle_lo, S.angle_hi)
            y_blk = S(S.scatter_volume(x))
            assert torch.equal(y_blk, y_full[:, blk, :])                    # FP per angle is independent: bit-exact
            slab = S.T(w[:, blk, :].contiguous())
            torch.testing.assert_close(S.gather_volume(slab), bp_full, rtol=1e-5, atol=1e-6)   # sum order differs
            for row, z0, z1 in S.slab_pieces():
                assert float(slab[row + z1 - z0: row + S.piece_nz].abs().sum()) == 0.0          # padding stays empty
            # fused residual == explicit three passes
            yb, Rb = y_full[:, blk, :].contiguous(), Rw[:, blk, :].contiguous()
            r_fused = S.residual(x.contiguous(), yb, Rb, torch.empty_like(yb))
            torch.testing.assert_close(r_fused, Rb * (S.local(x) - yb), rtol=1e-5, atol=1e-6)
            recs.append(S.gather_volume(sirt(S, yb, 5)))
            # A.T(A(x)): with the fused exchange the backprojection finds its rows already in place
            count = lambda: S.local.astra_projector.info().kernel_launches
            y_blk = S(S.scatter_volume(x))
            n0 = count()
            slab = S.T(y_blk)
            n1 = count()
            torch.testing.assert_close(S.gather_volume(slab), bp_of_fp, rtol=1e-5, atol=1e-5 * float(bp_of_fp.max()))
            y_blk = S(S.scatter_volume(x))
            y_blk *= 2.0                                                    # modified since: must be exchanged again
            torch.testing.assert_close(S.gather_volume(S.T(y_blk)), 2.0 * bp_of_fp, rtol=1e-5, atol=2e-5 * float(bp_of_fp.max()))
            if mode == "rows" and (not env or "TSP_SHARD_NO_FP_PUSH" in env):
                assert S._peer, "the peer-memory exchange was not used"
                assert n1 - n0 == (0 if not env else 1)                     # tsp_push_rows only without the fused store
                n0 = count()
                S.T(w[:, blk, :].contiguous())
                assert count() == n0 + 1                                    # a foreign array: tsp_push_rows
                S.close()
                assert S._peer is None
            elif mode == "rows":
                assert not S._peer
        assert float(torch.linalg.vector_norm(recs[0] - recs[1]) / torch.linalg.vector_norm(recs[0])) < 1e-5
        for key in env_keys:
            os.environ.pop(key, None)
        for other in recs[2:]:
            assert float(torch.linalg.vector_norm(recs[0] - other) / torch.linalg.vector_norm(recs[0])) < 1e-5
        rec = recs[1]
        if rank == 0:
            torch.save(rec.cpu(), os.path.join(tmpdir, "rec.pt"))
            # single-GPU SIRT with the same loop

            class Single:
                proj_shape, slab_shape = tuple(A.range_shape), tuple(vg.shape)
                T = A.T

                def __call__(self, v, out=None):
                    return A(v, out=out)

            torch.save(sirt(Single(), y_full, 5).cpu(), os.path.join(tmpdir, "ref.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_operator_nccl(tmp_path):
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rec, ref = torch.load(tmp_path / "rec.pt"), torch.load(tmp_path / "ref.pt")
    assert float(torch.linalg.vector_norm(rec - ref) / torch.linalg.vector_norm(ref)) < 1e-5
