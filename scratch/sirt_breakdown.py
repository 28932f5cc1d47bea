"""Where a sharded SIRT iteration spends its time (run under torchrun): CUDA-event timings, max over ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import tomosipo_b200 as ts
from tomosipo_b200.distributed import ShardedOperator, sirt, sirt_weights
from bench import workload

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
vg, pg = workload()

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

out = {}
for K in (1, 4):
    S = ShardedOperator(vg, pg, chunks=K)
    x = torch.rand(S.slab_shape, device=dev); y = torch.rand(S.proj_shape, device=dev)
    R = torch.rand(S.proj_shape, device=dev); C = torch.rand(S.slab_shape, device=dev)
    ytmp = torch.empty_like(y); xb = torch.empty_like(x)
    full = S._full_volume(y); partial = S._partial_volume(y)
    out[f"K{K} fp_plain_local"] = timed(lambda: S.local(full[: S.vol_shape[0]], out=ytmp))
    out[f"K{K} fp_fused_local"] = timed(lambda: S.residual(full[: S.vol_shape[0]], y, R, ytmp))
    out[f"K{K} all_gather"] = timed(lambda: [S._all_gather_chunk(full, x, c) for c in range(K)])
    out[f"K{K} bp_compute_only"] = timed(lambda: S._bp_chunks(y, partial, lambda c: None))
    out[f"K{K} reduce_scatter_only"] = timed(lambda: [S._reduce_scatter_chunk(S._piece_view(xb, c), partial, c) for c in range(K)])
    out[f"K{K} bp_full"] = timed(lambda: S.T(y, out=xb))
    out[f"K{K} fp_full"] = timed(lambda: S(x, out=ytmp))
    W = (R, C)
    xs = torch.zeros(S.slab_shape, device=dev)
    out[f"K{K} sirt_iter"] = timed(lambda: sirt(S, y, 1, x=xs, weights=W))
    out[f"K{K} sirt_5iters/5"] = timed(lambda: sirt(S, y, 5, x=xs, weights=W), reps=2) / 5
    del S, x, y, R, C, ytmp, xb, full, partial
if rank == 0:
    for k, v in out.items(): print(f"{k:28s} {v:8.3f} ms")
dist.destroy_process_group()
