"""Supersampling 2x at 256^3: staged-kernel path vs the direct kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tomosipo_b200 as ts
n = 256
vg = ts.volume(shape=n, size=1)
pg = ts.cone(angles=180, shape=(n, 3 * n // 2), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6)
A = ts.operator(vg, pg, voxel_supersampling=2, detector_supersampling=2)
x = torch.rand(tuple(A.domain_shape), device="cuda"); y = torch.empty(tuple(A.range_shape), device="cuda"); xb = torch.empty_like(x)
for _ in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); A(x, out=y); e[1].record(); A.T(y, out=xb); e[2].record(); torch.cuda.synchronize()
    print(f"ss=2 {os.environ.get('TSP_SS_DIRECT','staged')}: fp {e[0].elapsed_time(e[1]):.2f} ms  bp {e[1].elapsed_time(e[2]):.2f} ms", flush=True)
