"""One FP + BP at cfg 3 (or a smaller size) through the C ABI: the command ncu wraps.
usage: python scratch/prof_step.py [n] [angles] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomosipo_b200 as ts
from tomosipo_b200 import _backend as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
na = int(sys.argv[2]) if len(sys.argv) > 2 else 720
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kind = sys.argv[4] if len(sys.argv) > 4 else "cone"
vg = ts.volume(shape=n, size=1)
if kind == "cone":
    pg = ts.cone(angles=na, shape=(n, 3 * n // 2), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
else:
    pg = ts.parallel(angles=na, shape=(n, 3 * n // 2), size=(1.25, 1.875)).to_vec()
A = ts.operator(vg, pg)
x = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data).cuda()
y = torch.empty(tuple(A.range_shape), device="cuda")
xb = torch.empty_like(x)
for _ in range(reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); A(x, out=y); e[1].record(); A.T(y, out=xb); e[2].record()
    torch.cuda.synchronize()
    print(f"{kind} n={n} angles={na}: fp {e[0].elapsed_time(e[1]):.2f} ms  bp {e[1].elapsed_time(e[2]):.2f} ms", flush=True)
print("checksums", float(y.double().sum()), float(xb.double().sum()))
