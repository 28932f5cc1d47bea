"""e2e (host pinned arrays) FP and BP timing through the operator API at cfg 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomosipo_b200 as ts
n=512
vg = ts.volume(shape=n, size=1)
pg = ts.cone(angles=720, shape=(n, 3*n//2), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
A = ts.operator(vg, pg)
xh = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data).pin_memory().numpy()
yh = torch.empty(tuple(A.range_shape), dtype=torch.float32).pin_memory().numpy()
xbh = torch.empty(tuple(A.domain_shape), dtype=torch.float32).pin_memory().numpy()
A(xh, out=yh); A.T(yh, out=xbh)
torch.cuda.synchronize()
tf=[];tb=[]
for _ in range(3):
    t0=time.perf_counter(); A(xh, out=yh); t1=time.perf_counter(); A.T(yh, out=xbh); t2=time.perf_counter()
    tf.append(t1-t0); tb.append(t2-t1)
print(f"chunks={os.environ.get('TSP_HOST_CHUNKS','8')} e2e fp {min(tf)*1e3:.1f} ms bp {min(tb)*1e3:.1f} ms  -> {2*n**3*720/(min(tf)+min(tb))/1e9:.0f} GUPS")
