"""Per-call timing of the small host-array path (README config: 128^3, 100 angles, 128x192)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomosipo_b200 as ts
from tomosipo_b200 import _backend as B

vg = ts.volume(shape=128)
pg = ts.cone(size=np.sqrt(2), cone_angle=1 / 2, angles=100, shape=(128, 192))
A = ts.operator(vg, pg)
P = A.astra_projector
x = np.random.default_rng(0).random(A.domain_shape).astype(np.float32)
y = np.zeros(A.range_shape, np.float32)
xp = torch.from_numpy(x).pin_memory().numpy(); yp = torch.zeros(tuple(A.range_shape)).pin_memory().numpy()

def t(f, n=20):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

print("raw ABI FP pageable  %.3f ms" % t(lambda: P.project(B.FP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)))
print("raw ABI FP pinned    %.3f ms" % t(lambda: P.project(B.FP, False, xp.ctypes.data, yp.ctypes.data, B.MEM_HOST, 0, 0)))
print("raw ABI BP pageable  %.3f ms" % t(lambda: P.project(B.BP, False, x.ctypes.data, y.ctypes.data, B.MEM_HOST, 0, 0)))
print("raw ABI BP pinned    %.3f ms" % t(lambda: P.project(B.BP, False, xp.ctypes.data, yp.ctypes.data, B.MEM_HOST, 0, 0)))
xd = torch.from_numpy(x).cuda(); yd = torch.empty(tuple(A.range_shape), device="cuda")
print("device FP            %.3f ms" % t(lambda: A(xd, out=yd)))
print("device BP            %.3f ms" % t(lambda: A.T(yd, out=xd)))
print("A(x) f32 pageable in, new out %.3f ms" % t(lambda: A(x)))
x64 = x.astype(np.float64)
import warnings; warnings.simplefilter("ignore")
print("A(x) f64 in, new out          %.3f ms" % t(lambda: A(x64)))
print("np.empty+fill 10MB   %.3f ms" % t(lambda: np.empty(A.range_shape, np.float32).fill(1.0)))
print("pinned_empty+fill    %.3f ms" % t(lambda: B.pinned_empty(A.range_shape).fill(1.0)))
t0 = time.perf_counter(); torch.cuda.synchronize(); print("sync %.3f ms" % ((time.perf_counter() - t0) * 1e3))
