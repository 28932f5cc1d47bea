"""cfg 5 (learned primal-dual slab): 1x256x256 slab, 256 angles, 384 det, batch 16, autograd fwd+bwd."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tomosipo_b200 as ts
from tomosipo_b200.torch_support import to_autograd

vg = ts.volume(shape=256, size=1)
pg = ts.parallel(angles=256, shape=(256, 384), size=(1, 1.5))
A = ts.operator(vg[:1], pg.to_vec()[:, :1, :])
f = to_autograd(A, is_2d=True, num_extra_dims=2)
fT = to_autograd(A.T, is_2d=True, num_extra_dims=2)
torch.manual_seed(0)
x = torch.randn(16, 1, 256, 256, device="cuda", requires_grad=True)
def step():
    y = f(x)
    z = fT(y)
    z.sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"cfg5: fwd(FP,BP)+bwd(FP,BP) batch 16: {dt*1e3:.3f} ms per step = {dt*1e6/64:.1f} us per projector application; launches so far {A.astra_projector.info().kernel_launches}")

# the same step captured once into a CUDA graph and replayed
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        x.grad = None
        step()
torch.cuda.current_stream().wait_stream(side)
x.grad = None
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
for _ in range(3): graph.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): graph.replay()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"cfg5 graph replay: {dt*1e3:.3f} ms per step")
