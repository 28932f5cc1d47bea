"""BASELINE.json configs[0], [1], [4] through the public API (configs[2], [3]: bench.py).  One JSON line each."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomosipo_b200 as ts
from tomosipo_b200.torch_support import to_autograd

def sync(): torch.cuda.synchronize()

# ---- configs[0]: README SIRT verbatim (README.md:139-164 of the reference): host float64 numpy arrays, 100 iterations
def cfg1():
    pg = ts.cone(size=np.sqrt(2), cone_angle=1 / 2, angles=100, shape=(128, 192))
    vg = ts.volume(shape=128)
    A = ts.operator(vg, pg)
    phantom = np.zeros(A.domain_shape); phantom[20:50, 20:50, 20:50] = 1.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = A(phantom)
        R = 1 / A(np.ones(A.domain_shape)); R = np.minimum(R, 1 / ts.epsilon)
        C = 1 / A.T(np.ones(A.range_shape)); C = np.minimum(C, 1 / ts.epsilon)
        def run(n):
            x = np.zeros(A.domain_shape)
            for _ in range(n):
                x += C * A.T(R * (y - A(x)))
            return x
        run(3); sync()
        t0 = time.perf_counter(); x = run(100); sync(); dt = time.perf_counter() - t0
        # the same loop with float32 torch CUDA tensors (README.md:170-184) and the fused library loop
        yt = torch.from_numpy(y.astype(np.float32)).cuda()
        from tomosipo_b200.algorithms import sirt
        sirt(A, yt, 3); sync()
        t0 = time.perf_counter(); xg = sirt(A, yt, 100); sync(); dtg = time.perf_counter() - t0
    err = float(np.linalg.norm(x - xg.cpu().numpy()) / np.linalg.norm(x))
    upd = 2 * 128 ** 3 * 100 * 100
    print(json.dumps({"config": "configs[0] README SIRT 128^3, 100 angles, 128x192, 100 iterations",
                      "host_numpy_float64_s": dt, "host_numpy_GUPS": upd / dt / 1e9,
                      "cuda_tensors_fused_s": dtg, "cuda_GUPS": upd / dtg / 1e9, "rel_diff_host_vs_cuda": err,
                      "residual": float(np.linalg.norm(x - phantom) / np.linalg.norm(phantom))}))

# ---- configs[1]: parallel 256^3, 180 angles, 256x256, torch CUDA, autograd forward + backward
def cfg2():
    A = ts.operator(ts.volume(shape=256), ts.parallel(angles=180, shape=(256, 256)))
    f = to_autograd(A)
    x = torch.from_numpy(ts.phantom.hollow_box(ts.data(A.domain)).data).cuda().requires_grad_(True)
    def step():
        x.grad = None
        y = f(x); y.backward(y)
    for _ in range(3): step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): step()
    e1.record(); sync()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"config": "configs[1] parallel3d 256^3, 180 angles, 256x256, autograd fwd+bwd (FP + BP)",
                      "ms_per_step": ms, "GUPS": 2 * 256 ** 3 * 180 / ms / 1e6}))

# ---- configs[4]: learned primal-dual slab, batch 16 (notebooks/learned_pd.py): 2 FP + 2 BP applications per step
def cfg5():
    vg = ts.volume(shape=256, size=1)
    pg = ts.parallel(angles=256, shape=(256, 384), size=(1, 1.5))
    A = ts.operator(vg[:1], pg.to_vec()[:, :1, :])
    f, fT = to_autograd(A, is_2d=True, num_extra_dims=2), to_autograd(A.T, is_2d=True, num_extra_dims=2)
    torch.manual_seed(0)
    x = torch.randn(16, 1, 256, 256, device="cuda", requires_grad=True)
    def step():
        x.grad = None
        z = fT(f(x)); z.backward(z)
    for _ in range(3): step()
    sync()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n): step()
    sync(); ms = (time.perf_counter() - t0) / n * 1e3
    upd = 16 * (2 * 256 * 384 * 256 + 2 * 256 * 256 * 256)   # ray-slices (FP) + voxel-angles (BP), fwd + bwd
    print(json.dumps({"config": "configs[4] learned-PD slab 1x256x256, 256 angles, 384 det, batch 16, fwd + bwd",
                      "ms_per_step": ms, "us_per_projector_application": ms * 1e3 / 64, "GUPS": upd / ms / 1e6}))

# ---- the reference's own published SIRT benchmark (notebooks/sirt_benchmark.py:29-34, notebooks/README.md:43-47):
#      200 iterations, ts.parallel 256^3 / 384 angles / 256 x 384 detector, torch CUDA tensors, in-place loop;
#      published: 17.914 s on an RTX 2080 Ti (ASTRA's own SIRT3D_CUDA: 19.288 s)
def published_sirt():
    A = ts.operator(ts.volume(shape=256), ts.parallel(angles=384, shape=(256, 384)))
    x = torch.from_numpy(ts.phantom.hollow_box(ts.data(A.domain)).data).cuda()
    y = A(x)
    from tomosipo_b200.algorithms import sirt, _weights
    R, C = _weights(A, y, ts.epsilon)
    def loop(n):                       # the reference's loop, line by line (sirt_benchmark.py:130-136)
        x_cur = torch.zeros_like(x); y_tmp = torch.empty_like(y); x_tmp = torch.empty_like(x)
        for _ in range(n):
            A(x_cur, out=y_tmp); y_tmp -= y; y_tmp *= R
            A.T(y_tmp, out=x_tmp); x_tmp *= C; x_cur -= x_tmp
        return x_cur
    loop(3); sync()
    t0 = time.perf_counter(); xa = loop(200); sync(); dt = time.perf_counter() - t0
    sirt(A, y, 3); sync()
    t0 = time.perf_counter(); xb = sirt(A, y, 200); sync(); dtf = time.perf_counter() - t0
    print(json.dumps({"config": "published analogue: SIRT 200 it, parallel 256^3, 384 angles, 256x384 (sirt_benchmark.py)",
                      "reference_loop_s": dt, "fused_s": dtf, "published_s_2080ti": 17.914,
                      "GUPS_reference_loop": 2 * 256 ** 3 * 384 * 200 / dt / 1e9,
                      "rel_diff_loop_vs_fused": float((xa - xb).norm() / xa.norm())}))

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cfg1(); cfg2(); cfg5(); published_sirt()
