#!/usr/bin/env python
"""Benchmark of the projection hot path (FP3D + BP3D) on B200.

Contract: ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line.

* workload   BASELINE.json configs[2]: cone_vec 512^3 volume, 720 angles,
             512 x 768 detector (SURVEY.md 8d cfg 3).  A "step" is one forward
             projection followed by one backprojection of the whole problem.
* metric     GUPS = voxels x angles / second / 1e9, summed over FP and BP.
* value      device-resident inputs, CUDA-event timed.
* e2e        the same step through the public operator API with pinned HOST
             arrays (H2D + kernels + D2H inside the timed region).
* roofline   for the dominant kernel: algorithmic bytes 4*(vol + proj) per
             launch over its measured duration, against the measured HBM peak;
             `traffic` = DRAM bytes per launch from the committed ncu capture
             (profiles/r02_traffic.json, `traffic_source` says so); `interp` adds
             the in-SM interpolation view (updates/clk/SM) that actually bounds
             these kernels.
* sirt       SIRT iterations/s on the same problem (BASELINE.json metric, second
             half): fused tsp_sirt at N = 1, sharded loop at N > 1.
* cfg4_sirt  SIRT ms / iteration at BASELINE configs[3] (1024^3 x 1440), the
             configuration the north_star's scaling target is written on.
* sharded_parity_rel_l2 (N > 1): the NCCL-sharded operator against the
             single-GPU operator on a 128^3 problem, max over ranks.
* cpu_baseline / --impl reference: the CPU restatement (oracle/, fp32,
             OpenMP, all host cores) on the headline volume and detector over a
             sample of the 720 angles (cost is linear in angles).
             ASTRA -- the reference's engine -- has no CPU 3-D projector and is
             not installable here, so the port is the CPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP3D+BP3D GUPS (voxels x angles / s), cone_vec 512^3 x 720 angles"
UNIT = "GUPS"
WORKLOAD_CFG3 = "cone_vec 512^3 vol, 720 angles, 512x768 det (BASELINE configs[2]), FP+BP per step"
WORKLOAD_CFG4 = "cone_vec 1024^3 vol, 1440 angles, 1024x1536 det (BASELINE configs[3]), FP+BP per step"


def workload(n=512, n_angles=720):
    """cfg 3: ts.volume(shape=n, size=1) and the cone_vec geometry
    ts.cone(angles, shape=(n, 1.5 n), size=(1.875, 2.8125), SOD 4, SDD 6).to_vec()."""
    import tomosipo_b200 as ts

    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=n_angles, shape=(n, 3 * n // 2), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
    return vg, pg


TRAFFIC_FILE = "profiles/r02_traffic.json"
TRAFFIC_SOURCE = (f"{TRAFFIC_FILE}: dram__bytes_read.sum + dram__bytes_write.sum per launch from the builder's committed "
                  "`ncu --set full` capture of this workload (not measured in this run)")


def load_traffic():
    """DRAM bytes per launch of the two hot kernels, from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, TRAFFIC_FILE)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []   # (arrival time, csv line)
        self.proc = None
        self.window = None  # (t0, t1) of the timed region; samples before it were taken under warm-up load

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [s for t, s in self.samples if self.window and self.window[0] <= t <= self.window[1] + 0.15]
        scope = "timed region"
        if len(lines) < 2:  # region shorter than the sampling period: use every sample taken under load (warm-up + timed)
            lines = [s for t, s in self.samples]
            scope = "warm-up + timed region"
        for s in lines:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


def omp_threads():
    """Threads the OpenMP oracle actually uses (torchrun exports OMP_NUM_THREADS=1 to its children)."""
    v = os.environ.get("OMP_NUM_THREADS")
    return int(v) if v and v.isdigit() else (os.cpu_count() or 1)


CPU_SAMPLE_ANGLES = 16


def cpu_sample(n=512, n_total_angles=720, n_sample=CPU_SAMPLE_ANGLES, repeats=1):
    """Bounded CPU sample of the SAME workload: the headline volume and detector (n^3, n x 1.5n) and every
    (n_total / n_sample)-th of its angles.  FP and BP cost are linear in the number of angles (each angle is an
    independent pass over the volume), so GUPS on the sample is GUPS on the whole angle set; the sampled angles
    cover the circle, hence the same mix of x- and y-marching rays."""
    from oracle import oracle as O

    det = (n, 3 * n // 2)
    every = max(1, n_total_angles // n_sample)
    t = np.linspace(0, 2 * np.pi, n_total_angles, endpoint=False)[::every][:n_sample]
    vec = O.cone_vectors(t, 2.8125 / det[1], 1.875 / det[0], 4.0, 2.0)
    Q = O.OracleProjector(O.CONE_VEC, (n, n, n), [-0.5] * 3, [0.5] * 3, det, vec)
    x = O.hollow_box(n)
    y = np.zeros(Q.proj_shape, np.float32)
    xb = np.zeros(Q.vol_shape, np.float32)
    t0 = time.perf_counter()
    for _ in range(repeats):
        Q.fp(x, out=y, dtype=np.float32)
        Q.bp(y, out=xb, dtype=np.float32)
    dt = (time.perf_counter() - t0) / repeats
    updates = 2.0 * n ** 3 * len(t)
    return updates / dt / 1e9, dt, (f"cone_vec {n}^3 vol, {n}x{3 * n // 2} det, {len(t)} of the {n_total_angles} angles "
                                    f"(every {every}th), FP+BP, fp32 OpenMP port of the same arithmetic")


def run_reference(args):
    """--impl reference: the CPU port on all host cores (rank 0 only), same volume / detector / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1); set before the OpenMP runtime starts
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cores = omp_threads()
    # One short pass warms the thread pool / page cache and measures the rate; the per-step sample is then sized so
    # that exactly --steps steps end within ~3 minutes on this host: 4 ... 16 of the 720 angles, a multiple of 4.
    _, dt4, _ = cpu_sample(n_sample=4)
    per_angle = dt4 / 4
    n_sample = int(180.0 / max(args.steps, 1) / per_angle) // 4 * 4
    n_sample = max(4, min(CPU_SAMPLE_ANGLES, n_sample))
    vals, ms = [], []
    for i in range(args.steps):
        g, dt, sample = cpu_sample(n_sample=n_sample)
        vals.append(g); ms.append(dt * 1e3)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_CFG3, "phantom": "hollow_box",
                   "sample": f"each step = FP+BP over {n_sample} of the 720 angles, spread over the circle (cost is linear in "
                             "angles); GUPS counts the sampled updates only",
                   "note": "ASTRA (the reference engine) is CUDA-only and absent; CPU port of the same arithmetic"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_host_to_gpu(torch, local):
    """Restrict this process to the CPUs NVML lists as local to its GPU (same NUMA node / PCIe root) while the pinned
    host buffers of the end-to-end leg are allocated and used: torchrun does not bind its workers, and pinned pages on
    the far socket cost host<->device bandwidth.  Returns (affinity to restore | None, description)."""
    if os.environ.get("TSP_BENCH_NO_BIND") or not hasattr(os, "sched_setaffinity"):
        return None, "none"
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        before = os.sched_getaffinity(0)
        cpus = near & before
        if not cpus or cpus == before:
            return None, f"unchanged ({len(before)} CPUs allowed, {len(near)} local to the GPU)"
        os.sched_setaffinity(0, cpus)
        return before, f"{len(cpus)} of {len(before)} allowed CPUs (NVML affinity of the GPU)"
    except Exception as exc:  # no NVML / no permission: run unbound
        return None, f"none ({type(exc).__name__})"


def sharded_parity_check(ts, ShardedOperator, dev, world):
    """The angle- / z-sharded operator (NCCL all_gather / reduce_scatter, chunked overlap) against the single-GPU
    operator on a 128^3 x 96-angle cone problem: relative L2 of this rank's FP angle block and BP z-slab, max over
    ranks.  Differences come only from the order of the fp32 sum over angle blocks (~1e-7)."""
    import torch
    import torch.distributed as dist

    n, na = 128, 96
    vg = ts.volume(shape=n, size=1)
    pg = ts.cone(angles=na, shape=(n, 3 * n // 2), size=(1.875, 2.8125), src_orig_dist=4, src_det_dist=6).to_vec()
    S = ShardedOperator(vg, pg)
    A = ts.operator(vg, pg)
    g = torch.Generator(device=dev).manual_seed(7)           # same seed on every rank: replicated inputs
    x_full = torch.rand(tuple(A.domain_shape), device=dev, generator=g)
    w_full = torch.rand(tuple(A.range_shape), device=dev, generator=g)
    y_ref = A(x_full)[:, S.angle_lo:S.angle_hi, :]
    y_blk = S(S.scatter_volume(x_full))
    e_fp = torch.linalg.vector_norm(y_blk - y_ref) / torch.linalg.vector_norm(y_ref)
    xb_ref = S.scatter_volume(A.T(w_full))
    xb_slab = S.T(w_full[:, S.angle_lo:S.angle_hi, :].contiguous())
    S.zero_padding_(xb_slab)
    e_bp = torch.linalg.vector_norm(xb_slab - xb_ref) / torch.linalg.vector_norm(xb_ref)
    t = torch.stack([e_fp, e_bp]).float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mode = S.bp_exchange
    S.close()                                                # band buffers of the row exchange (collective)
    return {"fp": float(t[0]), "bp": float(t[1]), "problem": f"cone {n}^3 x {na} angles x {n}x{3 * n // 2}, {S.chunks} z-chunks, bp_exchange={mode}",
            "ranks": world}


def cfg4_sirt(ts, ShardedOperator, dev, world, local, iterations=3):
    """SIRT ms / iteration at BASELINE configs[3] (1024^3, 1440 angles, 1024 x 1536) on `world` GPUs: fused tsp_sirt at
    N = 1, the sharded overlapped loop at N > 1; one warm-up iteration, weights are set-up (untimed)."""
    import torch
    import torch.distributed as dist

    vg, pg = workload(1024, 1440)
    S = ShardedOperator(vg, pg)
    y = torch.empty(S.proj_shape, device=dev, dtype=torch.float32)
    g = torch.Generator(device=dev).manual_seed(11)
    y.uniform_(0.0, 1.0, generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world == 1:
        from tomosipo_b200.algorithms import _weights

        A = S.local
        R_, C_ = _weights(A, y, ts.epsilon)
        xs = torch.zeros(tuple(A.domain_shape), device=dev)
        y_tmp = torch.empty_like(y)
        strm = torch.cuda.current_stream().cuda_stream
        run = lambda k: A.astra_projector.sirt(xs.data_ptr(), y.data_ptr(), R_.data_ptr(), C_.data_ptr(), y_tmp.data_ptr(), k,
                                               device=local, stream=strm)
    else:
        from tomosipo_b200.distributed import sirt as sirt_sharded, sirt_weights

        W_ = sirt_weights(S, dev)
        xs = torch.zeros(S.slab_shape, device=dev)
        run = lambda k: sirt_sharded(S, y, k, x=xs, weights=W_)
    run(1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    run(iterations)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iterations
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"workload": "cone 1024^3 vol, 1440 angles, 1024x1536 det (BASELINE configs[3]), SIRT", "ms_per_iter": ms,
            "iters_per_s": 1e3 / ms, "iterations": iterations, "n_gpus": world,
            "gups": 2.0 * 1024.0 ** 3 * 1440 / (ms * 1e-3) / 1e9}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import tomosipo_b200 as ts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: tomosipo_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    big = args.workload == "cfg4"
    vg, pg = workload(1024, 1440) if big else workload()
    n = vg.shape[0]
    n_angles = pg.num_angles
    # N > 1: projections sharded by angle, volume sharded in z-slabs.
    #   A(x):   all_gather of the z-slabs, then FP of the rank's angle block
    #   A.T(y): BP of the rank's angle block, then reduce_scatter(sum) into z-slabs
    from tomosipo_b200.distributed import ShardedOperator

    S = ShardedOperator(vg, pg)
    A = S.local
    P = A.astra_projector
    a_lo, a_hi = S.angle_lo, S.angle_hi
    x_full = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data).to(dev)
    x = S.scatter_volume(x_full)                        # this rank's z-slab (whole volume when N == 1)
    del x_full
    y = torch.empty(S.proj_shape, device=dev, dtype=torch.float32)
    xb = torch.empty(S.slab_shape, device=dev, dtype=torch.float32)

    def fp():
        S(x, out=y)

    def bp():
        S.T(y, out=xb)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        fp()
        bp()
    torch.cuda.synchronize()

    def launches_so_far():
        """Kernels launched by this rank's projectors: the local operator and, at N > 1, the z-chunk sub-operators."""
        ops = {id(S.local): S.local}
        if world > 1:
            ops.update({id(op): op for op in S.bp_operators()})
        return sum(op.astra_projector.info().kernel_launches for op in ops.values())

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = launches_so_far()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    t_begin.record()
    for i in range(args.steps):
        ev[i][0].record()
        fp()
        ev[i][1].record()
        bp()
        ev[i][2].record()
    t_end.record()
    torch.cuda.synchronize()
    sampler.window = (wall0, time.time())
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    launches = launches_so_far() - launches0
    fp_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))

    updates_step = 2.0 * n ** 3 * n_angles  # whole job, FP + BP
    value = updates_step * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end to end: the call a user makes, A(x) / A.T(y) on pinned HOST arrays
    nvox, npix = int(np.prod(A.domain_shape)), int(np.prod(A.range_shape))
    e2e = None
    if not args.skip_e2e:
        restore, binding = bind_host_to_gpu(torch, local)
        try:
            e2e_steps = max(10, min(args.steps, 20))
            if world == 1:
                # one GPU: numpy arrays straight into A(x) / A.T(y) (the library's chunked copy / compute pipeline)
                xh = torch.from_numpy(ts.phantom.hollow_box(ts.data(vg)).data).pin_memory().numpy()
                yh = torch.empty(tuple(A.range_shape), dtype=torch.float32).pin_memory().numpy()
                xbh = torch.empty(tuple(A.domain_shape), dtype=torch.float32).pin_memory().numpy()

                def e2e_step():
                    A(xh, out=yh)
                    A.T(yh, out=xbh)
                    return float(xbh[n // 2, n // 2, n // 2])

                h2d = 4 * (nvox + npix)   # FP: volume in; BP: projections in
                d2h = 4 * (npix + nvox)   # FP: projections out; BP: volume out
            else:
                # N GPUs: every rank keeps its shard of each array in pinned host memory.  Per step and rank: slab up ->
                # all_gather + FP -> angle block down, and (independent data, as for a user holding measured projections)
                # block up -> BP + reduce_scatter -> slab down.  The uploads run on a copy-in stream, the FP result's
                # download on a copy-out stream behind the BP kernels; the step ends when both results are on the host.
                # The whole job moves each array once per direction.
                xs_h = torch.empty(S.slab_shape, dtype=torch.float32).pin_memory()
                xs_h.copy_(x)
                w_h = torch.empty(S.proj_shape, dtype=torch.float32).pin_memory()
                w_h.copy_(y)
                yb_h = torch.empty(S.proj_shape, dtype=torch.float32).pin_memory()
                xb_h = torch.empty(S.slab_shape, dtype=torch.float32).pin_memory()
                xd, wd = torch.empty_like(x), torch.empty_like(y)
                s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
                ev_x, ev_w, ev_fp = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

                def e2e_step():
                    cur = torch.cuda.current_stream()
                    s_in.wait_stream(cur)
                    with torch.cuda.stream(s_in):
                        xd.copy_(xs_h, non_blocking=True)    # H2D: this rank's slab
                        ev_x.record()
                        wd.copy_(w_h, non_blocking=True)     # H2D: this rank's angle block (behind the FP kernels)
                        ev_w.record()
                    cur.wait_event(ev_x)
                    S(xd, out=y)                             # all_gather + FP of the rank's angle block
                    ev_fp.record(cur)
                    s_out.wait_event(ev_fp)
                    with torch.cuda.stream(s_out):
                        yb_h.copy_(y, non_blocking=True)     # D2H: FP result (behind the BP kernels)
                    cur.wait_event(ev_w)
                    S.T(wd, out=xb)                          # BP + reduce_scatter
                    xb_h.copy_(xb, non_blocking=True)        # D2H: slab
                    cur.wait_stream(s_out)
                    torch.cuda.synchronize()                 # both results are on the host: end of the step
                    return float(xb_h[0, 0, 0]) + float(yb_h[0, 0, 0])

                h2d = 4 * (int(np.prod(S.slab_shape)) + int(np.prod(S.proj_shape))) * world
                d2h = h2d

            e2e_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([e2e_s], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            e2e = {"value": updates_step * e2e_steps / e2e_s / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                   "path": "A(x) / A.T(y) on pinned numpy arrays (library host pipeline)" if world == 1 else
                           "pinned host shards <-> ShardedOperator (slab up -> FP -> block down | block up -> BP -> slab down; "
                           "copies on side streams, one synchronize per step)",
                   "host_cpu_binding": binding}
        except Exception as exc:  # keep the device-resident line even if the host leg cannot run (e.g. no pinned memory)
            print(f"[bench] host-array leg failed on rank {rank}: {exc!r}", file=sys.stderr)
            e2e = None
        finally:
            if restore is not None:
                os.sched_setaffinity(0, restore)

    # ---- SIRT iterations / s on the same problem (device-resident)
    sirt_iters = max(2, min(args.steps, 5))
    if world == 1:
        from tomosipo_b200.algorithms import _weights

        # weights R = 1/A(1), C = 1/A^T(1) are set-up (notebooks/sirt_benchmark.py:116-128); iterations are timed
        R_, C_ = _weights(A, y, ts.epsilon)
        xs = torch.zeros_like(xb); y_tmp = torch.empty_like(y)
        strm = torch.cuda.current_stream().cuda_stream
        P.sirt(xs.data_ptr(), y.data_ptr(), R_.data_ptr(), C_.data_ptr(), y_tmp.data_ptr(), 1, device=local, stream=strm)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        P.sirt(xs.data_ptr(), y.data_ptr(), R_.data_ptr(), C_.data_ptr(), y_tmp.data_ptr(), sirt_iters, device=local,
               stream=strm)
        e1.record()
        torch.cuda.synchronize()
        sirt_ms = e0.elapsed_time(e1) / sirt_iters
        del xs, y_tmp, R_, C_
    else:
        from tomosipo_b200.distributed import sirt as sirt_sharded, sirt_weights

        W_ = sirt_weights(S, dev)                       # set-up, untimed (as at N = 1)
        xs = torch.zeros(S.slab_shape, device=dev)
        sirt_sharded(S, y, 1, x=xs, weights=W_)
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        sirt_sharded(S, y, sirt_iters, x=xs, weights=W_)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sirt_ms = float(t.item()) / sirt_iters
        del xs, W_

    # kernel names from the projectors that actually ran: FP on the rank-local operator, BP on it too at N = 1 but on the
    # z-chunk sub-operators at N > 1 (ShardedOperator._bp_chunks)
    info = P.info()
    bp_infos = [info] if world == 1 else [op.astra_projector.info() for op in S.bp_operators()]
    bp_name = "bp_tma_kernel" if all(i.bp_uses_tma for i in bp_infos) else "bp_kernel"
    fp_name = "fp_tma_kernel" if info.fp_uses_tma else "fp_cols_kernel"
    n_chunks = S.chunks
    if world > 1 and S.bp_exchange == "rows":
        lo_, hi_ = S.row_bounds[rank]
        how = ("exchanged by NCCL all_to_all" if not S._peer else
               "stored into the peers' buffers over NVLink by the forward projector's own stores (tsp_fp_push, CUDA IPC buffers; "
               "tsp_push_rows for arrays that did not come out of A(x))" if S._peer["fp_push"] else
               "stored by the ranks into each other's buffers over NVLink (tsp_push_rows, CUDA IPC buffers)")
        bp_scheme = (f"detector row bands (rank 0: rows {lo_}:{hi_} of {S.proj_shape[0]}) {how} -> BP of all angles "
                     "into the rank's z-slab")
        sirt_scheme = ("sharded: all_gather of the slabs -> fused residual FP -> row bands to the peers -> BP of all angles "
                       "into the rank's slab -> local update")
    else:
        bp_scheme = f"BP in {n_chunks} z-chunks -> NCCL reduce_scatter per chunk (overlapped)"
        sirt_scheme = (f"sharded: fused residual FP; BP in {n_chunks} z-chunks, each chunk's NCCL reduce_scatter + update + "
                       "all_gather on a side stream behind the next chunk's kernel")

    # ---- N > 1: the sharded operator against the single-GPU operator on a small problem (every rank, max over ranks)
    sharded_parity = None
    if world > 1:
        try:
            sharded_parity = sharded_parity_check(ts, ShardedOperator, dev, world)
        except Exception as exc:
            print(f"[bench] sharded parity check failed on rank {rank}: {exc!r}", file=sys.stderr)

    # ---- BASELINE configs[3] (cone 1024^3 x 1440 angles x 1024x1536): SIRT ms / iteration, the configuration the
    # north_star's scaling target is written on; carried in every line so that the 1 -> 8 GPU curve is driver-visible
    cfg4 = None
    if not big and not args.skip_cfg4:
        del x, y, xb
        S = A = P = None
        torch.cuda.empty_cache()
        try:
            cfg4 = cfg4_sirt(ts, ShardedOperator, dev, world, local, iterations=3)
        except Exception as exc:
            print(f"[bench] cfg4 SIRT leg failed on rank {rank}: {exc!r}", file=sys.stderr)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = load_peaks()
    traffic = load_traffic()
    b_alg = 4.0 * (nvox + npix)  # bytes per FP or per BP launch (SET mode), SURVEY.md 8d
    dom = bp_name if bp_ms >= fp_ms else fp_name
    dom_ms = max(bp_ms, fp_ms)
    achieved = b_alg / (dom_ms * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    upd = float(n) ** 3 * (a_hi - a_lo)
    interp = {
        "fp_gups": upd / (fp_ms * 1e-3) / 1e9, "bp_gups": upd / (bp_ms * 1e-3) / 1e9,
        "fp_updates_per_clk_per_sm": upd / (fp_ms * 1e-3) / (sm_mhz * 1e6) / 148,
        "bp_updates_per_clk_per_sm": upd / (bp_ms * 1e-3) / (sm_mhz * 1e6) / 148,
        "ceiling_updates_per_clk_per_sm": 8.9,
        "ceiling_note": "measured on B200 (scratch/ubench/gather_ceiling.cu, tex_arm.cu; profiles/r02_texture_vs_shared.md): "
                        "shared memory, 4 conflict-free LDS.32 taps per update and nothing else = 8.91 updates/clk/SM; with scalar "
                        "bilinear arithmetic 6.33; 3-row BP loop 9.52; texture unit with hardware bilinear (9-bit weights) 4.17, "
                        "with TLD4 + exact fp32 lerp 2.00",
        "fp_kernel": fp_name, "bp_kernel": bp_name,
    }
    t_dom = traffic.get(dom) if world == 1 and not big else None  # the ncu capture is of cfg 3 on one GPU
    cpu_base = None
    if world == 1 and not big:  # rank 0 at N = 1 only
        cpu_gups, cpu_dt, sample = cpu_sample()
        cpu_base = {"value": cpu_gups, "unit": UNIT, "cores": omp_threads(), "kind": "port", "sample": sample}
    line = {
        "metric": METRIC.replace("512^3 x 720", "1024^3 x 1440") if big else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_CFG4 if big else WORKLOAD_CFG3,
                   "phantom": "hollow_box", "l2": ("inputs (4.3 GB + 9.1 GB) larger than L2" if big else
                                                   "inputs (537 MB + 1132 MB) larger than L2"),
                   "parallelism": "single GPU" if world == 1 else
                   f"angle-sharded x{world}, z-sharded volume: NCCL all_gather -> FP; {bp_scheme}"},
        "fp_ms": fp_ms, "bp_ms": bp_ms,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": t_dom,
                     "traffic_source": (TRAFFIC_SOURCE if t_dom is not None else None), "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": b_alg, "interp": interp},
        "cpu_baseline": cpu_base,
        "e2e": e2e,
        "sirt": {"iters_per_s": 1e3 / sirt_ms, "ms_per_iter": sirt_ms, "iterations": sirt_iters,
                 "path": "tsp_sirt (fused epilogues)" if world == 1 else
                 sirt_scheme},
        "cfg4_sirt": cfg4,
        "sharded_parity_rel_l2": sharded_parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg3 = BASELINE configs[2] (the headline, default); cfg4 = configs[3], 1024^3 x 1440 (scaling study)")
    ap.add_argument("--skip-e2e", action="store_true", help="leave out the host-array leg (scaling study at cfg4)")
    ap.add_argument("--skip-cfg4", action="store_true", help="leave out the 1024^3 x 1440 SIRT sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
