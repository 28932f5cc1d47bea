set -x
timeout 400 python -m pytest tests/test_distributed_gpu.py tests/test_gpu_variants.py -m gpu -q -k "nccl or several_gpus" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r02_bench_n2.err | grep '^{' > gpurun_out/r02_bench_n2.json; tail -2 gpurun_out/r02_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','sharded_parity_rel_l2','gpu_launches'): print(k, d.get(k))
print(d['config']['parallelism'])"
