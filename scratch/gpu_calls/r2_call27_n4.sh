set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/r02_bench_n4.err | grep '^{' > gpurun_out/r02_bench_n4.json; tail -2 gpurun_out/r02_bench_n4.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n4.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','sharded_parity_rel_l2','gpu_launches'): print(k, d.get(k))
print(d['config']['parallelism'])"
