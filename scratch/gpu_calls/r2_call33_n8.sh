set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r02_bench_n8.err | grep '^{' > gpurun_out/r02_bench_n8.json; tail -2 gpurun_out/r02_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n8.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','sharded_parity_rel_l2','gpu_launches'): print(k, d.get(k))
print(d['config']['parallelism'])"
TSP_SHARD_NO_FP_PUSH=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 --skip-cfg4 --skip-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no_fp_push', d['value'], d['ms_per_step'], d['fp_ms'], d['bp_ms'], d['sirt']['ms_per_iter'])"
