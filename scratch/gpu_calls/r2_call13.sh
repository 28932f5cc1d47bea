set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_distributed_gpu.py tests/test_gpu_variants.py -m gpu -q -k "nccl or several_gpus or interior" 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','sharded_parity_rel_l2','gpu_launches'): print(k, d.get(k))
print(d['roofline']['kernel'], d['roofline']['interp']['bp_kernel'])"
TSP_SHARD_NO_FP_BLOCKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cfg4 --skip-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no fp blocks:', d['value'], d['fp_ms'], d['bp_ms'])"
