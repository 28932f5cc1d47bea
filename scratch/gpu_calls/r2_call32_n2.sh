set -x
timeout 400 python -m pytest tests/test_distributed_gpu.py -m gpu -q 2>&1 | tail -15
for nf in "" 1; do
TSP_SHARD_NO_FP_PUSH=$nf timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cfg4 2> gpurun_out/r02_n2_fpp$nf.err | grep '^{' > gpurun_out/r02_n2_fpp$nf.json; tail -3 gpurun_out/r02_n2_fpp$nf.err
python -c "
import json; d=json.load(open('gpurun_out/r02_n2_fpp$nf.json'))
print('no_fp_push=$nf', d['value'], d['ms_per_step'], d['fp_ms'], d['bp_ms'], d['sirt']['ms_per_iter'], d['e2e']['value'], d['sharded_parity_rel_l2'], d['gpu_launches'])"
done
