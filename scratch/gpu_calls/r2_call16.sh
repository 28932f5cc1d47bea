set -x
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q 2>&1 | tail -4
for e in "" "TSP_SHARD_NO_PRETRANSPOSE=1"; do
env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cfg4 --skip-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e', d['value'], d['fp_ms'], d['bp_ms'], d['sirt']['ms_per_iter'], d['gpu_launches'], d['sharded_parity_rel_l2'])"
done
