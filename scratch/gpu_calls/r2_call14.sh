set -x
for e in "" "TSP_SHARD_NO_FP_BLOCKS=1" "TSP_HOST_CHUNKS=4" "TSP_SHARD_CHUNKS=8"; do
env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cfg4 --skip-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e', d['value'], d['fp_ms'], d['bp_ms'], d['sirt']['ms_per_iter'], d['gpu_launches'])"
done
