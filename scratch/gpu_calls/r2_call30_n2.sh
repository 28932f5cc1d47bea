set -x
timeout 300 python -m pytest tests/test_distributed_gpu.py -m gpu -q 2>&1 | tail -15
for np2p in "" 1; do
TSP_SHARD_NO_P2P=$np2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cfg4 --skip-e2e 2> gpurun_out/r02_n2_p2p$np2p.err | grep '^{' > gpurun_out/r02_n2_p2p$np2p.json; tail -3 gpurun_out/r02_n2_p2p$np2p.err
python -c "
import json; d=json.load(open('gpurun_out/r02_n2_p2p$np2p.json'))
print('no_p2p=$np2p', d['value'], d['ms_per_step'], d['fp_ms'], d['bp_ms'], d['sirt']['ms_per_iter'], d['sharded_parity_rel_l2'], d['gpu_launches'])"
done
