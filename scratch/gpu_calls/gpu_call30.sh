TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29571 scratch/sirt_breakdown.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -20
$TR --master-port 29572 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_n8_cfg3_k4p.err | grep '^{' > gpurun_out/bench_n8_cfg3_k4p.json; cat gpurun_out/bench_n8_cfg3_k4p.json | cut -c1-300; tail -3 gpurun_out/bench_n8_cfg3_k4p.err
