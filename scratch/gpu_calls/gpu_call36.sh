TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29581 bench.py --gpus 8 --workload cfg4 --skip-e2e --steps 3 --warmup 3 2> gpurun_out/bench_n8_cfg4_final.err | grep '^{' > gpurun_out/bench_n8_cfg4_final.json; cat gpurun_out/bench_n8_cfg4_final.json | cut -c1-200; tail -3 gpurun_out/bench_n8_cfg4_final.err
