set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_n8_cfg3_k4.err | grep '^{' > gpurun_out/bench_n8_cfg3_k4.json; cat gpurun_out/bench_n8_cfg3_k4.json; tail -3 gpurun_out/bench_n8_cfg3_k4.err
$TR --master-port 29543 bench.py --gpus 8 --workload cfg4 --skip-e2e --steps 3 --warmup 3 2> gpurun_out/bench_n8_cfg4_k4.err | grep '^{' > gpurun_out/bench_n8_cfg4_k4.json; cat gpurun_out/bench_n8_cfg4_k4.json; tail -3 gpurun_out/bench_n8_cfg4_k4.err
