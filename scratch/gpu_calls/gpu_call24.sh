set -x
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_cfg3.json 2> gpurun_out/bench_n8_cfg3.err; cat gpurun_out/bench_n8_cfg3.json; tail -3 gpurun_out/bench_n8_cfg3.err
TSP_SHARD_NO_PIPELINE=1 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_cfg3_nopipe.json 2> gpurun_out/bench_n8_cfg3_nopipe.err; cat gpurun_out/bench_n8_cfg3_nopipe.json; tail -3 gpurun_out/bench_n8_cfg3_nopipe.err
$TR --master-port 29523 bench.py --gpus 8 --workload cfg4 --skip-e2e --steps 3 --warmup 3 > gpurun_out/bench_n8_cfg4.json 2> gpurun_out/bench_n8_cfg4.err; cat gpurun_out/bench_n8_cfg4.json; tail -3 gpurun_out/bench_n8_cfg4.err
TSP_SHARD_NO_PIPELINE=1 $TR --master-port 29524 bench.py --gpus 8 --workload cfg4 --skip-e2e --steps 3 --warmup 3 > gpurun_out/bench_n8_cfg4_nopipe.json 2> gpurun_out/bench_n8_cfg4_nopipe.err; cat gpurun_out/bench_n8_cfg4_nopipe.json; tail -3 gpurun_out/bench_n8_cfg4_nopipe.err
