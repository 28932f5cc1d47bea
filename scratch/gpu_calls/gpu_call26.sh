set -x
TSPROJ_LIB=$PWD/scratch/lib_u8.so python scratch/bench_cfg5.py 2>&1 | tail -2
python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2_v26.err | grep '^{' > gpurun_out/bench_n2_v26.json; cat gpurun_out/bench_n2_v26.json; tail -3 gpurun_out/bench_n2_v26.err
TSP_SHARD_CHUNKS=1 $TR --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2_v26_k1.err | grep '^{' > gpurun_out/bench_n2_v26_k1.json; cat gpurun_out/bench_n2_v26_k1.json; tail -3 gpurun_out/bench_n2_v26_k1.err
