set -x
nvidia-smi -L
python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_v20.json 2> gpurun_out/bench_n2_v20.err; cat gpurun_out/bench_n2_v20.json; tail -5 gpurun_out/bench_n2_v20.err
TSP_SHARD_NO_PIPELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_v20_nopipe.json 2> gpurun_out/bench_n2_v20_nopipe.err; cat gpurun_out/bench_n2_v20_nopipe.json; tail -5 gpurun_out/bench_n2_v20_nopipe.err
