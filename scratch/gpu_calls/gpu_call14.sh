set -x
nvidia-smi -L
python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_v14.json 2> gpurun_out/bench_n2_v14.err; cat gpurun_out/bench_n2_v14.json; tail -3 gpurun_out/bench_n2_v14.err
