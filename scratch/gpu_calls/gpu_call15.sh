set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err; cat gpurun_out/bench_v15.json; tail -3 gpurun_out/bench_v15.err
