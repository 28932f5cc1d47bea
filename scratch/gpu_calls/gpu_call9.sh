set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; cat gpurun_out/bench_v9.json; tail -3 gpurun_out/bench_v9.err
python scratch/bench_cfg5.py 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
