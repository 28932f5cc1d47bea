set -x
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err; cat gpurun_out/bench_v19.json; tail -3 gpurun_out/bench_v19.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v19.json 2>&1; cat gpurun_out/bench_ref_v19.json
python scratch/prof_step.py 1024 1440 2 2>&1 | tail -3
python scratch/bench_cfg5.py 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v19.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b19.log 2>&1
