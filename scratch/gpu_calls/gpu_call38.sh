set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v38.json 2> gpurun_out/bench_v38.err; cat gpurun_out/bench_v38.json; tail -3 gpurun_out/bench_v38.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v38.json 2>&1; cut -c1-300 gpurun_out/bench_ref_v38.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v38.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b38.log 2>&1
