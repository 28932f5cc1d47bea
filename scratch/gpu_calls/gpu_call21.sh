set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v21.json 2> gpurun_out/bench_v21.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v21.json')); print(d['value'], d['e2e'], d['fp_ms'], d['bp_ms'], d['sirt'])"; tail -3 gpurun_out/bench_v21.err
python scratch/bench_cfg5.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg5.csv python scratch/bench_cfg5.py > gpurun_out/cfg5_ncu.log 2>&1
python bench.py --workload cfg4 --skip-e2e --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; cat gpurun_out/bench_cfg4_n1.json; tail -3 gpurun_out/bench_cfg4_n1.err
