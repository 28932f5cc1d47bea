set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','roofline','gpu_launches','clocks'): print(k, d.get(k))"
timeout 300 python scratch/bench_configs.py 2>/dev/null > gpurun_out/r02_bench_configs.jsonl; cut -c1-250 gpurun_out/r02_bench_configs.jsonl
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cfg4 > gpurun_out/r02_launches_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
