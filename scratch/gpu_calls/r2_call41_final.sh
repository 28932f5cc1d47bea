set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json'))
for k in ('value','ms_per_step','fp_ms','bp_ms','e2e','sirt','cfg4_sirt','gpu_launches','clocks'): print(k, d.get(k))"
timeout 300 python scratch/bench_configs.py 2>/dev/null > gpurun_out/r02_bench_configs.jsonl; cut -c1-200 gpurun_out/r02_bench_configs.jsonl
timeout 300 ncu --clock-control none -k regex:fp_tma -c 6 --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --csv --log-file gpurun_out/r02_cfg4_fp_seg_metrics.csv python scratch/prof_step.py 1024 1440 1 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
