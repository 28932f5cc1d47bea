set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -3 gpurun_out/r02_bench_ref.err; cut -c1-400 gpurun_out/r02_bench_ref.json
timeout 300 python scratch/bench_configs.py 2>/dev/null > gpurun_out/r02_bench_configs.jsonl; cut -c1-250 gpurun_out/r02_bench_configs.jsonl
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1
TSP_FP_SPS=1 timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | tail -2
timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cfg4 > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_tma' -c 3 -o gpurun_out/r02_prof_final python scratch/prof_step.py 512 720 1 > gpurun_out/r02_prof_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
