python scratch/bench_configs.py 2>&1 | tail -5
