set -x
TSP_DEBUG=1 timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | grep -v "host\|plan" | tail -10
TSP_FP_NO_CLASSES=1 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
TSP_FP_CLASSES=2 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 512 720 2 par 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 256 180 2 par 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err; cut -c1-1500 gpurun_out/r02_bench_a.json
