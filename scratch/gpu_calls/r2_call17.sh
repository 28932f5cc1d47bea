set -x
TSP_BP_ROWS=4 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
TSP_BP_ROWS=4 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_variants.py -m gpu -q -x -k "bp or BP or matches" 2>&1 | tail -3
for c in 4 6 8 12; do TSP_HOST_CHUNKS=$c timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1; done
