set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k supersampling 2>&1 | tail -5
timeout 120 python scratch/ss_bench.py 2>&1 | tail -2
TSP_SS_DIRECT=1 timeout 300 python scratch/ss_bench.py 2>&1 | tail -2
timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
