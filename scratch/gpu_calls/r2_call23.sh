set -x
timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1
TSP_HOST_NO_GRADED=1 timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_gpu_variants.py -m gpu -q -x -k "configs2 or host" 2>&1 | tail -3
