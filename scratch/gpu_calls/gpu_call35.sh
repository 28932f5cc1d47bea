python scratch/bench_e2e.py 2>&1 | tail -1
TSP_HOST_UNIFORM=1 python scratch/bench_e2e.py 2>&1 | tail -1
TSP_DEBUG=1 python scratch/bench_e2e.py 2>&1 | grep "host" | head -24
python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "host" 2>&1 | tail -2
