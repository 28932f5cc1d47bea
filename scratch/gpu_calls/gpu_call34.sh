for k in 4 8 12 16; do TSP_HOST_CHUNKS=$k python scratch/bench_e2e.py 2>&1 | tail -1; done
TSP_HOST_BP_INORDER=1 python scratch/bench_e2e.py 2>&1 | tail -1
python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "host" 2>&1 | tail -2
