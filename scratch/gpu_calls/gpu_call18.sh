for k in 2 4 6 8 12 16; do TSP_HOST_CHUNKS=$k python scratch/bench_e2e.py 2>&1 | tail -1; done
TSP_HOST_NO_PIPELINE=1 python scratch/bench_e2e.py 2>&1 | tail -1
