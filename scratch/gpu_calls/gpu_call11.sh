set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/bench_cfg5.py 2>&1 | tail -1
TSP_NO_THIN=1 python scratch/bench_cfg5.py 2>&1 | tail -1
