set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
TSP_DEBUG=1 python scratch/prof_step.py 512 720 3 2>&1 | tail -4
