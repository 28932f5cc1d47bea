set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/prof_step.py 512 720 3 2>&1 | tail -3
python scratch/prof_step.py 512 720 2 par 2>&1 | tail -2
TSP_BP_ZPT=16 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
