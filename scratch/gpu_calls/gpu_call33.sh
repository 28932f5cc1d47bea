python scratch/prof_step.py 256 384 3 par 2>&1 | tail -2
python scratch/prof_step.py 256 384 3 cone 2>&1 | tail -2
python scratch/prof_step.py 512 720 3 par 2>&1 | tail -2
TSP_DEBUG=1 python scratch/prof_step.py 256 384 1 par 2>&1 | grep "tsp" | head
