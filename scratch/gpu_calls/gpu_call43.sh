python -m pytest tests/test_gpu_kernels.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -3
python scratch/prof_step.py 512 720 3 2>&1 | tail -3 | head -2
TSP_NO_EARLY_POLL=1 python scratch/prof_step.py 512 720 3 2>&1 | tail -3 | head -2
python scratch/prof_step.py 256 384 2 par 2>&1 | tail -2 | head -1
