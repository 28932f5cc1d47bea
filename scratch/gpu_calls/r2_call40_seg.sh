set -x
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "fp_variant or segmented" 2>&1 | tail -4
timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | grep "fp " | tail -1
timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | grep "fp " | tail -1
