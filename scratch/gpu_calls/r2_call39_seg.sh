set -x
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "fp_variant or segmented" 2>&1 | tail -4
for sgm in 1 2 3 4 6; do
TSP_FP_SEGMENTS=$sgm timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | grep "fp " | tail -1
done
for sgm in 1 2; do
TSP_FP_SEGMENTS=$sgm timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | grep "fp " | tail -1
done
