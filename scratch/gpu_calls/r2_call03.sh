set -x
for v in NOSETUP ST8; do TSPROJ_LIB=scratch/libtsproj_$v.so timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3; done
TSP_BP_ZPT=24 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -3
