set -x
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
TSPROJ_LIB=scratch/libtsproj_THINNP.so timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -k "thin or batch" 2>&1 | tail -5
