set -x
timeout 120 python scratch/diag_host.py 2>&1 | head -8
timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | tail -3
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
TSPROJ_LIB=scratch/libtsproj_THINNP.so timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
timeout 300 python scratch/bench_configs.py 2>&1 | grep -E "configs\[0\]|configs\[4\]" | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
