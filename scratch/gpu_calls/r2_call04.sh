set -x
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -25
for v in NOSETUP ST8; do TSPROJ_LIB=scratch/libtsproj_$v.so timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3; done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
