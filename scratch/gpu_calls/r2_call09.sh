set -x
for v in NO_FOLD_DZ NO_OPAQUE NO_BOTH; do TSPROJ_LIB=scratch/libtsproj_$v.so timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -2; done
timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -2
timeout 300 python scratch/bench_configs.py 2>&1 | grep -E "configs\[0\]|configs\[4\]" | cut -c1-300
TSP_NO_PINNED=1 timeout 300 python scratch/bench_configs.py 2>&1 | grep -E "configs\[0\]" | cut -c1-300
