set -x
for v in NOWAIT NOSETUP ST8; do TSPROJ_LIB=scratch/libtsproj_$v.so python scratch/prof_step.py 512 720 2 2>&1 | tail -3; done
TSP_BP_ZPT=24 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
python scratch/prof_step.py 512 720 2 2>&1 | tail -3
