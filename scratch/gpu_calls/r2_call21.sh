set -x
TSPROJ_LIB=scratch/libtsproj_NREG.so timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | tail -3
TSPROJ_LIB=scratch/libtsproj_NREG.so timeout 120 python scratch/prof_step.py 512 720 2 par 2>&1 | tail -2
timeout 120 python scratch/prof_step.py 512 720 2 par 2>&1 | tail -2
