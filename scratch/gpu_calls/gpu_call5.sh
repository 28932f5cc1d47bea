set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scratch/prof_step.py 512 720 3 2>&1 | tail -3
TSP_FP_STAGES=4 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
TSP_FP_NO_TMA=1 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
python scratch/prof_step.py 512 720 2 par 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:'fp_tma' -c 1 -o gpurun_out/prof_fp_v5 python scratch/prof_step.py 512 720 1 > gpurun_out/prof_fp_v5.log 2>&1
