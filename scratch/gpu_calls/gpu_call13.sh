set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/prof_step.py 512 720 3 2>&1 | tail -3
TSP_BP_NO_ROWS3=1 python scratch/prof_step.py 512 720 2 2>&1 | tail -2
python scratch/prof_step.py 512 720 2 par 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_tma' -c 3 -o gpurun_out/prof_v13 python scratch/prof_step.py 512 720 1 > gpurun_out/prof_v13.log 2>&1
