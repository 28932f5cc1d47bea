set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/prof_step.py 512 720 3 2>&1 | tail -3
TSP_BP_ZPT=32 python scratch/prof_step.py 512 720 3 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:'bp_tma' -c 1 -o gpurun_out/prof_bp_v5 python scratch/prof_step.py 512 720 1 > gpurun_out/prof_bp_v5.log 2>&1
