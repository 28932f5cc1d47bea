set -x
nvidia-smi -L
for c in 0 1 2 -3 5; do timeout 60 scratch/tma_test 0 $c 4; done
scratch/ubench/tex_arm | tee gpurun_out/r02_tex_arm.txt
scratch/ubench/gather_ceiling | tee gpurun_out/r02_gather_ceiling.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TSP_BP_ROWS=2 python scratch/prof_step.py 512 720 3 2>&1 | tail -4
python scratch/prof_step.py 512 720 3 2>&1 | tail -4
TSP_BP_ZPT=16 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
python scratch/prof_step.py 512 720 2 par 2>&1 | tail -3
ncu --set full --clock-control none -k regex:'^k' -c 9 -o gpurun_out/r02_prof_tex scratch/ubench/tex_arm > gpurun_out/r02_prof_tex.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bp_tma' -c 1 -o gpurun_out/r02_prof_bp_rows python scratch/prof_step.py 512 720 1 > gpurun_out/r02_prof_bp_rows.log 2>&1
