set -x
python scratch/prof_step.py 512 720 3 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_v3.csv python scratch/prof_step.py 512 720 2 > gpurun_out/launches_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_cols' -c 3 -o gpurun_out/prof_v3 python scratch/prof_step.py 512 720 1 > gpurun_out/prof_v3.log 2>&1
ls -la gpurun_out
