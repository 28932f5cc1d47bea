set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_v10.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_v10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_tma' -c 3 -o gpurun_out/prof_v10 python scratch/prof_step.py 512 720 1 > gpurun_out/prof_v10.log 2>&1
