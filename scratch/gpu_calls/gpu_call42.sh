ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_tma' -c 3 -o gpurun_out/prof_final python scratch/prof_step.py 512 720 1 > gpurun_out/prof_final.log 2>&1
ls -la gpurun_out/
