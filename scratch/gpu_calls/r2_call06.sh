set -x
timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | tail -4
TSP_DEBUG=1 TSP_FP_SPS=2 timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | grep -v host | tail -6
TSP_FP_SPS=2 TSP_FP_R=4 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
TSP_FP_R=4 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py tests/test_operator_gpu.py -m gpu -q -x 2>&1 | tail -4
TSP_FP_SPS=2 timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:'bp_tma|fp_tma' -c 3 -o gpurun_out/r02_prof_v2 python scratch/prof_step.py 512 720 1 > gpurun_out/r02_prof_v2.log 2>&1
