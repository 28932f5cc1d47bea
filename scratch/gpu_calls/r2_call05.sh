set -x
TSP_DEBUG=1 timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | grep -v "host" | tail -8
TSP_FP_SPS=1 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
TSPROJ_LIB=scratch/libtsproj_PUB.so timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 512 720 2 par 2>&1 | tail -3
timeout 120 python scratch/bench_cfg5.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py tests/test_operator_gpu.py -m gpu -q -x 2>&1 | tail -8
TSPROJ_LIB=scratch/libtsproj_PUB.so timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
