python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py tests/test_operator_gpu.py -m gpu -x -q -k "not 1024 and not full_size" 2>&1 | tail -3
python scratch/bench_cfg5.py 2>&1 | tail -2
TSP_THIN_NO_STAGE=1 python scratch/bench_cfg5.py 2>&1 | tail -2
