set -x
for l in scratch/lib_u1.so tomosipo_b200/libtsproj.so scratch/lib_u4.so; do TSPROJ_LIB=$PWD/$l python scratch/bench_cfg5.py 2>&1 | tail -2; done
python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
