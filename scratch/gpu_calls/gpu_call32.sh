python -m pytest tests/test_operator_gpu.py -m gpu -x -q -k "1024 or full_size" 2>&1 | tail -4
python scratch/bench_configs.py 2>&1 | grep "published"
