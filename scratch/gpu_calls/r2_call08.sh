set -x
TSP_DEBUG=1 timeout 120 python scratch/prof_step.py 512 720 3 2>&1 | grep -v "host\|plan" | tail -9
TSP_FP_NO_CLASSES=1 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
TSP_FP_CLASSES=2 timeout 120 python scratch/prof_step.py 512 720 2 2>&1 | tail -3
timeout 120 python scratch/prof_step.py 1024 1440 2 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_kernels.py tests/test_operator_gpu.py tests/test_fdk.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cfg4 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b.json')); print(d['value'], d['fp_ms'], d['bp_ms'], d['e2e'], d['sirt'])"
timeout 300 python scratch/bench_configs.py 2>&1 | tail -12
