set -x
timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_full_size_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cfg4 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['fp_ms'], d['bp_ms'], d['e2e']['value'], d['sirt']['ms_per_iter'])"
