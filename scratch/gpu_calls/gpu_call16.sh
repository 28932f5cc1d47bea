set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err; cat gpurun_out/bench_v16.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['fp_ms'], d['bp_ms'])"; tail -3 gpurun_out/bench_v16.err
TSP_HOST_NO_PIPELINE=1 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no pipeline', d['value'], d['e2e'])"
