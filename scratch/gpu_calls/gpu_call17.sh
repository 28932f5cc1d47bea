set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TSP_DEBUG=1 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v17.json 2> gpurun_out/bench_v17.err; cat gpurun_out/bench_v17.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['fp_ms'], d['bp_ms'])"; grep "host" gpurun_out/bench_v17.err | head -20
