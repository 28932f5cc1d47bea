set -x
for nb in "" 1; do
TSP_BENCH_NO_BIND=$nb timeout 300 python bench.py --steps 5 --warmup 3 --skip-cfg4 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no_bind=$nb', d['value'], d['e2e']['value'], d['e2e'].get('host_cpu_binding'))"
done
timeout 120 python scratch/bench_e2e.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cfg4 > gpurun_out/r02_launches_bench.log 2>&1
