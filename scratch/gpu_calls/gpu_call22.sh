set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scratch/bench_cfg5.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg5_v22.csv python scratch/bench_cfg5.py > gpurun_out/cfg5_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:thin -s 8 -c 4 -o gpurun_out/prof_thin_v22 python scratch/bench_cfg5.py > gpurun_out/prof_thin_v22.log 2>&1
ls -la gpurun_out/
