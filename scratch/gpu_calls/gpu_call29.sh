python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scratch/sirt_breakdown.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -25
