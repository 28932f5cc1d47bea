python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scratch/sirt_breakdown.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -25
