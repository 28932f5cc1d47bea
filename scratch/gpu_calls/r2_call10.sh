set -x
timeout 120 python scratch/diag_host.py 2>&1 | tail -14
TSP_NO_PINNED=1 timeout 120 python scratch/diag_host.py 2>&1 | tail -5
TSP_POOL_KEEP_MB=4096 timeout 120 python scratch/diag_host.py 2>&1 | head -5
