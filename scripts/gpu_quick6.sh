#!/bin/bash
timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -3
SUST=K2,K1 python scripts/sustained.py 2>&1 | tail -20
timeout 200 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --images 31250 --iters 20 2>&1 | tail -2 | head -1
timeout 200 python scripts/quick_scan.py --nq 1 --k 10 --mode 1 --excl 50 --images 10000000 --patches 1 --iters 10 2>&1 | tail -2 | head -1
