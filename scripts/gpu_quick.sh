#!/bin/bash
# quick check: scan tests + K2/K1 timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 0 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 1 --mode 2 --excl 0 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 10 --mode 2 --excl 500 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --dim 768 --images 160000 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 2>&1 | tail -2
