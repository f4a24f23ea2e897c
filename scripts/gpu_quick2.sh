#!/bin/bash
timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --dim 768 --images 160000 2>&1 | tail -2
SSW_TC768_NT64=1 timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --dim 768 --images 160000 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --dim 256 --images 500000 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --dim 768 --images 160000 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 8 --k 50 --mode 2 --excl 50 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --images 120000 --patches 41 2>&1 | tail -2
