#!/bin/bash
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
for img in 500 31250; do
timeout 100 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --images $img --iters 30 2>&1 | tail -2 | head -1
timeout 100 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --images $img --iters 30 2>&1 | tail -2 | head -1
done
