#!/bin/bash
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
for img in 31250 250000; do
for k in 1 50; do
timeout 200 python scripts/quick_scan.py --nq 64 --k $k --mode 2 --excl 50 --images $img --iters 20 2>&1 | tail -2 | head -1
done; done
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 3000 --images 250000 --iters 20 2>&1 | tail -2 | head -1
