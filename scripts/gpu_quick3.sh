#!/bin/bash
for img in 31250 62500 125000; do
for k in 1 50; do
timeout 200 python scripts/quick_scan.py --nq 64 --k $k --mode 2 --excl 50 --images $img --iters 20 2>&1 | tail -2 | head -1
done; done
timeout 200 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --images 31250 --iters 20 2>&1 | tail -2 | head -1
