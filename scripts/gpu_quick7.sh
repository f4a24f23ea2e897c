#!/bin/bash
echo "== config-5 shard: 12.5M x 768 fp16 (19.2 GB) on one GPU"
timeout 300 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --dim 768 --images 312500 --iters 10 2>&1 | tail -2
timeout 300 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --dim 768 --images 312500 --iters 10 2>&1 | tail -2 | head -1
echo "== fixed overhead probe: tiny databases (one or two tiles per CTA)"
for img in 500 4000 16000; do
timeout 100 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --images $img --iters 30 2>&1 | tail -2 | head -1
timeout 100 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --images $img --iters 30 2>&1 | tail -2 | head -1
done
