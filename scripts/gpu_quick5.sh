#!/bin/bash
mkdir -p gpurun_out
for k in 1 50; do
timeout 200 python scripts/quick_scan.py --nq 64 --k $k --mode 2 --excl 50 --images 31250 --iters 20 2>&1 | tail -2 | head -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/l_small_k$k.csv python scripts/quick_scan.py --nq 64 --k $k --mode 2 --excl 50 --images 31250 --iters 5 > /dev/null 2>&1
done
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --images 250000 --iters 20 2>&1 | tail -2 | head -1
