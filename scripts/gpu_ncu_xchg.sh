#!/bin/bash
# ncu --set full of the side-SM exchange kernel (world = 1, 1.25M-row shard, 64 queries, k = 50)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:exchange_slim_kernel -s 5 -c 1 -o gpurun_out/r02_xchg_slim_side_full -f python scripts/step_breakdown.py --images 31250 --iters 3 --pipeline 4 > gpurun_out/r02_ncu_xchg_slim.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/r02_ncu_xchg_slim.log
ls -la gpurun_out/*.ncu-rep
