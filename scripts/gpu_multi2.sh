#!/bin/bash
# 2-GPU box, trace build: timeline of the pipelined step with a real peer (1.25M rows per rank = the 8-GPU shard size)
mkdir -p gpurun_out
cp seesaw_b200/libseesaw_b200_trace.so seesaw_b200/libseesaw_b200.so
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/trace_pipe.py --side 4 0 --trials 2 > gpurun_out/r02_trace_world2.log 2>&1
grep "^rank" gpurun_out/r02_trace_world2.log || tail -20 gpurun_out/r02_trace_world2.log
