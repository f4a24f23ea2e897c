#!/bin/bash
# trace build (clock64): what does a K2 CTA wait for at boost and at power-capped clocks?
mkdir -p gpurun_out
cp seesaw_b200/libseesaw_b200_trace.so seesaw_b200/libseesaw_b200.so
timeout 150 python scripts/trace_k2.py 250000 2.0 > gpurun_out/r02_k2_waits.log 2>&1; cat gpurun_out/r02_k2_waits.log
