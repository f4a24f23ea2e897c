#!/bin/bash
# Full single-GPU round: parity tests, smoke, bench; A/B of the sharded step's shapes on an 8-GPU-sized shard
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -E "float32 kNN|plain fp16|stage-1 scans of|exact mode|fp16 storage alone|float32 storage" gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
if [ "$1" = "ref" ]; then
  timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  cut -c1-400 gpurun_out/bench_ref.json
fi
timeout 120 python scripts/ab_pipeline.py --shapes inline 0 4 --rounds 4 > gpurun_out/r02_ab_pipeline_shapes.log 2>&1; cat gpurun_out/r02_ab_pipeline_shapes.log
