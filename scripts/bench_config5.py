#!/usr/bin/env python
"""BASELINE.json configs[4]: ViT-L/14 scale — 100M x 768 fp16 multiscale patches (153.6 GB) row-sharded over 8
B200 (12.5M rows = 19.2 GB per GPU), single-query latency and 64-query batched throughput.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_config5.py

Every rank generates its shard in HBM (counter-based generator, shard-indexed); timing = CUDA events around the
whole step (scan + fused peer exchange + merge), max over ranks.  Prints one JSON line on rank 0."""
import json
import os
import sys

# one JSON line on the real stdout; whatever libraries print (NCCL's version banner) goes to stderr
os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.stdout.flush()
_real_stdout = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import synth  # noqa: E402
from seesaw_b200.sharded import ShardedPatchDatabase  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
rows_total = int(os.environ.get("SSW_C5_ROWS", 100_000_000)) // 8 * world       # 12.5M rows per GPU
DIM, PATCHES, NQ, K = 768, 40, 64, 50
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
counts = np.full(rows_total // PATCHES, PATCHES, np.int64)
sdb = ShardedPatchDatabase.synthetic(counts, DIM, seed=7, rank=rank, world_size=world, device=local)
sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
q = torch.from_numpy(synth.unit_queries(NQ, DIM, 1)).to(dev)
rng = np.random.default_rng(2)
ex = [rng.choice(len(counts), size=50, replace=False) for _ in range(NQ)]
bits = sdb.local.build_exclude_bits(ex, NQ)


def timed(fn, steps, warm=5):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


ms_batch = timed(lambda: sdb.scan_topk_device(q, K, d_exclude_bits=bits), 50)
q1, b1 = q[:1].contiguous(), bits[:1]
ms_single = timed(lambda: sdb.scan_topk_device(q1, K, d_exclude_bits=b1), 50)
res = sdb.scan_topk_device(q, K, d_exclude_bits=bits)
torch.cuda.synchronize()
if rank == 0:
    gb = rows_total * DIM * 2 / 1e9
    _real_stdout.write(json.dumps({"config": f"{rows_total} x {DIM} fp16 ({gb:.1f} GB) over {world} GPU(s), {rows_total // world} rows per GPU",
                      "batched_64": {"ms_per_batch": ms_batch, "queries_per_s": NQ / ms_batch * 1e3, "aggregate_hbm_gbs": gb / ms_batch * 1e3},
                      "single_query": {"ms": ms_single, "aggregate_hbm_gbs": gb / ms_single * 1e3},
                      "top1_dbidx_q0": int(res["dbidx"][0, 0]), "count_q0": int(res["count"][0])}) + "\n")
    _real_stdout.flush()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
