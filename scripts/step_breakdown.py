"""Dev helper: where does a batched-scan step spend its time on a SMALL shard (the 8-GPU case: 1.25M rows)?
Prints, per shard size: whole step (CUDA events), the scan kernel alone (library events), the rest; the same
through the fused-exchange path with world = 1; and the host-buffer call (wall clock)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import synth  # noqa: E402
from seesaw_b200.sharded import ShardedPatchDatabase  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, nargs="+", default=[31250, 62500, 250000])
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--pipeline", type=int, nargs="*", default=None,
                help="only the device API and the PIPELINED sharded step (world = 1) with these numbers of side SMs")
args = ap.parse_args()
NQ, K, PATCHES = 64, 50, 40
peak = 6516.7

for n_img in args.images:
    sdb = ShardedPatchDatabase.synthetic(np.full(n_img, PATCHES, np.int64), args.dim, seed=4, rank=0, world_size=1, device=0)
    db = sdb.local
    q_host = synth.unit_queries(NQ, args.dim, 1)
    rng = np.random.default_rng(2)
    ex = [np.sort(rng.choice(n_img, size=50, replace=False)).astype(np.int32) for _ in range(NQ)]
    d_q = torch.from_numpy(q_host).cuda()
    bits = db.build_exclude_bits(ex, NQ)
    ideal_us = db.n_rows * args.dim * 2 / (peak * 1e9) * 1e6

    def run(fn, label):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()

        def loop(profile):
            if profile:
                db.profile(True)
                db.profile_read()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / args.iters * 1e6
            step = e0.elapsed_time(e1) / args.iters * 1e3
            kern = 0.0
            if profile:
                kms, kn = db.profile_read()
                db.profile(False)
                kern = kms / max(kn, 1) * 1e3
            return step, wall, kern

        step, wall, _ = loop(False)            # programmatic dependent launches are off while profiling
        _, _, kern = loop(True)
        print(f"rows={db.n_rows:>9} {label:<28} step {step:7.1f} us  scan kernel {kern:7.1f} us  rest {step - kern:6.1f} us  "
              f"wall {wall:7.1f} us  HBM-ideal {ideal_us:6.1f} us  ideal/step {ideal_us / step:.3f}", flush=True)

    run(lambda: db.scan_topk_device(d_q, K, bits, decoded=True), "device API (merge kernel)")
    db.scan_stats(True)
    for _ in range(10):
        db.scan_topk_device(d_q, K, bits, decoded=True)
    upd, off = db.scan_stats(False)
    print(f"rows={db.n_rows:>9} per query and CTA and launch: {upd / 10 / NQ / 148:.1f} list updates, {off / 10 / NQ / 148:.1f} images offered", flush=True)
    sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
    run(lambda: sdb.scan_topk_device(d_q, K, d_exclude_bits=bits), "fused exchange, world=1")
    if args.pipeline is not None:
        for side in args.pipeline:
            sdb.set_side_sms(side)
            run(lambda: sdb.scan_topk_device(d_q, K, d_exclude_bits=bits, pipelined=True), f"pipelined, side SMs = {side}")
            sdb.drain()
            torch.cuda.synchronize()
        sdb.close()
        torch.cuda.empty_cache()
        continue
    run(lambda: sdb.scan_topk(q_host, K, exclude=ex), "host API (sharded, world=1)")
    run(lambda: db.scan_topk(q_host, K, exclude=ex), "host API ssw_scan_topk")
    from seesaw_b200.engine import exclude_lists_to_csr
    ids, offs = exclude_lists_to_csr(ex, NQ)
    run(lambda: db.scan_topk_csr(q_host, K, ids, offs), "host API, CSR excludes")
    sdb.close()
    torch.cuda.empty_cache()
