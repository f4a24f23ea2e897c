"""Dev helper (library built with `make -C seesaw_b200/csrc EXTRA=-DSSW_TRACE=2`): does the exchange of step i really run
under the scan of step i+1, how long does it take, how long does it wait for its peers?  %globaltimer stamps of the last
scan launch (every CTA) and of the last two exchange launches (first 8 blocks), printed relative to the first CTA
entry of the last scan.  One GPU (world = 1) or under torchrun (one rank per GPU, --images per rank)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import _lib, synth  # noqa: E402
from seesaw_b200.sharded import ShardedPatchDatabase  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=31250, help="images per rank (40 rows each)")
ap.add_argument("--side", type=int, nargs="+", default=[4, 2, 0])
ap.add_argument("--trials", type=int, default=3)
ap.add_argument("--steps", type=int, default=100)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NQ, K = 64, 50
n_img = args.images * world
sdb = ShardedPatchDatabase.synthetic(np.full(n_img, 40, np.int64), 512, seed=4, rank=rank, world_size=world, device=local)
db = sdb.local
d_q = torch.from_numpy(synth.unit_queries(NQ, 512, 1)).cuda()
rng = np.random.default_rng(2)
bits = db.build_exclude_bits([np.sort(rng.choice(n_img, size=50, replace=False)).astype(np.int32) for _ in range(NQ)], NQ)
sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
db.scan_stats(True)
f = _lib.lib.ssw_scan_trace_read
f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int]
for side in args.side:
    sdb.set_side_sms(side)
    for trial in range(args.trials):
        for _ in range(10):
            sdb.scan_topk_device(d_q, K, d_exclude_bits=bits, pipelined=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sdb.scan_topk_device(d_q, K, d_exclude_bits=bits, pipelined=True)
        e1.record()
        sdb.drain()
        torch.cuda.synchronize()
        step = e0.elapsed_time(e1) / args.steps * 1e3
        out = np.zeros((164, 16), np.int64)
        assert f(db._h, out.ctypes.data_as(C.c_void_p), 164) == 0
        grid = 148 - side
        sc = out[:grid].astype(np.float64)
        t0 = sc[:, 0].min()
        last = sdb._xchg["epoch"]
        nb = side if side > 0 else 8

        def xch(epoch):
            p = epoch & 1
            se = (out[160 + p].astype(np.float64).reshape(8, 2)[:nb] - t0) / 1e3
            pw = out[162 + p].astype(np.float64).reshape(8, 2)[:nb]
            return se, (pw[:, 0] - t0) / 1e3, pw[:, 1] / 1e3

        (xn, pn, wn), (xp, pp, wp) = xch(last), xch(last - 1)
        us = lambda col: (sc[:, col] - t0) / 1e3
        print(f"rank {rank}/{world} side={side} trial={trial} step {step:6.1f} us | last scan: entry 0..{us(0).max():5.1f}  first tile ready "
              f"{np.median(us(9)):5.1f}  exit {np.median(us(13)):6.1f} (min {us(13).min():6.1f} max {us(13).max():6.1f}) | exchange before it: start "
              f"{xp[:, 0].min():6.1f}..{xp[:, 0].max():6.1f} pushed {pp.max():6.1f} end {xp[:, 1].max():6.1f} waited {wp.max():5.1f} | its own exchange: "
              f"start {xn[:, 0].min():6.1f} pushed {pn.max():6.1f} end {xn[:, 1].max():6.1f} waited {wn.max():5.1f}", flush=True)
if world > 1:
    dist.barrier()
sdb.close()
if world > 1:
    dist.destroy_process_group()
