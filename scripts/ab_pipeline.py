"""Dev helper: A/B of the sharded step's shapes on ONE GPU (world = 1) under the same clock conditions — the boards
power-cap their SM clock under sustained load and the batched scan's epilogue follows the clock, so shapes timed one
after the other are not comparable.  Interleaves the shapes round after round after a warm-up and samples the SM
clock while each measurement runs."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import synth  # noqa: E402
from seesaw_b200.sharded import ShardedPatchDatabase  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=31250)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--shapes", nargs="+", default=["inline", "0", "4", "2", "8"])
ap.add_argument("--rounds", type=int, default=5)
ap.add_argument("--steps", type=int, default=150)
ap.add_argument("--warm-s", type=float, default=1.5)
args = ap.parse_args()
import pynvml  # noqa: E402
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)
NQ, K = 64, 50
sdb = ShardedPatchDatabase.synthetic(np.full(args.images, 40, np.int64), args.dim, seed=4, rank=0, world_size=1, device=0)
db = sdb.local
d_q = torch.from_numpy(synth.unit_queries(NQ, args.dim, 1)).cuda()
rng = np.random.default_rng(2)
bits = db.build_exclude_bits([np.sort(rng.choice(args.images, size=50, replace=False)).astype(np.int32) for _ in range(NQ)], NQ)
sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
ideal_us = db.n_rows * (args.dim * 2 + 1 / 8) / 6516.7e9 * 1e6


def step(shape):
    if shape == "inline":
        return sdb.scan_topk_device(d_q, K, d_exclude_bits=bits)
    return sdb.scan_topk_device(d_q, K, d_exclude_bits=bits, pipelined=True)


import time  # noqa: E402
t0 = time.perf_counter()
while time.perf_counter() - t0 < args.warm_s:
    for _ in range(50):
        step("inline")
    torch.cuda.synchronize()
res = {s: [] for s in args.shapes}
for rnd in range(args.rounds):
    for shape in args.shapes:
        if shape != "inline":
            sdb.set_side_sms(int(shape))
        for _ in range(10):
            step(shape)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(shape)
        if shape != "inline":
            sdb.drain()
        e1.record()
        mhz = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)       # the GPU is still working through the queue
        torch.cuda.synchronize()
        res[shape].append((e0.elapsed_time(e1) / args.steps * 1e3, mhz))
print(f"rows={db.n_rows} dim={args.dim} HBM-ideal {ideal_us:.1f} us; step in us (SM MHz while it ran), one column per round")
for shape in args.shapes:
    cells = "  ".join(f"{us:6.1f} ({mhz})" for us, mhz in res[shape])
    med = float(np.median([us for us, _ in res[shape]]))
    label = "exchange in line" if shape == "inline" else f"pipelined, side SMs = {shape}"
    print(f"{label:<26} {cells}   median {med:6.1f}  ideal/step {ideal_us / med:.3f}")
sdb.close()
