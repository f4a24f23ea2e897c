"""Dev helper: time the stage-1 scan kernels on a synthetic database (CUDA events)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import synth  # noqa: E402
from seesaw_b200.engine import PatchDatabase  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=250000)
ap.add_argument("--patches", type=int, default=40)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--store", default="f16")
ap.add_argument("--nq", type=int, default=1)
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--excl", type=int, default=50)
args = ap.parse_args()

counts = np.full(args.images, args.patches, np.int64)
dbidx = synth.dbidx_of_rows(counts)
t0 = time.time()
db = PatchDatabase.synthetic(dbidx, args.dim, seed=4, kind="tri", store=args.store)
db.set_scan_mode(args.mode)
print(f"db {db.n_rows} x {db.dim} {args.store}: created in {time.time()-t0:.2f}s")
q = torch.from_numpy(synth.unit_queries(args.nq, args.dim, 1)).cuda()
rng = np.random.default_rng(0)
ex = [rng.choice(args.images, size=args.excl, replace=False) for _ in range(args.nq)]
bits = db.build_exclude_bits(ex, args.nq) if args.excl else None
for _ in range(3):
    keys, ids = db.scan_topk_device(q, args.k, bits)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
ev[0].record()
for i in range(args.iters):
    db.scan_topk_device(q, args.k, bits, keys, ids)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters)])
esz = 2 if args.store == "f16" else 4
gb = db.n_rows * args.dim * esz / 1e9
print(f"nq={args.nq} k={args.k} mode={args.mode}: median {np.median(ms):.3f} ms  min {ms.min():.3f} ms  "
      f"-> {gb / (np.median(ms) / 1e3) * (1 if args.nq >= 8 and args.mode != 1 else args.nq):.0f} GB/s algorithmic "
      f"({gb:.2f} GB per pass), {args.nq / (np.median(ms) / 1e3):.1f} q/s")
print("top ids q0:", ids[0, :8].tolist())
