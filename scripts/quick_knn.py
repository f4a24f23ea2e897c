"""Dev helper: time the tcgen05 kNN-graph build on synthetic fp16 vectors (CUDA events)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200.knn_graph import knn_candidates_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--rows", type=int, default=0, help="only the first ROWS output rows (0 = all)")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(5)
v = torch.randn(args.n, args.dim, device="cuda", generator=g)
v = (v / v.norm(dim=1, keepdim=True)).half().contiguous()
rows = (0, args.rows if args.rows else args.n)
knn_candidates_device(v, args.k, rows=rows)
torch.cuda.synchronize()
ms = []
for _ in range(args.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx, dist = knn_candidates_device(v, args.k, rows=rows)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = np.array(ms)
flops = 2.0 * (rows[1] - rows[0]) * args.n * args.dim
print(f"n={args.n} rows={rows[1]-rows[0]} dim={args.dim} k={args.k}: median {np.median(ms):.2f} ms min {ms.min():.2f} ms "
      f"-> {flops / (np.median(ms) * 1e-3) / 1e12:.1f} TFLOP/s")
# spot check against torch
r = torch.randint(0, rows[1], (4,)).tolist()
ref = (1.0 - v[r].float() @ v.float().T)
print("spot check idx equal:", bool((torch.sort(ref, dim=1, stable=True).indices[:, :args.k + 1].int() == idx[r]).all()))
