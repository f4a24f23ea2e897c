"""Dev helper: sustained (power-capped) behaviour of the scan kernels next to a plain device copy."""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import synth  # noqa: E402
from seesaw_b200.engine import PatchDatabase  # noqa: E402


def clocks():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    return out


def run(name, fn, gb, iters=300, chunk=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for c in range(iters // chunk):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(chunk):
            fn()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / chunk
        print(f"{name:10s} chunk {c}: {ms:.3f} ms  {gb / ms * 1e3:7.0f} GB/s   [{clocks()}]", flush=True)


WHAT = os.environ.get("SUST", "copy,K1,K2").split(",")
if "copy" in WHAT:
    a = torch.empty(5 * 1024 ** 3 // 2, dtype=torch.float16, device="cuda")
    b = torch.empty_like(a)
    run("copy", lambda: b.copy_(a), 2 * a.numel() * 2 / 1e9)
    del a, b
counts = np.full(250000, 40, np.int64)
db = PatchDatabase.synthetic(synth.dbidx_of_rows(counts), 512, seed=4, kind="tri", store="f16")
q = torch.from_numpy(synth.unit_queries(64, 512, 1)).cuda()
rng = np.random.default_rng(0)
bits = db.build_exclude_bits([rng.choice(250000, size=50, replace=False) for _ in range(64)], 64)
gb = db.n_rows * 512 * 2 / 1e9
q1, b1 = q[:1].contiguous(), bits[:1]
for what in WHAT:
    if what == "K1":
        db.set_scan_mode(1)
        run("K1", lambda: db.scan_topk_device(q1, 50, b1), gb, iters=400)
    elif what == "K2":
        db.set_scan_mode(2)
        run("K2", lambda: db.scan_topk_device(q, 50, bits), gb, iters=600)
