"""Dev helper: sustained tensor throughput of cuBLAS (bf16/fp16 8192^3) next to the kNN-graph kernel."""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200.knn_graph import knn_candidates_device  # noqa: E402


def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                           "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()


def loop(name, fn, flops, seconds=2.5):
    fn()
    torch.cuda.synchronize()
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = fn()
        e1.record()
        e1.synchronize()
        print(f"{name:8s} {flops * n / (e0.elapsed_time(e1) * 1e-3) / 1e12:8.1f} TFLOP/s  [{clocks()}]", flush=True)


for dt in (torch.float16, torch.bfloat16):
    a = torch.randn(8192, 8192, device="cuda", dtype=dt)
    b = torch.randn(8192, 8192, device="cuda", dtype=dt)

    def mm():
        for _ in range(40):
            torch.matmul(a, b)
        return 40
    loop(f"cublas {str(dt)[-4:]}", mm, 2 * 8192 ** 3)
g = torch.Generator(device="cuda").manual_seed(5)
v = torch.randn(1_000_000, 512, device="cuda", generator=g)
v = (v / v.norm(dim=1, keepdim=True)).half().contiguous()
rows = 74 * 256 * 8


def knn():
    knn_candidates_device(v, 10, rows=(0, rows))
    return 1
loop("knn3", knn, 2.0 * rows * 1_000_000 * 512, seconds=3.0)
