"""Dev helper (library built with `make -C seesaw_b200/csrc EXTRA=-DSSW_TRACE`): per-CTA timeline of one K2 launch in SM
cycles, and the cycles each role spent WAITING — the producer for a free stage, the MMA warp for a free accumulator and
for a loaded stage, an epilogue warp for a finished accumulator — once right after start-up (boost clock) and once after
sustained load (power-capped clock): what a CTA waits for is what bounds it."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import _lib, synth  # noqa: E402
from seesaw_b200.engine import PatchDatabase  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
load_s = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
import pynvml  # noqa: E402
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)
dbidx = synth.dbidx_of_rows(np.full(n_img, 40, np.int64))
db = PatchDatabase.synthetic(dbidx, 512, seed=4, kind="tri", store="f16")
q = torch.from_numpy(synth.unit_queries(64, 512, 1)).cuda()
rng = np.random.default_rng(2)
bits = db.build_exclude_bits([np.sort(rng.choice(n_img, size=50, replace=False)) for _ in range(64)], 64)
db.set_scan_mode(2)
f = _lib.lib.ssw_scan_trace_read
f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int]
names = {0: "entry", 1: "setup done", 2: "pdl_wait done", 3: "producer start", 5: "producer last issued", 6: "mma: A ready",
         8: "mma: last committed", 9: "epi: tile0 ready", 12: "epi: mid tile", 10: "epi: last tile done", 11: "epi: published", 13: "exit"}
waits = {4: "producer waited for a free stage", 7: "mma waited for a free accumulator", 14: "mma waited for a loaded stage",
         15: "epilogue warp 2 waited for a finished accumulator"}


def traced(label):
    db.scan_stats(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        db.scan_topk_device(q, 50, bits)
    e0.record()
    db.scan_topk_device(q, 50, bits)
    e1.record()
    mhz = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
    torch.cuda.synchronize()
    out = np.zeros((148, 16), np.int64)
    assert f(db._h, out.ctypes.data_as(C.c_void_p), 148) == 0
    db.scan_stats(False)
    life = (out[:, 13] - out[:, 0]).astype(np.float64)
    print(f"== {label}: SM clock {mhz} MHz, step {e0.elapsed_time(e1) * 1e3:.1f} us, CTA lifetime median {np.median(life):.0f} cycles "
          f"(= {np.median(life) / mhz:.1f} us at that clock), max {life.max():.0f}")
    rel = (out - out[:, :1]).astype(np.float64)
    for i, nme in names.items():
        c = rel[:, i]
        print(f"   stamp {i:2d} {nme:24s} median {np.median(c):10.0f}  min {c.min():10.0f}  max {c.max():10.0f} cycles")
    for i, nme in waits.items():
        c = out[:, i].astype(np.float64)
        print(f"   wait  {i:2d} {nme:52s} median {np.median(c):10.0f} cycles = {100 * np.median(c / life):5.1f} % of the CTA's life "
              f"(min {100 * (c / life).min():5.1f} %, max {100 * (c / life).max():5.1f} %)")


traced("right after start-up")
t0 = time.perf_counter()
while time.perf_counter() - t0 < load_s:
    for _ in range(100):
        db.scan_topk_device(q, 50, bits)
    torch.cuda.synchronize()
traced(f"after {load_s:.1f} s of back-to-back launches")
t0 = time.perf_counter()
while time.perf_counter() - t0 < load_s:
    for _ in range(100):
        db.scan_topk_device(q, 50, bits)
    torch.cuda.synchronize()
traced(f"after {2 * load_s:.1f} s")
