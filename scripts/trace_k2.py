"""Dev helper (library built with `make -C seesaw_b200/csrc EXTRA=-DSSW_TRACE`): per-CTA timeline of one K2 launch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seesaw_b200 import _lib, synth  # noqa: E402
from seesaw_b200.engine import PatchDatabase  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 31250
dbidx = synth.dbidx_of_rows(np.full(n_img, 40, np.int64))
db = PatchDatabase.synthetic(dbidx, 512, seed=4, kind="tri", store="f16")
q = torch.from_numpy(synth.unit_queries(64, 512, 1)).cuda()
rng = np.random.default_rng(2)
bits = db.build_exclude_bits([np.sort(rng.choice(n_img, size=50, replace=False)) for _ in range(64)], 64)
db.set_scan_mode(2)
for _ in range(5):
    db.scan_topk_device(q, 50, bits)
torch.cuda.synchronize()
db.scan_stats(True)
db.scan_topk_device(q, 50, bits)
torch.cuda.synchronize()
out = np.zeros((148, 16), np.int64)
f = _lib.lib.ssw_scan_trace_read
f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int]
assert f(db._h, out.ctypes.data_as(C.c_void_p), 148) == 0
names = ["entry", "setup done", "pdl_wait done", "producer start", "producer tile0 issued", "producer last issued", "mma: A ready",
         "mma: tile0 committed", "mma: last committed", "epi: tile0 ready", "epi: last tile done", "epi: published", "epi: mid tile",
         "exit"]
rel = (out[:, :14] - out[:, :1]).astype(np.float64)
clk = 1.9e3  # cycles per us (approx; the SM clock under load may be lower)
print("stamp (cycles since CTA entry): median / min / max over 148 CTAs, and median in us at 1.9 GHz")
for i, nme in enumerate(names):
    c = rel[:, i]
    print(f"{i:2d} {nme:24s} {np.median(c):10.0f} {c.min():10.0f} {c.max():10.0f}   {np.median(c) / clk:8.2f} us")
