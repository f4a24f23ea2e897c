#!/usr/bin/env python
"""Latency / throughput of the cross-process batching front end (seesaw_b200/service.py) under 64 concurrent session
PROCESSES, each issuing single-query stage-1 scans back to back against one GPU-owning process (10M x 512 fp16).
Prints one JSON object: p50 / p99 request latency, achieved batch size and queries/s for several ``max_wait_s``."""
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N_IMAGES, PATCHES, DIM, K = int(os.environ.get("SSW_SVC_IMAGES", 250_000)), 40, 512, 50


def server(address, ready, max_wait_s):
    from seesaw_b200 import synth
    from seesaw_b200.engine import PatchDatabase
    from seesaw_b200.service import ScanServer
    db = PatchDatabase.synthetic(synth.dbidx_of_rows(np.full(N_IMAGES, PATCHES, np.int64)), DIM, seed=4, kind="tri", store="f16")
    srv = ScanServer(db, address, max_batch=64, max_wait_s=max_wait_s)
    ready.set()
    srv.serve_forever()


def session(address, i, n_req, start, out):
    from seesaw_b200 import synth
    from seesaw_b200.service import ScanClient
    c = ScanClient(address)
    q = synth.unit_queries(1, DIM, 100 + i)
    ex = [np.sort(np.random.default_rng(i).choice(N_IMAGES, size=50, replace=False))]
    c.scan_topk(q, K, exclude=ex)                     # connection + first batch warm
    start.wait()
    lat = np.empty(n_req)
    for r in range(n_req):
        t0 = time.perf_counter()
        c.scan_topk(q, K, exclude=ex)
        lat[r] = time.perf_counter() - t0
    out.put(lat)
    c.close()


def run(n_sessions, n_req, max_wait_s):
    from seesaw_b200.service import ScanClient
    ctx = mp.get_context("spawn")
    address = os.path.join(tempfile.mkdtemp(), "ssw.sock")
    ready, start = ctx.Event(), ctx.Event()
    srv = ctx.Process(target=server, args=(address, ready, max_wait_s), daemon=True)
    srv.start()
    assert ready.wait(600)
    out = ctx.Queue()
    ps = [ctx.Process(target=session, args=(address, i, n_req, start, out)) for i in range(n_sessions)]
    [p.start() for p in ps]
    time.sleep(3.0 + 0.05 * n_sessions)               # every session connected and warmed
    c = ScanClient(address)
    s0 = c.stats()
    t0 = time.perf_counter()
    start.set()
    lats = np.concatenate([out.get(timeout=600) for _ in ps])
    wall = time.perf_counter() - t0
    s1 = c.stats()
    [p.join(60) for p in ps]
    c.shutdown_server()
    srv.join(60)
    nb, nqs = s1["batches_issued"] - s0["batches_issued"], s1["queries_served"] - s0["queries_served"]
    return {"sessions": n_sessions, "requests_per_session": n_req, "max_wait_ms": max_wait_s * 1e3,
            "latency_ms_p50": float(np.percentile(lats, 50) * 1e3), "latency_ms_p99": float(np.percentile(lats, 99) * 1e3),
            "latency_ms_mean": float(lats.mean() * 1e3), "queries_per_s": float(len(lats) / wall),
            "gpu_passes": int(nb), "mean_batch": float(nqs / max(nb, 1))}


if __name__ == "__main__":
    res = {"workload": f"{N_IMAGES * PATCHES} x {DIM} fp16, top-{K}, 50 excluded ids per session; one request in flight per session process",
           "runs": []}
    for n_sessions, wait in ((1, 0.002), (8, 0.002), (64, 0.0), (64, 0.0005), (64, 0.002)):
        res["runs"].append(run(n_sessions, 200 if n_sessions > 1 else 100, wait))
        print(json.dumps(res["runs"][-1]), file=sys.stderr, flush=True)
    print(json.dumps(res))
