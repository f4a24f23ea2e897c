#!/usr/bin/env python
"""Benchmark of the SeeSaw vector-search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the metric is quoted on): a batch of 64 concurrent session
queries, each with its own exclusion set, against a synthetic 10M x 512 fp16 multiscale patch
database (250k images x 40 patches), per-image max + exclusion + top-50, row-sharded by image over
the N GPUs with one fused peer-memory exchange + merge kernel per step (strong scaling).
One step = one batch.  `value` = queries/s with inputs resident in HBM; `e2e` = the same through the
host-buffer entry point (C ABI `ssw_scan_topk` / `ssw_scan_topk_sharded`) with host<->device copies in the
timed region.  `--impl reference` times the reference's own CPU path (oracle/_ref, the verbatim copy of the
reference package made by oracle/build_ref.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "patch_scan_queries_per_s"
UNIT = "queries/s"
N_IMAGES, PATCHES, DIM, NQ, TOPK, N_EXCL = 250_000, 40, 512, 64, 50, 50
DB_SEED, Q_SEED, X_SEED = 4, 1, 2
CPU_FALLBACK_IMAGES = 25_000        # 1M rows: the sample the CPU arm falls back to when host RAM is short
K2_NCU = os.path.join(ROOT, "profiles", "r02_k2_scan_tc_ncu_full.txt")


def workload_config(n_gpus, exchange=None):
    return {"workload": "batched 64-session patch scan, 10M x 512 fp16 (250k images x 40 patches), "
                        "per-image max + per-query exclusion (50 ids) + top-50",
            "n_rows": N_IMAGES * PATCHES, "dim": DIM, "batch": NQ, "topk": TOPK, "exclude_per_query": N_EXCL,
            "storage": "fp16", "sharding": f"rows by image over {n_gpus} GPU(s)" + (f"; {exchange}" if exchange and n_gpus > 1 else ""),
            "l2": "inputs (10.24 GB) are larger than L2 (126 MB); no flush needed"}


def ncu_traffic_bytes(path=K2_NCU):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K2 launch on this exact workload, from the committed
    `ncu --set full` summary (scripts/ncu_summary.py); None when the file is missing."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for cand in (path, os.path.join(ROOT, "profiles", "r01_k2_scan_tc_ncu_full.txt")):
        total, seen = 0.0, 0
        try:
            for line in open(cand):
                f = line.split()
                if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in unit:
                    total += float(f[1]) * unit[f[2]]
                    seen += 1
        except OSError:
            continue
        if seen == 2:
            return int(total), os.path.relpath(cand, ROOT)
    return None, None


def make_queries_and_excludes():
    from seesaw_b200 import synth
    q = synth.unit_queries(NQ, DIM, Q_SEED)
    rng = np.random.default_rng(X_SEED)
    ex = [np.sort(rng.choice(N_IMAGES, size=N_EXCL, replace=False)).astype(np.int32) for _ in range(NQ)]
    return q, ex


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) on the host cores; the oracle port when it is absent
# ------------------------------------------------------------------------------------------
def blas_threads(n=None):
    """Pin the BLAS / OpenMP pools to n threads (default: every host core) whatever OMP_NUM_THREADS says —
    torchrun exports OMP_NUM_THREADS=1 — and return (limiter to keep alive, threads in use)."""
    n = n or os.cpu_count()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        lim = threadpool_limits(limits=n)
        used = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return lim, int(used)
    except Exception:      # noqa: BLE001
        return None, int(os.environ.get("OMP_NUM_THREADS", n))


def host_database(n_images, threads):
    """fp32 copy (the reference's storage type, multiscale_tools.py:200) of rows [0, n_images*40) of the SAME synthetic
    database the GPU arm scans, generated on the host cores."""
    from concurrent.futures import ThreadPoolExecutor

    from seesaw_b200 import synth
    n = n_images * PATCHES
    out = np.empty((n, DIM), np.float32)
    step = 50_000

    def fill(a):
        b = min(a + step, n)
        out[a:b] = synth.synth_rows(a, b - a, DIM, DB_SEED, "tri", np.float32)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(fill, range(0, n, step)))
    return out


def load_reference():
    """The unmodified reference's modules (oracle/_ref on the GPU box, /root/reference in the build container), or
    None when neither is there."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import refstubs
        return refstubs.import_reference()
    except Exception:      # noqa: BLE001
        return None


def cpu_scan_arm(steps, warmup, sample_images=None):
    """One step = ONE query of the 64-query batch answered by the reference's MultiscaleIndex._query_prelim
    (multiscale_index.py:291-312: sgemv + argsort of all rows + pandas isin + np.unique) on the full 10M x 512 fp32
    database — no extrapolation — when the host has the memory (>= 48 GB free), else on the first 1M rows scaled x10
    (flagged).  The reference has no batching: a 64-query batch costs 64 such calls."""
    import pandas as pd

    from seesaw_b200 import synth
    lim, threads = blas_threads()
    try:
        import psutil
        free_gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:      # noqa: BLE001
        free_gb = 0.0
    full = free_gb >= 48 and not sample_images
    n_images = N_IMAGES if full else (sample_images or CPU_FALLBACK_IMAGES)
    t0 = time.perf_counter()
    vecs = host_database(n_images, threads)
    gen_s = time.perf_counter() - t0
    dbidx = synth.dbidx_of_rows(np.full(n_images, PATCHES, np.int64))
    q, ex = make_queries_and_excludes()
    ex = [e[e < n_images] for e in ex]
    ref = load_reference()
    if ref is not None:
        kind = "reference"
        meta = pd.DataFrame({"dbidx": dbidx.astype(np.int64)})
        idx = ref.multiscale.MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, vec_index=None)

        def one(qi):
            return idx._query_prelim(vector=q[qi], topk_dbidx=TOPK, exclude_dbidx=ref.BitMap(ex[qi]))["dbidx"].values
        what = "seesaw.indices.multiscale.MultiscaleIndex._query_prelim of the unmodified reference (oracle/_ref; pyroaring replaced by a set-backed stub)"
    else:
        kind = "port"
        import seesaw_oracle as orc

        def one(qi):
            return orc.query_prelim(vecs, dbidx, q[qi], TOPK, exclude=ex[qi])["dbidx"]
        what = "oracle.query_prelim (numpy/pandas port of multiscale_index.py:291-312)"
    times, last = [], None
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        last = one(s % NQ)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    per_query = float(np.mean(times))
    scale = N_IMAGES // n_images
    out = {"value": 1.0 / (per_query * scale), "unit": UNIT, "cores": threads, "host_cores": os.cpu_count(), "kind": kind,
           "extrapolated": not full,
           "sample": f"{what}: 1 query per step on {'the full' if full else 'the first'} {n_images * PATCHES} x {DIM} fp32 rows "
                     f"({steps} timed steps after {warmup} warm-up, {per_query:.2f} s per query"
                     + ("" if full else f", scaled x{scale} to 10M rows (" + ("sample size forced on the command line" if sample_images
                                                                                 else f"only {free_gb:.0f} GB of host memory are free") + ")") + ")",
           "s_per_query": per_query * scale, "effective_gbs": n_images * PATCHES * DIM * 4 / per_query / 1e9,
           "host_generation_s": gen_s, "top1_dbidx_last_query": int(last[0]) if len(last) else None}
    del vecs
    return out


def cpu_knn_arm(budget_s=20.0):
    """kNN-graph CPU baseline: the reference's compute_exact_knn literally at N = 4k (and 16k when the
    budget allows), and at N = 1M the blockwise restatement (row blocks of 1024: 1 - Vb @ V.T, argpartition + stable
    sort of the k+1 smallest) on as many blocks as fit the time budget, scaled to N rows."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seesaw_oracle as orc
    lim, threads = blas_threads()
    ref = load_reference()
    rng = np.random.default_rng(5)
    out = {"cores": threads, "kind": "reference" if ref is not None else "port", "literal": []}
    fn = (lambda v, k: ref.knn_graph.compute_exact_knn(v, n_neighbors=k)) if ref is not None else orc.compute_exact_knn
    spent = 0.0
    for n in (4096, 16384):
        if n == 16384 and spent > 2.5:        # ~16x the 4k time: keep the whole bench within minutes
            break
        v = rng.standard_normal((n, DIM)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        t0 = time.perf_counter()
        fn(v, 10)
        dt = time.perf_counter() - t0
        spent += dt
        out["literal"].append({"n": n, "seconds": dt, "gflops_equivalent": 2.0 * n * n * DIM / dt / 1e9})
    n = 1_000_000
    v = rng.standard_normal((n, DIM)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    blocks, t_blocks = 0, 0.0
    while blocks < 8 and (blocks == 0 or t_blocks + t_blocks / blocks < budget_s):
        t0 = time.perf_counter()
        orc.exact_knn_candidates_blockwise(v, 10, block=1024, rows=(blocks * 1024, (blocks + 1) * 1024))
        t_blocks += time.perf_counter() - t0
        blocks += 1
    est = t_blocks / (blocks * 1024) * n
    out["blockwise_1m"] = {"blocks_timed": blocks, "rows_timed": blocks * 1024, "seconds_timed": t_blocks,
                           "seconds_for_1m_rows": est, "extrapolated": True, "gflops_equivalent": 2.0 * n * n * DIM / est / 1e9,
                           "what": "oracle.exact_knn_candidates_blockwise (1 - Vb @ V.T per 1024-row block, argpartition + stable sort), "
                                   "scaled by 1M / rows timed"}
    return out


class ClockSampler:
    """nvidia-smi in the background (it needs about a second to start, so it is launched at program start);
    samples are kept when their timestamp falls inside the timed window."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.t0, self.t1 = gpu_index, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= (self.t1 or r[0]) + 0.02]
        how = "timestamps inside the timed legs"
        if not inside and rows:       # clock skew / too short a window: fall back to the samples taken under load
            top = max(r[1] for r in rows)
            inside, how = [r for r in rows if r[1] >= 0.6 * top], "no sample inside the window: samples with the SM clock above idle"
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": reasons, "samples": len(sm), "selection": how,
                "window": "first warm-up step of the HBM-resident leg .. last timed step of the end-to-end leg"}


def main():
    # The driver reads ONE JSON line from stdout: send everything else libraries print (NCCL's version
    # banner, warnings) to stderr and keep the real stdout for the result line.
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-data-sweep", action="store_true", help="skip roofline_by_data (N=1)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 100M x 768 leg (N=8)")
    ap.add_argument("--nccl-exchange", action="store_true", help="N>1: use NCCL all-gather instead of the fused exchange kernel")
    ap.add_argument("--no-pipeline", action="store_true", help="N>1: do not overlap a step's exchange with the next step's scan")
    ap.add_argument("--side-sms", type=int, default=-1,
                    help="N>1, pipelined: SMs left to the exchange blocks (the scan runs on the others); 0 = exchange blocks next to "
                         "the scan CTAs, -1 = the library's choice by shard size")
    ap.add_argument("--cpu-sample-images", type=int, default=0, help="CPU arm on the first N images only (tests; default: the full database)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_scan_arm(steps=max(1, min(args.steps, 4)), warmup=1, sample_images=args.cpu_sample_images or None)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": NQ / r["value"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.gpus), "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "one timed step = one query (the reference has no batching; ms_per_step is 64 x the measured time per query)"}
        emit(line)
        return

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    import torch
    import torch.distributed as dist
    from seesaw_b200 import _lib, synth
    from seesaw_b200.engine import PatchDatabase, exclude_lists_to_csr
    from seesaw_b200.sharded import ShardedPatchDatabase

    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows_per_image = np.full(N_IMAGES, PATCHES, np.int64)
    sdb = ShardedPatchDatabase.synthetic(rows_per_image, DIM, seed=DB_SEED, rank=rank, world_size=world,
                                         device=local_rank)
    db = sdb.local
    exchange = ""
    if world > 1 and not args.nccl_exchange:
        sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
        side = sdb.set_side_sms(args.side_sms)
        exchange = ("fused peer-memory exchange + merge kernel, pipelined: the exchange of step i runs under the scan of step i+1 "
                    "(ssw_scan_topk_sharded_pipelined_device), " +
                    (f"on {side} SMs of its own (the scan on the other {148 - side})" if side > 0 else "exchange blocks next to the scan CTAs"))
    elif world > 1:
        exchange = "NCCL all-gather + merge kernel"
    if world > 1 and args.no_pipeline and not args.nccl_exchange:
        exchange = "fused peer-memory exchange + merge kernel (ssw_scan_topk_sharded_device)"
    q_host, ex_host = make_queries_and_excludes()
    d_q = torch.from_numpy(q_host).to(dev)
    d_bits = db.build_exclude_bits(ex_host, NQ)

    pipelined = world > 1 and not args.nccl_exchange and not args.no_pipeline

    def step_resident():
        # N > 1: pipelined steps — the exchange of step i (peer stores, flag wait, world merge) runs on a second stream
        # under the scan of step i+1; drained inside the timed region
        return sdb.scan_topk_device(d_q, TOPK, d_exclude_bits=d_bits, pipelined=pipelined)

    # ---- e2e leg: host buffers in, host results out, every step (exclude lists in the C ABI's CSR form)
    q_pinned = torch.from_numpy(q_host).pin_memory()
    ex_ids, ex_off = exclude_lists_to_csr(ex_host, NQ)
    h2d_bytes = q_host.nbytes + ex_ids.nbytes + ex_off.nbytes
    d2h_bytes = NQ * TOPK * (4 + 4 + 8) + NQ * 4

    def step_e2e():
        if world == 1:
            return db.scan_topk_csr(q_host, TOPK, ex_ids, ex_off)       # C ABI ssw_scan_topk: H2D + kernels + D2H
        if not args.nccl_exchange:
            return sdb.scan_topk_csr(q_host, TOPK, ex_ids, ex_off)      # C ABI ssw_scan_topk_sharded, host buffers
        dq = q_pinned.to(dev, non_blocking=True)
        bits = db.build_exclude_bits(ex_host, NQ)
        out = sdb.scan_topk_device(dq, TOPK, d_exclude_bits=bits)
        return {k: v.cpu() for k, v in out.items() if k in ("dbidx", "score", "row", "count")}

    def timed(fn, steps, profile=0, before=None, target=None, drain=None):
        target = target or db
        if before is not None:
            before()
        for _ in range(warmup):
            fn()
        if drain is not None:
            drain()
        barrier()
        if profile:
            target.profile(True, every=profile)
            target.profile_read()
        launches0 = _lib.kernel_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()                  # the last step's exchange completes inside the timed region
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        ms = max(dev_ms, 0.0)
        launches = _lib.kernel_launch_count() - launches0
        kern = target.profile_read() if profile else (0.0, 0)
        if profile:
            target.profile(False)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), launches, kern

    # the timed region: prep -> scan -> merge chained by programmatic dependent launches; every 8th step carries a
    # CUDA-event pair around its scan-kernel launch (on the kernel's stream) for the roofline, the other 7 run as in
    # production
    PROF_EVERY = 8 if args.steps >= 16 else 1
    dev_ms, wall_ms, launches, (kern_ms, kern_n) = timed(step_resident, args.steps, profile=PROF_EVERY,
                                                        before=sampler.window_begin if rank == 0 else None,
                                                        drain=sdb.drain if pipelined else None)
    e2e_steps = max(3, min(args.steps, 50))
    _, e2e_wall_ms, _, _ = timed(step_e2e, e2e_steps)
    if rank == 0:
        sampler.window_end()
    clocks = sampler.stop() if rank == 0 else None

    res = step_resident()
    if pipelined:
        sdb.drain()
    torch.cuda.synchronize()         # the two legs share the handle's workspace: one stream at a time
    res_e2e = step_e2e()
    torch.cuda.synchronize()

    # ---- parity of the sharded result: rank 0 rebuilds the WHOLE database on its own GPU and answers the same
    #      batch in one shard; ids, rows and scores of all [64, 50] results must be identical
    parity = None
    if rank == 0:
        if world == 1:
            whole = {k: res[k].cpu().numpy() for k in ("dbidx", "row", "score")}
        else:
            one = PatchDatabase.synthetic(synth.dbidx_of_rows(rows_per_image), DIM, seed=DB_SEED, kind="tri", store="f16",
                                          device=local_rank)
            bits1 = one.build_exclude_bits(ex_host, NQ)
            w = one.scan_topk_device(d_q, TOPK, bits1, decoded=True)
            torch.cuda.synchronize()
            whole = {k: w[k].cpu().numpy() for k in ("dbidx", "row", "score")}
            one.close()
            del one, w, bits1
            torch.cuda.empty_cache()
        same_dev = all((res[k].cpu().numpy() == whole[k]).all() for k in ("dbidx", "row", "score"))
        same_e2e = all((np.asarray(res_e2e[k].numpy() if hasattr(res_e2e[k], "numpy") else res_e2e[k]) == whole[k]).all()
                       for k in ("dbidx", "row", "score"))
        parity = bool(same_dev and same_e2e)
    if world > 1:
        dist.barrier()

    line = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if rank == 0:
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        esz = 2
        bytes_per_launch = db.n_rows * DIM * esz + db.n_rows // 8   # vectors + 1 boundary bit per row (this rank's shard)
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = bytes_per_launch / (kern_avg_ms * 1e-3) / 1e9 if kern_n else None
        ms_per_step = dev_ms / args.steps
        value = NQ / (ms_per_step * 1e-3)
        e2e_value = NQ / (e2e_wall_ms / e2e_steps * 1e-3)
        traffic, traffic_file = ncu_traffic_bytes() if world == 1 else (None, None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic", "config": workload_config(world, exchange),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps,
                        "api": "ssw_scan_topk (C ABI, host buffers)" if world == 1 else
                               ("ssw_scan_topk_sharded (C ABI, host buffers, every rank)" if not args.nccl_exchange else
                                "ShardedPatchDatabase.scan_topk_device from pinned host buffers + .cpu()")},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "ssw::scan_tc8_kernel<512,128,10,2> (K2, tcgen05 batched scan, 8 epilogue warps)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None,
                             "traffic": traffic,
                             "traffic_source": f"ncu --set full, one launch on this workload ({traffic_file})" if traffic_file else None,
                             "algorithmic_bytes_per_launch": int(bytes_per_launch),
                             "kernel_ms_avg": kern_avg_ms, "kernel_launches_timed": int(kern_n), "peak_source": peak_src,
                             "timing": f"CUDA-event pair on the kernel's stream around every {PROF_EVERY}th launch INSIDE the timed region "
                                       "(an event between two kernels rules out their programmatic dependent launch, so the "
                                       "sampled steps run without that overlap and the others with it)",
                             "kernel_share_of_step": (kern_avg_ms / ms_per_step) if ms_per_step else None},
                "clocks": clocks,
                "hbm_gbs_whole_step": N_IMAGES * PATCHES * DIM * esz / (ms_per_step * 1e-3) / 1e9,
                "parity_vs_n1": parity,
                "parity_how": "all [64,50] dbidx / rows / scores of the device-resident and the host-buffer step equal a one-shard "
                              "scan of the whole database recomputed on rank 0" if world > 1 else
                              "N = 1 is the one-shard scan itself (device-resident and host-buffer step compared)",
                "top1_dbidx_q0": int(res["dbidx"][0, 0])}
    # ---- secondary numbers (rank 0, N=1 only): single-query scan, kNN-graph build, CPU baseline
    if rank == 0 and world == 1:
        db.set_scan_mode(1)
        d_q1 = d_q[:1].contiguous()
        for _ in range(3):
            db.scan_topk_device(d_q1, TOPK, d_bits[:1])
        torch.cuda.synchronize()
        db.profile(True)
        db.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            db.scan_topk_device(d_q1, TOPK, d_bits[:1])
        e1.record()
        torch.cuda.synchronize()
        k1_ms, k1_n = db.profile_read()
        db.profile(False)
        db.set_scan_mode(0)
        step_ms = e0.elapsed_time(e1) / 10
        bytes1 = db.n_rows * (DIM * 2) + db.n_rows // 8
        line["single_query"] = {"kernel": "ssw::scan1_kernel<__half,2,0> (K1, streaming scan)", "ms_per_query": step_ms,
                                "queries_per_s": 1e3 / step_ms, "kernel_ms_avg": k1_ms / max(k1_n, 1),
                                "achieved_gbs": bytes1 / (k1_ms / max(k1_n, 1) * 1e-3) / 1e9,
                                "frac_of_peak": bytes1 / (k1_ms / max(k1_n, 1) * 1e-3) / 1e9 / line["roofline"]["peak"]}
    # ---- roofline under hostile and realistic data (N=1): the epilogue's work depends on the score order
    if rank == 0 and world == 1 and not args.no_data_sweep:
        sdb.close()
        del d_bits
        torch.cuda.empty_cache()
        line["roofline_by_data"] = data_sweep(torch, dev, peak, q_host, ex_host)
        sdb = None
    # ---- BASELINE configs[1]: 120k images x ~40 patches x 512 on one GPU, through the reference-facing class
    if rank == 0 and world == 1:
        line["config2_multiscale_120k_images"] = config2_leg(torch, dev, local_rank, q_host, d_q)
    # ---- BASELINE configs[0]: coarse index, 10k images x 512, one vector per image, top-10 with 300 excluded
    if rank == 0 and world == 1:
        import pandas as pd
        from seesaw_b200.indices import B200CoarseIndex, BitMap as _BitMap
        vc = synth.synth_rows(0, 10_000, DIM, 0, "tri", np.float32)
        cidx = B200CoarseIndex(embedding=None, vectors=vc, vector_meta=pd.DataFrame({"dbidx": np.arange(10_000, dtype=np.int64)}),
                               device=local_rank)
        exc = np.sort(np.random.default_rng(2).choice(10_000, size=300, replace=False))
        exb = _BitMap(exc)
        for _ in range(3):
            rc_ = cidx.query(topk=10, vector=q_host[0], exclude=exb)
        t0 = time.perf_counter()
        for i in range(50):
            rc_ = cidx.query(topk=10, vector=q_host[i % NQ], exclude=exb)
        ms_ours = (time.perf_counter() - t0) / 50 * 1e3
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import seesaw_oracle as _orc
        t0 = time.perf_counter()
        for i in range(20):
            ro_ = _orc.coarse_query(vc, np.arange(10_000), q_host[i % NQ], 10, exclude=exc)
        ms_cpu = (time.perf_counter() - t0) / 20 * 1e3
        same = bool((np.asarray(rc_["dbidxs"]) == _orc.coarse_query(vc, np.arange(10_000), q_host[49 % NQ], 10, exclude=exc)["dbidxs"]).all())
        line["config1_coarse_10k"] = {"query_call_ms": ms_ours, "cpu_oracle_ms": ms_cpu, "ids_equal_to_oracle": same,
                                      "query_call": "B200CoarseIndex.query(topk=10, vector, exclude=300 ids): host buffers in, result dict out",
                                      "cpu": f"oracle.coarse_query (numpy port of coarse_index.py:57-96) on {os.cpu_count()} host cores"}
        cidx.close()
    # ---- kNN-graph build (BASELINE config 4): k=10 exact graph over 1M x 512, output rows split over the ranks
    if not args.no_knn:
        from seesaw_b200.knn_graph import knn_candidates_device
        from seesaw_b200.sharded import knn_candidates_sharded, knn_row_ranges
        if sdb is not None and world == 1:
            sdb.close()
            sdb = None
            torch.cuda.empty_cache()
        n_knn = 1_000_000
        g = torch.Generator(device=dev).manual_seed(5)             # same seed on every rank: V is replicated
        v = torch.randn(n_knn, DIM, device=dev, generator=g)
        v = (v / v.norm(dim=1, keepdim=True)).half().contiguous()
        lo = int(knn_row_ranges(n_knn, world)[rank])
        knn_candidates_device(v, 10, rows=(lo, min(lo + 148 * 128, n_knn)))     # warm-up: one wave of row blocks
        knn_candidates_sharded(v[:65536], 10, rank=rank, world_size=world)      # ... and the all-gather path
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, _ = knn_candidates_sharded(v, 10, rank=rank, world_size=world)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        self_ok = bool((idx[::9973, 0].cpu() == torch.arange(0, n_knn, 9973, dtype=torch.int32)).all())
        if rank == 0:
            tf = 2.0 * n_knn * n_knn * DIM / (ms * 1e-3) / 1e12
            tpeak = float(peaks.get("bf16_tflops", 1590.0)) * world
            line["knn_build"] = {"kernel": "ssw::knn3_kernel<512> (K3: tcgen05 cta_group::2, M256 x N256 SS MMAs, fused row top-11)", "n": n_knn,
                                 "dim": DIM, "k": 10, "seconds": ms * 1e-3, "flops": 2.0 * n_knn * n_knn * DIM,
                                 "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                              "frac": tf / tpeak,
                                              "frac_of_sustained": tf / (float(peaks.get("bf16_tflops_sustained", tpeak / world)) * world)},
                                 "sharding": f"output rows over {world} GPU(s), V replicated, final all-gather of [N,11] ids+distances included",
                                 "self_is_nearest_spot_check": self_ok}
        if rank == 0 and world == 1:
            # the reference-facing call with HOST vectors in and the edge-table DataFrame out
            from seesaw_b200.knn_graph import compute_exact_knn
            v_host = v.cpu().numpy()
            del idx
            torch.cuda.empty_cache()
            t0 = time.perf_counter()
            df = compute_exact_knn(v_host, 10, device=local_rank)
            line["knn_build"]["e2e"] = {"seconds": time.perf_counter() - t0, "edges": int(len(df)),
                                        "api": "seesaw_b200.knn_graph.compute_exact_knn (C ABI ssw_knn_graph): H2D of the vectors, "
                                               "candidates + post_process_graph_df on the device, D2H of the edge table"}
            del df, v_host
            idx = None
        del v, idx
        torch.cuda.empty_cache()
    # ---- BASELINE configs[4] (N = 8): 100M x 768 fp16 (153.6 GB), 12.5M rows per GPU
    if world == 8 and not args.no_config5:
        if sdb is not None:
            sdb.close()
            sdb = None
        torch.cuda.empty_cache()
        c5 = config5_leg(torch, dist, dev, rank, world, local_rank, peak)
        if rank == 0:
            line["config5_100Mx768"] = c5
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if sdb is not None:
            sdb.close()
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_scan_arm(steps=2, warmup=1, sample_images=args.cpu_sample_images or None)
        if not args.no_knn:
            line["knn_build"]["cpu_baseline"] = cpu_knn_arm()
    if rank == 0:
        line.setdefault("cpu_baseline", None)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _kernel_frac(torch, db, d_q, k, bits, peak, mode, reps):
    """(kernel ms, fraction of the HBM peak, list updates and images offered per query and launch) for one scan kernel."""
    nq = d_q.shape[0]
    db.set_scan_mode(mode)
    for _ in range(3):
        db.scan_topk_device(d_q, k, bits)
    torch.cuda.synchronize()
    db.scan_stats(True)
    db.profile(True)
    db.profile_read()
    for _ in range(reps):
        db.scan_topk_device(d_q, k, bits)
    torch.cuda.synchronize()
    ms, n = db.profile_read()
    db.profile(False)
    upd, offered = db.scan_stats(False)
    db.set_scan_mode(0)
    ms /= max(n, 1)
    gbs = (db.n_rows * db.dim * 2 + db.n_rows // 8) / (ms * 1e-3) / 1e9
    return {"kernel_ms": ms, "achieved_gbs": gbs, "frac": gbs / peak,
            "list_updates_per_query": upd / max(n, 1) / (nq if mode == 2 else 1),
            "images_offered_per_query": offered / max(n, 1) / (nq if mode == 2 else 1)}


def data_sweep(torch, dev, peak, q_host, ex_host):
    """K1 and K2 roofline fraction at 10M x 512 fp16 on data that stresses the fused top-k epilogue (its work — list
    updates — depends on the ORDER of the scores): the benchmark's i.i.d. rows, rows sorted by ascending score of
    query 0 (every image beats everything before it), all-equal scores (every image passes the threshold vote and
    ties on the row), and unit-norm clustered rows (low-rank centres + noise, CLIP-like)."""
    from seesaw_b200 import synth
    from seesaw_b200.engine import PatchDatabase
    n_rows = N_IMAGES * PATCHES
    dbidx = synth.dbidx_of_rows(np.full(N_IMAGES, PATCHES, np.int64))
    d_q = torch.from_numpy(q_host).to(dev)
    out = {}

    def measure(name, db, note):
        bits = db.build_exclude_bits(ex_host, NQ)
        r = {"data": note,
             "K2_batch64": _kernel_frac(torch, db, d_q, TOPK, bits, peak, 2, 10),
             "K1_single": _kernel_frac(torch, db, d_q[:1].contiguous(), TOPK, bits[:1], peak, 1, 5)}
        out[name] = r
        db.close()
        torch.cuda.empty_cache()

    measure("iid", PatchDatabase.synthetic(dbidx, DIM, seed=DB_SEED, kind="tri", store="f16", device=dev.index),
            "the benchmark's database: i.i.d. triangular noise")
    g = torch.Generator(device=dev).manual_seed(11)
    # unit-norm clustered: 2000 centres in a 64-dimensional subspace + isotropic noise, rows L2-normalised
    basis = torch.randn(64, DIM, device=dev, generator=g)
    centres = torch.randn(2000, 64, device=dev, generator=g) @ basis / 8.0
    v = torch.empty((n_rows, DIM), dtype=torch.float16, device=dev)
    chunk = 1_000_000
    for a in range(0, n_rows, chunk):
        b = min(a + chunk, n_rows)
        # an image's patches share a centre (40 consecutive rows)
        c = torch.randint(0, 2000, ((b - a) // PATCHES + 1,), device=dev, generator=g).repeat_interleave(PATCHES)[: b - a]
        x = centres[c] + 0.6 * torch.randn(b - a, DIM, device=dev, generator=g)
        v[a:b] = (x / x.norm(dim=1, keepdim=True)).half()
        del x, c
    measure("clustered_unit_norm", PatchDatabase.from_device_tensor(v, dbidx, store="f16"),
            "unit-norm rows: 2000 cluster centres in a 64-d subspace + noise, one centre per image (CLIP-like)")
    # ascending score for query 0: the same clustered rows, sorted by their score under query 0
    s0 = torch.empty(n_rows, dtype=torch.float32, device=dev)
    q0 = torch.from_numpy(q_host[0]).to(dev).half()
    for a in range(0, n_rows, chunk):
        s0[a:a + chunk] = (v[a:a + chunk] @ q0).float()
    order = torch.argsort(s0)
    del s0
    vs = torch.empty_like(v)
    for a in range(0, n_rows, chunk):
        vs[a:a + chunk] = v[order[a:a + chunk]]
    del v, order
    torch.cuda.empty_cache()
    measure("ascending_for_query0", PatchDatabase.from_device_tensor(vs, dbidx, store="f16"),
            "the clustered rows sorted by ASCENDING score of query 0: in every CTA's range each image beats all before it")
    vs[:] = vs[0]
    measure("all_equal_scores", PatchDatabase.from_device_tensor(vs, dbidx, store="f16"),
            "every row identical: all scores tie, every image passes the threshold vote and loses on the row index")
    del vs
    torch.cuda.empty_cache()
    worst = min(min(r["K2_batch64"]["frac"], r["K1_single"]["frac"]) for r in out.values())
    out["worst_frac"] = worst
    return out


def config2_leg(torch, dev, local_rank, q_host, d_q):
    """120k images x 20..60 patches (4.8M x 512): K1 single-query scan, and the whole reference-facing call
    B200MultiscaleIndex.query(avg_score) on (a) the fp16-valued synthetic database and (b) REAL float32 unit vectors
    with the tiling pipeline's float32 boxes in exact mode (fp16 scan + float32 copy, certified re-ranking)."""
    from seesaw_b200 import synth
    from seesaw_b200.engine import PatchDatabase
    from seesaw_b200.indices import B200MultiscaleIndex, BitMap
    counts2 = synth.patches_per_image(120_000, 20, 60, 3)
    meta2 = synth.synth_vector_meta(counts2, 4)
    db2 = PatchDatabase.synthetic(meta2["dbidx"].to_numpy().astype(np.int32), DIM, seed=4, kind="tri", store="f16",
                                  device=local_rank)
    idx2 = B200MultiscaleIndex.from_database(db2, meta2)
    seen = BitMap(np.random.default_rng(5).choice(120_000, size=30, replace=False))

    def time_query(idx, reps=20):
        for _ in range(3):
            idx.query(vector=q_host[0], topk=3, shortlist_size=50, exclude=seen, agg_method="avg_score")
        t0 = time.perf_counter()
        for i in range(reps):
            out = idx.query(vector=q_host[i % NQ], topk=3, shortlist_size=50, exclude=seen, agg_method="avg_score")
        return (time.perf_counter() - t0) / reps * 1e3, out

    ms_query, out2 = time_query(idx2)
    d_q2 = d_q[:1].contiguous()
    bits2 = db2.build_exclude_bits([np.asarray(list(seen))], 1)
    db2.set_scan_mode(1)
    for _ in range(3):
        db2.scan_topk_device(d_q2, TOPK, bits2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        db2.scan_topk_device(d_q2, TOPK, bits2)
    e1.record()
    torch.cuda.synchronize()
    ms_k1 = e0.elapsed_time(e1) / 20
    r = {"n_rows": int(db2.n_rows), "dim": DIM, "storage": "fp16",
         "single_query_scan_ms": ms_k1, "single_query_scan_gbs": db2.n_rows * DIM * 2 / (ms_k1 * 1e-3) / 1e9,
         "query_call_ms": ms_query, "query_call_qps": 1e3 / ms_query,
         "query_call": "B200MultiscaleIndex.query(vector, topk=3, shortlist_size=50, exclude=30 seen ids, agg_method='avg_score'): "
                       "stage 1 (K1) + stage 2 (K7) on the GPU, host buffers in, result dict with activation DataFrames out",
         "returned": [int(x) for x in out2["dbidxs"]]}
    idx2.close()
    del db2, idx2
    torch.cuda.empty_cache()
    # (b) a real-index shape: float32 unit vectors (not fp16-representable), float32 pyramid boxes, exact mode.
    #     1.2M rows keep the host-side generation (numpy Gaussians) within seconds; the scan itself is size-linear.
    meta3, counts3 = synth.synth_pyramid_meta(30_000, 6)
    vecs3 = synth.unit_rows(int(counts3.sum()), DIM, 7)
    idx3 = B200MultiscaleIndex(embedding=None, vectors=vecs3, vector_meta=meta3, device=local_rank, store="f16")   # exact="auto"
    seen = BitMap(np.random.default_rng(5).choice(30_000, size=30, replace=False))
    ms3, out3 = time_query(idx3)
    info = idx3.db.exact_info()
    r["real_float32_index"] = {"n_rows": int(len(vecs3)), "n_images": 30_000, "vectors": "float32 unit Gaussians, not fp16-representable",
                               "boxes": "float32, pyramid tiling (multiscale_tools.py:96-117)", "storage": "fp16 scan copy + float32 copy (exact mode)",
                               "query_call_ms": ms3, "query_call_qps": 1e3 / ms3, "device_path": bool(idx3._store_exact and idx3._boxes_on_device),
                               "exact_queries": info["queries"], "float32_rescans": info["rescans"], "rho": info["rho"],
                               "returned": [int(x) for x in out3["dbidxs"]]}
    idx3.close()
    del idx3, vecs3
    torch.cuda.empty_cache()
    return r


def config5_leg(torch, dist, dev, rank, world, local_rank, peak):
    """100M x 768 fp16 over 8 GPUs: 64-query batch throughput and single-query latency, device time, max over ranks;
    K2<768> roofline fraction from event pairs around its launches on rank 0."""
    from seesaw_b200 import synth
    from seesaw_b200.sharded import ShardedPatchDatabase
    dim, rows_total = 768, 100_000_000
    counts = np.full(rows_total // PATCHES, PATCHES, np.int64)
    sdb = ShardedPatchDatabase.synthetic(counts, dim, seed=7, rank=rank, world_size=world, device=local_rank)
    sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
    q = torch.from_numpy(synth.unit_queries(NQ, dim, 1)).to(dev)
    rng = np.random.default_rng(2)
    ex = [rng.choice(len(counts), size=50, replace=False) for _ in range(NQ)]
    bits = sdb.local.build_exclude_bits(ex, NQ)

    def timed(fn, steps, warm=5, profile=False):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        if profile:
            sdb.local.profile(True)
            sdb.local.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        kern = sdb.local.profile_read() if profile else (0.0, 0)
        if profile:
            sdb.local.profile(False)
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), kern

    ms_batch, _ = timed(lambda: sdb.scan_topk_device(q, TOPK, d_exclude_bits=bits), 50)
    _, (kms, kn) = timed(lambda: sdb.scan_topk_device(q, TOPK, d_exclude_bits=bits), 20, profile=True)
    q1, b1 = q[:1].contiguous(), bits[:1]
    ms_single, _ = timed(lambda: sdb.scan_topk_device(q1, TOPK, d_exclude_bits=b1), 50)
    res = sdb.scan_topk_device(q, TOPK, d_exclude_bits=bits)
    torch.cuda.synchronize()
    gb = rows_total * dim * 2 / 1e9
    shard_bytes = sdb.local.n_rows * dim * 2 + sdb.local.n_rows // 8
    kern_ms = kms / max(kn, 1)
    out = {"config": f"{rows_total} x {dim} fp16 ({gb:.1f} GB) over {world} GPUs, {rows_total // world} rows per GPU",
           "batched_64": {"ms_per_batch": ms_batch, "queries_per_s": NQ / ms_batch * 1e3, "aggregate_hbm_gbs": gb / ms_batch * 1e3,
                          "ideal_ms_at_peak": shard_bytes / (peak * 1e9) * 1e3, "frac_of_peak_whole_step": shard_bytes / (peak * 1e9) * 1e3 / ms_batch},
           "roofline_rank0": {"kernel": "ssw::scan_tc8_kernel<768,128,10,1> (one accumulator, drained to registers)", "kernel_ms_avg": kern_ms,
                              "achieved_gbs": shard_bytes / (kern_ms * 1e-3) / 1e9 if kn else None,
                              "frac": shard_bytes / (kern_ms * 1e-3) / 1e9 / peak if kn else None},
           "single_query": {"ms": ms_single, "aggregate_hbm_gbs": gb / ms_single * 1e3},
           "top1_dbidx_q0": int(res["dbidx"][0, 0]), "count_q0": int(res["count"][0])}
    sdb.close()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
