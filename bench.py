#!/usr/bin/env python
"""Benchmark of the SeeSaw vector-search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the metric is quoted on): a batch of 64 concurrent session
queries, each with its own exclusion set, against a synthetic 10M x 512 fp16 multiscale patch
database (250k images x 40 patches), per-image max + exclusion + top-50, row-sharded by image over
the N GPUs with an NCCL all-gather + merge of the per-shard top-k lists (strong scaling).
One step = one batch.  `value` = queries/s with inputs resident in HBM; `e2e` = the same through the
host-buffer entry point (C ABI `ssw_scan_topk` at N=1) with host<->device copies in the timed region.
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
CPU_SAMPLE_IMAGES = 25_000          # 1M rows: the bounded sample the CPU arm runs on


def workload_config(n_gpus, exchange=None):
    return {"workload": "batched 64-session patch scan, 10M x 512 fp16 (250k images x 40 patches), "
                        "per-image max + per-query exclusion (50 ids) + top-50",
            "n_rows": N_IMAGES * PATCHES, "dim": DIM, "batch": NQ, "topk": TOPK, "exclude_per_query": N_EXCL,
            "storage": "fp16", "sharding": f"rows by image over {n_gpus} GPU(s)" + (f"; {exchange}" if exchange and n_gpus > 1 else ""),
            "l2": "inputs (10.24 GB) are larger than L2 (126 MB); no flush needed"}


def ncu_traffic_bytes(path=os.path.join(ROOT, "profiles", "r01_k2_scan_tc_ncu_full.txt")):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K2 launch on this exact workload, from the committed
    `ncu --set full` summary (scripts/ncu_summary.py); None when the file is missing."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        for line in open(path):
            f = line.split()
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in unit:
                total += float(f[1]) * unit[f[2]]
                seen += 1
    except OSError:
        return None
    return int(total) if seen == 2 else None


def make_queries_and_excludes():
    from seesaw_b200 import synth
    q = synth.unit_queries(NQ, DIM, Q_SEED)
    rng = np.random.default_rng(X_SEED)
    ex = [np.sort(rng.choice(N_IMAGES, size=N_EXCL, replace=False)).astype(np.int32) for _ in range(NQ)]
    return q, ex


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (numpy/pandas), all host threads numpy uses
# ------------------------------------------------------------------------------------------
def cpu_reference_arm(steps, warmup, n_queries_per_step=2):
    """Times oracle.query_prelim (reference: multiscale_index.py:291-312) on a bounded sample:
    the first 1M rows of the same synthetic database as an fp32 copy (the reference stores fp32,
    multiscale_tools.py:200), a few of the 64 queries per step, and scales rows linearly to 10M
    (argsort is N log N, so this favours the CPU)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seesaw_oracle as orc
    from seesaw_b200 import synth
    n_rows = CPU_SAMPLE_IMAGES * PATCHES
    vecs = synth.synth_rows(0, n_rows, DIM, DB_SEED, "tri", np.float32)
    dbidx = synth.dbidx_of_rows(np.full(CPU_SAMPLE_IMAGES, PATCHES, np.int64))
    q, ex = make_queries_and_excludes()
    ex = [e[e < CPU_SAMPLE_IMAGES] for e in ex]
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for j in range(n_queries_per_step):
            qi = (s * n_queries_per_step + j) % NQ
            orc.query_prelim(vecs, dbidx, q[qi], TOPK, exclude=ex[qi])
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    per_query_1m = float(np.mean(times)) / n_queries_per_step
    per_query_full = per_query_1m * (N_IMAGES / CPU_SAMPLE_IMAGES)
    return {"value": 1.0 / per_query_full, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle.query_prelim (numpy sgemv + stable argsort + isin + unique) on the first "
                      f"{n_rows} rows as fp32, {n_queries_per_step} queries/step x {steps} steps, "
                      f"{per_query_1m * 1e3:.1f} ms/query at 1M rows, scaled x{N_IMAGES // CPU_SAMPLE_IMAGES} to 10M rows",
            "ms_per_query_sample": per_query_1m * 1e3}


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
    ap.add_argument("--nccl-exchange", action="store_true", help="N>1: use NCCL all-gather instead of the fused exchange kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(steps=max(1, min(args.steps, 8)), warmup=1)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": NQ / r["value"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.gpus), "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    import torch
    import torch.distributed as dist
    from seesaw_b200 import _lib, synth
    from seesaw_b200.engine import merge_topk_device
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
    exchange = "fused peer-memory exchange + merge kernel (ssw_scan_topk_sharded_device)"
    if world > 1 and not args.nccl_exchange:
        sdb.enable_fused_exchange(nq_cap=NQ, k_cap=64)
    elif world > 1:
        exchange = "NCCL all-gather + merge kernel"
    q_host, ex_host = make_queries_and_excludes()
    d_q = torch.from_numpy(q_host).to(dev)
    d_bits = db.build_exclude_bits(ex_host, NQ)

    def step_resident():
        return sdb.scan_topk_device(d_q, TOPK, d_exclude_bits=d_bits)

    # ---- e2e leg: host buffers in, host results out, every step
    q_pinned = torch.from_numpy(q_host).pin_memory()
    ex_ids = np.concatenate(ex_host).astype(np.int32)
    ex_off = np.zeros(NQ + 1, np.int64)
    ex_off[1:] = np.cumsum([len(e) for e in ex_host])
    h2d_bytes = q_host.nbytes + ex_ids.nbytes + ex_off.nbytes
    d2h_bytes = NQ * TOPK * (4 + 4 + 8) + NQ * 4

    def step_e2e():
        if world == 1:
            return db.scan_topk(q_host, TOPK, exclude=ex_host)          # C ABI ssw_scan_topk: H2D + kernels + D2H
        if not args.nccl_exchange:
            return sdb.scan_topk(q_host, TOPK, exclude=ex_host)          # C ABI ssw_scan_topk_sharded, host buffers
        dq = q_pinned.to(dev, non_blocking=True)
        bits = db.build_exclude_bits(ex_host, NQ)
        out = sdb.scan_topk_device(dq, TOPK, d_exclude_bits=bits)
        return {k: v.cpu() for k, v in out.items() if k in ("dbidx", "score", "row", "count")}

    def timed(fn, steps, profile=False, before=None):
        if before is not None:
            before()
        for _ in range(warmup):
            fn()
        barrier()
        if profile:
            db.profile(True)
            db.profile_read()
        launches0 = _lib.kernel_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        ms = max(dev_ms, 0.0)
        launches = _lib.kernel_launch_count() - launches0
        kern = db.profile_read() if profile else (0.0, 0)
        if profile:
            db.profile(False)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), launches, kern

    dev_ms, wall_ms, launches, (kern_ms, kern_n) = timed(step_resident, args.steps, profile=True,
                                                        before=sampler.window_begin if rank == 0 else None)
    e2e_steps = max(3, min(args.steps, 50))
    _, e2e_wall_ms, _, _ = timed(step_e2e, e2e_steps)
    if rank == 0:
        sampler.window_end()
    clocks = sampler.stop() if rank == 0 else None

    # parity spot check inside the bench: shard-merged result == single-call result of rank 0's view
    res = step_resident()
    torch.cuda.synchronize()

    line = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    if rank == 0:
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        esz = 2
        bytes_per_launch = db.n_rows * DIM * esz + db.n_rows // 8   # vectors + 1 boundary bit per row (this rank's shard)
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = bytes_per_launch / (kern_avg_ms * 1e-3) / 1e9 if kern_n else None
        ms_per_step = dev_ms / args.steps
        value = NQ / (ms_per_step * 1e-3)
        e2e_value = NQ / (e2e_wall_ms / e2e_steps * 1e-3)
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
                             "traffic": ncu_traffic_bytes() if world == 1 else None,
                             "traffic_source": "ncu --set full, one launch on this workload (profiles/r01_k2_scan_tc_ncu_full.txt)",
                             "algorithmic_bytes_per_launch": int(bytes_per_launch),
                             "kernel_ms_avg": kern_avg_ms, "kernel_launches_timed": int(kern_n), "peak_source": peak_src,
                             "kernel_share_of_step": (kern_avg_ms / ms_per_step) if ms_per_step else None},
                "clocks": clocks,
                "hbm_gbs_whole_step": N_IMAGES * PATCHES * DIM * esz / (ms_per_step * 1e-3) / 1e9,
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
        bytes1 = db.n_rows * (DIM * 2 + 4)
        line["single_query"] = {"kernel": "ssw::scan1_kernel<__half,2,0> (K1, streaming scan)", "ms_per_query": step_ms,
                                "queries_per_s": 1e3 / step_ms, "kernel_ms_avg": k1_ms / max(k1_n, 1),
                                "achieved_gbs": bytes1 / (k1_ms / max(k1_n, 1) * 1e-3) / 1e9,
                                "frac_of_peak": bytes1 / (k1_ms / max(k1_n, 1) * 1e-3) / 1e9 / line["roofline"]["peak"]}
    # ---- BASELINE configs[1]: 120k images x ~40 patches x 512 fp16 on one GPU, through the reference-facing class
    if rank == 0 and world == 1:
        from seesaw_b200.engine import PatchDatabase
        from seesaw_b200.indices import B200MultiscaleIndex, BitMap
        counts2 = synth.patches_per_image(120_000, 20, 60, 3)
        meta2 = synth.synth_vector_meta(counts2, 4)
        db2 = PatchDatabase.synthetic(meta2["dbidx"].to_numpy().astype(np.int32), DIM, seed=4, kind="tri", store="f16",
                                      device=local_rank)
        idx2 = B200MultiscaleIndex.from_database(db2, meta2)
        seen = BitMap(np.random.default_rng(5).choice(120_000, size=30, replace=False))
        for _ in range(3):
            idx2.query(vector=q_host[0], topk=3, shortlist_size=50, exclude=seen, agg_method="avg_score")
        t0 = time.perf_counter()
        reps = 20
        for i in range(reps):
            out2 = idx2.query(vector=q_host[i % NQ], topk=3, shortlist_size=50, exclude=seen, agg_method="avg_score")
        ms_query = (time.perf_counter() - t0) / reps * 1e3
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
        line["config2_multiscale_120k_images"] = {
            "n_rows": int(db2.n_rows), "dim": DIM, "storage": "fp16",
            "single_query_scan_ms": ms_k1, "single_query_scan_gbs": db2.n_rows * DIM * 2 / (ms_k1 * 1e-3) / 1e9,
            "query_call_ms": ms_query, "query_call_qps": 1e3 / ms_query,
            "query_call": "B200MultiscaleIndex.query(vector, topk=3, shortlist_size=50, exclude=30 seen ids, agg_method='avg_score'): "
                          "stage 1 (K1) + stage 2 (K7) on the GPU, host buffers in, result dict with activation DataFrames out",
            "returned": [int(x) for x in out2["dbidxs"]]}
        idx2.close()
        del db2, idx2
        torch.cuda.empty_cache()
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sdb.close()
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_reference_arm(steps=4, warmup=1)
    if rank == 0:
        line.setdefault("cpu_baseline", None)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
