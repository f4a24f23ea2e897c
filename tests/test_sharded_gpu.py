"""GPU tests of the row-sharded scan: world size 1 always; world size 2 over NCCL when the box has two GPUs
(the driver's 1-GPU box skips that one; `gpurun --gpus 2 -- pytest tests/test_sharded_gpu.py -m gpu` runs it)."""
import os
import socket

import numpy as np
import pytest

import seesaw_oracle as orc
from seesaw_b200 import synth

pytestmark = pytest.mark.gpu


def _problem():
    counts = synth.patches_per_image(6000, 1, 40, 8)
    dbidx = synth.dbidx_of_rows(counts)
    n = int(counts.sum())
    qs = synth.lattice_queries(20, 512, 9)
    ids = np.unique(dbidx)
    rng = np.random.default_rng(3)
    excl = [rng.choice(ids, size=int(s), replace=False) for s in rng.choice([0, 5, 300], size=len(qs))]
    return counts, dbidx, n, qs, excl


def _check(res, counts, dbidx, n, qs, excl, k):
    vecs = synth.synth_rows(0, n, 512, 23, "lattice", np.float32)
    got = {name: t.cpu().numpy() for name, t in res.items()}
    for i in range(len(qs)):
        o = orc.query_prelim(vecs, dbidx, qs[i], k, exclude=excl[i])
        m = len(o["dbidx"])
        assert got["count"][i] == m
        assert (got["dbidx"][i, :m] == o["dbidx"]).all(), i
        assert (got["row"][i, :m] == o["best_row"]).all(), i
        assert (got["score"][i, :m] == o["max_score"]).all(), i


def _exact_sharded_check(rank, world):
    """Exact mode across shards: float32 unit vectors (not fp16-representable) in fp16 storage + float32 copy; every
    shard certifies its own float32 top-k, the fused exchange merges them — the float32 oracle's ids, 1e-5 scores."""
    from seesaw_b200.engine import PatchDatabase
    from seesaw_b200.sharded import ShardedPatchDatabase, shard_image_ranges
    counts = synth.patches_per_image(3000, 2, 20, 31)
    dbidx = synth.dbidx_of_rows(counts, dbidx_start=4, dbidx_stride=3)
    n = int(counts.sum())
    vecs = synth.unit_rows(n, 512, 32)
    qs = synth.unit_queries(9, 512, 33)
    ids = np.unique(dbidx)
    excl = [ids[i::13] for i in range(len(qs))]
    bounds = shard_image_ranges(counts, world)
    cum = np.concatenate([[0], np.cumsum(counts)])
    r0, r1 = int(cum[bounds[rank]]), int(cum[bounds[rank + 1]])
    local = PatchDatabase.from_arrays(vecs[r0:r1], dbidx[r0:r1], store="f16", device=rank, global_row_base=r0, exact=True)
    sdb = ShardedPatchDatabase(local, rank=rank, world_size=world)
    sdb.enable_fused_exchange(nq_cap=32, k_cap=64)
    for batch in (slice(0, 9), slice(0, 1)):
        res = sdb.scan_topk(qs[batch], 50, exclude=excl[batch])
        for j, i in enumerate(range(*batch.indices(len(qs)))):
            o = orc.query_prelim(vecs, dbidx, qs[i], 50, exclude=excl[i])
            np.testing.assert_allclose(res["score"][j], o["max_score"], rtol=1e-5, atol=1e-7)
            if not (res["dbidx"][j] == o["dbidx"]).all():          # only float32 near-ties may swap
                s64 = orc.scores_f64(vecs, qs[i])
                d, s, _ = orc.per_image_best(s64, dbidx, excl[i])
                best = dict(zip(d.tolist(), s.tolist()))
                for a, b in zip(res["dbidx"][j].tolist(), o["dbidx"].tolist()):
                    assert a == b or abs(best[a] - best[b]) <= 1e-5 * abs(best[b]), (i, a, b)
            assert (dbidx[res["row"][j]] == res["dbidx"][j]).all()          # rows are GLOBAL original rows
    assert local.exact_info()["queries"] == 10
    sdb.close()


def _pipelined_check(sdb, counts, dbidx, n, qs, excl):
    """Both shapes of the pipelined step: the exchange on SMs of its own (the scan on SM count - 4 or - 2 CTAs with a
    partition of its own) and next to the scan CTAs."""
    for side in (4, 0, 2, -1):
        sdb.set_side_sms(side)
        _pipelined_steps(sdb, counts, dbidx, n, qs, excl)


def _pipelined_steps(sdb, counts, dbidx, n, qs, excl):
    """Pipelined steps: the exchange of step i runs under the scan of step i+1 on alternating workspaces.  Different
    queries / k every step; every step's outputs must equal the oracle once the pipeline is drained."""
    import torch
    outs = []
    for step in range(7):
        sel = np.roll(np.arange(len(qs)), step)[: 3 + 2 * step]
        k = (50, 3, 17)[step % 3]
        bits = sdb.local.build_exclude_bits([excl[i] for i in sel], len(sel))
        res = sdb.scan_topk_device(torch.from_numpy(qs[sel]).cuda(), k, d_exclude_bits=bits, pipelined=True)
        outs.append((sel, k, res, bits))
    sdb.drain()
    torch.cuda.synchronize()
    for sel, k, res, _ in outs:
        _check(res, counts, dbidx, n, qs[sel], [excl[i] for i in sel], k)


def test_sharded_world1():
    import torch
    from seesaw_b200.sharded import ShardedPatchDatabase
    counts, dbidx, n, qs, excl = _problem()
    sdb = ShardedPatchDatabase.synthetic(counts, 512, seed=23, rank=0, world_size=1, device=0, kind="lattice")
    for fused in (False, True):
        if fused:
            sdb.enable_fused_exchange(nq_cap=32, k_cap=64)
        for k in (50, 3):
            res = sdb.scan_topk_device(torch.from_numpy(qs).cuda(), k, exclude=excl)
            torch.cuda.synchronize()
            _check(res, counts, dbidx, n, qs, excl, k)
    _pipelined_check(sdb, counts, dbidx, n, qs, excl)
    sdb.close()
    _exact_sharded_check(0, 1)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from seesaw_b200.sharded import ShardedPatchDatabase
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    counts, dbidx, n, qs, excl = _problem()
    sdb = ShardedPatchDatabase.synthetic(counts, 512, seed=23, rank=rank, world_size=world, device=rank, kind="lattice")
    for fused in (False, True):
        if fused:
            sdb.enable_fused_exchange(nq_cap=32, k_cap=64)
        for rep in range(3):                      # several epochs: both buffer parities are reused
            for k in (50, 3):
                res = sdb.scan_topk_device(torch.from_numpy(qs).cuda(), k, exclude=excl)
                torch.cuda.synchronize()
                _check(res, counts, dbidx, n, qs, excl, k)
    # host-buffer form of the fused step
    for k in (50, 3):
        res = sdb.scan_topk(qs, k, exclude=excl)
        _check({n: torch.from_numpy(v) for n, v in res.items()}, counts, dbidx, n, qs, excl, k)
    _pipelined_check(sdb, counts, dbidx, n, qs, excl)
    dist.barrier()
    sdb.close()
    _exact_sharded_check(rank, world)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_sharded_world2_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
