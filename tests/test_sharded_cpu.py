"""world_size-2 test of the row-sharded path on CPU (gloo): image-aligned sharding, the all-gather of
per-shard top-k lists and the merge give the same answer as one unsharded oracle scan.  The per-shard
scanner is injected (the oracle plays the GPU's part here); on GPUs the same ShardedPatchDatabase code
runs with PatchDatabase + NCCL (tests/test_sharded_gpu.py, bench.py)."""
import os
import socket

import numpy as np
import pytest

import seesaw_oracle as orc
from seesaw_b200 import sharded, synth


def test_shard_image_ranges_are_balanced_and_aligned():
    counts = synth.patches_per_image(1000, 20, 60, 3)
    for w in (1, 2, 3, 8):
        b = sharded.shard_image_ranges(counts, w)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        rows = np.add.reduceat(counts, b[:-1])[: w] if w > 1 else [counts.sum()]
        assert max(rows) - min(rows) <= 2 * 60
    # more ranks than images: empty shards are allowed
    b = sharded.shard_image_ranges(np.array([5, 5]), 4)
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b) >= 0).all()


def test_knn_row_ranges():
    for n, w in ((1000, 3), (128, 8), (1, 2), (100000, 8)):
        b = sharded.knn_row_ranges(n, w)
        assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
        assert all(x % 128 == 0 or x == n for x in b)


def test_key_roundtrip_and_order():
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.standard_normal(1000).astype(np.float32), np.float32([0.0, -0.0, 1.0, -1.0])])
    r = rng.integers(0, 2 ** 32 - 2, size=s.shape[0])
    k = sharded.encode_keys(s, r)
    s2, r2 = sharded.decode_keys(k)
    assert (s2 == s).all() and (r2 == r).all()
    order = np.argsort(~k, kind="stable")                       # descending keys
    want = np.lexsort((r, -(s + np.float32(0))))                # score desc, row asc
    assert (order == want).all()
    e = sharded.decode_keys(np.zeros(3, np.uint64))
    assert np.isinf(e[0]).all() and (e[1] == -1).all()


class OracleShard:
    """Stands in for PatchDatabase on a CPU rank: scans its rows with the oracle, emits keyed lists."""

    def __init__(self, vecs, dbidx, row_base):
        self.vecs, self.dbidx, self.row_base = vecs, dbidx, row_base

    def build_exclude_bits(self, exclude, nq):
        return exclude

    def scan_topk_device(self, queries, k, exclude):
        import torch
        nq = queries.shape[0]
        keys = np.zeros((nq, k), np.uint64)
        ids = np.full((nq, k), -1, np.int32)
        for i in range(nq):
            o = orc.query_prelim(self.vecs, self.dbidx, queries[i].numpy(), k,
                                 exclude=None if exclude is None else exclude[i])
            n = len(o["dbidx"])
            keys[i, :n] = sharded.encode_keys(o["max_score"], o["best_row"] + self.row_base)
            ids[i, :n] = o["dbidx"]
        return torch.from_numpy(keys.view(np.int64)), torch.from_numpy(ids)


def _merge_host(all_k, all_d, k):
    keys, ids = sharded.merge_candidates_host(all_k.numpy().view(np.uint64), all_d.numpy(), k)
    score, row = sharded.decode_keys(keys)
    return dict(key=keys, dbidx=ids, score=score, row=row, count=(keys != 0).sum(axis=1))


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = synth.patches_per_image(300, 1, 12, 5)
    dbidx = synth.dbidx_of_rows(counts, 4, 3)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 512, 17, "lattice", np.float32)     # exact arithmetic, many ties
    qs = synth.lattice_queries(5, 512, 18)
    ids = np.unique(dbidx)
    excl = [ids[::5], None, ids[:3], ids, ids[10:200]]
    b = sharded.shard_image_ranges(counts, world)
    r0, r1 = int(counts[: b[rank]].sum()), int(counts[: b[rank + 1]].sum())
    local = OracleShard(vecs[r0:r1], dbidx[r0:r1], r0)
    sdb = sharded.ShardedPatchDatabase(local, rank=rank, world_size=world, merge=_merge_host)
    for k in (1, 10, 60):
        res = sdb.scan_topk_device(torch.from_numpy(qs), k, exclude=excl)
        for i in range(len(qs)):
            o = orc.query_prelim(vecs, dbidx, qs[i], k, exclude=excl[i])
            m = len(o["dbidx"])
            assert res["count"][i] == m, (k, i)
            assert (res["dbidx"][i, :m] == o["dbidx"]).all(), (k, i)
            assert (res["row"][i, :m] == o["best_row"]).all(), (k, i)
            assert (res["score"][i, :m] == o["max_score"]).all(), (k, i)
            assert (res["dbidx"][i, m:] == -1).all()
    # kNN graph split by output rows + all-gather (the oracle's blockwise builder plays the kernel's part)
    v = synth.synth_rows(0, 333, 256, 19, "lattice", np.float32) * np.float32(0.25)

    def cand(vec, k, rows):
        i, d = orc.exact_knn_candidates_blockwise(vec.numpy(), k, block=64, rows=rows)
        return torch.from_numpy(i), torch.from_numpy(d)

    gi, gd = sharded.knn_candidates_sharded(torch.from_numpy(v), 10, rank=rank, world_size=world, candidates=cand)
    oi, od = orc.exact_knn_candidates(v, 10)
    assert (gi.numpy() == oi).all() and (gd.numpy() == od).all()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
