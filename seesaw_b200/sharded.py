"""Row-sharded patch database: one process per GPU, one NCCL all-gather per query batch.

The reference has no multi-device path (its only scaling axes are annoy and subsetting,
multiscale_index.py:304-308, 364-376).  Here the database is split by IMAGE into contiguous ranges of
about equal row count, so every image's patches live on one GPU and the per-image max stays local
(SURVEY.md §8e).  Each rank scans its shard for the same query batch, emits [nq, k] (key, dbidx)
candidates whose keys embed GLOBAL row numbers (tie-breaking is therefore shard-invariant), the
lists are all-gathered (nq*k*12 bytes per rank) and every rank merges them with the K4 kernel.

The collective plumbing is backend-agnostic: with the ``gloo`` backend and an injected scanner the
same code runs on CPU tensors, which is how tests/test_sharded_cpu.py covers world_size 2."""
from __future__ import annotations

import numpy as np


def shard_image_ranges(rows_per_image, world_size):
    """Contiguous image ranges with ~equal rows: returns int64 [world_size+1] image boundaries."""
    rows_per_image = np.asarray(rows_per_image, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(rows_per_image)])
    total = int(cum[-1])
    targets = (np.arange(world_size + 1, dtype=np.float64) * total / world_size).astype(np.int64)
    bounds = np.searchsorted(cum, targets, side="left")
    bounds[0], bounds[-1] = 0, len(rows_per_image)
    return np.maximum.accumulate(bounds).astype(np.int64)


def merge_candidates_host(keys, dbidx, k):
    """numpy statement of the merge (K4) for CPU tensors: keys uint64 [n_lists, nq, k], 0 = empty.
    Returns (key, dbidx) [nq, k] best-first.  Used by the gloo path and as the checker of K4."""
    n_lists, nq, kk = keys.shape
    flat_k = np.transpose(keys, (1, 0, 2)).reshape(nq, n_lists * kk)
    flat_d = np.transpose(dbidx, (1, 0, 2)).reshape(nq, n_lists * kk)
    order = np.argsort(~flat_k, axis=1, kind="stable")[:, :k]        # descending by key
    out_k = np.take_along_axis(flat_k, order, axis=1)
    out_d = np.where(out_k != 0, np.take_along_axis(flat_d, order, axis=1), -1)
    return out_k, out_d.astype(np.int32)


def encode_keys(score, row):
    """(score fp32, global row) -> uint64 keys, the host statement of make_key (csrc/ssw_common.cuh):
    order-preserving score bits in the high word, ~row in the low word, so a larger key is a better
    candidate (higher score, then lower row); -0.0 ties with +0.0."""
    s = np.asarray(score, dtype=np.float32) + np.float32(0.0)
    b = s.view(np.uint32)
    u = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint64)
    return (u << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(row).astype(np.uint64))


def decode_keys(keys):
    """uint64 keys -> (score fp32, global row int64); empty slots -> (-inf, -1)."""
    keys = np.asarray(keys, dtype=np.uint64)
    u = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(u & np.uint32(0x80000000), u & np.uint32(0x7FFFFFFF), ~u).astype(np.uint32)
    score = bits.view(np.float32).copy()
    row = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = keys == 0
    score[empty] = -np.inf
    row[empty] = -1
    return score, row


class ShardedPatchDatabase:
    """One rank's view of a row-sharded database.

    ``local`` must offer ``scan_topk_device(d_queries, k, d_exclude_bits) -> (keys, dbidx)`` and
    ``build_exclude_bits(exclude, nq)``: a :class:`seesaw_b200.engine.PatchDatabase` on GPUs."""

    def __init__(self, local, *, rank, world_size, group=None, merge=None):
        self.local, self.rank, self.world_size, self.group = local, rank, world_size, group
        self._merge = merge
        self._xchg = None            # fused peer-memory exchange (enable_fused_exchange)

    # ---- fused exchange over peer memory --------------------------------------------------
    def enable_fused_exchange(self, nq_cap=64, k_cap=64):
        """Replace [local merge -> NCCL all-gather -> merge] by ONE kernel per step that merges the shard's
        lists, stores them into every rank's exchange buffer over NVLink and merges the world's lists
        (C ABI ``ssw_scan_topk_sharded_device``).  Collective: every rank must call it.  The buffers are
        plain cudaMalloc memory shared through CUDA IPC handles, which travel over the process group."""
        import ctypes as C

        import torch.distributed as dist

        from ._lib import check, lib
        dev = self.local.device
        buf, nbytes = C.c_void_p(), C.c_int64()
        handle = C.create_string_buffer(64)
        check(lib.ssw_xchg_create(dev, self.world_size, nq_cap, k_cap, C.byref(buf), handle, C.byref(nbytes)))
        handles = [None] * self.world_size
        if self.world_size > 1:
            dist.all_gather_object(handles, handle.raw, group=self.group)
        peers = (C.c_void_p * self.world_size)()
        for r in range(self.world_size):
            if r == self.rank:
                peers[r] = buf.value
            else:
                p = C.c_void_p()
                check(lib.ssw_xchg_open(dev, C.create_string_buffer(handles[r], 64), C.byref(p)))
                peers[r] = p.value
        if self.world_size > 1:
            dist.barrier(group=self.group)
        self._xchg = dict(buf=buf, peers=peers, nq_cap=nq_cap, k_cap=k_cap, epoch=0)

    def close(self):
        if self._xchg is not None:
            from ._lib import lib
            dev = self.local.device
            for r in range(self.world_size):
                if r != self.rank:
                    lib.ssw_xchg_close(dev, self._xchg["peers"][r])
            lib.ssw_xchg_destroy(dev, self._xchg["buf"])
            self._xchg = None
        self.local.close()

    def scan_topk(self, queries, k, exclude=None):
        """Host-buffer form (numpy in, numpy out) of the sharded step through the fused exchange: what a
        session front end calls on every rank.  Same dict as PatchDatabase.scan_topk."""
        from .engine import exclude_lists_to_csr
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, self.local.dim))
        ids, offsets = exclude_lists_to_csr(exclude, q.shape[0])
        return self.scan_topk_csr(q, k, ids, offsets)

    def scan_topk_csr(self, q, k, exclude_ids=None, exclude_offsets=None):
        """:meth:`scan_topk` with the exclude lists already in the C ABI's CSR form (see PatchDatabase.scan_topk_csr)."""
        from ._lib import check, lib, ptr
        assert self._xchg is not None, "call enable_fused_exchange() first"
        x = self._xchg
        nq = q.shape[0]
        out = dict(dbidx=np.empty((nq, k), np.int32), score=np.empty((nq, k), np.float32),
                   row=np.empty((nq, k), np.int64), count=np.empty(nq, np.int32))
        x["epoch"] += 1
        check(lib.ssw_scan_topk_sharded(self.local._h, ptr(q), nq, int(k), ptr(exclude_ids), ptr(exclude_offsets), x["peers"],
                                        self.world_size, self.rank, x["nq_cap"], x["k_cap"], x["epoch"],
                                        ptr(out["dbidx"]), ptr(out["score"]), ptr(out["row"]), ptr(out["count"])))
        return out

    def drain(self, stream=None):
        """After pipelined steps: make ``stream`` wait for the last exchange (its outputs are complete from there)."""
        import ctypes as C

        import torch

        from ._lib import check, lib
        s = torch.cuda.current_stream(self.local.device) if stream is None else stream
        check(lib.ssw_scan_pipeline_drain(self.local._h, C.c_void_p(s.cuda_stream)))

    def set_side_sms(self, side_sms):
        """SMs the pipelined step leaves to its exchange blocks (-1 = automatic by shard size, the default; 0 = small
        exchange blocks next to the scan CTAs).  Call between steps, after :meth:`drain`.  Returns the number in force."""
        import ctypes as C

        from ._lib import check, lib
        check(lib.ssw_scan_pipeline_side_sms(self.local._h, int(side_sms)))
        n = C.c_int(0)
        check(lib.ssw_scan_pipeline_side_sms_in_use(self.local._h, C.byref(n)))
        return n.value

    def _scan_fused(self, d_queries, k, d_exclude_bits, stream=None, pipelined=False):
        import ctypes as C

        import torch

        from ._lib import check, lib
        x = self._xchg
        nq = d_queries.shape[0]
        dev = d_queries.device
        out = dict(key=torch.empty((nq, k), dtype=torch.int64, device=dev),
                   dbidx=torch.empty((nq, k), dtype=torch.int32, device=dev),
                   score=torch.empty((nq, k), dtype=torch.float32, device=dev),
                   row=torch.empty((nq, k), dtype=torch.int64, device=dev),
                   count=torch.empty((nq,), dtype=torch.int32, device=dev))
        x["epoch"] += 1
        s = torch.cuda.current_stream(dev) if stream is None else stream
        bits = None if d_exclude_bits is None else C.c_void_p(d_exclude_bits.data_ptr())
        fn = lib.ssw_scan_topk_sharded_pipelined_device if pipelined else lib.ssw_scan_topk_sharded_device
        check(fn(
            self.local._h, C.c_void_p(d_queries.data_ptr()), nq, int(k), bits, x["peers"], self.world_size, self.rank,
            x["nq_cap"], x["k_cap"], x["epoch"], C.c_void_p(out["key"].data_ptr()), C.c_void_p(out["dbidx"].data_ptr()),
            C.c_void_p(out["score"].data_ptr()), C.c_void_p(out["row"].data_ptr()), C.c_void_p(out["count"].data_ptr()),
            C.c_void_p(s.cuda_stream)))
        return out

    @classmethod
    def synthetic(cls, rows_per_image, dim, *, seed, rank, world_size, device, kind="tri", store="f16", group=None):
        """Every rank calls this with the SAME rows_per_image; rank r materialises only its images."""
        from .engine import PatchDatabase
        from .synth import dbidx_of_rows
        rows_per_image = np.asarray(rows_per_image, dtype=np.int64)
        bounds = shard_image_ranges(rows_per_image, world_size)
        i0, i1 = int(bounds[rank]), int(bounds[rank + 1])
        row_base = int(rows_per_image[:i0].sum())
        dbidx = dbidx_of_rows(rows_per_image[i0:i1], dbidx_start=i0)
        local = PatchDatabase.synthetic(dbidx, dim, seed=seed, kind=kind, store=store, device=device,
                                        global_row_base=row_base)
        return cls(local, rank=rank, world_size=world_size, group=group)

    def scan_topk_device(self, d_queries, k, exclude=None, d_exclude_bits=None, pipelined=False):
        """Same result on every rank: dict(key, dbidx [nq,k], and on GPUs score/row/count).  ``pipelined=True`` (fused
        exchange only): this step's exchange runs under the NEXT step's scan; its outputs are complete on the current
        stream once the next pipelined step has been enqueued or after :meth:`drain`."""
        import torch
        import torch.distributed as dist
        nq = d_queries.shape[0]
        if d_exclude_bits is None and exclude is not None:
            d_exclude_bits = self.local.build_exclude_bits(exclude, nq)
        if self._xchg is not None and nq <= self._xchg["nq_cap"] and k <= self._xchg["k_cap"]:
            return self._scan_fused(d_queries, k, d_exclude_bits, pipelined=pipelined)
        if self.world_size == 1 and self._merge is None:
            return self.local.scan_topk_device(d_queries, k, d_exclude_bits, decoded=True)
        keys, dbidx = self.local.scan_topk_device(d_queries, k, d_exclude_bits)
        w = self.world_size
        all_k = torch.empty((w * nq, k), dtype=keys.dtype, device=keys.device)
        all_d = torch.empty((w * nq, k), dtype=dbidx.dtype, device=dbidx.device)
        dist.all_gather_into_tensor(all_k, keys.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_d, dbidx.contiguous(), group=self.group)
        all_k, all_d = all_k.view(w, nq, k), all_d.view(w, nq, k)
        if self._merge is not None:
            return self._merge(all_k, all_d, k)
        from .engine import merge_topk_device
        return merge_topk_device(all_k, all_d, k)


# ---------------------------------------------------------------------------------------------
# exact kNN graph across ranks: V replicated on every GPU, rank r computes output rows
# [lo_r, hi_r) against all columns, no communication until the final all-gather of [N, k1] (idx, dist)
# ---------------------------------------------------------------------------------------------
def knn_row_ranges(n, world_size, block=128):
    """Row range of every rank, aligned to the kernel's 128-row blocks: int64 [world_size+1]."""
    blocks = -(-n // block)
    per = -(-blocks // world_size)
    return np.minimum(np.arange(world_size + 1, dtype=np.int64) * per * block, n)


def knn_candidates_sharded(d_vectors, n_neighbors, *, rank, world_size, group=None, candidates=None):
    """Every rank passes the SAME [N, dim] tensor (fp16 CUDA tensor on GPUs) and receives the full
    candidate table (idx int32 [N,k1], dist fp32 [N,k1]) — compute_exact_knn's matmul + argsort
    (seesaw/knn_graph.py:170-182) split by output rows.  ``candidates(vectors, k, rows=(lo, hi))`` is the
    per-rank builder (default: the tcgen05 kernel); injected on CPU for the gloo test."""
    import torch
    import torch.distributed as dist
    if candidates is None:
        from .knn_graph import knn_candidates_device as candidates
    n = d_vectors.shape[0]
    k1 = min(int(n_neighbors) + 1, n)
    bounds = knn_row_ranges(n, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    idx, dist_ = candidates(d_vectors, n_neighbors, rows=(lo, hi))
    if world_size == 1:
        return idx, dist_
    per = int(bounds[1] - bounds[0])                          # equal-size slots, the last ones may be partly empty
    pad_i = torch.full((per, k1), -1, dtype=idx.dtype, device=idx.device)
    pad_d = torch.full((per, k1), float("inf"), dtype=dist_.dtype, device=dist_.device)
    pad_i[: hi - lo], pad_d[: hi - lo] = idx, dist_
    all_i = torch.empty((world_size * per, k1), dtype=idx.dtype, device=idx.device)
    all_d = torch.empty((world_size * per, k1), dtype=dist_.dtype, device=dist_.device)
    dist.all_gather_into_tensor(all_i, pad_i, group=group)
    dist.all_gather_into_tensor(all_d, pad_d, group=group)
    return all_i[:n], all_d[:n]
