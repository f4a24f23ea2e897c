"""Host-side handle over the C ABI: a patch database resident in the HBM of one B200.

``PatchDatabase`` is what the reference-facing classes in ``seesaw_b200.indices`` hold instead of
scanning ``self.vectors`` with numpy (indices/multiscale/multiscale_index.py:170-199)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib  # noqa: F401
from ._lib import SSW_BOX_F32, SSW_BOX_F64, SSW_BOX_I32, SSW_F16, SSW_F32, check, lib, ptr
from .synth import kind_id

_DT = {np.dtype(np.float32): SSW_F32, np.dtype(np.float16): SSW_F16}


def _dtype_id(d):
    if isinstance(d, str):
        d = {"f32": np.float32, "fp32": np.float32, "float32": np.float32,
             "f16": np.float16, "fp16": np.float16, "float16": np.float16}[d]
    return _DT[np.dtype(d)]


def _ids(e):
    if isinstance(e, np.ndarray):
        return e.astype(np.int32).reshape(-1)
    return np.fromiter((int(v) for v in e), dtype=np.int32)


def exclude_lists_to_csr(exclude, nq):
    """None | list of nq iterables of dbidx -> (ids int32, offsets int64 [nq+1])."""
    if exclude is None:
        return None, None
    assert len(exclude) == nq, "need one exclude list per query"
    arrs = [_ids(e) if e is not None else np.zeros(0, np.int32) for e in exclude]
    offsets = np.zeros(nq + 1, np.int64)
    offsets[1:] = np.cumsum([len(a) for a in arrs])
    ids = np.concatenate(arrs) if arrs else np.zeros(0, np.int32)
    return np.ascontiguousarray(ids, np.int32), offsets


class PatchDatabase:
    """[n_rows, dim] patch vectors + image id per row, resident on one GPU."""

    def __init__(self, handle):
        self._h = handle
        n_rows, n_images = C.c_int64(), C.c_int64()
        dim, dt, dev = C.c_int(), C.c_int(), C.c_int()
        check(lib.ssw_db_info(self._h, C.byref(n_rows), C.byref(n_images), C.byref(dim), C.byref(dt), C.byref(dev)))
        self.n_rows, self.n_images, self.dim = n_rows.value, n_images.value, dim.value
        self.dtype = np.float16 if dt.value == SSW_F16 else np.float32
        self.device = dev.value
        w = C.c_int64()
        check(lib.ssw_exclude_words(self._h, C.byref(w)))
        self.exclude_words = w.value

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, vectors, dbidx_per_row, *, store="f16", device=0, global_row_base=0, exact=False):
        """``exact=True`` (fp16 storage of float32 vectors only): keep the float32 rows in HBM as well, so that
        scans return the float32 top-k the reference computes (certified re-ranking, see ssw_db_attach_exact)."""
        vectors = np.ascontiguousarray(vectors)
        if vectors.dtype not in _DT:
            vectors = vectors.astype(np.float32)
        dbidx = np.ascontiguousarray(dbidx_per_row, dtype=np.int32).reshape(-1)
        assert vectors.ndim == 2 and vectors.shape[0] == dbidx.shape[0]
        h = C.c_void_p()
        check(lib.ssw_db_create(C.byref(h), device, ptr(vectors), _DT[vectors.dtype], _dtype_id(store),
                                vectors.shape[0], vectors.shape[1], ptr(dbidx), global_row_base))
        db = cls(h)
        if exact:
            assert vectors.dtype == np.float32 and _dtype_id(store) == SSW_F16, "exact mode: float32 vectors, fp16 storage"
            db.attach_exact(vectors)
        return db

    def attach_exact(self, vectors_f32):
        v = np.ascontiguousarray(vectors_f32, dtype=np.float32)
        assert v.shape == (self.n_rows, self.dim)
        check(lib.ssw_db_attach_exact(self._h, ptr(v)))

    def exact_info(self):
        """dict(attached, rho, vmax, queries, rescans) — see ssw_db_exact_info."""
        a, rho, vmax, nq, nr = C.c_int(), C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
        check(lib.ssw_db_exact_info(self._h, C.byref(a), C.byref(rho), C.byref(vmax), C.byref(nq), C.byref(nr)))
        return dict(attached=bool(a.value), rho=rho.value, vmax=vmax.value, queries=nq.value, rescans=nr.value)

    @classmethod
    def from_device_tensor(cls, d_vectors, dbidx_per_row, *, store="f16", global_row_base=0):
        """Vectors already on the GPU (a float16 / float32 CUDA tensor [n_rows, dim], rows grouped by ascending
        dbidx): copied / converted device to device, no trip through the host."""
        import torch
        assert d_vectors.is_cuda and d_vectors.is_contiguous() and d_vectors.dtype in (torch.float16, torch.float32)
        dbidx = np.ascontiguousarray(dbidx_per_row, dtype=np.int32).reshape(-1)
        assert d_vectors.ndim == 2 and d_vectors.shape[0] == dbidx.shape[0]
        h = C.c_void_p()
        check(lib.ssw_db_create_device(C.byref(h), d_vectors.device.index or 0, C.c_void_p(d_vectors.data_ptr()),
                                       SSW_F16 if d_vectors.dtype == torch.float16 else SSW_F32, _dtype_id(store),
                                       d_vectors.shape[0], d_vectors.shape[1], ptr(dbidx), global_row_base))
        return cls(h)

    @classmethod
    def synthetic(cls, dbidx_per_row, dim, *, seed, kind="tri", store="f16", device=0, global_row_base=0):
        dbidx = np.ascontiguousarray(dbidx_per_row, dtype=np.int32).reshape(-1)
        h = C.c_void_p()
        check(lib.ssw_db_create_synthetic(C.byref(h), device, _dtype_id(store), dbidx.shape[0], dim, ptr(dbidx),
                                          global_row_base, seed, kind_id(kind)))
        return cls(h)

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib.ssw_db_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scan_mode(self, mode):
        """0 auto, 1 streaming SIMT kernel, 2 tcgen05 batched kernel."""
        check(lib.ssw_set_scan_mode(self._h, int(mode)))

    def profile(self, on=True, every=1):
        """Bracket every ``every``-th scan-kernel launch with CUDA events (read back with :meth:`profile_read`)."""
        check(lib.ssw_profile_enable(self._h, int(every) if on else 0))

    def profile_read(self):
        """(summed scan-kernel milliseconds, launches) since the last read."""
        ms, n = C.c_double(), C.c_int64()
        check(lib.ssw_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def scan_stats(self, enable=True):
        """(list updates, images offered to a list) counted by the scan kernels since the last call; see ssw_scan_stats."""
        u, o = C.c_int64(), C.c_int64()
        check(lib.ssw_scan_stats(self._h, int(bool(enable)), C.byref(u), C.byref(o)))
        return u.value, o.value

    # ---- host-buffer API (what the reference-facing classes call) --------------------------
    def scan_topk(self, queries, k, exclude=None):
        """queries [nq, dim] or [dim] fp32; exclude: list of nq iterables of dbidx (or None).
        Returns dict(dbidx int32 [nq,k], score fp32, row int64, count int32 [nq])."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, self.dim))
        ids, offsets = exclude_lists_to_csr(exclude, q.shape[0])
        return self.scan_topk_csr(q, k, ids, offsets)

    def scan_topk_csr(self, q, k, exclude_ids=None, exclude_offsets=None):
        """:meth:`scan_topk` with the per-query exclude lists already in the C ABI's form: ``exclude_ids`` int32
        (all lists back to back) and ``exclude_offsets`` int64 [nq+1]; ``q`` float32 C-contiguous [nq, dim]."""
        nq = q.shape[0]
        ids, offsets = exclude_ids, exclude_offsets
        out_dbidx = np.empty((nq, k), np.int32)
        out_score = np.empty((nq, k), np.float32)
        out_row = np.empty((nq, k), np.int64)
        out_count = np.empty(nq, np.int32)
        check(lib.ssw_scan_topk(self._h, ptr(q), nq, int(k), ptr(ids), ptr(offsets), ptr(out_dbidx),
                                ptr(out_score), ptr(out_row), ptr(out_count)))
        return dict(dbidx=out_dbidx, score=out_score, row=out_row, count=out_count)

    def set_boxes(self, x1, y1, x2, y2, zoom_level):
        """vector_meta's box columns per ORIGINAL row, for the device stage 2 ('avg_score').  The IoU of the
        self-join is computed in the columns' own type like the reference's (box_utils.py:336-350): float32
        columns (what the tiling pipeline writes) stay float32, float64 stay float64, integer columns go as
        int32 (exact integer intersection / union, float32 quotient)."""
        boxes = [np.asarray(c).reshape(-1) for c in (x1, y1, x2, y2)]
        common = np.result_type(*boxes)                 # np.stack of the four columns, as df2tensor does
        if np.issubdtype(common, np.integer):
            kind, dt = SSW_BOX_I32, np.int32
            assert all(np.abs(c).max(initial=0) < 2 ** 31 for c in boxes), "integer box coordinates must fit int32"
        elif common == np.float32:
            kind, dt = SSW_BOX_F32, np.float32
        else:
            kind, dt = SSW_BOX_F64, np.float64
        cols = [np.ascontiguousarray(c.astype(dt)) for c in boxes]
        zoom = np.ascontiguousarray(np.asarray(zoom_level).astype(np.int32).reshape(-1))
        assert all(c.shape[0] == self.n_rows for c in cols) and zoom.shape[0] == self.n_rows
        check(lib.ssw_db_set_boxes_typed(self._h, kind, *[ptr(c) for c in cols], ptr(zoom)))
        self.has_boxes = True

    _AGG = {"plain_score": 0, "avg_score": 1}
    _AUG = {"all": 0, "greater": 1, "adjacent": 2}

    def rescore(self, query, cand_dbidx, *, query2=None, agg_method="avg_score", aug_larger="all"):
        """Stage 2 for the candidate images (rescore_candidates / score_frame2, multiscale_index.py:379-403,
        112-150): per candidate the aggregated score of its best patch (float64) and that patch's original
        row.  Scores are vectors.q [- vectors.query2]."""
        q = np.ascontiguousarray(np.asarray(query, dtype=np.float32).reshape(-1))
        q2 = None if query2 is None else np.ascontiguousarray(np.asarray(query2, dtype=np.float32).reshape(-1))
        ids = np.ascontiguousarray(np.asarray(cand_dbidx).astype(np.int32).reshape(-1))
        score = np.empty(len(ids), np.float64)
        row = np.empty(len(ids), np.int64)
        check(lib.ssw_rescore(self._h, ptr(q), ptr(q2), ptr(ids), len(ids), self._AGG[agg_method], self._AUG[aug_larger],
                              ptr(score), ptr(row)))
        return score, row

    def topk_from_scores(self, scores, k, exclude=None, row_mask=None):
        """Per-image max of a caller-supplied score per row (original order) + exclusion + top-k: the
        ``_get_top_dbidxs`` step of KnnProp2.next_batch (loops/graph_based.py:97-99).  ``row_mask``: rows
        that take part (default all).  Returns dict(dbidx, score, row) trimmed to the count."""
        sc = np.ascontiguousarray(np.asarray(scores, dtype=np.float32).reshape(-1))
        assert sc.shape[0] == self.n_rows
        mask = None if row_mask is None else np.ascontiguousarray(np.asarray(row_mask).astype(np.uint8).reshape(-1))
        ids = np.zeros(0, np.int32) if exclude is None else _ids(exclude)
        out_dbidx, out_score = np.empty(k, np.int32), np.empty(k, np.float32)
        out_row, cnt = np.empty(k, np.int64), np.zeros(1, np.int32)
        check(lib.ssw_topk_from_scores(self._h, ptr(sc), ptr(mask), int(k), ptr(ids), len(ids), ptr(out_dbidx),
                                       ptr(out_score), ptr(out_row), ptr(cnt)))
        n = int(cnt[0])
        return dict(dbidx=out_dbidx[:n], score=out_score[:n], row=out_row[:n])

    def topk_from_order(self, row_order, k, exclude=None):
        """Top-k images from rows ALREADY ranked by the caller (best first): the input ``_get_top_dbidxs`` itself
        takes.  Returns dict(dbidx, pos, row) trimmed to the count; ``pos`` indexes ``row_order``."""
        order = np.ascontiguousarray(np.asarray(row_order, dtype=np.int64).reshape(-1))
        ids = np.zeros(0, np.int32) if exclude is None else _ids(exclude)
        out_dbidx, out_pos = np.empty(k, np.int32), np.empty(k, np.int64)
        out_row, cnt = np.empty(k, np.int64), np.zeros(1, np.int32)
        check(lib.ssw_topk_from_order(self._h, ptr(order), len(order), int(k), ptr(ids), len(ids), ptr(out_dbidx), ptr(out_pos),
                                      ptr(out_row), ptr(cnt)))
        n = int(cnt[0])
        return dict(dbidx=out_dbidx[:n], pos=out_pos[:n], row=out_row[:n])

    def score_all(self, query):
        q = np.ascontiguousarray(np.asarray(query, dtype=np.float32).reshape(-1))
        assert q.shape[0] == self.dim
        out = np.empty(self.n_rows, np.float32)
        check(lib.ssw_score_all(self._h, ptr(q), ptr(out)))
        return out

    # ---- device-tensor API (torch tensors on this GPU; asynchronous on the given stream) ---
    def _stream(self, stream):
        import torch
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    def build_exclude_bits(self, exclude, nq, stream=None):
        """Device bitmaps [nq, exclude_words] (int32 tensor holding the uint32 words)."""
        import torch
        ids, offsets = exclude_lists_to_csr(exclude, nq)
        dev = torch.device("cuda", self.device)
        d_ids = torch.from_numpy(ids if len(ids) else np.zeros(1, np.int32)).to(dev)
        d_off = torch.from_numpy(offsets).to(dev)
        bits = torch.empty((nq, self.exclude_words), dtype=torch.int32, device=dev)
        check(lib.ssw_exclude_build_device(self._h, C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_off.data_ptr()), nq,
                                           int(offsets[-1]), C.c_void_p(bits.data_ptr()), self._stream(stream)))
        return bits

    def scan_topk_device(self, d_queries, k, d_exclude_bits=None, out_key=None, out_dbidx=None, stream=None,
                         decoded=False):
        """d_queries: float32 CUDA tensor [nq, dim].  Returns (keys [nq,k] int64 holding the uint64
        keys, dbidx int32 [nq,k]), or with ``decoded=True`` the dict key/dbidx/score/row/count;
        asynchronous."""
        import torch
        nq = d_queries.shape[0]
        assert d_queries.is_cuda and d_queries.dtype == torch.float32 and d_queries.is_contiguous()
        dev = d_queries.device
        if out_key is None:
            out_key = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if out_dbidx is None:
            out_dbidx = torch.empty((nq, k), dtype=torch.int32, device=dev)
        bits = None if d_exclude_bits is None else C.c_void_p(d_exclude_bits.data_ptr())
        score = row = count = None
        if decoded:
            score = torch.empty((nq, k), dtype=torch.float32, device=dev)
            row = torch.empty((nq, k), dtype=torch.int64, device=dev)
            count = torch.empty((nq,), dtype=torch.int32, device=dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        check(lib.ssw_scan_topk_device(self._h, C.c_void_p(d_queries.data_ptr()), nq, int(k), bits,
                                       p(out_key), p(out_dbidx), p(score), p(row), p(count), self._stream(stream)))
        if decoded:
            return dict(key=out_key, dbidx=out_dbidx, score=score, row=row, count=count)
        return out_key, out_dbidx

    def score_all_device(self, d_query, out=None, stream=None):
        import torch
        if out is None:
            out = torch.empty(self.n_rows, dtype=torch.float32, device=d_query.device)
        check(lib.ssw_score_all_device(self._h, C.c_void_p(d_query.data_ptr()), C.c_void_p(out.data_ptr()),
                                       self._stream(stream)))
        return out


def merge_topk_device(d_keys, d_dbidx, k, stream=None):
    """d_keys/d_dbidx: CUDA tensors [n_lists, nq, k] (int64 holding uint64 keys, int32).
    Returns dict of CUDA tensors key/dbidx/score/row/count for the merged top-k."""
    import torch
    n_lists, nq, kk = d_keys.shape
    assert kk == k and d_keys.is_contiguous() and d_dbidx.is_contiguous()
    dev = d_keys.device
    out = dict(key=torch.empty((nq, k), dtype=torch.int64, device=dev),
               dbidx=torch.empty((nq, k), dtype=torch.int32, device=dev),
               score=torch.empty((nq, k), dtype=torch.float32, device=dev),
               row=torch.empty((nq, k), dtype=torch.int64, device=dev),
               count=torch.empty((nq,), dtype=torch.int32, device=dev))
    s = torch.cuda.current_stream(dev) if stream is None else stream
    check(lib.ssw_merge_topk_device(dev.index or 0, C.c_void_p(d_keys.data_ptr()), C.c_void_p(d_dbidx.data_ptr()),
                                    n_lists, nq, k, C.c_void_p(out["key"].data_ptr()),
                                    C.c_void_p(out["dbidx"].data_ptr()), C.c_void_p(out["score"].data_ptr()),
                                    C.c_void_p(out["row"].data_ptr()), C.c_void_p(out["count"].data_ptr()),
                                    C.c_void_p(s.cuda_stream)))
    return out
