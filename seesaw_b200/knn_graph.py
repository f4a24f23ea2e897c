"""Exact kNN-graph construction on the B200, behind the reference's graph interfaces.

  compute_exact_knn      <->  seesaw/knn_graph.py:170-191   (matmul + row argsort -> tcgen05 kernel K3)
  post_process_graph_df  <->  seesaw/knn_graph.py:142-168   (unchanged semantics, vectorised)
  KNNGraph               <->  seesaw/knn_graph.py:246-286   plus the ``from_vectors`` / ``save`` entry points
                              that scripts/make_knn_graph.py:49-50 calls and the reference library lacks
                              (save format = ``{path}/forward.parquet``: scripts/make_knn_graphs_lvis.py:28-30)

The edge table is the reference's: columns src_vertex:int32, dst_vertex:int32, distance:float32,
dst_rank:int32, sorted by (src_vertex, dst_rank), one rank-0 zero-distance self edge per vertex."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pandas as pd

from ._lib import SSW_F16, SSW_F32, SSW_MAX_KNN_K1, check, lib, ptr


def knn_candidates(vectors, n_neighbors, *, device=0, rows=None):
    """The k1 = min(n_neighbors+1, N) columns minimising (fp32(1 - dot), column) for every row in
    ``rows`` (default: all), self included when it ranks.  The tensor cores multiply fp16 values; float32 input
    that is not fp16-representable is re-ranked / certified from the float32 rows on the device, so the result is the
    float32 neighbour list (see :func:`knn_exact_stats`).  Returns (idx int32 [rows,k1], dist fp32 [rows,k1])."""
    v = np.ascontiguousarray(vectors)
    if v.dtype not in (np.float16, np.float32):
        v = v.astype(np.float32)
    n, dim = v.shape
    k1 = min(int(n_neighbors) + 1, n)
    if k1 > SSW_MAX_KNN_K1:
        raise ValueError(f"n_neighbors+1 = {k1} exceeds the fused epilogue's limit {SSW_MAX_KNN_K1}")
    lo, hi = (0, n) if rows is None else rows
    idx = np.empty((hi - lo, k1), np.int32)
    dist = np.empty((hi - lo, k1), np.float32)
    check(lib.ssw_knn_build(device, ptr(v), SSW_F16 if v.dtype == np.float16 else SSW_F32, n, dim, k1, lo, hi,
                            ptr(idx), ptr(dist)))
    return idx, dist


def knn_candidates_device(d_vectors_f16, n_neighbors, *, rows=None, stream=None):
    """Device-resident variant: ``d_vectors_f16`` is a CUDA fp16 tensor [N, dim]; returns CUDA tensors."""
    import torch
    n, dim = d_vectors_f16.shape
    assert d_vectors_f16.dtype == torch.float16 and d_vectors_f16.is_contiguous()
    k1 = min(int(n_neighbors) + 1, n)
    lo, hi = (0, n) if rows is None else rows
    dev = d_vectors_f16.device
    idx = torch.empty((hi - lo, k1), dtype=torch.int32, device=dev)
    dist = torch.empty((hi - lo, k1), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev) if stream is None else stream
    check(lib.ssw_knn_build_device(dev.index or 0, C.c_void_p(d_vectors_f16.data_ptr()), n, dim, k1, lo, hi,
                                   C.c_void_p(idx.data_ptr()), C.c_void_p(dist.data_ptr()),
                                   C.c_void_p(s.cuda_stream)))
    return idx, dist


def edges_from_candidates(idx, dist, nvec, src_offset=0):
    """post_process_graph_df (knn_graph.py:142-168) for a candidate table whose rows are already
    ordered by (distance, column): clip at 0, drop self edges, dst_rank = 1.. in table order
    (== rank('first') on the clipped distances, which are non-decreasing along a row), add the
    rank-0 self edge of every vertex, order by (src_vertex, dst_rank).
    ``src_offset``/partial tables: self edges are added for vertices [src_offset, src_offset+rows)."""
    rows, k1 = idx.shape
    src_ids = np.arange(rows, dtype=np.int64) + src_offset
    keep = (idx != src_ids[:, None].astype(idx.dtype)) & (idx >= 0)
    m = keep.sum(axis=1)
    starts = np.cumsum(m + 1) - (m + 1)                      # row i occupies [starts[i], starts[i] + 1 + m[i])
    total = int((m + 1).sum())
    src = np.empty(total, np.int32)
    dst = np.empty(total, np.int32)
    dis = np.zeros(total, np.float32)
    rank = np.zeros(total, np.int32)
    src[:] = np.repeat(src_ids, m + 1).astype(np.int32)
    dst[starts] = src_ids.astype(np.int32)                   # self edge: distance 0, rank 0
    r = np.cumsum(keep, axis=1)                              # 1-based rank among kept entries
    pos = (starts[:, None] + r)[keep]
    dst[pos] = idx[keep]
    dis[pos] = np.clip(dist[keep].astype(np.float32), 0.0, None)
    rank[pos] = r[keep].astype(np.int32)
    return pd.DataFrame({"src_vertex": src, "dst_vertex": dst, "distance": dis, "dst_rank": rank})


def post_process_graph_df(df, nvec):
    """The contract of the reference's post_process_graph_df (knn_graph.py:142-168) for an ARBITRARY edge table (e.g.
    from an approximate method): int32 / float32 columns, distances clipped at 0, self edges dropped, ``dst_rank`` =
    1.. by (distance, table order) within each source, one rank-0 zero-distance self edge per vertex, rows ordered
    by (src_vertex, dst_rank).  Array code: one stable lexsort instead of a pandas groupby-rank."""
    src = df["src_vertex"].to_numpy().astype(np.int32)
    dst = df["dst_vertex"].to_numpy().astype(np.int32)
    dist = np.clip(df["distance"].to_numpy().astype(np.float32), 0.0, None)
    keep = src != dst
    src, dst, dist = src[keep], dst[keep], dist[keep]
    order = np.lexsort((np.arange(len(src)), dist, src))          # by source, then distance, then table order
    src, dst, dist = src[order], dst[order], dist[order]
    first = np.concatenate([[0], np.flatnonzero(np.diff(src)) + 1]) if len(src) else np.zeros(0, np.int64)
    run = np.repeat(first, np.diff(np.append(first, len(src))))
    rank = (np.arange(len(src)) - run + 1).astype(np.int32)
    me = np.arange(nvec, dtype=np.int32)
    src = np.concatenate([src, me])
    dst = np.concatenate([dst, me])
    dist = np.concatenate([dist, np.zeros(nvec, np.float32)])
    rank = np.concatenate([rank, np.zeros(nvec, np.int32)])
    order = np.lexsort((rank, src))
    return pd.DataFrame({"src_vertex": src[order], "dst_vertex": dst[order], "distance": dist[order], "dst_rank": rank[order]})


def knn_exact_stats():
    """After compute_exact_knn / knn_candidates on float32 vectors: dict(rows_refined, rows_rescanned, rho) — how many
    rows had their tensor-core candidates re-ranked in float32, how many of those needed the full float32 re-scan, and
    the largest fp16 rounding residual of a vector (0: the fp16 pass was exact).  See ssw_knn_exact_stats."""
    a, b, rho = C.c_int64(), C.c_int64(), C.c_double()
    check(lib.ssw_knn_exact_stats(C.byref(a), C.byref(b), C.byref(rho)))
    return dict(rows_refined=a.value, rows_rescanned=b.value, rho=rho.value)


def compute_exact_knn(vectors, n_neighbors, *, device=0):
    """compute_exact_knn (knn_graph.py:170-191): tcgen05 candidates + the edge table of post_process_graph_df,
    both on the device (C ABI ``ssw_knn_graph``); the host only wraps the four columns in a DataFrame."""
    v = np.ascontiguousarray(vectors)
    if v.dtype not in (np.float16, np.float32):
        v = v.astype(np.float32)
    n, dim = v.shape
    k1 = min(int(n_neighbors) + 1, n)
    if k1 > SSW_MAX_KNN_K1:
        raise ValueError(f"n_neighbors+1 = {k1} exceeds the fused epilogue's limit {SSW_MAX_KNN_K1}")
    cap = n * (k1 + 1)
    src, dst = np.empty(cap, np.int32), np.empty(cap, np.int32)
    dis, rank = np.empty(cap, np.float32), np.empty(cap, np.int32)
    total = C.c_int64()
    check(lib.ssw_knn_graph(device, ptr(v), SSW_F16 if v.dtype == np.float16 else SSW_F32, n, dim, int(n_neighbors),
                            ptr(src), ptr(dst), ptr(dis), ptr(rank), cap, C.byref(total)))
    t = total.value
    return pd.DataFrame({"src_vertex": src[:t], "dst_vertex": dst[:t], "distance": dis[:t], "dst_rank": rank[:t]})


class _DistanceKernel:
    """Edge weight as a function of the cosine distance — the ``kfun`` argument of get_weight_matrix."""

    def __init__(self, edist, hard):
        assert edist > 0
        self.edist, self.hard = float(edist), hard

    def __call__(self, distances):
        d = np.asarray(distances)
        if self.hard:                                     # knn_graph.py:24-30: 0/1 weights, cut at edist
            return (d <= self.edist).astype("float32")
        assert d.min() >= -0.0001 and d.max() <= 2.0001   # cosine distances (knn_graph.py:15-16)
        return np.exp(-(d.astype("float64") * (1.0 / self.edist)))      # knn_graph.py:8-22


def rbf_kernel(edist):
    """exp(-distance / edist) in float64 (knn_graph.py:8-22)."""
    return _DistanceKernel(edist, hard=False)


def knn_kernel(edist=2.1):
    """1 for neighbours within ``edist``, else 0 (knn_graph.py:24-30)."""
    return _DistanceKernel(edist, hard=True)


def compute_knn_from_nndescent(vectors, *, n_neighbors, n_jobs=-1, low_memory=False, device=0, **kwargs):
    """Drop-in for the reference's NN-descent graph builder (knn_graph.py:193-211), same signature and the same edge
    table.  pynndescent is approximate (and unpinned in the reference's lock file); the tensor-core build is exact at
    a fraction of its cost, so this routes to :func:`compute_exact_knn` — recall 1.0.  ``n_jobs`` / ``low_memory`` /
    pynndescent keywords are accepted and ignored."""
    return compute_exact_knn(vectors, n_neighbors, device=device)


def get_weight_matrix_device(df, *, kfun, device=0, return_weight_sum=False):
    """The matrix label propagation runs over — get_weight_matrix(df, kfun=, self_edges=False, normalized=False,
    symmetric=True) (knn_graph.py:31-104, as loops/graph_based.py:36-43 calls it) — built on the GPU (C ABI
    ``ssw_weight_matrix``): mirror look-ups, row sizes, scan, scatter and per-row ordering for the ~2 k N entries run as
    six small kernels; ``kfun`` is evaluated here with numpy, so the weights are the reference's own values and the CSR
    arrays come out bit-identical to scipy's.  Returns a scipy ``csr_array`` (and W.sum(0) when asked)."""
    import scipy.sparse as sp
    src = np.ascontiguousarray(df["src_vertex"].to_numpy(), dtype=np.int32)
    dst = np.ascontiguousarray(df["dst_vertex"].to_numpy(), dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(kfun(df["distance"].to_numpy()), dtype=np.float64))
    assert (w >= 0).all(), "edge weights must be non-negative"
    n_edges = len(src)
    n = int(src[-1]) + 1 if n_edges else 0
    indptr = np.empty(n + 1, np.int64)
    indices, data = np.empty(2 * n_edges, np.int32), np.empty(2 * n_edges, np.float64)
    wsum = np.empty(n, np.float64)
    nnz = C.c_int64()
    check(lib.ssw_weight_matrix(int(device), ptr(src), ptr(dst), ptr(w), n_edges, n, ptr(indptr), ptr(indices), ptr(data),
                                2 * n_edges, C.byref(nnz), ptr(wsum)))
    W = sp.csr_array((data[:nnz.value].copy(), indices[:nnz.value].copy(), indptr), shape=(n, n))
    return (W, wsum) if return_weight_sum else W


def get_weight_matrix(df, *, kfun, self_edges=False, normalized, laplacian=False, symmetric=True):
    """Weight matrix / graph Laplacian of a kNN edge table — get_weight_matrix (knn_graph.py:31-104), all variants,
    on the host (scipy); the variant label propagation uses has a device form, :func:`get_weight_matrix_device`.  Same semantics: weights kfun(distance), an
    edge listed from both ends gets the mean of its two weights, one listed from one end keeps its weight, the
    diagonal is stored as explicit zeros (the reference's setdiag(0.) keeps them), CSR with sorted indices;
    ``laplacian`` gives D - W (optionally D^-1/2 (D - W) D^-1/2)."""
    import scipy.sparse as sp
    assert not self_edges
    src, dst = df.src_vertex.values.astype(np.int64), df.dst_vertex.values.astype(np.int64)
    n = np.unique(src).shape[0]
    assert int((src == dst).sum()) == n, "one self edge per vertex expected"
    w = kfun(df.distance.values)
    assert (w >= 0).all(), "edge weights must be non-negative"
    w = np.asarray(w, dtype=np.float64)
    if symmetric:
        # every listed edge votes once for (i, j) and once for (j, i); weight = sum of listed weights / votes
        i2, j2 = np.concatenate([src, dst]), np.concatenate([dst, src])
        votes = sp.coo_array((np.ones(i2.shape[0]), (i2, j2)), shape=(n, n)).tocsr()
        pos_w = w > 0
        ww = np.where(pos_w, w, 0.0)
        wsum = sp.coo_array((np.concatenate([ww, ww]), (i2, j2)), shape=(n, n)).tocsr()
        votes.sum_duplicates(), wsum.sum_duplicates()
        votes.sort_indices(), wsum.sort_indices()
        assert np.array_equal(votes.indptr, wsum.indptr) and np.array_equal(votes.indices, wsum.indices)
        out = sp.csr_array((wsum.data / votes.data, wsum.indices.copy(), wsum.indptr.copy()), shape=(n, n))
        assert np.isclose(out.diagonal(), 1.0, atol=1e-5).all()       # kfun(0) == 1 on the self edges
    else:
        keep = w > 0
        out = sp.coo_array((w[keep], (src[keep], dst[keep])), shape=(n, n)).tocsr()
        out.sum_duplicates()
        out.sort_indices()
    rows = np.repeat(np.arange(n), np.diff(out.indptr))
    out.data[rows == out.indices] = 0.0                               # setdiag(0.): stored zeros stay
    degree = np.asarray(out.sum(axis=1)).reshape(-1)
    assert (degree > 0).all(), "no zero degree nodes allowed"
    if laplacian:
        assert symmetric
        out = -out
        out.setdiag(degree)
        if normalized:
            inv_sqrt = sp.dia_array((1.0 / np.sqrt(degree), 0), shape=(n, n))
            out = inv_sqrt @ (out @ inv_sqrt)
    out = sp.csr_array(out)
    out.sum_duplicates()
    out.sort_indices()
    return out


def get_lookup_ranges(sorted_col, nvecs):
    """CSR row pointer over a sorted vertex column (knn_graph.py:136-140)."""
    counts = np.bincount(np.asarray(sorted_col, dtype=np.int64), minlength=nvecs)
    return np.concatenate([[0], np.cumsum(counts)])


def factor_neighbors(knng, idx, k_intra):
    """factor_neighbors (knn_graph.py:213-242): split a patch-level graph into edges BETWEEN images — per source the
    nearest patch of every other image, re-ranked 0.. by distance — and edges WITHIN the source's own image (rank
    1..k_intra by distance; the rank-0 self edge has distance 0 and is rank 1 there, as in the reference).  ``idx`` is
    an index whose ``vector_meta.dbidx`` maps vertices to images.  Array code over the (src, rank)-sorted edge table."""
    dbidx = idx.vector_meta["dbidx"].to_numpy().astype(np.int32)
    df = knng.knn_df
    src, dst = df["src_vertex"].to_numpy(), df["dst_vertex"].to_numpy()
    dist, pos = df["distance"].to_numpy(), np.arange(len(df))
    sdb, ddb = dbidx[src], dbidx[dst]
    same = sdb == ddb

    def first_rank(keys, sel):
        """1-based rank of every selected edge within its key group by (distance, table order)."""
        order = np.lexsort((pos[sel], dist[sel]) + tuple(k[sel] for k in reversed(keys)))
        ks = [k[sel][order] for k in keys]
        new = np.ones(len(order), bool)
        if len(order) > 1:
            new[1:] = np.logical_or.reduce([k[1:] != k[:-1] for k in ks])
        start = np.maximum.accumulate(np.where(new, np.arange(len(order)), 0))
        rank = np.empty(len(order), np.int64)
        rank[order] = np.arange(len(order)) - start + 1
        return rank

    inter_sel = np.flatnonzero(~same)
    edge_rank = first_rank([src, ddb], inter_sel)
    keep = inter_sel[edge_rank <= 1]                                   # nearest patch per (source, other image)
    inter = df.iloc[keep].assign(src_dbidx=sdb[keep], dst_dbidx=ddb[keep])
    inter = inter.assign(dst_rank=(first_rank([src], keep) - 1).astype("int"))
    intra_sel = np.flatnonzero(same)
    r = first_rank([src], intra_sel)
    keep2 = intra_sel[r <= k_intra]
    intra = df.iloc[keep2].assign(src_dbidx=sdb[keep2], dst_dbidx=ddb[keep2], dst_rank=r[r <= k_intra].astype("int"))
    return pd.concat([inter, intra], ignore_index=True)


def _graph_base():
    """Inside the reference's environment KNNGraph IS the reference's class (plus the entry points its CLI expects);
    standalone, a minimal container with the same attributes."""
    try:
        from seesaw.knn_graph import KNNGraph as ref      # type: ignore
        return ref
    except Exception:      # noqa: BLE001
        return None


_RefKNNGraph = _graph_base()

if _RefKNNGraph is not None:
    class KNNGraph(_RefKNNGraph):
        """seesaw.knn_graph.KNNGraph plus ``from_vectors`` / ``save``, which scripts/make_knn_graph.py:49-50 calls and the
        reference class lacks."""

        @staticmethod
        def from_vectors(vectors, *, n_neighbors, device=0, **_ignored):
            return KNNGraph(compute_exact_knn(vectors, n_neighbors, device=device)), None

        def save(self, path, overwrite=False):
            os.makedirs(path, exist_ok=overwrite)
            self.knn_df.to_parquet(f"{path}/forward.parquet")

        @staticmethod
        def from_file(path):
            return KNNGraph(pd.read_parquet(f"{path}/forward.parquet"))
else:
    class KNNGraph:
        """Container over the edge table with the attributes the reference's consumers read (knn_graph.py:246-286):
        ``knn_df``, ``k`` / ``maxk`` (min / median over vertices of the largest rank), ``nvecs``, ``ind_ptr`` (CSR over
        src_vertex), ``restrict_k``, ``rev_lookup``; ``from_vectors`` / ``save`` are the entry points of
        scripts/make_knn_graph.py:49-50, ``from_file`` reads ``{path}/forward.parquet`` (:272-283)."""

        def __init__(self, knn_df, nvecs=None):
            self.knn_df = knn_df
            src = knn_df["src_vertex"].to_numpy().astype(np.int64)
            n = int(src.max()) + 1 if len(src) else 0
            top = np.zeros(n, np.int64)
            np.maximum.at(top, src, knn_df["dst_rank"].to_numpy().astype(np.int64))
            present = np.bincount(src, minlength=n) > 0
            self._top_rank = pd.Series(top[present], index=np.flatnonzero(present))
            self.k = self._top_rank.min()
            self.maxk = self._top_rank.median()
            self.nvecs = int(present.sum())
            self.ind_ptr = get_lookup_ranges(src, self.nvecs)

        def _check_rep(self):
            """Invariants the self edges guarantee (knn_graph.py:258-262): every destination is also a source, and the
            sources are exactly 0 .. nvecs-1."""
            seen_src = np.zeros(self.nvecs, bool)
            seen_src[self.knn_df["src_vertex"].to_numpy()] = True
            dst = self.knn_df["dst_vertex"].to_numpy()
            assert dst.max(initial=-1) < self.nvecs and seen_src.all() and seen_src[dst].all(), "self edges should guarantee this"

        @staticmethod
        def from_vectors(vectors, *, n_neighbors, device=0, **_ignored):
            """Returns (graph, auxiliary index or None) as the CLI unpacks it."""
            return KNNGraph(compute_exact_knn(vectors, n_neighbors, device=device)), None

        def save(self, path, overwrite=False):
            os.makedirs(path, exist_ok=overwrite)
            self.knn_df.to_parquet(f"{path}/forward.parquet")

        @staticmethod
        def from_file(path):
            return KNNGraph(pd.read_parquet(f"{path}/forward.parquet"))

        def restrict_k(self, *, k):
            if k > self.maxk:
                raise AssertionError(f"can only do up to k={self.k} neighbors based on input df")
            if k == self.maxk:
                return self
            return KNNGraph(self.knn_df[self.knn_df["dst_rank"] < k].reset_index(drop=True))

        def rev_lookup(self, dst_vertex) -> pd.DataFrame:
            lo, hi = self.ind_ptr[dst_vertex], self.ind_ptr[dst_vertex + 1]
            return self.knn_df.iloc[lo:hi]
