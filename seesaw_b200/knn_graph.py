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
    ``rows`` (default: all), self included when it ranks.  fp32 input is rounded to fp16 for the
    tensor cores (exact for fp16-valued data).  Returns (idx int32 [rows,k1], dist fp32 [rows,k1])."""
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
    """Generic form for an arbitrary edge table (e.g. from an approximate method), same contract
    as the reference function."""
    df = df.assign(src_vertex=df.src_vertex.astype("int32"), dst_vertex=df.dst_vertex.astype("int32"),
                   distance=np.clip(df.distance.values.astype("float32"), 0.0, None))
    df = df[df.src_vertex != df.dst_vertex]
    df = df.assign(dst_rank=df.groupby("src_vertex").distance.rank("first").astype("int32"))
    me = np.arange(nvec, dtype=np.int32)
    selfs = pd.DataFrame({"src_vertex": me, "dst_vertex": me, "distance": np.zeros(nvec, np.float32),
                          "dst_rank": np.zeros(nvec, np.int32)})
    return pd.concat([df, selfs], ignore_index=True).sort_values(["src_vertex", "dst_rank"]).reset_index(drop=True)


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


def rbf_kernel(edist):
    """knn_graph.py:8-22: cosine distance -> weight exp(-distance / edist), float64."""
    assert edist > 0
    spread = 1.0 / edist

    def kernel(arr):
        assert arr.min() >= -0.0001 and arr.max() <= 2.0001
        return np.exp(-(arr.astype("float64") * spread))

    return kernel


def knn_kernel(edist=2.1):
    """knn_graph.py:24-30: 0/1 weights, neighbours beyond ``edist`` discarded."""
    assert edist > 0.0

    def kernel(arr):
        return (arr <= edist).astype("float32")

    return kernel


def get_weight_matrix(df, *, kfun, self_edges=False, normalized, laplacian=False, symmetric=True):
    """Weight matrix / graph Laplacian of a kNN edge table — get_weight_matrix (knn_graph.py:31-104), the input of
    label propagation.  One-off host step (scipy); stated here so the graph chain compute_exact_knn ->
    get_weight_matrix -> B200LabelPropagation lives in one package.  Same semantics: weights kfun(distance), an
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


class KNNGraph:
    def __init__(self, knn_df, nvecs=None):
        self.knn_df = knn_df
        ks = knn_df.groupby("src_vertex").dst_rank.max()
        self._ks = ks
        self.k = ks.min()
        self.maxk = ks.median()
        self.nvecs = ks.shape[0]
        self.ind_ptr = get_lookup_ranges(knn_df.src_vertex, self.nvecs)

    def _check_rep(self):
        """knn_graph.py:258-262: with one self edge per vertex the sources and destinations coincide."""
        srcs, dsts = np.unique(self.knn_df.src_vertex.values), np.unique(self.knn_df.dst_vertex.values)
        assert np.array_equal(srcs, dsts), "self edges should guarantee this"
        assert self._ks.index.max() + 1 == len(srcs), "self edges guarantee this"

    @staticmethod
    def from_vectors(vectors, *, n_neighbors, device=0, **_ignored):
        """Entry point of scripts/make_knn_graph.py:49 — returns (graph, auxiliary index or None)."""
        return KNNGraph(compute_exact_knn(vectors, n_neighbors, device=device)), None

    def save(self, path, overwrite=False):
        os.makedirs(path, exist_ok=overwrite)
        self.knn_df.to_parquet(f"{path}/forward.parquet")

    @staticmethod
    def from_file(path):
        return KNNGraph(pd.read_parquet(f"{path}/forward.parquet"))

    def restrict_k(self, *, k):
        if k < self.maxk:
            return KNNGraph(self.knn_df[self.knn_df.dst_rank < k].reset_index(drop=True))
        if k > self.maxk:
            raise AssertionError(f"can only do up to k={self.k} neighbors based on input df")
        return self

    def rev_lookup(self, dst_vertex) -> pd.DataFrame:
        return self.knn_df.iloc[self.ind_ptr[dst_vertex]:self.ind_ptr[dst_vertex + 1]]
