"""GPU parity tests of the tcgen05 kNN-graph build (K3) against the reference outputs and the oracle."""
import numpy as np
import pandas as pd
import pytest

import cases
import seesaw_oracle as orc
from seesaw_b200 import synth
from test_oracle import assert_graph_equal_mod_ties

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kg():
    from seesaw_b200 import knn_graph
    return knn_graph


def _ref_df(golden, name):
    return pd.DataFrame({c: golden[f"{name}/{c}"] for c in ("src_vertex", "dst_vertex", "distance", "dst_rank")})


def check_candidates(idx, dist, v, k, rel=1e-5):
    """ids equal to the oracle's, or differing only between entries whose fp64 distances are within
    rel (tensor-core vs BLAS summation order); distances within 1e-5 absolute (the north-star bound: scores within 1e-5 at fp32)."""
    oi, od = orc.exact_knn_candidates_blockwise(v, k, block=512)
    np.testing.assert_allclose(dist, od, rtol=0, atol=1e-5)
    bad = np.argwhere(idx != oi)
    if len(bad):
        v64 = v.astype(np.float64)
        for r, c in bad:
            da = 1.0 - v64[r] @ v64[idx[r, c]]
            db = 1.0 - v64[r] @ v64[oi[r, c]]
            assert abs(da - db) <= rel * max(abs(da), abs(db), 1e-3), (r, c, da, db)
    return len(bad)


@pytest.mark.parametrize("name", ["knn_600", "knn_small_n"])
def test_knn_vs_reference_golden(kg, golden, name):
    c = cases.KNN[name]
    v = cases.knn_inputs(c)
    df = kg.compute_exact_knn(v, c["k"])
    ref = _ref_df(golden, name)
    assert list(df.columns) == list(ref.columns) and [str(t) for t in df.dtypes] == [str(t) for t in ref.dtypes]
    assert len(df) == len(ref)
    np.testing.assert_allclose(df.distance.values, ref.distance.values, rtol=0, atol=1e-5)
    if not (df.dst_vertex.values == ref.dst_vertex.values).all():
        idx, dist = kg.knn_candidates(v, c["k"])
        check_candidates(idx, dist, v, c["k"])


def test_knn_duplicates_mod_ties(kg, golden):
    c = cases.KNN["knn_dups"]
    v = cases.knn_inputs(c)
    idx, dist = kg.knn_candidates(v, c["k"])
    check_candidates(idx, dist, v, c["k"])
    # duplicates: self is not always the first column, and exact ties break by ascending column
    assert (idx[:, 0] != np.arange(len(v))).any()


@pytest.mark.parametrize("n,dim", [(3000, 512), (1000, 256), (1500, 768), (129, 512), (128, 512)])
def test_knn_lattice_bit_exact(kg, n, dim):
    """Exact arithmetic, massive exact ties: ids and distances must be bit-identical to the oracle."""
    v = synth.synth_rows(0, n, dim, 21, "lattice", np.float32) * np.float32(0.25)
    idx, dist = kg.knn_candidates(v, 10)
    oi, od = orc.exact_knn_candidates_blockwise(v, 10, block=512)
    assert (dist == od).all()
    assert (idx == oi).all()
    df = kg.edges_from_candidates(idx, dist, n)
    pd.testing.assert_frame_equal(df, orc.post_process_graph(oi, od, n))


def test_knn_gaussian_and_row_ranges(kg):
    n = 5000
    v = synth.synth_rows(0, n, 512, 33, "tri", np.float32)
    v = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float16).astype(np.float32)
    idx, dist = kg.knn_candidates(v, 10)
    check_candidates(idx, dist, v, 10)
    # a row range (the multi-GPU sharding unit) reproduces the same rows; fp16 input == fp32 input
    i2, d2 = kg.knn_candidates(v.astype(np.float16), 10, rows=(1000, 1777))
    assert (i2 == idx[1000:1777]).all() and (d2 == dist[1000:1777]).all()
    g, _ = kg.KNNGraph.from_vectors(v[:700], n_neighbors=5)
    assert g.nvecs == 700 and g.k >= 5


@pytest.mark.parametrize("n,dup", [(3000, False), (1025, True), (5, False)])
def test_edge_table_on_device_equals_post_process_graph_df(kg, n, dup):
    """ssw_knn_graph (candidates + post_process_graph_df on the device) == oracle edge table, frame-equal:
    self edges, dropped self columns, ranks, clipping, order; duplicates make rows with k1 instead of k1-1 edges."""
    v = synth.synth_rows(0, n, 512, 27, "lattice", np.float32) * np.float32(0.25)
    if dup:
        v[1::7] = v[0::7][: len(v[1::7])]
    df = kg.compute_exact_knn(v, 10)
    want = orc.compute_exact_knn(v, 10)
    pd.testing.assert_frame_equal(df, want)
    g = kg.KNNGraph(df)
    assert g.nvecs == n and (np.diff(g.ind_ptr) >= 1).all()


@pytest.mark.parametrize("n,dim,k", [(2, 512, 10), (257, 256, 1), (700, 512, 40), (900, 256, 63), (1300, 768, 33),
                                     (513, 512, 32), (2049, 512, 5)])
def test_knn_variants_and_limits(kg, n, dim, k):
    """Every launch path: CTA pairs with N=256 (k1 <= 33 at dim 512), the TS pair kernel (larger k1), the single-CTA
    kernel (dim 768); k1 up to the limit of 64; n not a multiple of any tile — bit-exact on lattice data."""
    v = synth.synth_rows(0, n, dim, 35, "lattice", np.float32) * np.float32(0.25)
    idx, dist = kg.knn_candidates(v, k)
    oi, od = orc.exact_knn_candidates_blockwise(v, k, block=256)
    assert idx.shape == oi.shape
    assert (dist == od).all() and (idx == oi).all()
    with pytest.raises(Exception):
        kg.knn_candidates(synth.synth_rows(0, 100, dim, 1, "lattice", np.float32), 64)     # k1 = 65 > SSW_MAX_KNN_K1


def test_float32_vectors_give_the_float32_graph():
    """Vectors that are NOT fp16-representable: the tensor cores only propose candidates, distances and
    ranks come from the float32 rows — the neighbour lists equal the float32 oracle's (a swap admissible only between
    neighbours whose float64 distances differ by < 1e-5 relative), the distances agree to 1e-6; the plain fp16 build of
    the same vectors differs visibly, which is what the refinement is for."""
    import ctypes as C
    from seesaw_b200 import knn_graph as kg, synth
    from seesaw_b200._lib import SSW_F16, check, lib, ptr
    n, k = 6000, 10
    v = synth.unit_rows(n, 512, 71)
    assert not (v.astype(np.float16).astype(np.float32) == v).all()
    idx, dist = kg.knn_candidates(v, k)
    stats = kg.knn_exact_stats()
    assert stats["rows_refined"] == n and stats["rho"] > 0
    oi, od = orc.exact_knn_candidates(v, k)
    np.testing.assert_allclose(dist, od, rtol=0, atol=2e-6)
    d64 = 1.0 - v.astype(np.float64) @ v.astype(np.float64).T
    bad_rows = np.flatnonzero((idx != oi).any(axis=1))
    for r in bad_rows:
        for a, b in zip(idx[r], oi[r]):
            if a != b:
                assert abs(d64[r, a] - d64[r, b]) <= 1e-5 * max(abs(d64[r, a]), 1e-3), (r, a, b)
    print(f"float32 kNN: {len(bad_rows)} of {n} rows differ from the float32 oracle (near-ties), "
          f"{stats['rows_rescanned']} rows needed the full float32 re-scan")
    # the edge table through the reference-facing call
    df = kg.compute_exact_knn(v, k)
    want = orc.compute_exact_knn(v, k)
    assert len(df) == len(want) and (df.src_vertex.values == want.src_vertex.values).all()
    assert (df.dst_vertex.values != want.dst_vertex.values).sum() <= (k + 1) * len(bad_rows)
    # fp16 rounding alone: neighbour lists change
    i16 = np.empty((n, k + 1), np.int32)
    d16 = np.empty((n, k + 1), np.float32)
    v16 = v.astype(np.float16)
    check(lib.ssw_knn_build(0, ptr(v16), SSW_F16, n, 512, k + 1, 0, n, ptr(i16), ptr(d16)))
    assert (i16 != oi).any(axis=1).sum() > len(bad_rows)
    print(f"plain fp16 build of the same vectors: {(i16 != oi).any(axis=1).sum()} rows differ")


def test_float32_near_duplicates_fall_back_to_the_full_rescan():
    """Clusters tighter than the fp16 rounding error cannot be certified from fp16 candidates: those rows are re-scanned
    against all columns in float32 and still equal the float64 ranking up to near-ties."""
    from seesaw_b200 import knn_graph as kg, synth
    rng = np.random.default_rng(5)
    centres = synth.unit_rows(20, 512, 72)
    n = 1500
    v = (centres[rng.integers(0, 20, size=n)] + rng.standard_normal((n, 512)).astype(np.float32) * np.float32(3e-6)).astype(np.float32)
    v[7] = v[3]                                       # and one exact duplicate pair
    idx, dist = kg.knn_candidates(v, 10)
    stats = kg.knn_exact_stats()
    assert stats["rows_rescanned"] > 0
    oi, od = orc.exact_knn_candidates(v, 10)
    np.testing.assert_allclose(dist, od, rtol=0, atol=2e-6)
    d64 = 1.0 - v.astype(np.float64) @ v.astype(np.float64).T
    for r in np.flatnonzero((idx != oi).any(axis=1)):
        for a, b in zip(idx[r], oi[r]):
            if a != b:
                assert abs(d64[r, a] - d64[r, b]) <= 1e-6, (r, a, b)       # fp32 resolution of a 512-term dot near 1
    # rows 3 and 7 hold the same vector: identical distance rows, and column tie-breaking does not depend on the row
    assert (idx[3] == idx[7]).all() and (dist[3] == dist[7]).all()
