"""Pins the CPU oracle (oracle/seesaw_oracle.py) against (a) the reference's own unit pins,
(b) the committed outputs of the unmodified reference (tests/golden/, made by
oracle/make_golden.py) and (c) the live reference when /root/reference is present."""
import numpy as np
import pandas as pd
import pytest

import cases
import seesaw_oracle as orc


def test_reference_pin_distinct_topk_positions(golden):
    # the reference's own unit test, multiscale_index.py:182-187
    ex = np.array([10, 11, 11, 12, 12, 12, 13, 13])
    assert (orc.distinct_topk_positions(ex, 2) == np.array([0, 1])).all()
    assert (orc.distinct_topk_positions(ex, 10) == np.array([0, 1, 3, 6])).all()
    assert (golden["pin/distinct_topk_positions"] == np.array([0, 1])).all()


def test_reference_pin_edge_table_schema():
    # knn_graph.py:109-134 pins the 4-column edge table; post_process_graph must produce it
    idx = np.array([[0, 1], [1, 0]], np.int32)
    dist = np.array([[0., 1.], [0., 1.]], np.float32)
    df = orc.post_process_graph(idx, dist, 2)
    assert list(df.columns) == ["src_vertex", "dst_vertex", "distance", "dst_rank"]
    assert df.src_vertex.tolist() == [0, 0, 1, 1] and df.dst_vertex.tolist() == [0, 1, 1, 0]
    assert df.dst_rank.tolist() == [0, 1, 0, 1] and df.distance.tolist() == [0., 1., 0., 1.]
    assert str(df.src_vertex.dtype) == "int32" and str(df.distance.dtype) == "float32"


@pytest.mark.parametrize("name", list(cases.CASES))
def test_prelim_matches_reference_golden(golden, name):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)
    dbidx = meta.dbidx.values
    for xname, ex in cases.exclude_sets(meta, c["seed"] + 7).items():
        for qi in range(2):
            r = orc.query_prelim(vecs, dbidx, qs[qi], 50, exclude=ex)
            key = f"{name}/prelim/{xname}/{qi}"
            assert (r["dbidx"] == golden[key + "/dbidx"]).all(), key
            assert (r["max_score"] == golden[key + "/score"]).all(), key   # same BLAS call: bit-equal
            # direct (sort-free) statement agrees with the sort-based one
            d2, s2, r2 = orc.per_image_best(vecs @ qs[qi], dbidx, ex)
            k = len(r["dbidx"])
            assert (d2[:k] == r["dbidx"]).all() and (r2[:k] == r["best_row"]).all()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_query_matches_reference_golden(golden, name):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)
    ex = cases.exclude_sets(meta, c["seed"] + 7)["some"]
    for agg, topk, use_v2 in (("plain_score", 1, False), ("plain_score", 3, True), ("avg_score", 3, False)):
        key = f"{name}/query/{agg}/{topk}/{int(use_v2)}"
        if key + "/dbidxs" not in golden:
            continue
        r = orc.multiscale_query(vecs, meta, qs[2], topk, 50, exclude=ex,
                                 vector2=qs[3] * 0.25 if use_v2 else None, agg_method=agg)
        assert (r["dbidxs"] == golden[key + "/dbidxs"]).all(), key
        sc = np.array([a.score.values[0] for a in r["activations"]])
        np.testing.assert_allclose(sc, golden[key + "/act_score"], rtol=1e-6, atol=1e-7)
        bx = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]])
        assert (bx == golden[key + "/act_box"]).all(), key


def test_coarse_matches_reference_golden(golden):
    c = cases.COARSE
    from seesaw_b200 import synth
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    q = synth.unit_queries(1, c["dim"], c["qseed"])[0]
    ex = np.sort(np.random.default_rng(c["xseed"]).choice(c["n"], size=c["n_excl"], replace=False))
    r = orc.coarse_query(v, np.arange(c["n"]), q, c["topk"], exclude=ex)
    assert (r["dbidxs"] == golden["coarse/dbidxs"]).all()
    assert (r["scores"].astype(np.float32) == golden["coarse/scores"]).all()
    assert r["nextstartk"] == int(golden["coarse/nextstartk"][0])
    assert orc.coarse_query(v, np.arange(c["n"]), q, 10, exclude=np.arange(c["n"])) is None


def assert_graph_equal_mod_ties(mine, ref):
    """Edge tables equal up to the order among exactly tied distances of one source; a
    neighbour may differ only if it is tied with the source's cut-off (largest) distance."""
    assert len(mine) == len(ref)
    assert (mine.src_vertex.values == ref.src_vertex.values).all()
    assert (mine.dst_rank.values == ref.dst_rank.values).all()
    assert (mine.distance.values == ref.distance.values).all()
    diff = np.flatnonzero(mine.dst_vertex.values != ref.dst_vertex.values)
    for src in np.unique(mine.src_vertex.values[diff]):
        a, b = mine[mine.src_vertex == src], ref[ref.src_vertex == src]
        cut = a.distance.max()
        for d in np.unique(a.distance.values):
            sa, sb = set(a.dst_vertex[a.distance == d]), set(b.dst_vertex[b.distance == d])
            assert sa == sb or d == cut, (src, d, sa, sb)


@pytest.mark.parametrize("name", list(cases.KNN))
def test_knn_matches_reference_golden(golden, name):
    c = cases.KNN[name]
    v = cases.knn_inputs(c)
    df = orc.compute_exact_knn(v, c["k"])
    ref = pd.DataFrame({col: golden[f"{name}/{col}"] for col in ("src_vertex", "dst_vertex", "distance", "dst_rank")})
    if c.get("dup"):
        # exact duplicate vectors give exactly tied distances; the reference's unstable argsort
        # leaves their order undefined, so compare modulo ties
        assert_graph_equal_mod_ties(df, ref)
    else:
        pd.testing.assert_frame_equal(df, ref)
    # blockwise restatement == full restatement
    i1, d1 = orc.exact_knn_candidates(v, c["k"])
    i2, d2 = orc.exact_knn_candidates_blockwise(v, c["k"], block=64)
    assert (i1 == i2).all() and (d1 == d2).all()


def test_live_reference_agrees_on_fresh_seed(reference):
    """Not a fixture replay: a seed the goldens never saw, reference vs oracle, same process."""
    c = dict(n_images=250, p_lo=2, p_hi=12, dim=512, seed=77, qseed=78)
    vecs, meta, qs = cases.ms_inputs(c)
    idx = reference.multiscale.MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, vec_index=None)
    ex = cases.exclude_sets(meta, 5)["some"]
    r = idx._query_prelim(vector=qs[0], topk_dbidx=40, exclude_dbidx=reference.BitMap(ex))
    o = orc.query_prelim(vecs, meta.dbidx.values, qs[0], 40, exclude=ex)
    assert (r["dbidx"].values == o["dbidx"]).all() and (r["max_score"].values == o["max_score"]).all()
    full = idx.query(vector=qs[1], topk=3, shortlist_size=40, exclude=reference.BitMap(ex),
                     agg_method="avg_score", aug_larger="greater", rescore_method=None)
    mine = orc.multiscale_query(vecs, meta, qs[1], 3, 40, exclude=ex, agg_method="avg_score", aug_larger="greater")
    assert (np.asarray(full["dbidxs"]) == mine["dbidxs"]).all()
    v = cases.knn_inputs(dict(n=333, dim=512, seed=79, k=7))
    a = reference.knn_graph.compute_exact_knn(v, n_neighbors=7)
    b = orc.compute_exact_knn(v, 7)
    pd.testing.assert_frame_equal(a, b)


def test_lattice_data_is_exact_in_any_order():
    """Hard part A: lattice dot products are exact in fp32, so fp32 == fp64 and a
    permuted summation order gives identical bits; ties are plentiful."""
    from seesaw_b200 import synth
    v = synth.synth_rows(0, 4000, 512, 3, "lattice", np.float32)
    q = synth.lattice_queries(1, 512, 4)[0]
    s32 = v @ q
    s64 = v.astype(np.float64) @ q.astype(np.float64)
    assert (s32.astype(np.float64) == s64).all()
    perm = np.random.default_rng(0).permutation(512)
    assert ((v[:, perm] @ q[perm]) == s32).all()
    assert len(np.unique(s32)) < len(s32) // 4


@pytest.mark.parametrize("name", list(cases.LP))
def test_label_propagation_matches_reference_golden(golden, name):
    """oracle.label_propagation_fit against the reference's LabelPropagation.fit_transform on the reference's
    own weight matrix (stored in the fixture): bit-identical float64."""
    import scipy.sparse as sp
    c = cases.LP[name]
    W = sp.csr_array((golden[f"{name}/W_data"], golden[f"{name}/W_indices"], golden[f"{name}/W_indptr"]),
                     shape=(c["n"], c["n"]))
    ids, vals, reg, start = cases.lp_inputs(c)
    got, it, conv = orc.label_propagation_fit(W, reg_lambda=c["reg_lambda"], max_iter=c["max_iter"], epsilon=c["epsilon"],
                                              label_ids=ids, label_values=vals, reg_values=reg, start_value=start)
    assert (got == golden[f"{name}/values"]).all()
    assert (got[ids[-1]] == 1.0) and it >= 1
    assert conv == (name != "lp_noreg")          # the fixture holds one run that hits max_iter
