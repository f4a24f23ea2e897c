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


@pytest.mark.parametrize("name", list(cases.CASES_F32))
def test_float32_index_matches_reference_golden(golden, name):
    """The data a real index holds — float32 unit vectors that are not fp16-representable and the tiling
    pipeline's float32 boxes: stage 1 and every stage-2 variant (float32 IoU self-join, per-level idxmax,
    float64-accumulated mean handed back as float32) equal the unmodified reference's outputs bit for bit."""
    c = cases.CASES_F32[name]
    vecs, meta, qs = cases.msf_inputs(c)
    assert vecs.dtype == np.float32 and not (vecs.astype(np.float16).astype(np.float32) == vecs).all()
    assert all(meta[col].dtype == np.float32 for col in ("x1", "y1", "x2", "y2")) and meta.zoom_level.dtype == np.int16
    xs = cases.exclude_sets(meta, c["seed"] + 7)
    for xname in ("none", "some"):
        for qi in range(2):
            r = orc.query_prelim(vecs, meta.dbidx.values, qs[qi], 50, exclude=xs[xname])
            key = f"{name}/prelim/{xname}/{qi}"
            assert (r["dbidx"] == golden[key + "/dbidx"]).all() and (r["max_score"] == golden[key + "/score"]).all(), key
    for agg, aug, topk, use_v2 in cases.MSF_QUERY_VARIANTS:
        key = f"{name}/query/{agg}/{aug}/{topk}/{int(use_v2)}"
        r = orc.multiscale_query(vecs, meta, qs[2], topk, 40, exclude=xs["some"], vector2=qs[3] * 0.25 if use_v2 else None,
                                 agg_method=agg, aug_larger=aug)
        assert (r["dbidxs"] == golden[key + "/dbidxs"]).all(), key
        sc = np.array([a.score.values[0] for a in r["activations"]], np.float64)
        assert (sc == golden[key + "/act_score"]).all(), key            # same arithmetic, same dtype: bit-equal
        bx = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.float32)
        assert (bx == golden[key + "/act_box"]).all(), key


def test_pyramid_tiling_equals_reference_pipeline(reference):
    """seesaw_b200.synth.pyramid_tiling (the float32 box generator of the fixtures) against the reference's
    generate_multiscale_tiling (multiscale_tools.py:96-117) on blank images of the same sizes."""
    import importlib
    import PIL.Image
    from seesaw_b200 import synth
    mt = importlib.import_module("seesaw.indices.multiscale.multiscale_tools")
    for (w, h) in [(640, 480), (333, 500), (1280, 960), (224, 224), (150, 100)]:
        for mts in (60, 224):
            df = mt.generate_multiscale_tiling(PIL.Image.new("RGB", (w, h)), tile_size=224, factor=.5, min_tile_size=mts)
            t = synth.pyramid_tiling(w, h, min_tile_size=mts)
            assert len(df) == len(t["x1"])
            for col in ("x1", "y1", "x2", "y2", "zoom_level"):
                assert df[col].dtype == t[col].dtype and (df[col].values == t[col]).all(), (w, h, mts, col)


def test_integer_box_iou_is_float32_quotient(reference):
    """Integer box columns: torchvision keeps intersection / union as int64 and ``inter / union`` is torch's
    float32 true division (box_utils.py:341-345) — the oracle's and the host mirror's IoU equal it bit for bit."""
    import importlib
    from seesaw_b200 import rescore
    bu = importlib.import_module("seesaw.box_utils")
    rng = np.random.default_rng(3)
    x1, y1 = rng.integers(0, 900, 40), rng.integers(0, 900, 40)
    df = pd.DataFrame({"x1": x1, "y1": y1, "x2": x1 + rng.integers(1, 700, 40), "y2": y1 + rng.integers(1, 700, 40)})
    want = bu.box_iou(df, df)
    assert want.dtype == np.float32
    b = df[["x1", "y1", "x2", "y2"]].to_numpy()
    assert (orc._pairwise_iou(b) == want).all() and (rescore._iou_matrix(b) == want).all()
    f = df.astype(np.float32) / np.float32(1.7)
    want = bu.box_iou(f, f)
    b = f[["x1", "y1", "x2", "y2"]].to_numpy()
    assert want.dtype == np.float32 and (orc._pairwise_iou(b) == want).all() and (rescore._iou_matrix(b) == want).all()


def test_pandas_float32_group_mean_is_kahan_float32():
    """The arithmetic score_frame2's ``groupby('iloc_left').score_right.mean()`` performs on the float32 score
    column (multiscale_index.py:142), pinned against pandas itself: float32 Kahan sum in row order / float32
    count.  (A float64 accumulation rounded to float32 differs on ~5% of random groups.)"""
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 6, size=3000)
    g = np.repeat(np.arange(len(sizes)), sizes)
    s = (rng.standard_normal(len(g)) * 0.1).astype(np.float32)
    want = pd.DataFrame({"g": g, "s": s}).groupby("g").s.mean().values
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    got = np.array([orc._pandas_group_mean_f32(s[a:a + k]) for a, k in zip(starts, sizes)], np.float32)
    assert want.dtype == np.float32 and (got == want).all()
