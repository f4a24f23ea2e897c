"""GPU parity tests of the reference-facing classes (seesaw_b200/indices.py) against the committed outputs
of the unmodified reference (tests/golden/) and the oracle: the calls a SeeSaw session makes."""
import json

import numpy as np
import pandas as pd
import pytest

import cases
import seesaw_oracle as orc
from seesaw_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ix():
    from seesaw_b200 import indices
    return indices


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("device_rescore", [True, False])
def test_multiscale_query_vs_reference_golden(ix, golden, name, device_rescore):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    ex = cases.exclude_sets(meta, c["seed"] + 7)["some"]
    for agg, topk, use_v2 in (("plain_score", 1, False), ("plain_score", 3, True), ("avg_score", 3, False)):
        key = f"{name}/query/{agg}/{topk}/{int(use_v2)}"
        if key + "/dbidxs" not in golden:
            continue
        r = idx.query(vector=qs[2], vector2=qs[3] * 0.25 if use_v2 else None, topk=topk, shortlist_size=50,
                      exclude=ix.BitMap(ex), agg_method=agg, aug_larger="all", rescore_method=None,
                      device_rescore=device_rescore)
        assert (np.asarray(r["dbidxs"]) == golden[key + "/dbidxs"]).all(), key
        sc = np.array([a.score.values[0] for a in r["activations"]], np.float64)
        np.testing.assert_allclose(sc, golden[key + "/act_score"], rtol=1e-5, atol=1e-6)
        box = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.int64)
        assert (box == golden[key + "/act_box"]).all(), key
        assert list(r["activations"][0].columns) == ["x1", "y1", "x2", "y2", "dbidx", "score"]
    idx.close()


def test_session_loop_and_interface(ix, tmp_path):
    """query_stateful accumulates `returned`, never repeats an image, matches the oracle step by step;
    plus the attribute / method surface callers read."""
    counts = synth.patches_per_image(900, 2, 20, 4)
    meta = synth.synth_vector_meta(counts, 5, dbidx_start=100, dbidx_stride=1)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 512, 6, "lattice", np.float32)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16", path=str(tmp_path))
    assert len(idx) == 900 and idx.vectors.shape == (n, 512) and idx.get_knng_path("x").endswith("/knn_graph/x")
    q = synth.lattice_queries(1, 512, 7)[0]
    assert (idx.score(q) == vecs @ q).all()
    iq = idx.new_query()
    seen = []
    for step in range(5):
        r = iq.query_stateful(vector=q, batch_size=3, shortlist_size=20)
        want = orc.multiscale_query(vecs, meta, q, 3, 20, exclude=np.array(seen, np.int64))
        assert (np.asarray(r["dbidxs"]) == want["dbidxs"]).all(), step
        assert not set(r["dbidxs"].tolist()) & set(seen)
        seen += r["dbidxs"].tolist()
    assert sorted(iq.returned) == sorted(seen)
    # subset: a new index over some images answers like the oracle on those rows
    keep = np.unique(meta.dbidx.values)[::3]
    sub = idx.subset(ix.BitMap(keep))
    mask = np.isin(meta.dbidx.values, keep)
    r = sub.query(vector=q, topk=4, shortlist_size=30, exclude=None)
    want = orc.multiscale_query(vecs[mask], meta[mask].reset_index(drop=True), q, 4, 30)
    assert (np.asarray(r["dbidxs"]) == want["dbidxs"]).all()
    sub.close()
    # everything excluded -> empty result instead of the reference's tuple crash
    r = idx.query(vector=q, topk=3, shortlist_size=10, exclude=ix.BitMap(np.unique(meta.dbidx.values)))
    assert len(r["dbidxs"]) == 0
    idx.close()


def test_constructor_registry_roundtrip(ix, tmp_path):
    """info.json -> AccessMethod.load -> from_path (seesaw/indices/interface.py:36-45) on a parquet index."""
    counts = synth.patches_per_image(200, 1, 6, 8)
    meta = synth.synth_vector_meta(counts, 9)
    vecs = synth.synth_rows(0, int(counts.sum()), 512, 10, "lattice", np.float32)
    d = tmp_path / "index"
    d.mkdir()
    meta.assign(vectors=list(vecs)).to_parquet(d / "vectors.sorted.cached")
    json.dump({"constructor": "seesaw_b200.indices.B200MultiscaleIndex", "model": "m", "dataset": "d"},
              open(d / "info.json", "w"))
    idx = ix.AccessMethod.load(str(d), options={}, exclude=None)
    q = synth.lattice_queries(1, 512, 11)[0]
    r = idx.query(vector=q, topk=5, shortlist_size=25)
    want = orc.multiscale_query(vecs, meta, q, 5, 25)
    assert (np.asarray(r["dbidxs"]) == want["dbidxs"]).all()
    idx.close()


def test_coarse_index_config1_golden_and_vector_index(ix, golden):
    c = cases.COARSE
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    cidx = ix.B200CoarseIndex(embedding=None, vectors=v, vector_meta=pd.DataFrame({"dbidx": np.arange(c["n"], dtype=np.int64)}))
    q = synth.unit_queries(1, c["dim"], c["qseed"])[0]
    ex = np.sort(np.random.default_rng(c["xseed"]).choice(c["n"], size=c["n_excl"], replace=False))
    r = cidx.query(topk=c["topk"], vector=q, exclude=ix.BitMap(ex))
    assert (np.asarray(r["dbidxs"]) == golden["coarse/dbidxs"]).all()
    assert r["nextstartk"] == int(golden["coarse/nextstartk"][0])
    np.testing.assert_allclose([a.score.values[0] for a in r["activations"]], golden["coarse/scores"], rtol=1e-5, atol=1e-6)
    assert r["activations"][0][["x1", "y1", "x2", "y2"]].values.tolist() == [[0, 0, 224, 224]]
    rr = cidx.query(topk=5, vector=None, exclude=ix.BitMap(ex))                 # random ranking branch (:70-71)
    assert len(rr["dbidxs"]) == 5 and not set(rr["dbidxs"].tolist()) & set(ex.tolist())
    empty = cidx.query(topk=5, vector=q, exclude=ix.BitMap(np.arange(c["n"])))   # :61-62
    assert isinstance(empty, tuple)
    cidx.close()
    # the ANN slot: exact, best first, same shape asserts as the reference wrapper
    vi = ix.B200VectorIndex(vectors=v)
    rows, scores = vi.query(q.reshape(1, -1), 20)
    o = np.argsort(-(v @ q), kind="stable")[:20]
    assert (rows == o).all()
    np.testing.assert_allclose(scores, (v @ q)[o], rtol=1e-5, atol=1e-6)
    with pytest.raises(AssertionError):
        vi.query(q[:100], 5)


def test_batcher_on_gpu(ix):
    """Concurrent sessions through the batching front end get exactly their own single-call results."""
    import threading
    from seesaw_b200.service import ScanBatcher
    counts = synth.patches_per_image(3000, 5, 30, 14)
    meta = synth.synth_vector_meta(counts, 15)
    vecs = synth.synth_rows(0, int(counts.sum()), 512, 16, "lattice", np.float32)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    qs = synth.lattice_queries(48, 512, 17)
    alone = [idx.query(vector=q, topk=3, shortlist_size=40, exclude=ix.BitMap(range(i, 3000, 11))) for i, q in enumerate(qs)]
    b = ScanBatcher(idx.db, max_batch=64, max_wait_s=0.05)
    idx.attach_batcher(b)
    got = [None] * len(qs)

    def session(i):
        got[i] = idx.query(vector=qs[i], topk=3, shortlist_size=40, exclude=ix.BitMap(range(i, 3000, 11)))

    th = [threading.Thread(target=session, args=(i,)) for i in range(len(qs))]
    [t.start() for t in th]
    [t.join() for t in th]
    b.close()
    assert b.batches_issued < len(qs)
    for a, g in zip(alone, got):
        assert (a["dbidxs"] == g["dbidxs"]).all()
    idx.attach_batcher(None)
    idx.close()


@pytest.mark.parametrize("aug", ["all", "greater", "adjacent"])
def test_device_stage2_avg_score_vs_oracle(ix, aug):
    """K7 (IoU join + per-zoom-level averaging on the device) against the oracle's score_frame2 restatement,
    all three aug_larger filters, images of 1..80 patches, exact lattice scores (ties between patches)."""
    counts = synth.patches_per_image(600, 1, 80, 21)
    meta = synth.synth_vector_meta(counts, 22, dbidx_start=7, dbidx_stride=2)
    vecs = synth.synth_rows(0, int(counts.sum()), 512, 23, "lattice", np.float32)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    qs = synth.lattice_queries(3, 512, 24)
    ex = np.unique(meta.dbidx.values)[::5]
    for q, v2 in ((qs[0], None), (qs[1], qs[2] * 0.5)):
        got = idx.query(vector=q, vector2=v2, topk=7, shortlist_size=40, exclude=ix.BitMap(ex), agg_method="avg_score",
                        aug_larger=aug)
        want = orc.multiscale_query(vecs, meta, q, 7, 40, exclude=ex, vector2=v2, agg_method="avg_score", aug_larger=aug)
        host = idx.query(vector=q, vector2=v2, topk=7, shortlist_size=40, exclude=ix.BitMap(ex), agg_method="avg_score",
                         aug_larger=aug, device_rescore=False)
        assert (np.asarray(got["dbidxs"]) == want["dbidxs"]).all() and (np.asarray(host["dbidxs"]) == want["dbidxs"]).all()
        gs = np.array([a.score.values[0] for a in got["activations"]])
        ws = np.array([a.score.values[0] for a in want["activations"]])
        np.testing.assert_allclose(gs, ws, rtol=1e-6, atol=1e-7)
        for a, b in zip(got["activations"], want["activations"]):
            assert (a[["x1", "y1", "x2", "y2"]].values == b[["x1", "y1", "x2", "y2"]].values).all()
    idx.close()


def test_serving_only_index_from_database(ix):
    """from_database (vectors generated in HBM, no host copy) answers like the host-constructed index."""
    from seesaw_b200.engine import PatchDatabase
    counts = synth.patches_per_image(800, 3, 30, 31)
    meta = synth.synth_vector_meta(counts, 32)
    n = int(counts.sum())
    db = PatchDatabase.synthetic(meta.dbidx.to_numpy().astype(np.int32), 512, seed=33, kind="lattice", store="f16")
    served = ix.B200MultiscaleIndex.from_database(db, meta)
    vecs = synth.synth_rows(0, n, 512, 33, "lattice", np.float32)
    q = synth.lattice_queries(1, 512, 34)[0]
    ex = np.unique(meta.dbidx.values)[::6]
    for agg in ("plain_score", "avg_score"):
        got = served.query(vector=q, topk=5, shortlist_size=30, exclude=ix.BitMap(ex), agg_method=agg)
        want = orc.multiscale_query(vecs, meta, q, 5, 30, exclude=ex, agg_method=agg)
        assert (np.asarray(got["dbidxs"]) == want["dbidxs"]).all()
    assert served.vectors is None and len(served) == 800
    assert (served.score(q) == vecs @ q).all()
    served.close()


def test_subset_as_device_mask(ix):
    """subset(share_device=True): no second copy in HBM, other images masked out of every scan; results equal the
    copying subset and the oracle on the masked data, rows renumbered like the reference's subset."""
    counts = synth.patches_per_image(1200, 2, 25, 41)
    meta = synth.synth_vector_meta(counts, 42, dbidx_start=3, dbidx_stride=3)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 512, 43, "lattice", np.float32)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    keep = np.unique(meta.dbidx.values)[1::4]
    view = idx.subset(ix.BitMap(keep), share_device=True)
    copy = idx.subset(ix.BitMap(keep))
    assert view.db is idx.db and copy.db is not idx.db
    mask = np.isin(meta.dbidx.values, keep)
    sub_v, sub_m = vecs[mask], meta[mask].reset_index(drop=True)
    assert len(view) == len(keep) and (view.vectors == sub_v).all() and view.vector_meta.equals(sub_m)
    qs = synth.lattice_queries(3, 512, 44)
    seen = keep[::9]
    for q in qs:
        assert (view.score(q) == sub_v @ q).all()
        for agg in ("plain_score", "avg_score"):
            a = view.query(vector=q, topk=4, shortlist_size=25, exclude=ix.BitMap(seen), agg_method=agg)
            b = copy.query(vector=q, topk=4, shortlist_size=25, exclude=ix.BitMap(seen), agg_method=agg)
            w = orc.multiscale_query(sub_v, sub_m, q, 4, 25, exclude=seen, agg_method=agg)
            assert (np.asarray(a["dbidxs"]) == w["dbidxs"]).all() and (np.asarray(b["dbidxs"]) == w["dbidxs"]).all()
            for x, y in zip(a["activations"], w["activations"]):
                assert (x[["x1", "y1", "x2", "y2"]].values == y[["x1", "y1", "x2", "y2"]].values).all()
        p = view._query_prelim(vector=q, topk_dbidx=30, exclude_dbidx=ix.BitMap(seen))
        o = orc.query_prelim(sub_v, sub_m.dbidx.values, q, 30, exclude=seen)
        assert (p.dbidx.values == o["dbidx"]).all() and (p.best_row.values == o["best_row"]).all()
    # everything of the subset excluded -> empty; a subset of the subset
    assert len(view.query(vector=qs[0], topk=3, shortlist_size=10, exclude=ix.BitMap(keep))["dbidxs"]) == 0
    inner = view.subset(ix.BitMap(keep[::2]))
    w = orc.multiscale_query(vecs[np.isin(meta.dbidx.values, keep[::2])],
                             meta[np.isin(meta.dbidx.values, keep[::2])].reset_index(drop=True), qs[1], 3, 20)
    assert (np.asarray(inner.query(vector=qs[1], topk=3, shortlist_size=20)["dbidxs"]) == w["dbidxs"]).all()
    view.close()
    copy.close()
    assert idx.query(vector=qs[0], topk=2, shortlist_size=10)["dbidxs"].shape == (2,)      # parent still alive
    idx.close()
