"""CPU tests of the reference-facing classes with the oracle standing in for the device database (tests/fake_db.py):
the host logic — clamps, exclusion bookkeeping, ordering rules, row renumbering of subsets, result shapes — is
exercised here on every CPU run; the same scenarios run against the real kernels in tests/test_indices_gpu.py."""
import numpy as np
import pandas as pd
import pytest

import cases
import seesaw_oracle as orc
from fake_db import FakeDB
from seesaw_b200 import synth


@pytest.fixture()
def ix(monkeypatch):
    from seesaw_b200 import indices
    monkeypatch.setattr(indices, "PatchDatabase", FakeDB)
    return indices


@pytest.mark.parametrize("name", ["ms_small", "ms_768"])
def test_query_vs_reference_golden_host_logic(ix, golden, name):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta)
    ex = cases.exclude_sets(meta, c["seed"] + 7)["some"]
    for device_rescore in (True, False):
        for agg, topk, use_v2 in (("plain_score", 1, False), ("plain_score", 3, True), ("avg_score", 3, False)):
            key = f"{name}/query/{agg}/{topk}/{int(use_v2)}"
            r = idx.query(vector=qs[2], vector2=qs[3] * 0.25 if use_v2 else None, topk=topk, shortlist_size=50,
                          exclude=ix.BitMap(ex), agg_method=agg, aug_larger="all", rescore_method=None,
                          device_rescore=device_rescore)
            assert (np.asarray(r["dbidxs"]) == golden[key + "/dbidxs"]).all(), (key, device_rescore)
            box = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.int64)
            assert (box == golden[key + "/act_box"]).all()
    for xname, ex in cases.exclude_sets(meta, c["seed"] + 7).items():
        p = idx._query_prelim(vector=qs[0], topk_dbidx=50, exclude_dbidx=ix.BitMap(ex))
        assert (p.dbidx.values == golden[f"{name}/prelim/{xname}/0/dbidx"]).all(), xname
    idx.close()
    assert idx.db.closed


def test_session_subsets_and_score_arrays_host_logic(ix):
    counts = synth.patches_per_image(300, 1, 14, 4)
    meta = synth.synth_vector_meta(counts, 5, dbidx_start=100, dbidx_stride=2)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 256, 6, "lattice", np.float32)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, excluded=[100, 102])
    assert len(idx) == 298                                                   # index-level excluded ids shrink all_indices
    q = synth.lattice_queries(1, 256, 7)[0]
    iq = idx.new_query()
    seen = []
    for _ in range(4):
        r = iq.query_stateful(vector=q, batch_size=3, shortlist_size=20, agg_method="avg_score")
        want = orc.multiscale_query(vecs, meta, q, 3, 20, exclude=np.array(seen, np.int64), agg_method="avg_score",
                                    index_excluded=[100, 102])
        assert (np.asarray(r["dbidxs"]) == want["dbidxs"]).all()
        seen += r["dbidxs"].tolist()
    # the clamp counts only images that exist and are eligible (foreign ids in the exclude set do not count)
    p = idx._query_prelim(vector=q, topk_dbidx=10 ** 6, exclude_dbidx=ix.BitMap([100, 104, 10 ** 7]))
    assert len(p) == 297            # 300 images - 2 index-level excluded - 1 excluded here (100 is both, 10**7 unknown)
    # device-mask subset and copying subset agree with the oracle on the masked rows, rows renumbered
    keep = np.unique(meta.dbidx.values)[::3]
    mask = np.isin(meta.dbidx.values, keep)
    sub_v, sub_m = vecs[mask], meta[mask].reset_index(drop=True)
    for share in (True, False):
        sub = idx.subset(ix.BitMap(keep), share_device=share)
        assert (sub.db is idx.db) == share
        pr = sub._query_prelim(vector=q, topk_dbidx=25, exclude_dbidx=ix.BitMap(keep[:4]))
        o = orc.query_prelim(sub_v, sub_m.dbidx.values, q, 25, exclude=keep[:4])
        assert (pr.dbidx.values == o["dbidx"]).all() and (pr.best_row.values == o["best_row"]).all(), share
        assert (sub.score(q) == sub_v @ q).all()
        r = sub.query(vector=q, topk=5, shortlist_size=25, agg_method="avg_score")
        w = orc.multiscale_query(sub_v, sub_m, q, 5, 25, agg_method="avg_score")
        assert (np.asarray(r["dbidxs"]) == w["dbidxs"]).all()
    # scores that are not a dot product (KnnProp2.next_batch)
    rng = np.random.default_rng(8)
    scores = (rng.integers(-9, 10, size=n) / 4.0).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False))
    order = rows[np.argsort(-scores[rows], kind="stable")]
    got = idx.top_dbidxs(vec_idxs=order, scores=scores[order], exclude=keep[:9], topk=30)
    d, s, r = orc.get_top_dbidxs(order, scores[order], meta.dbidx.values, keep[:9], 30)
    assert (got.dbidx.values == d).all() and (got.best_row.values == r).all()
    view = idx.subset(ix.BitMap(keep), share_device=True)
    sub_scores = scores[mask]
    sub_order = np.argsort(-sub_scores, kind="stable")
    got = view.top_dbidxs(vec_idxs=sub_order, scores=sub_scores[sub_order], exclude=None, topk=12)
    d, s, r = orc.get_top_dbidxs(sub_order, sub_scores[sub_order], sub_m.dbidx.values, None, 12)
    assert (got.dbidx.values == d).all() and (got.best_row.values == r).all()


def test_coarse_and_vector_index_host_logic(ix, golden):
    c = cases.COARSE
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    cidx = ix.B200CoarseIndex(embedding=None, vectors=v, vector_meta=pd.DataFrame({"dbidx": np.arange(c["n"], dtype=np.int64)}))
    q = synth.unit_queries(1, c["dim"], c["qseed"])[0]
    ex = np.sort(np.random.default_rng(c["xseed"]).choice(c["n"], size=c["n_excl"], replace=False))
    r = cidx.query(topk=c["topk"], vector=q, exclude=ix.BitMap(ex))
    assert (np.asarray(r["dbidxs"]) == golden["coarse/dbidxs"]).all() and r["nextstartk"] == int(golden["coarse/nextstartk"][0])
    assert isinstance(cidx.query(topk=3, vector=q, exclude=ix.BitMap(np.arange(c["n"]))), tuple)
    assert len(cidx.query(topk=10 ** 6, vector=q, exclude=ix.BitMap(ex))["dbidxs"]) == c["n"] - c["n_excl"]
    sub = cidx.subset(ix.BitMap(np.arange(0, c["n"], 50)))
    assert len(sub) == c["n"] // 50
    vi = ix.B200VectorIndex(vectors=v)
    rows, scores = vi.query(q, 7)
    assert (rows == np.argsort(-(v @ q), kind="stable")[:7]).all()
