"""GPU parity tests of the stage-1 scan (K1 streaming kernel, K4 merge) through the C ABI,
against the CPU oracle and the committed reference outputs."""
import numpy as np
import pytest

import cases
import seesaw_oracle as orc
from seesaw_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from seesaw_b200 import engine
    return engine


def admissible(res_dbidx, res_score, vecs, dbidx_of_row, q, exclude, k, rel=1e-5):
    """ids must equal the fp32 oracle's, or differ only where the fp64 scores of the swapped
    entries are within rel (summation-order noise, SURVEY.md §7 hard part B)."""
    o = orc.query_prelim(vecs, dbidx_of_row, q, k, exclude=exclude)
    kk = len(o["dbidx"])
    assert len(res_dbidx) == kk
    np.testing.assert_allclose(res_score, o["max_score"], rtol=1e-5, atol=1e-6)
    if (res_dbidx == o["dbidx"]).all():
        return 0
    s64 = orc.scores_f64(vecs, q)
    d, s, _ = orc.per_image_best(s64, dbidx_of_row, exclude)
    best = dict(zip(d.tolist(), s.tolist()))
    bad = 0
    for a, b in zip(res_dbidx.tolist(), o["dbidx"].tolist()):
        if a != b:
            assert abs(best[a] - best[b]) <= rel * max(abs(best[a]), abs(best[b])), (a, b, best[a], best[b])
            bad += 1
    return bad


@pytest.mark.parametrize("store", ["f32", "f16"])
@pytest.mark.parametrize("name", list(cases.CASES))
def test_prelim_vs_reference_golden(eng, golden, name, store):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)          # fp16-exact values: f16 storage is lossless here
    dbidx = meta.dbidx.values
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store=store)
    assert db.n_rows == len(vecs) and db.n_images == len(np.unique(dbidx))
    for xname, ex in cases.exclude_sets(meta, c["seed"] + 7).items():
        r = db.scan_topk(qs[:2], 50, exclude=[ex, ex])
        for qi in range(2):
            key = f"{name}/prelim/{xname}/{qi}"
            n = r["count"][qi]
            gd, gs = golden[key + "/dbidx"], golden[key + "/score"]
            assert n == len(gd), key
            np.testing.assert_allclose(r["score"][qi, :n], gs, rtol=1e-5, atol=1e-6)
            if not (r["dbidx"][qi, :n] == gd).all():
                admissible(r["dbidx"][qi, :n], r["score"][qi, :n], vecs, dbidx, qs[qi], ex, 50)
            assert (r["dbidx"][qi, n:] == -1).all() and (r["row"][qi, n:] == -1).all()
            # returned row is the best row of that image
            rows = r["row"][qi, :n]
            assert (dbidx[rows] == r["dbidx"][qi, :n]).all()
    db.close()


@pytest.mark.parametrize("store", ["f32", "f16"])
def test_lattice_bit_exact_with_ties(eng, store):
    """Exact arithmetic + massive ties: ids, rows and scores must be bit-identical to the oracle."""
    counts = synth.patches_per_image(3000, 1, 30, 5)
    dbidx = synth.dbidx_of_rows(counts, 100, 2)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 512, 9, "lattice", np.float32)
    qs = synth.lattice_queries(3, 512, 10)
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store=store)
    ex = np.unique(dbidx)[::7]
    for k in (1, 10, 50, 333):
        r = db.scan_topk(qs, k, exclude=[ex, None, ex[:3]])
        for qi, e in enumerate([ex, None, ex[:3]]):
            o = orc.query_prelim(vecs, dbidx, qs[qi], k, exclude=e)
            kk = len(o["dbidx"])
            assert r["count"][qi] == kk
            assert (r["dbidx"][qi, :kk] == o["dbidx"]).all()
            assert (r["row"][qi, :kk] == o["best_row"]).all()
            assert (r["score"][qi, :kk] == o["max_score"]).all()
    db.close()


def test_synthetic_db_matches_host_generator(eng):
    counts = synth.patches_per_image(500, 2, 9, 1)
    dbidx = synth.dbidx_of_rows(counts)
    n = int(counts.sum())
    for kind in ("tri", "lattice"):
        db = eng.PatchDatabase.synthetic(dbidx, 512, seed=77, kind=kind, store="f16", global_row_base=1000)
        host = synth.synth_rows(1000, n, 512, 77, kind, np.float32)
        q = synth.unit_queries(1, 512, 3)[0] if kind == "tri" else synth.lattice_queries(1, 512, 3)[0]
        s = db.score_all(q)
        np.testing.assert_allclose(s, host @ q, rtol=1e-5, atol=1e-6)
        if kind == "lattice":
            assert (s == host @ q).all()
        r = db.scan_topk(q, 20)
        o = orc.query_prelim(host, dbidx, q, 20)
        assert (r["row"][0] == o["best_row"] + 1000).all() or kind == "tri"
        db.close()


def test_unsorted_rows_and_row_mapping(eng):
    """Rows not grouped by image (multiscale_index.py:255 has the order assert commented out)."""
    rng = np.random.default_rng(0)
    counts = synth.patches_per_image(700, 1, 12, 2)
    dbidx = synth.dbidx_of_rows(counts, 3, 5)
    n = int(counts.sum())
    perm = rng.permutation(n)
    vecs = synth.synth_rows(0, n, 512, 13, "lattice", np.float32)[perm]
    dbidx = dbidx[perm]
    q = synth.lattice_queries(1, 512, 14)[0]
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
    r = db.scan_topk(q, 40, exclude=[np.unique(dbidx)[:50]])
    o = orc.query_prelim(vecs, dbidx, q, 40, exclude=np.unique(dbidx)[:50])
    assert (r["dbidx"][0] == o["dbidx"]).all() and (r["row"][0] == o["best_row"]).all()
    assert (db.score_all(q) == vecs @ q).all()
    db.close()


def test_edge_cases(eng):
    vecs = synth.synth_rows(0, 5, 512, 1, "lattice", np.float32)
    q = synth.lattice_queries(1, 512, 2)[0]
    # single image, k larger than the database, everything excluded
    db = eng.PatchDatabase.from_arrays(vecs, np.zeros(5, np.int32), store="f32")
    r = db.scan_topk(q, 10)
    assert r["count"][0] == 1 and r["dbidx"][0, 0] == 0 and r["score"][0, 0] == (vecs @ q).max()
    r = db.scan_topk(q, 10, exclude=[[0]])
    assert r["count"][0] == 0 and (r["dbidx"][0] == -1).all() and np.isinf(r["score"][0]).all()
    db.close()
    # one row per image (coarse layout), dims 256/768/1024
    for dim in (256, 768, 1024):
        v = synth.synth_rows(0, 2000, dim, 3, "lattice", np.float32)
        qq = synth.lattice_queries(1, dim, 4)[0]
        for store in ("f16", "f32"):
            db = eng.PatchDatabase.from_arrays(v, np.arange(2000), store=store)
            r = db.scan_topk(qq, 10, exclude=[np.arange(0, 2000, 3)])
            o = orc.coarse_query(v, np.arange(2000), qq, 10, exclude=np.arange(0, 2000, 3))
            assert (r["dbidx"][0] == o["dbidxs"]).all() and (r["score"][0] == o["scores"]).all()
            db.close()
    with pytest.raises(Exception):
        eng.PatchDatabase.from_arrays(np.zeros((4, 100), np.float32), np.arange(4))
    db = eng.PatchDatabase.from_arrays(vecs, np.arange(5), store="f32")
    with pytest.raises(Exception):
        db.scan_topk(q, 0)
    with pytest.raises(Exception):
        db.scan_topk(q, 100000)
    db.close()


def test_config1_coarse_golden(eng, golden):
    c = cases.COARSE
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    q = synth.unit_queries(1, c["dim"], c["qseed"])[0]
    ex = np.sort(np.random.default_rng(c["xseed"]).choice(c["n"], size=c["n_excl"], replace=False))
    db = eng.PatchDatabase.from_arrays(v, np.arange(c["n"]), store="f16")
    r = db.scan_topk(q, c["topk"], exclude=[ex])
    assert (r["dbidx"][0] == golden["coarse/dbidxs"]).all()
    np.testing.assert_allclose(r["score"][0], golden["coarse/scores"], rtol=1e-5, atol=1e-6)
    db.close()


# ------------------------------------------------------------------ K2: tcgen05 batched scan
def test_batched_tc_lattice_bit_exact(eng):
    """64-query batch on the tensor-core kernel: exact arithmetic + ties, per-query exclude sets."""
    counts = synth.patches_per_image(4000, 1, 40, 6)
    dbidx = synth.dbidx_of_rows(counts, 7, 3)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 512, 19, "lattice", np.float32)
    qs = synth.lattice_queries(64, 512, 20)
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
    db.set_scan_mode(2)
    ids = np.unique(dbidx)
    rng = np.random.default_rng(1)
    ex = [rng.choice(ids, size=s, replace=False) for s in rng.choice([0, 5, 50, 500], size=64)]
    for k in (50, 7):
        r = db.scan_topk(qs, k, exclude=ex)
        for qi in range(64):
            o = orc.query_prelim(vecs, dbidx, qs[qi], k, exclude=ex[qi])
            kk = len(o["dbidx"])
            assert r["count"][qi] == kk
            assert (r["dbidx"][qi, :kk] == o["dbidx"]).all(), (k, qi)
            assert (r["row"][qi, :kk] == o["best_row"]).all(), (k, qi)
            assert (r["score"][qi, :kk] == o["max_score"]).all(), (k, qi)
    # partial batch, and more than 64 queries (two passes)
    for nq in (3, 100):
        qq = synth.lattice_queries(nq, 512, 30 + nq)
        r = db.scan_topk(qq, 10)
        for qi in range(nq):
            o = orc.query_prelim(vecs, dbidx, qq[qi], 10)
            assert (r["dbidx"][qi] == o["dbidx"]).all() and (r["row"][qi] == o["best_row"]).all()
    db.close()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_batched_tc_vs_reference_golden(eng, golden, name):
    c = cases.CASES[name]
    vecs, meta, qs = cases.ms_inputs(c)
    dbidx = meta.dbidx.values
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
    db.set_scan_mode(2)
    sets = cases.exclude_sets(meta, c["seed"] + 7)
    names = list(sets)
    # one batch holding every (exclude set, query) pair of the golden file
    batch_q = np.stack([qs[qi] for x in names for qi in range(2)])
    batch_ex = [sets[x] for x in names for qi in range(2)]
    r = db.scan_topk(batch_q, 50, exclude=batch_ex)
    j = 0
    for x in names:
        for qi in range(2):
            key = f"{name}/prelim/{x}/{qi}"
            gd, gs = golden[key + "/dbidx"], golden[key + "/score"]
            n = r["count"][j]
            assert n == len(gd), key
            np.testing.assert_allclose(r["score"][j, :n], gs, rtol=1e-5, atol=1e-6)
            if not (r["dbidx"][j, :n] == gd).all():
                admissible(r["dbidx"][j, :n], r["score"][j, :n], vecs, dbidx, qs[qi], sets[x], 50)
            j += 1
    db.close()


def test_batched_matches_streaming_kernel(eng):
    """K2 and K1 must return identical ids on Gaussian data (scores within 1e-5 relative)."""
    counts = synth.patches_per_image(20000, 20, 60, 3)
    dbidx = synth.dbidx_of_rows(counts)
    db = eng.PatchDatabase.synthetic(dbidx, 512, seed=4, kind="tri", store="f16")
    qs = synth.unit_queries(64, 512, 5)
    ex = [np.arange(i, 20000, 97) for i in range(64)]
    db.set_scan_mode(1)
    a = db.scan_topk(qs, 50, exclude=ex)
    db.set_scan_mode(2)
    b = db.scan_topk(qs, 50, exclude=ex)
    np.testing.assert_allclose(a["score"], b["score"], rtol=1e-5, atol=1e-6)
    mism = (a["dbidx"] != b["dbidx"]).sum()
    assert mism <= 4, mism      # only adjacent near-ties may swap
    db.close()


# ------------------------------------------------------------------ K4: merge of candidate lists
@pytest.mark.parametrize("n_lists,nq,k", [(1, 3, 50), (8, 64, 50), (148, 2, 64), (40, 3, 2048), (5, 1, 1), (300, 2, 700)])
def test_merge_kernel_vs_host_statement(eng, n_lists, nq, k):
    """ssw_merge_topk_device against the numpy statement of the merge (all three size regimes: survivors
    fit the sort buffer, need the shared-memory radix select, need the global radix select)."""
    import torch
    from seesaw_b200 import sharded
    rng = np.random.default_rng(n_lists * 1000 + k)
    score = rng.standard_normal((n_lists, nq, k)).astype(np.float32)
    score[rng.random(score.shape) < 0.3] = np.float32(0.25)            # exact score ties -> row order decides
    rows = rng.permutation(n_lists * nq * k).reshape(n_lists, nq, k)
    keys = sharded.encode_keys(score, rows)
    ids = rng.integers(0, 10 ** 6, size=keys.shape).astype(np.int32)
    empty = rng.random(keys.shape) < 0.2
    keys[empty] = 0
    ids[empty] = -1
    if n_lists == 5:
        keys[:] = 0                                                     # nothing at all
        ids[:] = -1
    out = eng.merge_topk_device(torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(ids).cuda(), k)
    hk, hd = sharded.merge_candidates_host(keys, ids, k)
    gk = out["key"].cpu().numpy().view(np.uint64)
    assert (gk == hk).all()
    assert (out["dbidx"].cpu().numpy() == hd).all()
    s, r = sharded.decode_keys(hk)
    assert (out["row"].cpu().numpy() == r).all()
    assert (out["count"].cpu().numpy() == (hk != 0).sum(axis=1)).all()
    got = out["score"].cpu().numpy()
    assert ((got == s) | (np.isinf(got) & np.isinf(s))).all()


# ------------------------------------------------------------------ more edge cases
def test_max_topk_and_ragged_images(eng):
    """k at the ABI limit (2048) on the streaming kernel, images from 1 to 300 rows, k > #images."""
    rng = np.random.default_rng(5)
    counts = rng.integers(1, 301, size=3000).astype(np.int64)
    counts[::50] = 1
    dbidx = synth.dbidx_of_rows(counts, 0, 7)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 256, 29, "lattice", np.float32)
    q = synth.lattice_queries(2, 256, 30)
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
    ids = np.unique(dbidx)
    for k in (2048, 3000 + 5):
        if k > 2048:
            with pytest.raises(Exception):
                db.scan_topk(q, k)
            continue
        r = db.scan_topk(q, k, exclude=[ids[::3], None])
        for qi, e in enumerate([ids[::3], None]):
            o = orc.query_prelim(vecs, dbidx, q[qi], k, exclude=e)
            kk = len(o["dbidx"])
            assert r["count"][qi] == kk
            assert (r["dbidx"][qi, :kk] == o["dbidx"]).all() and (r["row"][qi, :kk] == o["best_row"]).all()
            assert (r["score"][qi, :kk] == o["max_score"]).all()
    db.close()


def test_empty_database_and_tiny_batches(eng):
    db = eng.PatchDatabase.from_arrays(np.zeros((0, 512), np.float32), np.zeros(0, np.int32), store="f16")
    assert db.n_rows == 0 and db.n_images == 0
    q = synth.lattice_queries(3, 512, 1)
    for mode in (1, 2):
        db.set_scan_mode(mode)
        r = db.scan_topk(q, 5, exclude=[[1, 2], None, []])
        assert (r["count"] == 0).all() and (r["dbidx"] == -1).all()
    assert db.score_all(q[0]).shape == (0,)
    db.close()
    # one row in total; batched kernel with k = 64 (its limit) and nq = 1
    v = synth.synth_rows(0, 1, 512, 2, "lattice", np.float32)
    db = eng.PatchDatabase.from_arrays(v, np.array([9], np.int32), store="f16")
    db.set_scan_mode(2)
    r = db.scan_topk(q[:1], 64)
    assert r["count"][0] == 1 and r["dbidx"][0, 0] == 9 and r["row"][0, 0] == 0 and r["score"][0, 0] == (v @ q[0])[0]
    with pytest.raises(Exception):
        db.scan_topk(q[:1], 65)          # beyond the batched kernel's list length: forced mode must refuse
    db.set_scan_mode(0)
    assert db.scan_topk(q[:1], 65)["count"][0] == 1      # auto mode falls back to the streaming kernel
    db.close()


# ------------------------------------------------------------------ K5: top-k from caller-supplied scores
@pytest.mark.parametrize("grouped", [True, False])
def test_topk_from_scores_matches_get_top_dbidxs(eng, grouped):
    """KnnProp2.next_batch path: arbitrary row scores (many exact ties), a subset of rows, exclusion."""
    from seesaw_b200.indices import B200MultiscaleIndex
    rng = np.random.default_rng(11)
    counts = synth.patches_per_image(5000, 1, 25, 12)
    meta = synth.synth_vector_meta(counts, 13, dbidx_start=3, dbidx_stride=2)
    n = int(counts.sum())
    vecs = synth.synth_rows(0, n, 256, 14, "lattice", np.float32)
    if not grouped:
        perm = rng.permutation(n)
        vecs, meta = vecs[perm], meta.iloc[perm].reset_index(drop=True)
    idx = B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    dbidx = meta.dbidx.to_numpy()
    scores = (rng.integers(-50, 51, size=n) / 8.0).astype(np.float32)          # heavy ties
    ids = np.unique(dbidx)
    for frac, ex, k in ((1.0, None, 50), (0.7, ids[::4], 50), (0.05, ids[:100], 2000), (1.0, ids, 10)):
        rows = np.sort(rng.choice(n, size=int(n * frac), replace=False)) if frac < 1 else np.arange(n)
        order = rows[np.argsort(-scores[rows], kind="stable")]                  # what top_k(k=None) hands over
        got = idx.top_dbidxs(vec_idxs=order, scores=scores[order], exclude=ex, topk=k)
        d, s, r = orc.get_top_dbidxs(order, scores[order], dbidx, ex, k)
        assert len(got) == len(d)
        assert (got.dbidx.values == d).all() and (got.max_score.values == s).all() and (got.best_row.values == r).all()
        # the score-keyed entry point (ssw_topk_from_scores): float32 scores per row, ties to the lower row
        dense, mask = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        dense[rows], mask[rows] = scores[rows], 1
        g2 = idx.db.topk_from_scores(dense, k, exclude=ex, row_mask=None if frac == 1 else mask)
        assert (g2["dbidx"] == d).all() and (g2["score"] == s).all() and (g2["row"] == r).all()
    # float64 propagation scores that differ only beyond float32 precision, handed over in the CALLER'S order (ties in
    # arbitrary order, as an unstable argsort leaves them): the result follows that order exactly
    s64 = 0.5 + rng.integers(0, 40, size=n) * 1e-13
    order = rng.permutation(n)
    order = order[np.argsort(-s64[order], kind="stable")]
    got = idx.top_dbidxs(vec_idxs=order, scores=s64[order], exclude=ids[::5], topk=60)
    d, s, r = orc.get_top_dbidxs(order, s64[order], dbidx, ids[::5], 60)
    assert got.max_score.dtype == np.float64 and len(np.unique(s64.astype(np.float32))) == 1
    assert (got.dbidx.values == d).all() and (got.max_score.values == s).all() and (got.best_row.values == r).all()
    idx.close()


# ------------------------------------------------------------------ randomized configurations
def test_random_configurations_both_kernels(eng):
    """Seeded sweep over shapes the fixed cases do not hit together: ragged images, tiny and odd sizes, all dims,
    k from 1 to 64, partial / multi-pass batches, permuted rows, both kernels — ids, rows and scores bit-exact
    against the oracle on exact-arithmetic data."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        dim = int(rng.choice([256, 512, 768]))
        n_images = int(rng.choice([1, 2, 37, 149, 600, 2500]))
        hi = int(rng.choice([1, 3, 17, 70]))
        counts = rng.integers(1, hi + 1, size=n_images).astype(np.int64)
        dbidx = synth.dbidx_of_rows(counts, int(rng.integers(0, 50)), int(rng.integers(1, 4)))
        n = int(counts.sum())
        vecs = synth.synth_rows(0, n, dim, 100 + trial, "lattice", np.float32)
        if rng.random() < 0.5:
            perm = rng.permutation(n)
            vecs, dbidx = vecs[perm], dbidx[perm]
        nq = int(rng.choice([1, 2, 9, 33, 64, 70]))
        k = int(rng.choice([1, 5, 50, 64]))
        mode = int(rng.choice([1, 2]))
        qs = synth.lattice_queries(nq, dim, 500 + trial)
        ids = np.unique(dbidx)
        ex = [None if rng.random() < 0.3 else rng.choice(ids, size=int(rng.integers(0, len(ids) + 1)), replace=False)
              for _ in range(nq)]
        db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
        db.set_scan_mode(mode)
        r = db.scan_topk(qs, k, exclude=ex)
        for qi in range(nq):
            o = orc.query_prelim(vecs, dbidx, qs[qi], k, exclude=ex[qi])
            kk = len(o["dbidx"])
            tag = (trial, dim, n_images, hi, nq, k, mode, qi)
            assert r["count"][qi] == kk, tag
            assert (r["dbidx"][qi, :kk] == o["dbidx"]).all(), tag
            assert (r["row"][qi, :kk] == o["best_row"]).all(), tag
            assert (r["score"][qi, :kk] == o["max_score"]).all(), tag
            assert (r["dbidx"][qi, kk:] == -1).all(), tag
        db.close()
