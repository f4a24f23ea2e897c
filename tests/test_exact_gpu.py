"""GPU parity on the data a real SeeSaw index holds: float32 unit vectors that are NOT fp16-representable
(multiscale_tools.py:200) and the tiling pipeline's float32 boxes (multiscale_tools.py:96-117).

Three storage modes of the same database are held against the reference / the oracle:
  store="f32"            the reference's own dtype: ids equal, scores within 1e-5 relative;
  store="f16" + exact    fp16 scan + certified float32 re-ranking (ssw_db_attach_exact): the SAME bits as
                         store="f32" (both score rows with the canonical float32 dot product);
  store="f16" alone      the stated looser bound: every score within rho*||q|| of the float64 score
                         (rho = max_i ||v_i - fp16(v_i)||, reported by the library), ids may differ from the
                         oracle's only between images whose float64 scores are closer than 2*rho*||q||.
"""
import threading

import numpy as np
import pytest

import cases
import seesaw_oracle as orc
from seesaw_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ix():
    from seesaw_b200 import indices
    return indices


@pytest.fixture(scope="module")
def eng():
    from seesaw_b200 import engine
    return engine


@pytest.mark.parametrize("name", list(cases.CASES_F32))
@pytest.mark.parametrize("store", ["f32", "f16"])
def test_float32_index_vs_reference_golden(ix, golden, name, store):
    """B200MultiscaleIndex on float32 vectors + float32 boxes against the unmodified reference's outputs:
    stage 1, and stage 2 for plain / avg aggregation with every aug_larger filter, on the DEVICE path."""
    c = cases.CASES_F32[name]
    vecs, meta, qs = cases.msf_inputs(c)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store=store)   # exact="auto"
    assert idx._store_exact and idx._boxes_on_device
    assert idx.db.exact_info()["attached"] == (store == "f16")
    xs = cases.exclude_sets(meta, c["seed"] + 7)
    for xname in ("none", "some"):
        for qi in range(2):
            p = idx._query_prelim(vector=qs[qi], topk_dbidx=50, exclude_dbidx=ix.BitMap(xs[xname]))
            key = f"{name}/prelim/{xname}/{qi}"
            assert (p.dbidx.values == golden[key + "/dbidx"]).all(), key
            np.testing.assert_allclose(p.max_score.values, golden[key + "/score"], rtol=1e-5, atol=1e-7)
    for device_rescore in (True, False):
        for agg, aug, topk, use_v2 in cases.MSF_QUERY_VARIANTS:
            key = f"{name}/query/{agg}/{aug}/{topk}/{int(use_v2)}"
            r = idx.query(vector=qs[2], vector2=qs[3] * 0.25 if use_v2 else None, topk=topk, shortlist_size=40,
                          exclude=ix.BitMap(xs["some"]), agg_method=agg, aug_larger=aug, rescore_method=None,
                          device_rescore=device_rescore)
            assert (np.asarray(r["dbidxs"]) == golden[key + "/dbidxs"]).all(), (key, device_rescore)
            sc = np.array([a.score.values[0] for a in r["activations"]], np.float64)
            np.testing.assert_allclose(sc, golden[key + "/act_score"], rtol=1e-5, atol=1e-7)
            bx = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.float32)
            assert (bx == golden[key + "/act_box"]).all(), (key, device_rescore)     # the same winning patch
    idx.close()


def _f64_image_best(vecs, dbidx, q, exclude, chunk=400_000):
    s = np.empty(len(vecs), np.float64)
    q64 = q.astype(np.float64)
    for a in range(0, len(vecs), chunk):
        s[a:a + chunk] = vecs[a:a + chunk].astype(np.float64) @ q64
    d, sc, rows = orc.per_image_best(s, dbidx, exclude)
    return d, sc, rows


def _check_against_f64(res_dbidx, res_score, d64, s64, k, score_tol, swap_tol):
    """Returned scores within score_tol of the float64 ones; ids equal to the float64 ranking's, or swapped only
    between images whose float64 scores are within swap_tol.  Returns the number of positions that differ."""
    best = dict(zip(d64.tolist(), s64.tolist()))
    got = np.array([best[i] for i in res_dbidx.tolist()])
    assert np.abs(got - res_score.astype(np.float64)).max() <= score_tol
    diff = 0
    kth = s64[k - 1]
    for pos, (a, b) in enumerate(zip(res_dbidx.tolist(), d64[:k].tolist())):
        if a != b:
            diff += 1
            assert abs(best[a] - best[b]) <= swap_tol, (pos, a, b, best[a], best[b])
    assert got.min() >= kth - swap_tol            # nothing returned that is clearly outside the true top-k
    return diff


def test_real_fp32_vectors_at_config2_scale(eng):
    """BASELINE configs[1] scale (120k images x 20..60 patches = 4.8M x 512) with TRUE float32 unit Gaussians —
    not fp16-representable — in all three storage modes, single query (K1) and a 64-query batch (K2)."""
    try:
        import psutil
        if psutil.virtual_memory().available < 48 * 2 ** 30:
            pytest.skip("needs ~40 GB of host memory for the float32 database and its float64 check")
    except ImportError:
        pass
    counts = synth.patches_per_image(120_000, 20, 60, 3)
    dbidx = synth.dbidx_of_rows(counts)
    n = int(counts.sum())
    vecs = synth.unit_rows(n, 512, 44)
    qs = synth.unit_queries(64, 512, 45)
    rng = np.random.default_rng(46)
    excl = [np.sort(rng.choice(120_000, size=30, replace=False)) for _ in range(64)]
    k = 50
    checked = [0, 1, 63]
    f64 = {qi: _f64_image_best(vecs, dbidx, qs[qi], excl[qi]) for qi in checked}

    db32 = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f32")
    r32 = db32.scan_topk(qs, k, exclude=excl)          # fp32 storage: the streaming kernel per query
    db32.close()
    n_diff32 = 0
    for qi in checked:
        d64, s64, _ = f64[qi]
        o = orc.query_prelim(vecs, dbidx, qs[qi], k, exclude=excl[qi])
        np.testing.assert_allclose(r32["score"][qi], o["max_score"], rtol=1e-5, atol=1e-7)
        tol = 1e-5 * abs(s64[k - 1])
        _check_against_f64(r32["dbidx"][qi], r32["score"][qi], d64, s64, k, score_tol=2e-6, swap_tol=tol)
        # against the float32 oracle itself: equal ids, a swap admissible only where float64 calls it a near-tie
        best = dict(zip(d64.tolist(), s64.tolist()))
        for a, b in zip(r32["dbidx"][qi].tolist(), o["dbidx"].tolist()):
            if a != b:
                n_diff32 += 1
                assert abs(best[a] - best[b]) <= tol, (qi, a, b)
    print(f"float32 storage: {n_diff32} of {len(checked) * k} positions differ from the float32 oracle (near-ties < 1e-5 rel)")

    dbx = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16", exact=True)
    info = dbx.exact_info()
    assert info["attached"] and 0 < info["rho"] < 2 ** -11 * 1.01 * info["vmax"]
    rx = dbx.scan_topk(qs, k, exclude=excl)            # K2 (fp16) + certified float32 re-ranking
    r1 = dbx.scan_topk(qs[:1], k, exclude=excl[:1])    # K1 (fp16) + the same
    info = dbx.exact_info()
    print(f"exact mode: rho={info['rho']:.3e} vmax={info['vmax']:.4f} queries={info['queries']} float32 re-scans={info['rescans']}")
    assert info["queries"] == 65
    # the SAME bits as float32 storage: both rank rows by the canonical float32 dot product
    assert (rx["dbidx"] == r32["dbidx"]).all() and (rx["row"] == r32["row"]).all() and (rx["score"] == r32["score"]).all()
    assert (r1["dbidx"][0] == r32["dbidx"][0]).all() and (r1["score"][0] == r32["score"][0]).all()
    dbx.close()

    db16 = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16")
    bound = info["rho"] * 1.0 + 1e-5                   # ||q|| = 1; + fp32 summation noise
    r16 = db16.scan_topk(qs, k, exclude=excl)
    db16.set_scan_mode(1)
    r16_1 = db16.scan_topk(qs[:1], k, exclude=excl[:1])
    db16.close()
    total_diff = 0
    for qi in checked:
        d64, s64, _ = f64[qi]
        total_diff += _check_against_f64(r16["dbidx"][qi], r16["score"][qi], d64, s64, k, score_tol=bound, swap_tol=2 * bound)
    _check_against_f64(r16_1["dbidx"][0], r16_1["score"][0], f64[0][0], f64[0][1], k, score_tol=bound, swap_tol=2 * bound)
    print(f"fp16 storage alone: {total_diff} of {len(checked) * k} top-{k} positions differ from the float64 ranking "
          f"(all within 2*rho*||q|| = {2 * bound:.2e})")


def test_exact_mode_falls_back_to_float32_scan_on_ties(eng):
    """Duplicated images tie exactly around the k-th place: no error bound can certify the fp16 candidates, so
    those queries are re-scanned over the float32 rows — and still equal the oracle under index tie-breaking.
    Lattice values: every dot product is exact in float32 in any summation order, so the oracle's BLAS and the
    kernels produce the same bits and the ties are real."""
    counts = synth.patches_per_image(600, 2, 9, 7)
    dbidx = synth.dbidx_of_rows(counts, dbidx_start=3, dbidx_stride=2)
    n = int(counts.sum())
    base = synth.synth_rows(0, 40, 512, 8, "lattice", np.float32)
    vecs = np.ascontiguousarray(base[np.random.default_rng(9).integers(0, 40, size=n)])   # every row is one of 40 vectors
    qs = synth.lattice_queries(6, 512, 10)
    ex = [np.unique(dbidx)[i::11] for i in range(6)]
    db = eng.PatchDatabase.from_arrays(vecs, dbidx, store="f16", exact=True)
    for batch in (qs, qs[:1]):
        r = db.scan_topk(batch, 50, exclude=ex[:len(batch)])
        for i, q in enumerate(batch):
            o = orc.query_prelim(vecs, dbidx, q, 50, exclude=ex[i])
            assert (r["dbidx"][i] == o["dbidx"]).all() and (r["row"][i] == o["best_row"]).all(), i
            assert (r["score"][i] == o["max_score"]).all(), i
    info = db.exact_info()
    assert info["rescans"] > 0 and info["queries"] == 7 and info["rho"] == 0.0
    # fewer eligible images than candidates asked for: everything eligible is re-scored, nothing to certify
    few = np.unique(dbidx)[5:]
    r = db.scan_topk(qs[:2], 50, exclude=[few, few])
    o = orc.query_prelim(vecs, dbidx, qs[0], 50, exclude=few)
    assert r["count"][0] == 5 and (r["dbidx"][0, :5] == o["dbidx"]).all() and (r["dbidx"][0, 5:] == -1).all()
    assert db.exact_info()["rescans"] == info["rescans"]
    # index.score reads the float32 rows
    assert (db.score_all(qs[0]) == vecs @ qs[0]).all()
    db.close()
    # near-duplicates that are NOT fp16-representable: clusters far tighter than the fp16 rounding error, so the
    # certificate fails and the float32 re-scan decides; the float64 ranking may only differ between near-ties
    rng = np.random.default_rng(12)
    centres = synth.unit_rows(30, 512, 11)
    pick = rng.integers(0, 30, size=n)
    v2 = centres[pick] + rng.standard_normal((n, 512)).astype(np.float32) * np.float32(2e-6)
    q2 = synth.unit_queries(3, 512, 13)
    db = eng.PatchDatabase.from_arrays(v2, dbidx, store="f16", exact=True)
    r = db.scan_topk(q2, 20, exclude=None)
    assert db.exact_info()["rescans"] > 0
    for i in range(3):
        d64, s64, _ = _f64_image_best(v2, dbidx, q2[i], None)
        _check_against_f64(r["dbidx"][i], r["score"][i], d64, s64, 20, score_tol=2e-6, swap_tol=1e-5 * abs(s64[19]) + 1e-7)
    db.close()


def test_concurrent_sessions_share_one_handle(ix):
    """Several session threads call query(agg_method='avg_score') — stage 1 through a shared ScanBatcher, stage 2
    (ssw_rescore) and score() directly — on ONE database handle: the host-buffer entry points serialise on the
    handle's mutex, so every result equals the single-threaded one (staging block race)."""
    from seesaw_b200.service import ScanBatcher
    c = cases.CASES_F32["msf_unit"]
    vecs, meta, _ = cases.msf_inputs(c)
    idx = ix.B200MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, store="f16")
    qs = synth.unit_queries(12, c["dim"], 99)
    seen = np.unique(meta.dbidx.values)[::7]
    want = [idx.query(vector=q, topk=4, shortlist_size=30, exclude=ix.BitMap(seen), agg_method="avg_score") for q in qs]
    want_scores = [idx.score(q) for q in qs]
    batcher = ScanBatcher(idx.db, max_batch=8)
    idx.attach_batcher(batcher)
    errors = []

    def session(i):
        try:
            for rep in range(6):
                j = (i + rep) % len(qs)
                r = idx.query(vector=qs[j], topk=4, shortlist_size=30, exclude=ix.BitMap(seen), agg_method="avg_score")
                assert (np.asarray(r["dbidxs"]) == want[j]["dbidxs"]).all(), (i, rep)
                got = [a.score.values[0] for a in r["activations"]]
                assert got == [a.score.values[0] for a in want[j]["activations"]], (i, rep)
                assert (idx.score(qs[j]) == want_scores[j]).all()
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=session, args=(i,)) for i in range(12)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    batcher.close()
    idx.attach_batcher(None)
    idx.close()
    assert not errors, errors[0]
