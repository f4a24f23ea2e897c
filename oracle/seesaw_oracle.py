"""TEST INFRASTRUCTURE — CPU oracle for the SeeSaw vector-search hot path.

A numpy/pandas restatement of the reference's algorithm (orm011/seesaw 1.3.0), used ONLY as
the checker: by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product (``seesaw_b200``) never imports it.

Parity status: PINNED.  ``tests/test_oracle.py`` runs the unmodified reference (imported through
``oracle/refstubs.py``, when ``/root/reference`` is present) against this restatement on seeded inputs
in the build container, ``oracle/make_golden.py`` stores reference outputs as fixtures under ``tests/golden/``
(those travel to the GPU box, the reference does not), and the reference's own unit pin
``test_distinct_topk_positions`` (multiscale_index.py:182-187) is reproduced in
``tests/test_oracle.py``.

The one deliberate difference from the literal reference: every ``np.argsort`` here is
``kind="stable"``.  The reference uses numpy's default (unstable) sort, so its order among
EXACTLY equal scores is undefined; the stable order (score desc, original row asc) is the
tie-breaking definition the CUDA kernels implement (SURVEY.md §7 "Reference tie order is
undefined").  On tie-free inputs the two agree and the tests assert that.

All ``file:line`` citations are relative to ``/root/reference/seesaw``.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

# --------------------------------------------------------------------------------------
# Stage 1: scan -> sorted rows -> exclusion -> first occurrence per image -> head(k)
# --------------------------------------------------------------------------------------


def get_top_exact(vector, vectors):
    """indices/multiscale/multiscale_index.py:170-175 — fp32 matvec, then ALL rows sorted by
    descending score (stable: equal scores keep ascending row order)."""
    scores = vectors @ np.asarray(vector).reshape(-1)
    order = np.argsort(-scores, kind="stable")
    return order, scores[order]


def distinct_topk_positions(dbidxs, topk):
    """multiscale_index.py:177-180 — positions of the first occurrence of every distinct value,
    ascending, truncated to ``topk``."""
    _, first = np.unique(np.asarray(dbidxs), return_index=True)
    return np.sort(first)[:topk]


def _as_id_array(ids):
    if ids is None:
        return np.zeros(0, dtype=np.int64)
    if isinstance(ids, np.ndarray):
        return ids.astype(np.int64).reshape(-1)
    return np.fromiter((int(v) for v in ids), dtype=np.int64)


def get_top_dbidxs(vec_idxs, scores, dbidx_of_row, exclude, topk):
    """multiscale_index.py:189-199 — walk the score-sorted rows, drop rows of excluded images,
    keep the first (= best) row of each remaining image, return the first ``topk`` images.
    Returns (dbidx, max_score, best_row)."""
    sorted_dbidx = np.asarray(dbidx_of_row)[vec_idxs]
    keep = ~np.isin(sorted_dbidx, _as_id_array(exclude))
    kept_dbidx, kept_scores, kept_rows = sorted_dbidx[keep], scores[keep], vec_idxs[keep]
    pos = distinct_topk_positions(kept_dbidx, topk)
    return kept_dbidx[pos], kept_scores[pos], kept_rows[pos]


def query_prelim(vectors, dbidx_of_row, vector, topk_dbidx, exclude=None, index_excluded=None):
    """MultiscaleIndex._query_prelim, exact branch (multiscale_index.py:291-312).

    ``index_excluded`` is ``MultiscaleIndex.excluded`` (:216,223): it only shrinks
    ``all_indices`` and hence the clamp of ``topk`` (:295-298); like the reference, rows of such
    images are NOT filtered from the scan itself unless they are also in ``exclude``.
    Returns dict(dbidx int64[k'], max_score fp32[k'], best_row int64[k']).  k' == 0 -> empties
    (the reference returns the tuple ``[], [], []`` there, :300-302)."""
    exclude = _as_id_array(exclude)
    present = np.unique(np.asarray(dbidx_of_row))
    all_indices = np.setdiff1d(present, _as_id_array(index_excluded))
    included = np.setdiff1d(all_indices, exclude)
    k = min(int(topk_dbidx), included.shape[0])
    if k == 0:
        return dict(dbidx=np.zeros(0, np.int64), max_score=np.zeros(0, np.float32),
                    best_row=np.zeros(0, np.int64))
    order, sorted_scores = get_top_exact(vector, vectors)
    d, s, r = get_top_dbidxs(order, sorted_scores, dbidx_of_row, exclude, k)
    return dict(dbidx=d.astype(np.int64), max_score=s, best_row=r.astype(np.int64))


# --------------------------------------------------------------------------------------
# Stage 2: rescoring of the shortlist (multiscale_index.py:341-352, 379-403, 112-150)
# --------------------------------------------------------------------------------------


def _pairwise_iou(boxes):
    """IoU of every box pair of one image (box_utils.py:336-350): df2tensor stacks the four columns
    (:329-334, numpy promotion), torchvision's _box_inter_union computes area / intersection / union in
    that dtype (float32 for the tiling pipeline's boxes, multiscale_tools.py:111; int64 for integer
    columns), and ``inter / union`` is a true division (float32 for integer tensors, torch's default).
    boxes [n,4] as x1,y1,x2,y2 in their own dtype."""
    integer = np.issubdtype(boxes.dtype, np.integer)
    boxes = boxes.astype(np.int64) if integer else boxes
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    iw = np.clip(np.minimum(x2[:, None], x2[None, :]) - np.maximum(x1[:, None], x1[None, :]), 0, None)
    ih = np.clip(np.minimum(y2[:, None], y2[None, :]) - np.maximum(y1[:, None], y1[None, :]), 0, None)
    inter = iw * ih
    union = (area[:, None] + area[None, :]) - inter
    with np.errstate(invalid="ignore", divide="ignore"):
        if integer:
            return inter.astype(np.float32) / union.astype(np.float32)
        return inter / union


def _pandas_group_mean_f32(values):
    """pandas' groupby(...).mean() of one group of a float32 column (multiscale_index.py:142): group_mean
    accumulates in the column's own float32 with Kahan compensation, in row order, and divides by the
    count in float32 (pandas/_libs/groupby.pyx; checked against pandas itself in tests/test_oracle.py)."""
    f = np.float32
    sumx, comp = f(0), f(0)
    for v in values:
        y = f(f(v) - comp)
        t = f(sumx + y)
        comp = f(f(t - sumx) - y)
        sumx = t
    return f(sumx / f(len(values)))


def frame_best_patch(frame: pd.DataFrame, agg_method="plain_score", aug_larger="all"):
    """score_frame2 (multiscale_index.py:112-150) for aug_weight='level_max'.
    Returns (iloc of the winning patch inside ``frame``, its score)."""
    s = frame["score"].to_numpy()
    if agg_method == "plain_score":
        # :117-118  first row (frame order) whose score equals the max
        return int(np.flatnonzero(s == s.max())[0]), s.max()
    assert agg_method == "avg_score"
    boxes = frame[["x1", "y1", "x2", "y2"]].to_numpy()     # common dtype of the four columns, as np.stack gives
    zoom = frame["zoom_level"].to_numpy()
    iou = _pairwise_iou(boxes)
    ok = iou > 0                                           # box_join(iou_gt=0), box_utils.py:357
    if aug_larger == "greater":
        ok &= zoom[None, :] >= zoom[:, None]               # :124-125
    elif aug_larger == "adjacent":
        ok &= zoom[None, :] == zoom[:, None]               # :126-127
    else:
        assert aug_larger == "all"
    n = len(frame)
    out = np.full(n, np.nan, dtype=s.dtype)                # groupby(...).mean() keeps the score column's dtype (:142)
    for left in range(n):                                  # groupby(['iloc_left','zoom_level_right']).iou.idxmax()
        rights = np.flatnonzero(ok[left])
        if rights.size == 0:
            continue
        picked = []
        for z in np.unique(zoom[rights]):
            cand = rights[zoom[rights] == z]
            picked.append(cand[np.argmax(iou[left, cand])])  # first max = lowest right index
        # :141-143 mean over zoom levels, ascending (groupby order), in pandas' arithmetic for the column dtype
        out[left] = _pandas_group_mean_f32(s[picked]) if s.dtype == np.float32 else np.mean(s[picked])
    # :149-150 rows whose aggregated score equals the max; NaN rows never compare equal
    best = np.nanmax(out)
    return int(np.flatnonzero(out == best)[0]), best


def rescore_candidates(meta: pd.DataFrame, scores, topk, agg_method="plain_score", aug_larger="all"):
    """rescore_candidates (multiscale_index.py:379-403): images visited in ascending dbidx
    (pandas groupby sorts), one winning patch per image, then argsort(-score)[:topk] (stable)."""
    meta = meta.reset_index(drop=True).assign(score=np.asarray(scores))
    ids, best_scores, acts = [], [], []
    for dbidx, frame in meta.groupby("dbidx"):
        i, sc = frame_best_patch(frame, agg_method=agg_method, aug_larger=aug_larger)
        row = frame.iloc[[i]][["x1", "y1", "x2", "y2", "dbidx"]].assign(score=sc)
        ids.append(dbidx)
        best_scores.append(sc)
        acts.append(row)
    order = np.argsort(-np.asarray(best_scores, dtype=np.float64), kind="stable")[:topk]
    return {"dbidxs": np.asarray(ids)[order].astype("int"),
            "activations": [acts[i] for i in order]}


def multiscale_query(vectors, vector_meta: pd.DataFrame, vector, topk, shortlist_size,
                     exclude=None, vector2=None, index_excluded=None,
                     agg_method="plain_score", aug_larger="all"):
    """MultiscaleIndex.query (multiscale_index.py:314-352), exact path."""
    if shortlist_size is None:
        shortlist_size = topk * 5                         # :325-326
    dbidx_of_row = vector_meta["dbidx"].to_numpy()
    short = query_prelim(vectors, dbidx_of_row, vector, shortlist_size, exclude, index_excluded)
    rows = np.flatnonzero(np.isin(dbidx_of_row, short["dbidx"]))          # :341-342
    sub = vectors[rows]
    scores = sub @ np.asarray(vector).reshape(-1)                          # :345
    if vector2 is not None:
        scores = scores - sub @ np.asarray(vector2).reshape(-1)           # :347-349
    return rescore_candidates(vector_meta.iloc[rows], scores, topk,
                              agg_method=agg_method, aug_larger=aug_larger)


# --------------------------------------------------------------------------------------
# Coarse index (indices/coarse/coarse_index.py:57-96)
# --------------------------------------------------------------------------------------


def coarse_query(vectors, dbidx_of_row, vector, topk, exclude=None, rng=None):
    """CoarseIndex.query: one row per image, rows ascending in dbidx (asserted at :49).
    Returns dict(dbidxs, scores, nextstartk) or None when nothing is included (:61-62, where the
    reference returns a pair of empty arrays)."""
    exclude = _as_id_array(exclude)
    dbidx_of_row = np.asarray(dbidx_of_row)
    included = np.setdiff1d(np.unique(dbidx_of_row), exclude)              # :60 (ascending)
    if included.shape[0] == 0:
        return None
    k = min(int(topk), included.shape[0])                                  # :64-65
    mask = np.isin(dbidx_of_row, included)                                 # :67
    vecs = vectors[mask]                                                   # :68
    if vector is None:
        scores = (rng or np.random).standard_normal(vecs.shape[0])        # :70-71
    else:
        scores = vecs @ np.asarray(vector).reshape(-1)                     # :73
    best = np.argsort(-scores, kind="stable")[:k]                          # :75
    return dict(dbidxs=included[best].astype(np.int64), scores=scores[best],
                nextstartk=int(exclude.shape[0] + k))                      # :76, :94


# --------------------------------------------------------------------------------------
# kNN graph (knn_graph.py:142-191)
# --------------------------------------------------------------------------------------


def exact_knn_candidates(vectors, n_neighbors):
    """compute_exact_knn up to the edge table (knn_graph.py:170-182): the k1 = min(n+1, N)
    columns of smallest fp32(1 - dot) per row, ties by ascending column.  Self is included
    whenever it ranks.  Returns (idx int32 [N,k1], dist fp32 [N,k1])."""
    n = vectors.shape[0]
    k1 = min(int(n_neighbors) + 1, n)
    all_pairs = 1.0 - (vectors @ vectors.T)
    idx = np.argsort(all_pairs, axis=-1, kind="stable")[:, :k1]
    dist = np.take_along_axis(all_pairs, idx, axis=-1)
    return idx.astype(np.int32), dist.astype(np.float32)


def exact_knn_candidates_blockwise(vectors, n_neighbors, block=1024, rows=None):
    """Same result as :func:`exact_knn_candidates` without the N x N matrix: row blocks of
    ``block``.  ``rows`` restricts the computation to a row range (used by the
    bounded CPU baseline and by big-N spot checks)."""
    n = vectors.shape[0]
    k1 = min(int(n_neighbors) + 1, n)
    lo, hi = (0, n) if rows is None else rows
    idx = np.empty((hi - lo, k1), np.int32)
    dist = np.empty((hi - lo, k1), np.float32)
    for b0 in range(lo, hi, block):
        b1 = min(b0 + block, hi)
        d = 1.0 - (vectors[b0:b1] @ vectors.T)
        if k1 < n:
            # candidates: everything <= the k1-th smallest value, then the stable order among them
            kth = np.partition(d, k1 - 1, axis=-1)[:, k1 - 1]
            for i in range(b1 - b0):
                cand = np.flatnonzero(d[i] <= kth[i])
                o = cand[np.argsort(d[i, cand], kind="stable")][:k1]
                idx[b0 - lo + i], dist[b0 - lo + i] = o, d[i, o]
        else:
            o = np.argsort(d, axis=-1, kind="stable")
            idx[b0 - lo:b1 - lo], dist[b0 - lo:b1 - lo] = o, np.take_along_axis(d, o, axis=-1)
    return idx, dist


def post_process_graph(idx, dist, nvec, src_offset=0):
    """post_process_graph_df (knn_graph.py:142-168) on a [rows,k1] candidate table: int32/fp32
    casts, distance clipped at 0, self edges dropped, dst_rank = 1.. by (distance, table order)
    within each source, one rank-0 zero-distance self edge re-added per vertex, sorted by
    (src_vertex, dst_rank)."""
    rows, k1 = idx.shape
    src = (np.repeat(np.arange(rows, dtype=np.int64), k1) + src_offset).astype(np.int32)
    df = pd.DataFrame({"src_vertex": src, "dst_vertex": idx.reshape(-1).astype(np.int32),
                       "distance": np.clip(dist.reshape(-1).astype(np.float32), 0.0, None)})
    df = df[df.src_vertex != df.dst_vertex]
    df = df.assign(dst_rank=df.groupby("src_vertex").distance.rank("first").astype("int32"))
    me = np.arange(nvec, dtype=np.int32)
    selfs = pd.DataFrame({"src_vertex": me, "dst_vertex": me,
                          "distance": np.zeros(nvec, np.float32), "dst_rank": np.zeros(nvec, np.int32)})
    df = pd.concat([df, selfs], ignore_index=True)
    return df.sort_values(["src_vertex", "dst_rank"], kind="stable").reset_index(drop=True)


def compute_exact_knn(vectors, n_neighbors):
    """compute_exact_knn (knn_graph.py:170-191) end to end."""
    idx, dist = exact_knn_candidates(vectors, n_neighbors)
    return post_process_graph(idx, dist, vectors.shape[0])


# --------------------------------------------------------------------------------------
# Label propagation over the kNN graph (label_propagation.py:6-83)
# --------------------------------------------------------------------------------------


def label_propagation_fit(weight_matrix, *, reg_lambda, max_iter, epsilon=1e-5, label_ids, label_values,
                          reg_values=None, start_value=None):
    """LabelPropagation(weight_matrix, reg_lambda=, max_iter=, epsilon=).fit_transform(label_ids=,
    label_values=, reg_values=, start_value=) — label_propagation.py:7-24 (constructor: weight_sum =
    W.sum(0)), :30-43 (_step) and :45-83 (loop).  Returns (values, iterations, converged); on convergence
    the PREVIOUS iterate is returned (:66-70, :83), otherwise the last one."""
    W = weight_matrix
    n = W.shape[0]
    weight_sum = np.asarray(W.sum(0)).reshape(-1)                                   # :24
    if reg_values is not None:
        assert reg_values.shape[0] == n
        reg = reg_values
    else:
        assert reg_lambda == 0                                                      # :50
        reg = np.zeros(n)
    if start_value is not None:
        old = start_value.copy()                                                    # :54
    elif reg_values is not None:
        old = reg_values.copy()                                                     # :56
    else:
        old = np.zeros(n)                                                           # :58
    old[label_ids] = label_values                                                   # :60
    converged, i = False, 0
    for i in range(1, max_iter + 1):
        new = (W @ old + (reg_lambda * reg)) / (weight_sum + reg_lambda)            # :31-32
        new[label_ids] = label_values                                               # :42
        if np.max((new - old) ** 2) < epsilon:                                      # :66
            converged = True
            break
        old = new
    return old, i, converged


# --------------------------------------------------------------------------------------
# helpers for the tolerance rule used by the floating-point parity tests
# --------------------------------------------------------------------------------------


def scores_f64(vectors, vector):
    return vectors.astype(np.float64) @ np.asarray(vector, dtype=np.float64).reshape(-1)


def per_image_best(scores, dbidx_of_row, exclude=None):
    """Direct (sort-free) statement of stage 1: for every non-excluded image the row minimising
    (-score, row).  Returns (dbidx, best_score, best_row) ranked by (-score, row)."""
    dbidx_of_row = np.asarray(dbidx_of_row)
    order = np.lexsort((np.arange(scores.shape[0]), -scores, dbidx_of_row))
    first = np.ones(order.shape[0], bool)
    first[1:] = dbidx_of_row[order][1:] != dbidx_of_row[order][:-1]
    rows = order[first]
    keep = ~np.isin(dbidx_of_row[rows], _as_id_array(exclude))
    rows = rows[keep]
    rank = np.lexsort((rows, -scores[rows]))
    rows = rows[rank]
    return dbidx_of_row[rows], scores[rows], rows
