"""Stage 2 of the multiscale query: rescoring of the <= shortlist_size candidate images.

Host-side mirror of ``rescore_candidates`` / ``score_frame2``
(seesaw/indices/multiscale/multiscale_index.py:379-403, 112-150) and of the IoU self-join they use
(seesaw/box_utils.py:336-372).  It touches at most shortlist_size x ~60 rows, so it stays on the
CPU; what changes is that the candidate rows come from the index's CSR ranges
instead of an O(N) ``vector_meta.dbidx.isin`` (multiscale_index.py:341-342).

Differences from the reference, both confined to exactly tied values where its unstable
``np.argsort`` (:399) leaves the order undefined: ties here go to the lower dbidx.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

_BOX = ["x1", "y1", "x2", "y2"]


def _iou_matrix(b):
    """Pairwise IoU of boxes [n,4] (x1,y1,x2,y2) as box_iou computes it (box_utils.py:336-350, torchvision's
    _box_inter_union followed by a true division) IN THE BOXES' OWN DTYPE: float32 boxes (what the tiling
    pipeline writes, multiscale_tools.py:111) give float32 areas / intersections / unions / quotients, integer
    boxes exact integers and a float32 quotient, float64 boxes float64 throughout."""
    integer = np.issubdtype(b.dtype, np.integer)
    if integer:
        b = b.astype(np.int64)
    elif b.dtype != np.float32:
        b = b.astype(np.float64)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    w = np.minimum(b[:, None, 2], b[None, :, 2]) - np.maximum(b[:, None, 0], b[None, :, 0])
    h = np.minimum(b[:, None, 3], b[None, :, 3]) - np.maximum(b[:, None, 1], b[None, :, 1])
    inter = np.clip(w, 0, None) * np.clip(h, 0, None)
    union = (area[:, None] + area[None, :]) - inter
    with np.errstate(invalid="ignore", divide="ignore"):
        if integer:
            return inter.astype(np.float32) / union.astype(np.float32)
        return inter / union


def aggregate_patch_scores(scores, boxes, zoom, aug_larger="all"):
    """'avg_score' with aug_weight='level_max' (multiscale_index.py:119-147): for every patch, the
    mean over zoom levels of the score of the best-overlapping patch of that level.
    Returns float32 [n] (pandas' groupby mean accumulates in float64 and hands back the score column's
    float32); NaN where a patch overlaps nothing under the aug_larger filter."""
    n = scores.shape[0]
    scores = np.asarray(scores, dtype=np.float32)
    iou = _iou_matrix(boxes)
    allowed = iou > 0
    if aug_larger == "greater":
        allowed &= zoom[None, :] >= zoom[:, None]
    elif aug_larger == "adjacent":
        allowed &= zoom[None, :] == zoom[:, None]
    elif aug_larger != "all":
        raise AssertionError(f"unknown aug_larger {aug_larger!r}")
    masked = np.where(allowed, iou, iou.dtype.type(-1.0))
    # pandas' group_mean on the float32 score column: Kahan-compensated float32 sum over the levels in
    # ascending order, float32 division by the count (vectorised over the left patches)
    f = np.float32
    sumx, comp, levels = np.zeros(n, f), np.zeros(n, f), np.zeros(n, f)
    for z in np.unique(zoom):
        cols = np.flatnonzero(zoom == z)
        sub = masked[:, cols]
        best = cols[np.argmax(sub, axis=1)]            # first maximum = lowest right position (idxmax)
        has = sub.max(axis=1) > 0
        y = scores[best] - comp
        t = sumx + y
        comp = np.where(has, (t - sumx) - y, comp)
        sumx = np.where(has, t, sumx)
        levels += has
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(levels > 0, sumx / levels, f(np.nan)).astype(f)


def best_patch(scores, boxes=None, zoom=None, agg_method="plain_score", aug_larger="all"):
    """(position of the winning patch within the image, its score) — score_frame2."""
    if agg_method == "plain_score":
        m = scores.max()
        return int(np.flatnonzero(scores == m)[0]), m
    if agg_method != "avg_score":
        raise NotImplementedError(f"agg_method {agg_method!r}")
    agg = aggregate_patch_scores(scores, boxes, zoom, aug_larger)
    m = np.nanmax(agg)
    return int(np.flatnonzero(agg == m)[0]), m


def rescore_candidates(row_groups, dbidxs, scores_per_group, meta_cols, topk, *, agg_method="plain_score",
                       aug_larger="all", **_ignored):
    """row_groups: list of original-row index arrays, one per candidate image, images in ASCENDING
    dbidx (pandas groupby order, multiscale_index.py:388); scores_per_group: matching score arrays.
    meta_cols: dict of numpy columns x1,y1,x2,y2,zoom_level of the whole index.
    Returns the reference's result dict {"dbidxs", "activations"}."""
    n = len(row_groups)
    best_score = np.zeros(n)
    best_row = np.zeros(n, np.int64)
    for i, (rows, sc) in enumerate(zip(row_groups, scores_per_group)):
        if agg_method == "plain_score":
            pos, val = best_patch(sc)
        else:
            boxes = np.stack([meta_cols[c][rows] for c in _BOX], axis=1)
            pos, val = best_patch(sc, boxes, meta_cols["zoom_level"][rows], agg_method, aug_larger)
        best_score[i], best_row[i] = val, rows[pos]
    order = np.argsort(-best_score, kind="stable")[:topk]
    acts = []
    for i in order:
        r = best_row[i]
        acts.append(pd.DataFrame({"x1": [meta_cols["x1"][r]], "y1": [meta_cols["y1"][r]],
                                  "x2": [meta_cols["x2"][r]], "y2": [meta_cols["y2"][r]],
                                  "dbidx": [dbidxs[i]], "score": [best_score[i]]}))
    return {"dbidxs": np.asarray(dbidxs)[order].astype("int"), "activations": acts}
