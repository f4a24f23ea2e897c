"""TEST INFRASTRUCTURE — an oracle-backed stand-in for seesaw_b200.engine.PatchDatabase, so the host logic of the
reference-facing classes (seesaw_b200/indices.py) runs in the CPU suite.  Never used by the product."""
import numpy as np
import pandas as pd

import seesaw_oracle as orc


class FakeDB:
    def __init__(self, vectors, dbidx):
        self.v = np.asarray(vectors, dtype=np.float32)
        self.dbidx = np.asarray(dbidx, dtype=np.int64)
        self.n_rows, self.dim = self.v.shape
        self.n_images = len(np.unique(self.dbidx))
        self.dtype, self.device, self.closed, self.boxes = np.float16, 0, False, None

    @classmethod
    def from_arrays(cls, vectors, dbidx_per_row, *, store="f16", device=0, global_row_base=0, exact=False):
        return cls(vectors, dbidx_per_row)

    def scan_topk(self, queries, k, exclude=None):
        q = np.asarray(queries, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        out = dict(dbidx=np.full((nq, k), -1, np.int32), score=np.full((nq, k), -np.inf, np.float32),
                   row=np.full((nq, k), -1, np.int64), count=np.zeros(nq, np.int32))
        for i in range(nq):
            o = orc.query_prelim(self.v, self.dbidx, q[i], k, exclude=None if exclude is None else exclude[i])
            n = len(o["dbidx"])
            out["dbidx"][i, :n], out["score"][i, :n], out["row"][i, :n], out["count"][i] = o["dbidx"], o["max_score"], o["best_row"], n
        return out

    def score_all(self, query):
        return self.v @ np.asarray(query, dtype=np.float32).reshape(-1)

    def set_boxes(self, x1, y1, x2, y2, zoom_level):
        self.boxes = pd.DataFrame({"x1": x1, "y1": y1, "x2": x2, "y2": y2, "zoom_level": zoom_level})

    def rescore(self, query, cand_dbidx, *, query2=None, agg_method="avg_score", aug_larger="all"):
        s = self.v @ np.asarray(query, dtype=np.float32).reshape(-1)
        if query2 is not None:
            s = s - self.v @ np.asarray(query2, dtype=np.float32).reshape(-1)
        score, row = np.empty(len(cand_dbidx)), np.empty(len(cand_dbidx), np.int64)
        for i, d in enumerate(cand_dbidx):
            rows = np.flatnonzero(self.dbidx == d)
            frame = (self.boxes.iloc[rows] if self.boxes is not None else pd.DataFrame(index=rows)).assign(score=s[rows])
            pos, val = orc.frame_best_patch(frame, agg_method=agg_method, aug_larger=aug_larger)
            score[i], row[i] = val, rows[pos]
        return score, row

    def topk_from_scores(self, scores, k, exclude=None, row_mask=None):
        rows = np.arange(self.n_rows) if row_mask is None else np.flatnonzero(row_mask)
        order = rows[np.argsort(-np.asarray(scores)[rows], kind="stable")]
        d, s, r = orc.get_top_dbidxs(order, np.asarray(scores)[order], self.dbidx, exclude, k)
        return dict(dbidx=d.astype(np.int32), score=s.astype(np.float32), row=r.astype(np.int64))

    def topk_from_order(self, row_order, k, exclude=None):
        order = np.asarray(row_order, dtype=np.int64)
        d, pos, r = orc.get_top_dbidxs(order, np.arange(len(order)), self.dbidx, exclude, k)
        return dict(dbidx=d.astype(np.int32), pos=pos.astype(np.int64), row=r.astype(np.int64))

    def close(self):
        self.closed = True
