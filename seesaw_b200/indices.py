"""Reference-facing index classes: same constructor registry, attributes, method names, keyword
arguments and result shapes as the reference's plug points, with the scan running on the B200.

  B200MultiscaleIndex  <->  seesaw/indices/multiscale/multiscale_index.py:201-376 (MultiscaleIndex)
  B200CoarseIndex      <->  seesaw/indices/coarse/coarse_index.py:16-108        (CoarseIndex)
  B200VectorIndex      <->  seesaw/vector_index.py:44-60                         (VectorIndex, the ANN slot)
  InteractiveQuery     <->  seesaw/query_interface.py:7-52

To switch an on-disk index over, set ``"constructor": "seesaw_b200.indices.B200MultiscaleIndex"`` in
its ``info.json`` (seesaw/indices/interface.py:36-45); see INTEGRATION.md.  None of these classes
scans on the CPU: without the CUDA library / a B200 they raise.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd

from .bitmap import BitMap, FrozenBitMap, as_id_array
from .engine import PatchDatabase
from .rescore import rescore_candidates

try:  # inside the reference's environment: be a real subclass of its plug-in base classes
    from seesaw.indices.interface import AccessMethod as _AccessMethodBase  # type: ignore
except Exception:  # pragma: no cover - standalone use
    class _AccessMethodBase:  # mirrors seesaw/indices/interface.py:10-45
        path: str = None

        def get_knng_path(self, name: str = None):
            return f"{self.path}/knn_graph/{'' if name is None else name}"

        @staticmethod
        def load(index_path: str, *, options: dict = None, exclude=None):
            meta = json.load(open(f"{index_path}/info.json", "r"))
            pieces = meta["constructor"].split(".")
            import importlib
            cons = getattr(importlib.import_module(".".join(pieces[:-1])), pieces[-1])
            return cons.from_path(index_path, **(options or {}), exclude=exclude)


AccessMethod = _AccessMethodBase


class InteractiveQuery:
    """Tracks what has been returned and passes it as ``exclude`` (query_interface.py:34-49)."""

    def __init__(self, index):
        self.index = index
        self.returned = BitMap()
        self.label_db = None

    def query_stateful(self, *args, **kwargs):
        batch_size = kwargs.pop("batch_size")
        res = self.index.query(*args, topk=batch_size, **kwargs, exclude=self.returned)
        self.returned.update(np.asarray(res["dbidxs"]).astype(np.int64).tolist())
        return res


def _read_vectors_parquet(path, columns=None):
    """``vectors.sorted.cached`` / ``vectors`` parquet dataset -> DataFrame (the reference goes
    through Ray: services.py:25-30, util.py:110-128)."""
    import pyarrow.parquet as pq
    return pq.read_table(path, columns=columns).to_pandas()


def _column_to_matrix(col) -> np.ndarray:
    if hasattr(col, "to_numpy") and getattr(col.dtype, "name", "") != "object":
        arr = col.to_numpy()
        if arr.ndim == 2:
            return np.ascontiguousarray(arr, dtype=np.float32)
    return np.ascontiguousarray(np.stack([np.asarray(v, dtype=np.float32) for v in col]))


_ROW0 = pd.RangeIndex(1)
_ACT_COLS = pd.Index(["x1", "y1", "x2", "y2", "dbidx", "score"])


def _frame_of_columns(arrays):
    """DataFrame(x1, y1, x2, y2, dbidx, score) over equally long 1-D arrays.  pandas' dict constructor spends ~110 us
    on sanitising six columns; its own array-level constructor (what its readers use) needs 20 us.  It is private API,
    so any surprise falls back to the public constructor."""
    try:
        return pd.DataFrame._from_arrays(arrays, columns=_ACT_COLS, index=pd.RangeIndex(len(arrays[0])), verify_integrity=False)
    except Exception:      # pragma: no cover - a pandas without that entry point
        return pd.DataFrame(dict(zip(_ACT_COLS, arrays)))


def _activation_frames(x1, y1, x2, y2, dbidx, score):
    """The reference returns one single-row DataFrame(x1, y1, x2, y2, dbidx, score) per hit
    (multiscale_index.py:392-397, coarse_index.py:87-92).  Building them one by one costs ~0.15 ms each in pandas
    (from_records); one frame for all hits, sliced into single rows that get a fresh RangeIndex (no per-row
    reset_index copy), gives the same frames at ~20 us each."""
    big = _frame_of_columns([np.ascontiguousarray(a).reshape(-1) for a in (x1, y1, x2, y2, dbidx, score)])
    out = []
    for i in range(len(big)):
        f = big.iloc[i:i + 1]
        f.index = _ROW0
        out.append(f)
    return out


class _GpuIndexMixin:
    """Device copy + host CSR shared by the multiscale and coarse classes."""

    def _init_device(self, device, store, db=None, exact="auto"):
        """``store``: HBM type of the scanned copy ("f16" halves the bytes of every scan, "f32" is the
        reference's own type).  ``exact`` (fp16 storage of float32 vectors): "auto" keeps the float32 rows in
        HBM as well whenever fp16 would change a value, so every result is the reference's float32 result
        (certified re-ranking, PatchDatabase.from_arrays); False scans and rescoring on the fp16 values alone
        (scores within ||v - fp16(v)||.||q|| <= 2^-11 ||v||.||q|| of the reference's, ids may differ between
        images closer than that — the "looser bound" of fp16 storage)."""
        dbidx = self.vector_meta["dbidx"].to_numpy()
        self._dbidx_of_row = dbidx.astype(np.int64)
        self.device = device
        self.store = store
        self._exact_opt = exact
        self._batcher = None
        self._all_ids = np.sort(as_id_array(self.all_indices))     # eligible image ids, for the top-k clamp
        if db is not None:
            # serving-only: the vectors already live in HBM (from_database); no host copy is kept
            assert db.n_rows == len(dbidx)
            self.db, self._store_exact = db, True
        else:
            v = self.vectors
            f32_store = store in ("f32", "fp32", "float32")
            # the HBM copy holds the reference's values exactly: fp32 storage, or fp16-representable data
            same = f32_store or v.dtype == np.float16 or bool(
                (v[: min(len(v), 4096)].astype(np.float16).astype(np.float32) == v[: min(len(v), 4096)]).all()
                and (v.astype(np.float16).astype(np.float32) == v).all())
            keep_f32 = (not same) and exact in ("auto", True, "device")
            self.db = PatchDatabase.from_arrays(v, dbidx.astype(np.int32), store=store, device=device, exact=keep_f32)
            self._store_exact = same or keep_f32
        # host CSR over ORIGINAL rows: rows of image i are _rows_sorted[_starts[i]:_starts[i+1]], ascending
        order = np.argsort(self._dbidx_of_row, kind="stable")
        sorted_ids = self._dbidx_of_row[order]
        self._img_ids, self._starts = np.unique(sorted_ids, return_index=True)
        self._starts = np.append(self._starts, len(order))
        self._rows_sorted = order

    def _rows_of(self, dbidx):
        i = np.searchsorted(self._img_ids, dbidx)
        return self._rows_sorted[self._starts[i]:self._starts[i + 1]]

    def attach_batcher(self, batcher):
        """Route stage-1 scans through a :class:`seesaw_b200.service.ScanBatcher` shared by concurrent
        sessions (one GPU pass per batch instead of one per session)."""
        self._batcher = batcher

    def _map_rows(self, rows):
        """device-reported ORIGINAL rows -> positions in this index's vectors / vector_meta (identity, except
        for a subset view that shares its parent's database)."""
        return rows

    def _always_excluded(self):
        return None

    def _scan_one(self, qvec, k, ex):
        extra = self._always_excluded()
        if extra is not None:
            ex = np.concatenate([np.asarray(ex, dtype=np.int64).reshape(-1), extra])
        if self._batcher is not None:
            r = self._batcher.scan_topk_one(qvec, k, ex)
            return dict(dbidx=r["dbidx"], score=r["score"], row=self._map_rows(r["row"]))
        r = self.db.scan_topk(qvec.reshape(1, -1), k, exclude=[ex])
        n = int(r["count"][0])
        return dict(dbidx=r["dbidx"][0, :n], score=r["score"][0, :n], row=self._map_rows(r["row"][0, :n]))

    def score(self, vec):
        """Full score vector in original row order (multiscale_index.py:284-285, coarse_index.py:37-38)."""
        return self.db.score_all(np.asarray(vec, dtype=np.float32).reshape(-1))

    def top_dbidxs(self, *, vec_idxs, scores, exclude=None, topk):
        """``_get_top_dbidxs(vec_idxs=, scores=, vector_meta=, exclude=, topk=)`` (multiscale_index.py:189-199)
        for scores that did not come from this index's scan — KnnProp2.next_batch passes the rows sorted by
        label-propagation score (loops/graph_based.py:97-99).  Like the reference it walks the rows IN THE GIVEN
        ORDER (``vec_idxs`` best first, ``scores`` aligned with it): per image the earliest-listed row, images ranked
        by that position, excluded images dropped, the first ``topk`` — on the device, independent of the scores'
        dtype (float64 propagation scores are not squeezed into float32).  Returns DataFrame(dbidx, max_score, best_row)."""
        vec_idxs = np.asarray(vec_idxs, dtype=np.int64).reshape(-1)
        scores = np.asarray(scores).reshape(-1)
        r = self.db.topk_from_order(self._unmap_rows(vec_idxs), int(topk), exclude=as_id_array(exclude))
        return pd.DataFrame({"dbidx": r["dbidx"].astype(np.int64), "max_score": scores[r["pos"]],
                             "best_row": self._map_rows(r["row"])})

    def _unmap_rows(self, rows):
        """positions in this index's vectors -> ORIGINAL rows of the database (identity, except for a shared subset)."""
        return rows

    def string2vec(self, string: str) -> np.ndarray:
        init_vec = self.embedding.from_string(string=string)
        return init_vec / np.linalg.norm(init_vec)

    def close(self):
        self.db.close()


class B200MultiscaleIndex(_GpuIndexMixin, AccessMethod):
    """Two-stage lookup (scan -> per-image max -> exclusion -> shortlist -> rescore) with stage 1 on
    the GPU.  Constructor and attributes follow MultiscaleIndex (multiscale_index.py:203-231)."""

    def __init__(self, *, embedding, vectors: np.ndarray, vector_meta: pd.DataFrame, vec_index=None,
                 min_zoom_level=1, path: str = None, excluded=None, device: int = 0, store: str = "f16",
                 exact="auto", _db=None):
        self.embedding = embedding
        self.path = path
        self.excluded = BitMap([]) if excluded is None else BitMap(as_id_array(excluded))
        if min_zoom_level != 1 and _db is None:   # multiscale_index.py:224-231
            keep = (vector_meta["zoom_level"] >= min_zoom_level).to_numpy()
            vector_meta = vector_meta[keep].reset_index(drop=True)
            vectors = vectors[keep]
        self.vectors = None if vectors is None else np.ascontiguousarray(vectors)
        self.vector_meta = vector_meta
        self.vec_index = vec_index          # accepted for interface parity; the exact GPU scan supersedes it
        self.all_indices = FrozenBitMap(self.vector_meta["dbidx"].to_numpy()) - self.excluded
        self._init_device(device, store, db=_db, exact=exact)
        self._meta_cols = {c: self.vector_meta[c].to_numpy() for c in ("x1", "y1", "x2", "y2", "zoom_level")
                           if c in self.vector_meta.columns}
        # boxes travel in their own dtype (float32 from the tiling pipeline, multiscale_tools.py:111): K7 repeats
        # the reference's IoU arithmetic in that type
        self._boxes_on_device = len(self._meta_cols) == 5 and all(
            np.issubdtype(v.dtype, np.number) for v in self._meta_cols.values())
        if self._boxes_on_device:
            self.db.set_boxes(*[self._meta_cols[c] for c in ("x1", "y1", "x2", "y2", "zoom_level")])

    @staticmethod
    def from_database(db: PatchDatabase, vector_meta: pd.DataFrame, *, embedding=None, path=None):
        """Serving-only index over vectors that already live in HBM (a :class:`PatchDatabase`, e.g. one shard
        generated or loaded on the device): both query stages run on the GPU, so no host copy of the
        vectors is kept (``vectors`` is None; ``get_data`` / host rescoring are unavailable)."""
        return B200MultiscaleIndex(embedding=embedding, vectors=None, vector_meta=vector_meta, path=path,
                                   device=db.device, store="f16" if db.dtype == np.float16 else "f32", _db=db)

    @staticmethod
    def from_path(index_path: str, *, use_vec_index=False, device=0, store="f16", exact="auto", exclude=None, **options):
        """multiscale_index.py:234-269; reads ``info.json`` and ``vectors.sorted.cached`` directly
        (no Ray).  The embedding model is resolved lazily by the caller through ``embedding=``."""
        info = json.load(open(f"{index_path}/info.json"))
        df = _read_vectors_parquet(f"{index_path}/vectors.sorted.cached").reset_index(drop=True)
        meta = df[["dbidx", "zoom_level", "x1", "y1", "x2", "y2"]]
        vectors = _column_to_matrix(df["vectors"])
        return B200MultiscaleIndex(embedding=options.get("embedding"), vectors=vectors, vector_meta=meta,
                                   path=index_path, excluded=info.get("excluded", None), device=device, store=store,
                                   exact=exact)

    def __len__(self):
        return len(self.all_indices)

    # ---- stage 1 ---------------------------------------------------------------------------
    def _prelim_arrays(self, qvec, topk_dbidx, exclude_dbidx):
        """Stage 1 as plain arrays: dict(dbidx, score, row) best first, or None when nothing is eligible.
        The clamp of multiscale_index.py:295-298 needs only the NUMBER of eligible images."""
        ex = as_id_array(exclude_dbidx)
        pos = np.searchsorted(self._all_ids, ex)
        pos[pos == len(self._all_ids)] = 0
        n_excluded = int(np.unique(ex[self._all_ids[pos] == ex]).shape[0]) if len(ex) and len(self._all_ids) else 0
        k = min(int(topk_dbidx), len(self._all_ids) - n_excluded)
        if k <= 0:
            return None
        return self._scan_one(qvec, k, ex)

    def _query_prelim(self, *, vector, topk_dbidx, exclude_dbidx=None, force_exact=False):
        """multiscale_index.py:291-312.  Returns DataFrame(dbidx, max_score) like the reference
        (plus best_row); an EMPTY frame when nothing is eligible (the reference returns the tuple
        ``[], [], []`` there, which its own caller cannot use)."""
        r = self._prelim_arrays(np.asarray(vector, dtype=np.float32).reshape(-1), topk_dbidx, exclude_dbidx)
        if r is None:
            return pd.DataFrame({"dbidx": np.zeros(0, np.int64), "max_score": np.zeros(0, np.float32),
                                 "best_row": np.zeros(0, np.int64)})
        return pd.DataFrame({"dbidx": r["dbidx"].astype(np.int64), "max_score": r["score"], "best_row": r["row"]})

    # ---- stage 1 + 2 -----------------------------------------------------------------------
    def query(self, *, vector, vector2=None, topk, shortlist_size, exclude=None, force_exact=False, **kwargs):
        """multiscale_index.py:314-352."""
        if shortlist_size is None:
            shortlist_size = topk * 5
        qvec = np.asarray(vector, dtype=np.float32).reshape(-1)
        cand = self._prelim_arrays(qvec, shortlist_size, exclude)
        if cand is None:
            return {"dbidxs": np.zeros(0, dtype="int"), "activations": []}
        agg_method = kwargs.get("agg_method", "plain_score")
        by_id = np.argsort(cand["dbidx"], kind="stable")               # the reference walks images in ascending dbidx (:388)
        ids = cand["dbidx"][by_id].astype(np.int64)

        def result(order, rows, scores):
            r = np.asarray(rows)[order]
            acts = _activation_frames(self._meta_cols["x1"][r], self._meta_cols["y1"][r], self._meta_cols["x2"][r],
                                      self._meta_cols["y2"][r], ids[order], np.asarray(scores)[order])
            return {"dbidxs": ids[order].astype("int"), "activations": acts}

        # Both stages run on the device.  With float32 vectors in fp16 storage the database keeps the float32 rows
        # too (exact mode): stage 1 is the certified float32 top-k and stage 2 reads the float32 rows, so scores
        # are the reference's.  ``device_rescore=False`` forces the host mirror of stage 2 (cross-check).
        on_device = kwargs.get("device_rescore", True)
        if agg_method == "plain_score" and vector2 is None and on_device:
            # 'plain_score' rescoring recomputes exactly what stage 1 already returned — per image the max
            # patch score and the first row attaining it (:117-118) — so the shortlist only needs the
            # reference's final ordering: ascending dbidx, then a stable sort by score (:388-399).
            sc = cand["score"][by_id]
            return result(np.argsort(-sc.astype(np.float64), kind="stable")[:topk], cand["row"][by_id], sc)
        aug_larger = kwargs.get("aug_larger", "all")
        if (agg_method in ("plain_score", "avg_score") and aug_larger in ("all", "greater", "adjacent")
                and on_device and (agg_method == "plain_score" or self._boxes_on_device)):
            # stage 2 on the device (K7): patch scores, IoU join and per-level averaging for the <= shortlist images
            sc, rows = self.db.rescore(qvec, ids, query2=vector2, agg_method=agg_method, aug_larger=aug_larger)
            rows = self._map_rows(rows)
            sc = sc.astype(np.float32)                                       # the reference's score column is float32
            return result(np.argsort(-sc.astype(np.float64), kind="stable")[:topk], rows, sc)   # stable by score over ascending dbidx (:388-399)
        assert self.vectors is not None, "host rescoring needs the host copy of the vectors (index built with from_database)"
        groups = [self._rows_of(d) for d in ids]                      # CSR ranges, not an O(N) isin
        rows = np.concatenate(groups)
        sub = self.vectors[rows]
        scores = sub @ qvec                                            # :345
        if vector2 is not None:
            scores = scores - sub @ np.asarray(vector2, dtype=np.float32).reshape(-1)   # :347-349
        cuts = np.cumsum([len(g) for g in groups])[:-1]
        return rescore_candidates(groups, ids, np.split(scores, cuts), self._meta_cols, topk,
                                  agg_method=kwargs.get("agg_method", "plain_score"),
                                  aug_larger=kwargs.get("aug_larger", "all"))

    def new_query(self):
        return InteractiveQuery(self)

    def get_data(self, dbidx) -> pd.DataFrame:
        rows = self._rows_of(dbidx)
        return self.vector_meta.iloc[rows].assign(vectors=list(self.vectors[rows]))

    def subset(self, indices, share_device=False):
        """multiscale_index.py:364-376 — a new index over the rows of the given images.  With
        ``share_device=True`` the subset keeps using THIS index's database in HBM and only masks the other
        images out of every scan (a device-side image mask instead of a second copy of the vectors); the
        host-side ``vectors`` / ``vector_meta`` of the subset are renumbered like the reference's."""
        mask = np.isin(self._dbidx_of_row, as_id_array(indices))
        if mask.all():
            return self
        if share_device:
            return _SharedSubset(self, mask)
        assert self.vectors is not None, "a copying subset needs the host vectors; use share_device=True"
        return B200MultiscaleIndex(embedding=self.embedding, vectors=self.vectors[mask],
                                   vector_meta=self.vector_meta[mask].reset_index(drop=True),
                                   device=self.device, store=self.store, exact=self._exact_opt)


class _SharedSubset(B200MultiscaleIndex):
    """Subset view over the parent's device database (see B200MultiscaleIndex.subset)."""

    def __init__(self, parent, row_mask):
        self._parent = parent
        self.embedding, self.path, self.vec_index = parent.embedding, parent.path, None
        self.excluded = parent.excluded
        self._parent_rows = np.flatnonzero(row_mask)                       # subset position -> parent row
        self.vectors = None if parent.vectors is None else parent.vectors[row_mask]
        self.vector_meta = parent.vector_meta[row_mask].reset_index(drop=True)
        self._dbidx_of_row = parent._dbidx_of_row[row_mask]
        keep = np.unique(self._dbidx_of_row)
        self.all_indices = FrozenBitMap(keep) - self.excluded
        self._all_ids = np.sort(as_id_array(self.all_indices))
        self._complement = np.setdiff1d(parent._img_ids, keep).astype(np.int64)   # masked out of every scan
        self.db, self.device, self.store = parent.db, parent.device, parent.store
        self._batcher, self._store_exact, self._exact_opt = None, parent._store_exact, parent._exact_opt
        order = np.argsort(self._dbidx_of_row, kind="stable")
        self._img_ids, self._starts = np.unique(self._dbidx_of_row[order], return_index=True)
        self._starts = np.append(self._starts, len(order))
        self._rows_sorted = order
        self._meta_cols = {c: v[row_mask] for c, v in parent._meta_cols.items()}
        self._boxes_on_device = parent._boxes_on_device

    def _map_rows(self, rows):
        return np.searchsorted(self._parent_rows, np.asarray(rows, dtype=np.int64))

    def _always_excluded(self):
        return self._complement

    def score(self, vec):
        return self._parent.score(vec)[self._parent_rows]

    def _unmap_rows(self, rows):
        return self._parent_rows[np.asarray(rows, dtype=np.int64)]

    def top_dbidxs(self, *, vec_idxs, scores, exclude=None, topk):
        ex = np.concatenate([as_id_array(exclude), self._complement])
        return super().top_dbidxs(vec_idxs=vec_idxs, scores=scores, exclude=ex, topk=topk)

    def subset(self, indices, share_device=True):
        keep_rows = np.isin(self._parent._dbidx_of_row, np.intersect1d(as_id_array(indices), self._img_ids))
        return _SharedSubset(self._parent, keep_rows)

    def close(self):          # the database belongs to the parent
        pass


class B200CoarseIndex(_GpuIndexMixin, AccessMethod):
    """One vector per image (coarse_index.py:16-108)."""

    def __init__(self, embedding, vectors: np.ndarray, vector_meta: pd.DataFrame, path: str = None,
                 device: int = 0, store: str = "f16"):
        self.path = path
        self.embedding = embedding
        self.vectors = np.ascontiguousarray(vectors)
        self.vector_meta = vector_meta
        self.all_indices = FrozenBitMap(self.vector_meta["dbidx"].to_numpy())
        self._init_device(device, store)

    @staticmethod
    def from_path(index_path: str, *, use_vec_index=False, device=0, store="f16", exclude=None, **options):
        df = _read_vectors_parquet(f"{index_path}/vectors")
        assert df.dbidx.is_monotonic_increasing, "sanity check"        # coarse_index.py:49
        return B200CoarseIndex(embedding=options.get("embedding"), vectors=_column_to_matrix(df["vectors"]),
                               vector_meta=df.drop("vectors", axis=1), path=index_path, device=device, store=store)

    def __len__(self):
        return len(self.all_indices)

    def query(self, *, topk, vector=None, exclude=None, startk=None, **kwargs):
        """coarse_index.py:57-96.  ``vector=None`` ranks by N(0,1) noise like the reference (:70-71);
        that branch needs no scan and is drawn on the host."""
        ex = as_id_array(exclude)
        pos = np.searchsorted(self._all_ids, ex)
        pos[pos == len(self._all_ids)] = 0
        n_excluded = int(np.unique(ex[self._all_ids[pos] == ex]).shape[0]) if len(ex) and len(self._all_ids) else 0
        n_included = len(self._all_ids) - n_excluded                      # |all_indices - exclude| (:60)
        if n_included == 0:
            return np.array([]), np.array([])                          # :61-62
        topk = min(int(topk), n_included)
        if vector is None:
            included = np.setdiff1d(self._all_ids, ex)
            scores = np.random.randn(included.shape[0])
            best = np.argsort(-scores, kind="stable")[:topk]
            ret, sc = included[best], scores[best]
        else:
            r = self._scan_one(np.asarray(vector, dtype=np.float32).reshape(-1), topk, ex)
            ret, sc = r["dbidx"].astype(np.int64), r["score"]
        assert ret.shape[0] == topk and len(set(ret.tolist())) == topk                 # :81-85
        assert np.intersect1d(ret, ex).shape[0] == 0
        n = len(ret)
        acts = _activation_frames(np.zeros(n, np.int64), np.zeros(n, np.int64), np.full(n, 224, np.int64),
                                  np.full(n, 224, np.int64), ret, sc)
        return {"dbidxs": ret, "nextstartk": len(ex) + ret.shape[0], "activations": acts}

    def new_query(self):
        return InteractiveQuery(self)

    def subset(self, indices):
        mask = np.isin(self._dbidx_of_row, as_id_array(indices))
        return B200CoarseIndex(embedding=self.embedding, vectors=self.vectors[mask],
                               vector_meta=self.vector_meta[mask].reset_index(drop=True),
                               device=self.device, store=self.store)


class B200VectorIndex:
    """Exact replacement for the annoy wrapper in the ``vec_index`` slot (vector_index.py:44-60):
    ``query(vector, top_k) -> (row indices, scores)`` best-first, recall 1.0, any supported dim."""

    def __init__(self, *, vectors=None, load_path=None, prefault=False, device=0, store="f16"):
        if vectors is None:
            vectors = np.load(load_path)
        self.dim = vectors.shape[1]
        self.db = PatchDatabase.from_arrays(vectors, np.arange(vectors.shape[0], dtype=np.int32),
                                            store=store, device=device)

    def ready(self):
        return True

    def query(self, vector, top_k):
        vector = np.asarray(vector, dtype=np.float32)
        assert vector.shape == (1, self.dim) or vector.shape == (self.dim,)      # vector_index.py:56
        r = self.db.scan_topk(vector.reshape(1, -1), int(top_k))
        n = int(r["count"][0])
        return r["row"][0, :n].copy(), r["score"][0, :n].copy()
