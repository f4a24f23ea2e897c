"""Label propagation over the kNN graph with the iteration on the B200.

  B200LabelPropagation  <->  seesaw/label_propagation.py:6-83 (LabelPropagation): same constructor
                             (weight_matrix: scipy CSR, reg_lambda, max_iter, epsilon, verbose) and the
                             same ``fit_transform(label_ids=, label_values=, reg_values=, start_value=)``.

The weight matrix is what the reference's ``get_weight_matrix`` (seesaw/knn_graph.py:31-104) builds from the
edge table of ``seesaw_b200.knn_graph.compute_exact_knn`` — that one-time scipy step stays on the host.  The
per-feedback-round loop (``_step``: one SpMV, a scale and a clamp per iteration, up to max_iter iterations
over ~2·k·N non-zeros) runs as CUDA kernel K6 in IEEE float64 with scipy's summation order, so the returned
vector is bit-identical to the reference's."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, ptr


class B200LabelPropagation:
    def __init__(self, weight_matrix, *, reg_lambda: float, max_iter: int, epsilon=1e-5, verbose=0, device=0):
        assert reg_lambda >= 0                                   # label_propagation.py:8
        csr = weight_matrix.tocsr() if hasattr(weight_matrix, "tocsr") else weight_matrix
        assert csr.has_sorted_indices                            # :21
        self.weight_matrix = csr
        self.n = csr.shape[0]
        self.epsilon, self.verbose, self.reg_lambda, self.max_iter = epsilon, verbose, reg_lambda, max_iter
        self.reg_values = None
        self.weight_sum = np.asarray(csr.sum(0)).reshape(-1)     # :24 (column sums, exactly as the reference)
        self.iterations, self.converged = 0, False
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        data = np.ascontiguousarray(csr.data, dtype=np.float64)
        wsum = np.ascontiguousarray(self.weight_sum, dtype=np.float64)
        self._h = C.c_void_p()
        check(lib.ssw_lp_create(C.byref(self._h), int(device), self.n, ptr(indptr), ptr(indices), ptr(data), ptr(wsum),
                                float(reg_lambda)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.ssw_lp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fit_transform(self, *, label_ids, label_values, reg_values=None, start_value=None):
        if reg_values is not None:
            assert reg_values.shape[0] == self.n                 # :47
            self.reg_values = reg_values
        else:
            assert self.reg_lambda == 0                          # :50
            self.reg_values = np.zeros(self.n)
        ids = np.ascontiguousarray(np.asarray(label_ids).reshape(-1), dtype=np.int64)
        vals = np.ascontiguousarray(np.broadcast_to(np.asarray(label_values, dtype=np.float64), ids.shape))
        reg = None if reg_values is None else np.ascontiguousarray(reg_values, dtype=np.float64)
        start = None if start_value is None else np.ascontiguousarray(start_value, dtype=np.float64)
        out = np.empty(self.n, np.float64)
        it, conv = C.c_int(), C.c_int()
        check(lib.ssw_lp_fit(self._h, ptr(ids), ptr(vals), len(ids), ptr(reg), ptr(start), int(self.max_iter),
                             float(self.epsilon), ptr(out), C.byref(it), C.byref(conv)))
        self.iterations, self.converged = it.value, bool(conv.value)
        if self.converged and self.verbose > 0:
            print(f"prop. converged after {it.value} iterations")
        if not self.converged:
            print(f"warning: did not converge after {it.value} iterations")   # :80-81
        return out


def normalize_scores(scores, epsilon):
    """research/knn_methods.py:87-95: affine map of the scores onto [epsilon, 1 - epsilon] (all equal -> 0.5)."""
    assert epsilon < 0.5
    lo, hi = scores.min(), scores.max()
    if hi - lo == 0:
        return scores - scores + 0.5
    return (scores - lo) / (hi - lo) * (1 - 2 * epsilon) + epsilon


class B200LabelPropagationRanker:
    """The ``knn_model`` of KnnProp2 (seesaw/loops/graph_based.py:70-113) — LabelPropagationRanker2 with its base
    class (research/knn_methods.py:97-199): prior scores from the text query, user labels clamped, propagation over
    the kNN graph on the GPU (:class:`B200LabelPropagation`), ``top_k`` over the unlabeled rows.  Same constructor
    keywords and methods; ``top_k`` breaks exact score ties by ascending row (the reference's argsort is unstable)."""

    def __init__(self, *, weight_matrix, normalize_scores, sigmoid_before_propagate, calib_a, calib_b, prior_weight,
                 normalize_epsilon=None, verbose=0, device=0, lp_factory=None, **other):
        self.nvecs = weight_matrix.shape[0]
        self.normalize_scores = normalize_scores
        if normalize_scores:
            assert normalize_epsilon is not None
            self.epsilon = normalize_epsilon
        self.calib_a, self.calib_b, self.prior_weight = calib_a, calib_b, prior_weight
        self.sigmoid_before_propagate = sigmoid_before_propagate
        self.is_labeled = np.zeros(self.nvecs)
        self.labels = np.zeros(self.nvecs)
        self.prior_scores = None
        self._current_scores = None
        self.weight_matrix = weight_matrix
        make = lp_factory or (lambda **kw: B200LabelPropagation(device=device, **kw))
        self.lp = make(reg_lambda=prior_weight, weight_matrix=weight_matrix, max_iter=300, verbose=verbose)   # :186-187

    def set_base_scores(self, init_scores):
        assert self.nvecs == init_scores.shape[0]
        if self.normalize_scores:
            init_scores = normalize_scores(init_scores, epsilon=self.epsilon)
        if self.sigmoid_before_propagate:
            from scipy.special import expit          # the reference's sigmoid (research/knn_methods.py:6)
            init_scores = expit(self.calib_a * (init_scores + self.calib_b))
        self.prior_scores = init_scores
        if self.is_labeled.sum() == 0:                      # no labels yet: nothing to propagate (:136-139)
            self._current_scores = self.prior_scores
        else:
            self._current_scores = self._propagate(self.prior_scores)

    def _propagate(self, scores):
        ids = np.flatnonzero(self.is_labeled.reshape(-1))
        return self.lp.fit_transform(label_ids=ids, label_values=self.labels.reshape(-1)[ids],
                                     reg_values=self.prior_scores, start_value=scores)

    def update(self, idxs, labels):
        for idx, label in zip(idxs, labels):
            label = float(label)
            assert np.isclose(label, 0) or np.isclose(label, 1)
            self.labels[int(idx)] = label
            self.is_labeled[int(idx)] = 1
        if (self.labels[self.is_labeled > 0] == 0).sum() > 0:    # the reference waits for a first negative (:153-158)
            self._current_scores = self._propagate(self.prior_scores)

    def current_scores(self):
        return self._current_scores

    def top_k(self, k, unlabeled_only=True):
        subset = np.flatnonzero(self.is_labeled < 1) if unlabeled_only else np.arange(self.nvecs)
        raw = self.current_scores()
        top = subset[np.argsort(-raw[subset], kind="stable")[:k]]
        return top, raw[top]
