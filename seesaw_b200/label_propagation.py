"""Label propagation over the kNN graph with the iteration on the B200.

  B200LabelPropagation  <->  seesaw/label_propagation.py:6-83 (LabelPropagation): same constructor
                             (weight_matrix: scipy CSR, reg_lambda, max_iter, epsilon, verbose) and the
                             same ``fit_transform(label_ids=, label_values=, reg_values=, start_value=)``.

The weight matrix is what the reference's ``get_weight_matrix`` (seesaw/knn_graph.py:31-104) builds from the
edge table of ``seesaw_b200.knn_graph.compute_exact_knn`` — that one-time scipy step stays on the host.  The
per-feedback-round loop (``_step``: one SpMV, a scale and a clamp per iteration, up to max_iter iterations
over ~2·k·N non-zeros) runs as CUDA kernel K6 in IEEE float64 with scipy's summation order, so the returned
vector is bit-identical to the reference's.  The rankers that drive it (research/knn_methods.py, out of this build's
scope) are used as they are: :func:`use_in_reference` rebinds the one class they instantiate."""
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
        """label_propagation.py:45-83.  The prior term ``reg_lambda * reg_values`` is formed HERE with numpy, i.e. in
        the dtype of ``reg_values`` exactly as the reference's ``_step`` forms it (:31 — float32 priors from
        ``index.score`` give a float32 product that is then added in float64), and handed to the kernel ready-made."""
        n = self.n
        if reg_values is not None:
            assert reg_values.shape[0] == n                      # :47
            self.reg_values = reg_values
        else:
            assert self.reg_lambda == 0                          # :50
            self.reg_values = np.zeros(n)
        if start_value is not None:                              # :53-58
            x0 = start_value.copy()
        elif reg_values is not None:
            x0 = self.reg_values.copy()
        else:
            x0 = np.zeros(n)
        ids = np.ascontiguousarray(np.asarray(label_ids).reshape(-1), dtype=np.int64)
        x0[ids] = label_values                                   # :60 (in x0's own dtype, like the reference)
        vals = np.ascontiguousarray(np.broadcast_to(np.asarray(label_values, dtype=np.float64), ids.shape))
        lreg = np.ascontiguousarray(np.asarray(self.reg_lambda * self.reg_values, dtype=np.float64).reshape(-1))
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1)
        out = np.empty(n, np.float64)
        it, conv = C.c_int(), C.c_int()
        check(lib.ssw_lp_fit_scaled(self._h, ptr(ids), ptr(vals), len(ids), ptr(lreg), ptr(x0), int(self.max_iter),
                                    float(self.epsilon), ptr(out), C.byref(it), C.byref(conv)))
        self.iterations, self.converged = it.value, bool(conv.value)
        # every iterate is a weighted average of neighbours and prior (:35-40); the reference asserts it per step
        low, high = min(0, self.reg_values.min()), max(1., self.reg_values.max())
        assert (out >= low).all(), "averaged scores should lie at or above 0"
        assert (out <= high).all(), "averaged scores should lie at or below 1"
        if self.converged and self.verbose > 0:
            print(f"prop. converged after {it.value} iterations")
        if not self.converged:
            print(f"warning: did not converge after {it.value} iterations")   # :80-81
        return out


def use_in_reference():
    """Inside the reference's environment: make its label-propagation rankers run their loop on the GPU.
    ``LabelPropagationRanker2`` (the knn_model of KnnProp2, seesaw/research/knn_methods.py:176-199) builds
    ``LabelPropagation(reg_lambda=, weight_matrix=, max_iter=, verbose=)`` from its module's namespace; rebinding that
    one name swaps the loop and nothing else.  Returns the previous binding."""
    import seesaw.research.knn_methods as km       # noqa: raises outside the reference's environment
    previous = km.LabelPropagation
    km.LabelPropagation = B200LabelPropagation
    return previous
