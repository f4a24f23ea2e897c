"""GPU parity of the label-propagation kernel (K6) through B200LabelPropagation: bit-identical float64 against
the reference's outputs (golden fixtures) and against the oracle on a larger graph built by the GPU kNN kernel."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
import seesaw_oracle as orc
from seesaw_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.LP))
def test_lp_vs_reference_golden(golden, name):
    from seesaw_b200.label_propagation import B200LabelPropagation
    c = cases.LP[name]
    W = sp.csr_array((golden[f"{name}/W_data"], golden[f"{name}/W_indices"], golden[f"{name}/W_indptr"]),
                     shape=(c["n"], c["n"]))
    ids, vals, reg, start = cases.lp_inputs(c)
    lp = B200LabelPropagation(W, reg_lambda=c["reg_lambda"], max_iter=c["max_iter"], epsilon=c["epsilon"])
    got = lp.fit_transform(label_ids=ids, label_values=vals, reg_values=reg, start_value=start)
    assert (got == golden[f"{name}/values"]).all()
    _, it, conv = orc.label_propagation_fit(W, reg_lambda=c["reg_lambda"], max_iter=c["max_iter"], epsilon=c["epsilon"],
                                            label_ids=ids, label_values=vals, reg_values=reg, start_value=start)
    assert lp.iterations == it and lp.converged == conv
    lp.close()


def test_lp_on_gpu_built_graph():
    """kNN graph from K3 -> symmetric rbf weights (host, scipy) -> propagation on device == oracle, bit for bit."""
    from seesaw_b200.knn_graph import compute_exact_knn
    from seesaw_b200.label_propagation import B200LabelPropagation
    n = 20000
    v = synth.synth_rows(0, n, 256, 61, "tri", np.float32)
    v = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float16).astype(np.float32)
    df = compute_exact_knn(v, 8)
    w = np.exp(-df.distance.values.astype(np.float64) / 0.5)
    A = sp.coo_array((w, (df.src_vertex.values, df.dst_vertex.values)), shape=(n, n)).tocsr()
    W = ((A + A.T) * 0.5).tocsr()
    W.setdiag(0.0)
    W.eliminate_zeros()
    W.sort_indices()
    rng = np.random.default_rng(1)
    ids = rng.choice(n, size=200, replace=False).astype(np.int64)
    vals = (rng.random(200) < 0.3).astype(np.float64)
    reg = rng.random(n)
    for lam, iters, eps in ((1.0, 30, 1e-8), (0.1, 7, 1e-12)):
        lp = B200LabelPropagation(W, reg_lambda=lam, max_iter=iters, epsilon=eps)
        got = lp.fit_transform(label_ids=ids, label_values=vals, reg_values=reg)
        want, it, conv = orc.label_propagation_fit(W, reg_lambda=lam, max_iter=iters, epsilon=eps, label_ids=ids,
                                                   label_values=vals, reg_values=reg)
        assert (got == want).all() and lp.iterations == it and lp.converged == conv
        lp.close()
