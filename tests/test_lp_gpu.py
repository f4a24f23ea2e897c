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


@pytest.mark.parametrize("name", list(cases.LP))
def test_device_weight_matrix_equals_reference_golden(golden, name):
    """ssw_weight_matrix (get_weight_matrix(symmetric=True) on the GPU) against the CSR arrays the unmodified reference
    produced (fixture) — indptr, indices and data bit for bit — and W.sum(0) as scipy computes it; then the whole graph
    chain on the device: exact kNN edge table -> weight matrix -> label propagation == the reference's scores."""
    import scipy.sparse as sp
    from seesaw_b200 import knn_graph as kg
    from seesaw_b200.label_propagation import B200LabelPropagation
    c = cases.LP[name]
    v = cases.lp_vectors(c)
    df = orc.compute_exact_knn(v, c["k"])
    W, wsum = kg.get_weight_matrix_device(df, kfun=kg.rbf_kernel(c["edist"]), return_weight_sum=True)
    assert np.array_equal(W.indptr, golden[f"{name}/W_indptr"]) and np.array_equal(W.indices, golden[f"{name}/W_indices"])
    assert np.array_equal(W.data, golden[f"{name}/W_data"]) and W.has_sorted_indices
    ref = sp.csr_array((golden[f"{name}/W_data"], golden[f"{name}/W_indices"], golden[f"{name}/W_indptr"]), shape=W.shape)
    assert np.array_equal(wsum, np.asarray(ref.sum(0)).reshape(-1))
    # hard-cut kernel: zero weights stay in the structure as explicit zeros, like the reference's assignment over symmetric_adj
    host = kg.get_weight_matrix(df, kfun=kg.knn_kernel(0.95), self_edges=False, normalized=False, symmetric=True)
    try:
        dev = kg.get_weight_matrix_device(df, kfun=kg.knn_kernel(0.95))
        assert np.array_equal(dev.indptr, host.indptr) and np.array_equal(dev.indices, host.indices) and np.array_equal(dev.data, host.data)
    except Exception as e:      # a vertex may lose all its weight under the hard cut: both sides refuse
        assert "zero degree" in str(e)
    # label propagation over the device-built matrix
    ids, vals, reg, start = cases.lp_inputs(c)
    lp = B200LabelPropagation(W, reg_lambda=c["reg_lambda"], max_iter=c["max_iter"], epsilon=c["epsilon"])
    got = lp.fit_transform(label_ids=ids, label_values=vals, reg_values=reg, start_value=start)
    assert (got == golden[f"{name}/values"]).all()
    lp.close()


def test_float32_priors_follow_the_reference_arithmetic():
    """KnnProp2 hands float32 priors (index.score) to the propagation: the reference multiplies reg_lambda * reg_values
    in float32 and adds in float64 (label_propagation.py:31).  With prior_weight not a power of two the float64 product
    differs, so the wrapper forms the prior term with numpy, like the reference."""
    import scipy.sparse as sp
    from seesaw_b200 import knn_graph as kg
    from seesaw_b200.label_propagation import B200LabelPropagation
    c = cases.LP["lp_reg"]
    W = kg.get_weight_matrix(orc.compute_exact_knn(cases.lp_vectors(c), c["k"]), kfun=kg.rbf_kernel(c["edist"]), self_edges=False,
                             normalized=False, symmetric=True)
    rng = np.random.default_rng(3)
    reg32 = rng.random(c["n"]).astype(np.float32)
    ids = rng.choice(c["n"], size=20, replace=False)
    vals = (rng.random(20) < 0.5).astype(np.float64)
    lam = 0.3
    lp = B200LabelPropagation(W, reg_lambda=lam, max_iter=60, epsilon=1e-9)
    got = lp.fit_transform(label_ids=ids, label_values=vals, reg_values=reg32, start_value=reg32)
    want, _, _ = orc.label_propagation_fit(W, reg_lambda=lam, max_iter=60, epsilon=1e-9, label_ids=ids, label_values=vals,
                                           reg_values=reg32, start_value=reg32)
    assert (got == want).all()
    assert not (np.float64(lam) * reg32.astype(np.float64) == (lam * reg32).astype(np.float64)).all()     # the case is not vacuous
    lp.close()
