"""TEST INFRASTRUCTURE — regenerate tests/golden/reference_outputs.npz.

Runs the UNMODIFIED reference (/root/reference, imported through oracle/refstubs.py) on seeded
synthetic inputs and stores its outputs.  Inputs are not stored: every case is a function of
(seed, shape) through seesaw_b200.synth, so the fixtures stay small and tests regenerate inputs.
Only tie-free float data is used here: the reference's np.argsort is unstable, so its output
on exactly tied scores is not a definition.

    python oracle/make_golden.py            # needs /root/reference; the GPU box never runs this
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from refstubs import import_reference  # noqa: E402
from seesaw_b200 import synth  # noqa: E402

from cases import (CASES, CASES_F32, COARSE, KNN, LP, MSF_QUERY_VARIANTS, RANKER_STEPS, ms_inputs, msf_inputs,  # noqa: E402
                   exclude_sets, knn_inputs, lp_vectors, lp_inputs)


def main():
    ref = import_reference()
    BitMap = ref.BitMap
    out = {}
    for name, c in CASES.items():
        vecs, meta, qs = ms_inputs(c)
        idx = ref.multiscale.MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, vec_index=None)
        for xname, ex in exclude_sets(meta, c["seed"] + 7).items():
            for qi in range(2):
                q = qs[qi]
                r = idx._query_prelim(vector=q, topk_dbidx=50, exclude_dbidx=BitMap(ex))
                key = f"{name}/prelim/{xname}/{qi}"
                if isinstance(r, tuple):                   # the reference's "[], [], []" (multiscale_index.py:302)
                    out[key + "/dbidx"] = np.zeros(0, np.int64)
                    out[key + "/score"] = np.zeros(0, np.float32)
                    continue
                out[key + "/dbidx"] = r["dbidx"].values.astype(np.int64)
                out[key + "/score"] = r["max_score"].values.astype(np.float32)
        for agg, topk, use_v2 in (("plain_score", 1, False), ("plain_score", 3, True), ("avg_score", 3, False)):
            if agg == "avg_score" and name == "ms_wide":
                continue                                   # slow and adds nothing
            ex = exclude_sets(meta, c["seed"] + 7)["some"]
            r = idx.query(vector=qs[2], vector2=qs[3] * 0.25 if use_v2 else None, topk=topk,
                          shortlist_size=50, exclude=BitMap(ex), agg_method=agg, aug_larger="all",
                          rescore_method=None)
            key = f"{name}/query/{agg}/{topk}/{int(use_v2)}"
            out[key + "/dbidxs"] = np.asarray(r["dbidxs"]).astype(np.int64)
            out[key + "/act_score"] = np.array([a.score.values[0] for a in r["activations"]], np.float64)
            out[key + "/act_box"] = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.int64)

    # float32 unit vectors + the tiling pipeline's float32 boxes: what a real index holds
    for name, c in CASES_F32.items():
        vecs, meta, qs = msf_inputs(c)
        idx = ref.multiscale.MultiscaleIndex(embedding=None, vectors=vecs, vector_meta=meta, vec_index=None)
        xs = exclude_sets(meta, c["seed"] + 7)
        for xname in ("none", "some"):
            for qi in range(2):
                r = idx._query_prelim(vector=qs[qi], topk_dbidx=50, exclude_dbidx=BitMap(xs[xname]))
                key = f"{name}/prelim/{xname}/{qi}"
                out[key + "/dbidx"] = r["dbidx"].values.astype(np.int64)
                out[key + "/score"] = r["max_score"].values.astype(np.float32)
        for agg, aug, topk, use_v2 in MSF_QUERY_VARIANTS:
            r = idx.query(vector=qs[2], vector2=qs[3] * 0.25 if use_v2 else None, topk=topk, shortlist_size=40,
                          exclude=BitMap(xs["some"]), agg_method=agg, aug_larger=aug, rescore_method=None)
            key = f"{name}/query/{agg}/{aug}/{topk}/{int(use_v2)}"
            out[key + "/dbidxs"] = np.asarray(r["dbidxs"]).astype(np.int64)
            out[key + "/act_score"] = np.array([a.score.values[0] for a in r["activations"]], np.float64)
            out[key + "/act_box"] = np.array([a[["x1", "y1", "x2", "y2"]].values[0] for a in r["activations"]], np.float32)
            assert all(a.score.dtype == np.float32 for a in r["activations"]), "score column dtype"

    c = COARSE
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    import pandas as pd
    cmeta = pd.DataFrame({"dbidx": np.arange(c["n"], dtype=np.int64)})
    cidx = ref.coarse.CoarseIndex(embedding=None, vectors=v, vector_meta=cmeta)
    q = synth.unit_queries(1, c["dim"], c["qseed"])[0]
    ex = np.sort(np.random.default_rng(c["xseed"]).choice(c["n"], size=c["n_excl"], replace=False))
    r = cidx.query(topk=c["topk"], vector=q, exclude=BitMap(ex))
    out["coarse/dbidxs"] = np.asarray(r["dbidxs"]).astype(np.int64)
    out["coarse/scores"] = np.array([a.score.values[0] for a in r["activations"]], np.float32)
    out["coarse/nextstartk"] = np.array([r["nextstartk"]])

    for name, c in KNN.items():
        v = knn_inputs(c)
        df = ref.knn_graph.compute_exact_knn(v, n_neighbors=c["k"])
        for col in ("src_vertex", "dst_vertex", "distance", "dst_rank"):
            out[f"{name}/{col}"] = df[col].values
    # label propagation: the reference's weight matrix (knn_graph.py:31-104) and its fit_transform
    import importlib
    lpmod = importlib.import_module("seesaw.label_propagation")
    for name, c in LP.items():
        v = lp_vectors(c)
        df = ref.knn_graph.compute_exact_knn(v, n_neighbors=c["k"])
        W = ref.knn_graph.get_weight_matrix(df, kfun=ref.knn_graph.rbf_kernel(c["edist"]), self_edges=False,
                                            normalized=False, symmetric=True)
        ids, vals, reg, start = lp_inputs(c)
        lp = lpmod.LabelPropagation(W, reg_lambda=c["reg_lambda"], max_iter=c["max_iter"], epsilon=c["epsilon"])
        res = lp.fit_transform(label_ids=ids, label_values=vals, reg_values=reg, start_value=start)
        out[f"{name}/W_indptr"], out[f"{name}/W_indices"], out[f"{name}/W_data"] = W.indptr, W.indices, W.data
        out[f"{name}/values"] = np.asarray(res, dtype=np.float64)

    # the knn_model of KnnProp2: LabelPropagationRanker2 over a short feedback session (research/knn_methods.py:97-199)
    km = importlib.import_module("seesaw.research.knn_methods")
    c = LP["lp_reg"]
    W = ref.knn_graph.get_weight_matrix(ref.knn_graph.compute_exact_knn(lp_vectors(c), n_neighbors=c["k"]),
                                        kfun=ref.knn_graph.rbf_kernel(c["edist"]), self_edges=False, normalized=False,
                                        symmetric=True)
    rk = km.LabelPropagationRanker2(weight_matrix=W, normalize_scores=False, sigmoid_before_propagate=True, calib_a=2.0,
                                    calib_b=-0.1, prior_weight=1.0)
    base = np.random.default_rng(9).standard_normal(c["n"])
    rk.set_base_scores(base.copy())
    for step, (idxs, labels) in enumerate(RANKER_STEPS):
        rk.update(idxs, labels)
        out[f"ranker/scores/{step}"] = np.asarray(rk.current_scores(), dtype=np.float64)

    # the reference's own unit pin (multiscale_index.py:182-187)
    out["pin/distinct_topk_positions"] = ref.multiscale.distinct_topk_positions(
        np.array([10, 11, 11, 12, 12, 12, 13, 13]), 2)

    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    np.savez_compressed(path, **out)
    with open(os.path.join(ROOT, "tests", "golden", "reference_cases.json"), "w") as f:
        json.dump(dict(multiscale=CASES, multiscale_f32=CASES_F32, coarse=COARSE, knn=KNN, label_propagation=LP,
                       note="outputs of the unmodified reference; regenerate with oracle/make_golden.py"), f, indent=1)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
