"""TEST INFRASTRUCTURE — seeded parity cases shared by oracle/make_golden.py and tests/.

Every case is a pure function of (seed, shape) through seesaw_b200.synth, so fixtures hold only
reference OUTPUTS and the tests regenerate the inputs."""
import numpy as np

from seesaw_b200 import synth

CASES = {
    # name: dict(kind of case + generator parameters)
    "ms_small": dict(n_images=400, p_lo=3, p_hi=9, dim=512, seed=11, qseed=12),
    "ms_wide": dict(n_images=1500, p_lo=20, p_hi=60, dim=512, seed=21, qseed=22),
    "ms_768": dict(n_images=300, p_lo=1, p_hi=5, dim=768, seed=31, qseed=32),
}
# real-index shape: float32 unit vectors that are NOT fp16-representable (multiscale_tools.py:200) and the tiling
# pipeline's float32 boxes (multiscale_tools.py:96-117) — the data the reference's own indices hold
CASES_F32 = {
    "msf_unit": dict(n_images=260, dim=512, seed=61, qseed=62, min_tile_size=60),
    "msf_unit_768": dict(n_images=120, dim=768, seed=63, qseed=64, min_tile_size=120),
}
COARSE = dict(n=10000, dim=512, seed=0, qseed=1, xseed=2, n_excl=300, topk=10)   # BASELINE config 1
KNN = {
    "knn_600": dict(n=600, dim=512, seed=41, k=10),
    "knn_small_n": dict(n=7, dim=512, seed=42, k=10),       # k+1 > N  -> k1 = N (knn_graph.py:172)
    "knn_dups": dict(n=300, dim=512, seed=43, k=5, dup=True),  # duplicate vectors: self may not rank first
}


def ms_inputs(c):
    counts = synth.patches_per_image(c["n_images"], c["p_lo"], c["p_hi"], c["seed"])
    meta = synth.synth_vector_meta(counts, c["seed"] + 1000, dbidx_start=5, dbidx_stride=3)
    vecs = synth.synth_rows(0, int(counts.sum()), c["dim"], c["seed"], "tri", np.float32)
    q = synth.unit_queries(4, c["dim"], c["qseed"])
    return vecs, meta, q


def msf_inputs(c):
    meta, counts = synth.synth_pyramid_meta(c["n_images"], c["seed"] + 1000, dbidx_start=7, dbidx_stride=2,
                                            min_tile_size=c["min_tile_size"],
                                            sizes=((640, 480), (500, 375), (480, 640), (1280, 960), (333, 500), (224, 224), (300, 260)))
    vecs = synth.unit_rows(int(counts.sum()), c["dim"], c["seed"])
    q = synth.unit_queries(4, c["dim"], c["qseed"])
    return vecs, meta, q


MSF_QUERY_VARIANTS = (("plain_score", "all", 1, False), ("plain_score", "all", 3, True), ("avg_score", "all", 3, False),
                      ("avg_score", "greater", 5, False), ("avg_score", "adjacent", 5, True))


def exclude_sets(meta, seed):
    ids = np.unique(meta.dbidx.values)
    rng = np.random.default_rng(seed)
    return {
        "none": np.zeros(0, np.int64),
        "some": np.sort(rng.choice(ids, size=min(30, len(ids) // 2), replace=False)),
        "most": np.sort(rng.choice(ids, size=len(ids) - 7, replace=False)),   # eligible < k
        "all": ids.copy(),                                                    # k' == 0
        "foreign": np.concatenate([ids[:5], np.array([10 ** 6, 10 ** 6 + 1])]),  # ids not in the DB
    }


def knn_inputs(c):
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    v = v / np.linalg.norm(v, axis=1, keepdims=True)
    v = v.astype(np.float16).astype(np.float32)            # fp16-valued, as BASELINE config 4
    if c.get("dup"):
        v[1::10] = v[0::10][: len(v[1::10])]                # exact duplicates
    return v




# label propagation over an exact kNN graph of `n` synthetic vectors (weights: rbf kernel, symmetric)
LP = {
    "lp_reg": dict(n=400, dim=64, seed=51, k=6, edist=0.5, reg_lambda=1.0, max_iter=40, epsilon=1e-7, n_labels=24, lseed=52),
    "lp_noreg": dict(n=250, dim=64, seed=53, k=5, edist=0.3, reg_lambda=0.0, max_iter=15, epsilon=1e-5, n_labels=10, lseed=54),
    "lp_start": dict(n=300, dim=64, seed=55, k=4, edist=1.0, reg_lambda=0.25, max_iter=500, epsilon=1e-9, n_labels=30, lseed=56,
                     start=True),
}


def lp_vectors(c):
    v = synth.synth_rows(0, c["n"], c["dim"], c["seed"], "tri", np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def lp_inputs(c):
    """label ids / values (with one duplicate id: the later value wins), prior scores, optional start vector."""
    rng = np.random.default_rng(c["lseed"])
    ids = rng.choice(c["n"], size=c["n_labels"], replace=False)
    ids = np.concatenate([ids, ids[:1]])
    vals = np.concatenate([(rng.random(c["n_labels"]) < 0.4).astype(np.float64), [1.0]])
    reg = rng.random(c["n"]) if c["reg_lambda"] > 0 else None
    start = rng.random(c["n"]) if c.get("start") else None
    return ids.astype(np.int64), vals, reg, start


# feedback session of the label-propagation ranker: (row ids, 0/1 labels) per round
RANKER_STEPS = [([3, 17], [1, 1]), ([40], [0]), ([5, 77, 120], [1, 0, 0])]
