"""CPU test of scripts/make_knn_graph.py (same command line as the reference's script, scripts/make_knn_graph.py:1-50)
with the oracle standing in for the GPU builder: argument handling, parquet in, forward.parquet out, refusal to
overwrite an existing output path."""
import importlib.util
import os

import numpy as np
import pandas as pd
import pytest

import seesaw_oracle as orc
from seesaw_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cli():
    spec = importlib.util.spec_from_file_location("make_knn_graph_cli", os.path.join(ROOT, "scripts", "make_knn_graph.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_make_knn_graph_cli(tmp_path, monkeypatch):
    from seesaw_b200 import knn_graph as kg
    monkeypatch.setattr(kg, "compute_exact_knn", lambda vectors, n_neighbors, device=0: orc.compute_exact_knn(vectors, n_neighbors))
    v = synth.synth_rows(0, 120, 256, 3, "tri", np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    inp, out = tmp_path / "vectors.parquet", tmp_path / "graph"
    pd.DataFrame({"dbidx": np.arange(120), "emb": list(v)}).to_parquet(inp)
    cli = load_cli()
    g = cli.main(["--inputpath", str(inp), "--outputpath", str(out), "--k", "4", "--column", "emb"])
    want = orc.compute_exact_knn(v, 4)
    pd.testing.assert_frame_equal(pd.read_parquet(out / "forward.parquet"), want)
    assert g.nvecs == 120 and g.k == 4
    pd.testing.assert_frame_equal(kg.KNNGraph.from_file(str(out)).knn_df, want)
    with pytest.raises(AssertionError):                    # the output path must not exist (reference script :44)
        cli.main(["--inputpath", str(inp), "--outputpath", str(out), "--k", "4", "--column", "emb"])
    with pytest.raises(SystemExit):
        cli.main(["--inputpath", str(inp)])                # --k and --outputpath are required
