#!/usr/bin/env python
"""Exact kNN graph of a parquet vector dataset on the B200 — same command line as the reference's
scripts/make_knn_graph.py:1-50 (--inputpath, --column, --k, --outputpath; the output path must not
exist), without Ray: the dataset is read with pyarrow and the graph is written as
{outputpath}/forward.parquet (the format KNNGraph.from_file reads, seesaw/knn_graph.py:273-283)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

parser = argparse.ArgumentParser(description="exact k-nn graph of a vector dataset (tcgen05 build on B200)")
parser.add_argument("--column", type=str, default="vectors", help="name of column within dataset")
parser.add_argument("--k", type=int, required=True, help="the k in k-nn: how many neighbors")
parser.add_argument("--inputpath", type=str, required=True, help="Parquet dataset with the vectors")
parser.add_argument("--outputpath", type=str, required=True, help="where to save this")
parser.add_argument("--device", type=int, default=0)


def main(argv=None):
    args = parser.parse_args(argv)
    import numpy as np
    import pyarrow.parquet as pq

    from seesaw_b200.knn_graph import KNNGraph

    inpath = os.path.expandvars(args.inputpath)
    assert os.path.exists(inpath)
    outpath = os.path.expandvars(args.outputpath)
    assert not os.path.exists(outpath), "output path already exists."
    # flatten the Arrow list column without materialising Python objects (a 1M x 512 column is 5e8 of them)
    col = pq.read_table(inpath, columns=[args.column]).column(args.column).combine_chunks()
    n = len(col)
    flat = col.flatten().to_numpy(zero_copy_only=False)
    assert n > 0 and flat.shape[0] % n == 0, "ragged vector column"
    vectors = np.ascontiguousarray(flat.reshape(n, flat.shape[0] // n), dtype=np.float32)
    knng, _ = KNNGraph.from_vectors(vectors, n_neighbors=args.k, device=args.device)
    knng.save(outpath, overwrite=True)
    return knng


if __name__ == "__main__":
    main()
