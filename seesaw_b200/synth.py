"""Deterministic synthetic patch databases (host side, numpy).

The same counter-based generator exists on the device (``csrc/synth.cu``,
``ssw_db_create_synthetic``) so that a 10M x 512 or 100M x 768 database can be produced
directly in HBM shard by shard while the CPU baseline / oracle sees bit-identical values on
whatever bounded sample it runs on.  Every value is exactly representable in fp16, so the fp16
device copy and the fp32 host copy the reference's numpy path uses hold the same numbers
(the reference stores fp32: indices/multiscale/multiscale_tools.py:200).

Element (row, 4c..4c+3) comes from one 64-bit splitmix finaliser of ``row * (dim/4) + c``:
byte pairs (b0+b1, b2+b3, b4+b5, b6+b7) of the hash, little-endian.
  kind "tri"     : (pair_sum - 255) / 2048   -> triangular, std 0.051, |row| ~ 1.15 at d=512
  kind "lattice" : (pair_sum % 9 - 4) / 8    -> entries k/8, every partial dot product of two
                   such rows is a multiple of 1/64 below 2^24/64, i.e. EXACT in fp32 in any
                   summation order; produces massive exact score ties (SURVEY.md §7 hard part A)
"""
from __future__ import annotations

import numpy as np

KIND_TRI = 0
KIND_LATTICE = 1
_KINDS = {"tri": KIND_TRI, "lattice": KIND_LATTICE}

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def kind_id(kind) -> int:
    return _KINDS[kind] if isinstance(kind, str) else int(kind)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def synth_rows(row_start: int, n_rows: int, dim: int, seed: int, kind="tri", dtype=np.float32):
    """Rows [row_start, row_start+n_rows) of the synthetic database, as ``dtype``."""
    assert dim % 4 == 0
    q = dim // 4
    with np.errstate(over="ignore"):
        ctr = (np.arange(row_start * q, (row_start + n_rows) * q, dtype=np.uint64)
               + np.uint64(seed) * _GOLDEN)
        h = _mix(ctr)
    b = h.view(np.uint8).reshape(n_rows, q, 4, 2).astype(np.int16)
    pair = b[..., 0] + b[..., 1]
    if kind_id(kind) == KIND_TRI:
        vals = (pair - 255).astype(np.float32) / np.float32(2048.0)
    else:
        vals = (pair % 9 - 4).astype(np.float32) / np.float32(8.0)
    return vals.reshape(n_rows, dim).astype(dtype)


def patches_per_image(n_images: int, lo: int, hi: int, seed: int):
    """Patch counts per image, uniform in [lo, hi] (integers(20, 61))."""
    if lo == hi:
        return np.full(n_images, lo, dtype=np.int64)
    return np.random.default_rng(seed).integers(lo, hi + 1, size=n_images).astype(np.int64)


def dbidx_of_rows(counts, dbidx_start=0, dbidx_stride=1):
    """Image id of every row, rows grouped by image, ids ascending."""
    ids = dbidx_start + dbidx_stride * np.arange(len(counts), dtype=np.int64)
    return np.repeat(ids, counts).astype(np.int32)


def unit_queries(n: int, dim: int, seed: int):
    """L2-normalised fp32 Gaussian query vectors (string2vec normalises: multiscale_index.py:279-282)."""
    q = np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)
    return q / np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)


def lattice_queries(n: int, dim: int, seed: int):
    """Queries with entries in {-4..4}/8 (exact-arithmetic companion of kind 'lattice')."""
    r = np.random.default_rng(seed).integers(-4, 5, size=(n, dim))
    return (r / 8.0).astype(np.float32)


def synth_vector_meta(counts, seed: int, dbidx_start=0, dbidx_stride=1):
    """A ``vector_meta`` frame (dbidx, zoom_level, x1, y1, x2, y2) shaped like the multiscale
    tiling output (multiscale_tools.py:96-117): per image a coarse-to-fine pyramid of square
    tiles on a 32-px grid, so boxes of different zoom levels overlap (needed by ``avg_score``)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    counts = np.asarray(counts, dtype=np.int64)
    n = int(counts.sum())
    dbidx = dbidx_of_rows(counts, dbidx_start, dbidx_stride)
    pos = np.arange(n) - np.repeat(np.cumsum(counts) - counts, counts)   # index within image
    zoom = (pos % 3 + 1).astype(np.int64)
    side = 224 * zoom
    x1 = rng.integers(0, 8, size=n) * 32
    y1 = rng.integers(0, 8, size=n) * 32
    return pd.DataFrame({"dbidx": dbidx.astype(np.int64), "zoom_level": zoom,
                         "x1": x1.astype(np.int64), "y1": y1.astype(np.int64),
                         "x2": (x1 + side).astype(np.int64), "y2": (y1 + side).astype(np.int64)})
