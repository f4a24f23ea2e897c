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


def pyramid_tiling(width: int, height: int, tile_size: int = 224, factor: float = 0.5, min_tile_size: int = 60):
    """Patch boxes of one image as the reference's tiling pipeline lays them out (multiscale_tools.py:16-117:
    ``pyramid`` -> ``strided_tiling`` per level -> pixel / scale_factor): a geometric pyramid of rescaled
    copies from the smallest scale (whole image ~ one tile) to the largest, each cut into tile_size squares on
    a half-tile stride.  Pure arithmetic on the image SIZE (no pixels); returns (x1, y1, x2, y2 as float32 in
    original-image pixels — the dtype the pipeline writes, :111 — and zoom_level as int16), rows in the
    pipeline's order.  tests/test_oracle.py checks it against the reference's generate_multiscale_tiling."""
    import math
    f = 1.0 / factor
    size = min(width, height)
    start_size = max(size, tile_size)
    start_scale, end_scale = start_size / size, tile_size / size
    ntimes = math.ceil(math.log(start_scale / end_scale) / math.log(f))
    start_size = math.ceil(math.exp(ntimes * math.log(f) + math.log(tile_size)))
    start_scale = start_size / size
    scales = np.geomspace(start=start_scale, stop=end_scale, num=ntimes + 1, endpoint=True).tolist()
    levels = sorted(range(len(scales)), key=lambda z: scales[z])          # ascending scale, zoom_level = original position
    cols = {"x1": [], "y1": [], "x2": [], "y2": [], "zoom_level": []}
    for pos, z in enumerate(levels):
        sf = scales[z]
        if not (224 / sf >= min_tile_size or pos == 0):                    # :99 keep the coarsest level at least
            continue
        w, h = max(math.floor(width * sf), tile_size), max(math.floor(height * sf), tile_size)
        half = tile_size // 2
        for sx in (0, half):
            for sy in (0, half):
                ii, jj = np.meshgrid(np.arange((h - sy) // tile_size), np.arange((w - sx) // tile_size), indexing="ij")
                x1 = jj.reshape(-1) * tile_size + sx
                y1 = ii.reshape(-1) * tile_size + sy
                for name, v in (("x1", x1), ("y1", y1), ("x2", x1 + tile_size), ("y2", y1 + tile_size)):
                    cols[name].append((v / sf).astype(np.float32))
                cols["zoom_level"].append(np.full(x1.shape[0], z, np.int16))
    return {k: np.concatenate(v) for k, v in cols.items()}


def synth_pyramid_meta(n_images: int, seed: int, dbidx_start=0, dbidx_stride=1, min_tile_size: int = 60,
                       sizes=((640, 480), (500, 375), (480, 640), (1024, 768), (333, 500), (800, 600), (224, 224), (300, 260))):
    """A ``vector_meta`` frame with the float32 boxes of :func:`pyramid_tiling` for ``n_images`` images whose
    sizes are drawn from ``sizes`` (typical photo shapes) — the box dtype and geometry of a real multiscale index."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(sizes), size=n_images)
    tilings = [pyramid_tiling(w, h, min_tile_size=min_tile_size) for (w, h) in sizes]
    counts = np.array([len(tilings[p]["x1"]) for p in pick], dtype=np.int64)
    ids = dbidx_start + dbidx_stride * np.arange(n_images, dtype=np.int64)
    out = {"dbidx": np.repeat(ids, counts)}
    for c in ("zoom_level", "x1", "y1", "x2", "y2"):
        out[c] = np.concatenate([tilings[p][c] for p in pick])
    return pd.DataFrame(out), counts


def unit_rows(n_rows: int, dim: int, seed: int):
    """L2-normalised float32 Gaussian rows — CLIP-like embeddings that are NOT fp16-representable
    (models/model.py:14-17 normalises; multiscale_tools.py:200 stores float32)."""
    v = np.random.default_rng(seed).standard_normal((n_rows, dim)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
