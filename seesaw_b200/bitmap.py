"""Id sets for ``exclude`` / ``all_indices``.

The reference passes ``pyroaring.BitMap`` objects (query_interface.py:19, multiscale_index.py:216,295).
When pyroaring is installed those are used unchanged; otherwise this numpy-backed class offers the
subset of the interface the hot path and its callers touch.  Anything iterable over ints is accepted
wherever an exclude set is expected (see :func:`as_id_array`)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from pyroaring import BitMap, FrozenBitMap  # type: ignore
    HAVE_PYROARING = True
except ImportError:
    HAVE_PYROARING = False

    class BitMap:
        """Sorted set of non-negative ints (numpy-backed)."""

        __slots__ = ("_a",)

        def __init__(self, values=()):
            self._a = np.unique(as_id_array(values))

        # -- container protocol: ascending iteration like a roaring bitmap
        def __len__(self):
            return int(self._a.shape[0])

        def __iter__(self):
            return iter(self._a.tolist())

        def __contains__(self, v):
            i = np.searchsorted(self._a, int(v))
            return bool(i < len(self._a) and self._a[i] == int(v))

        def __array__(self, dtype=None, copy=None):
            return self._a.astype(dtype if dtype is not None else np.uint32)

        def __eq__(self, other):
            return isinstance(other, BitMap) and np.array_equal(self._a, other._a)

        def __repr__(self):
            return f"BitMap({self._a.tolist()[:8]}{'...' if len(self) > 8 else ''})"

        def _new(self, a):
            out = BitMap.__new__(type(self) if type(self) is BitMap else BitMap)
            out._a = a
            return out

        def difference(self, *others):
            a = self._a
            for o in others:
                a = np.setdiff1d(a, as_id_array(o), assume_unique=False)
            return self._new(a)

        def union(self, *others):
            a = self._a
            for o in others:
                a = np.union1d(a, as_id_array(o))
            return self._new(a)

        def intersection(self, *others):
            a = self._a
            for o in others:
                a = np.intersect1d(a, as_id_array(o))
            return self._new(a)

        __sub__ = difference
        __or__ = union
        __and__ = intersection

        def intersection_cardinality(self, other):
            return int(np.intersect1d(self._a, as_id_array(other)).shape[0])

        def rank(self, value):
            """number of members <= value"""
            return int(np.searchsorted(self._a, int(value), side="right"))

        def update(self, *others):
            for o in others:
                self._a = np.union1d(self._a, as_id_array(o))

        def add(self, v):
            self.update([v])

    class FrozenBitMap(BitMap):
        def update(self, *a):
            raise AttributeError("FrozenBitMap is immutable")

        add = update


def as_id_array(ids) -> np.ndarray:
    """Any id collection (None, BitMap, pyroaring bitmap, ndarray, list, set) -> int64 ndarray."""
    if ids is None:
        return np.zeros(0, np.int64)
    if isinstance(ids, np.ndarray):
        return ids.astype(np.int64).reshape(-1)
    a = getattr(ids, "_a", None)
    if isinstance(a, np.ndarray):
        return a.astype(np.int64)
    try:
        n = len(ids)
    except TypeError:
        n = -1
    if n == 0:
        return np.zeros(0, np.int64)
    return np.fromiter((int(v) for v in ids), dtype=np.int64, count=n)
