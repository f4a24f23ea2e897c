"""TEST INFRASTRUCTURE — not product code.

Stub modules that let the UNMODIFIED reference (``/root/reference``) be imported in the
build container, where ``pyroaring``, ``ray``, ``annoy`` and ``pynndescent`` are not
installed and pydantic is v2 (the reference is written against v1).  Used only by
``oracle/make_golden.py`` (fixture generation) and by the CPU tests that validate the
numpy restatement in ``oracle/seesaw_oracle.py`` against the real reference code.
Nothing here is imported by ``seesaw_b200``; ``/root/reference`` does not exist on the GPU
box, so nothing that runs there may call :func:`import_reference`.

Stub surface:
  pyroaring.BitMap / FrozenBitMap  -> sorted-iteration set with the methods the hot path calls
      (multiscale_index.py:216,223,295  coarse_index.py:28,60,76,81-85  query_interface.py:19,48)
  ray.data.extensions.TensorArray  -> thin ndarray wrapper (multiscale_index.py:351,359)
  annoy.AnnoyIndex, pynndescent.NNDescent -> import-only dummies (vector_index.py:2, knn_graph.py:4)
  pydantic -> pydantic.v1 (basic_types.py:1)
"""
import os
import sys
import types

import numpy as np

def _reference_root():
    """/root/reference in the build container; on the GPU box the verbatim copy oracle/build_ref.py made (oracle/_ref)."""
    env = os.environ.get("SEESAW_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/seesaw"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()


class BitMap(set):
    """Set of non-negative ints whose iteration / array conversion is ascending, like a
    roaring bitmap.  ``coarse_index.py:76`` relies on ``np.array(bitmap)`` being sorted."""

    def __init__(self, values=()):
        if isinstance(values, np.ndarray):
            values = values.reshape(-1).tolist()
        super().__init__(int(v) for v in values)

    def __iter__(self):
        return iter(sorted(set.__iter__(self)))

    def __array__(self, dtype=None, copy=None):
        a = np.fromiter(sorted(set.__iter__(self)), dtype=np.uint32, count=len(self))
        return a if dtype is None else a.astype(dtype)

    def _wrap(self, s):
        return type(self)(set.__iter__(s)) if not isinstance(s, BitMap) else s

    def difference(self, *others):
        return BitMap(set.difference(set(set.__iter__(self)), *[set(o) for o in others]))

    def union(self, *others):
        return BitMap(set.union(set(set.__iter__(self)), *[set(o) for o in others]))

    def intersection(self, *others):
        return BitMap(set.intersection(set(set.__iter__(self)), *[set(o) for o in others]))

    def __sub__(self, other):
        return self.difference(other)

    def __or__(self, other):
        return self.union(other)

    def __and__(self, other):
        return self.intersection(other)

    def update(self, *others):
        for o in others:
            if isinstance(o, np.ndarray):
                o = o.reshape(-1).tolist()
            set.update(self, (int(v) for v in o))

    def intersection_cardinality(self, other):
        return len(set.intersection(set(set.__iter__(self)), set(other)))

    def rank(self, value):
        return sum(1 for v in set.__iter__(self) if v <= value)

    def __contains__(self, v):
        try:
            return set.__contains__(self, int(v))
        except (TypeError, ValueError):
            return False


class FrozenBitMap(BitMap):
    def update(self, *a):  # pragma: no cover - mirrors pyroaring
        raise AttributeError("FrozenBitMap is immutable")


class TensorArray:
    def __init__(self, a):
        self._a = np.asarray(a)

    def to_numpy(self, *a, **k):
        return self._a

    def __len__(self):
        return len(self._a)

    def __getitem__(self, i):
        return self._a[i]

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Anything()


def install_stubs():
    """Inject the stub modules.  Idempotent."""
    if "pyroaring" not in sys.modules or not hasattr(sys.modules["pyroaring"], "_ssw_stub"):
        try:
            import pyroaring  # noqa: F401  (a real install wins)
        except ImportError:
            _module("pyroaring", BitMap=BitMap, FrozenBitMap=FrozenBitMap, _ssw_stub=True)
    if "ray" not in sys.modules:
        try:
            import ray  # noqa: F401
        except ImportError:
            def remote(*a, **k):
                if len(a) == 1 and callable(a[0]) and not k:
                    return a[0]
                return lambda f: f

            ray = _module("ray", remote=remote, init=lambda *a, **k: None, get=lambda x: x,
                          put=lambda x: x, ObjectRef=_Anything, get_actor=_Anything(),
                          is_initialized=lambda: False)
            data = _module("ray.data", Dataset=_Anything, read_parquet=_Anything(),
                           ActorPoolStrategy=_Anything)
            ext = _module("ray.data.extensions", TensorArray=TensorArray,
                          TensorDtype=_Anything)
            actor = _module("ray.actor", ActorHandle=_Anything)
            ray.data, data.extensions, ray.actor = data, ext, actor
            _module("ray.data.datasource", FastFileMetadataProvider=_Anything)
            _module("ray.data.datasource.file_meta_provider", FastFileMetadataProvider=_Anything)
            ray.data.datasource = sys.modules["ray.data.datasource"]
    for name, attr in (("annoy", "AnnoyIndex"), ("pynndescent", "NNDescent")):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                _module(name, **{attr: _Anything})
    import pydantic
    if not hasattr(pydantic, "_ssw_stub") and int(pydantic.VERSION.split(".")[0]) >= 2:
        import pydantic.v1 as v1
        v1._ssw_stub = True
        sys.modules["pydantic"] = v1


def import_reference():
    """Return the reference's hot-path modules (multiscale_index, coarse_index, knn_graph,
    query_interface), importing them from REFERENCE_ROOT under the stubs."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "seesaw")):
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ms = importlib.import_module("seesaw.indices.multiscale.multiscale_index")
    co = importlib.import_module("seesaw.indices.coarse.coarse_index")
    kg = importlib.import_module("seesaw.knn_graph")
    qi = importlib.import_module("seesaw.query_interface")
    return types.SimpleNamespace(multiscale=ms, coarse=co, knn_graph=kg, query_interface=qi,
                                 BitMap=sys.modules["pyroaring"].BitMap)
