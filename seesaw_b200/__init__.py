"""seesaw_b200 — B200-native vector-search hot path of orm011/seesaw behind the reference's interfaces.

Submodules: ``indices`` (B200MultiscaleIndex, B200CoarseIndex, B200VectorIndex), ``knn_graph``, ``label_propagation``,
``service`` (ScanBatcher), ``sharded`` (one process per GPU), ``engine`` (PatchDatabase over the C ABI, ``_lib``).
Importing any of them loads ``libseesaw_b200.so``; there is no CPU fallback."""

__version__ = "1.0.0"
