import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared library is a build artefact (git-ignored): build it in-tree when a fresh checkout has none
    if not os.path.exists(os.path.join(ROOT, "seesaw_b200", "libseesaw_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference, importable only in the build container."""
    import refstubs
    if not os.path.isdir(os.path.join(refstubs.REFERENCE_ROOT, "seesaw")):
        pytest.skip("reference tree not present (GPU box)")
    return refstubs.import_reference()
