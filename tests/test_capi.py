"""CPU checks of the drop-in boundary: the shared library loads, exports every function that
include/seesaw_b200.h declares, the ctypes table binds each of them, and — with no GPU — every
compute entry point fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "seesaw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from seesaw_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} is declared in the header but not exported by the library"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.lib.ssw_version() >= 100


def test_constants_match_header():
    from seesaw_b200 import _lib
    text = open(os.path.join(ROOT, "include", "seesaw_b200.h")).read()
    for name in ("SSW_MAX_TOPK", "SSW_MAX_BATCH", "SSW_MAX_KNN_K1"):
        assert int(re.search(rf"#define {name} (\d+)", text).group(1)) == getattr(_lib, name)


def test_no_cpu_fallback():
    """Without an sm_100 device the product must raise (SSW_ERR_NO_DEVICE), never compute on the CPU."""
    from seesaw_b200 import _lib, engine, knn_graph
    if _lib.device_count() > 0:
        pytest.skip("a B200 is present")
    v = np.zeros((8, 512), np.float32)
    with pytest.raises(_lib.SeesawB200Error) as e:
        engine.PatchDatabase.from_arrays(v, np.arange(8))
    assert e.value.code == 2
    with pytest.raises(_lib.SeesawB200Error):
        knn_graph.knn_candidates(v, 3)


def test_argument_validation_needs_no_device():
    from seesaw_b200 import _lib
    h = C.c_void_p()
    # unsupported dim / null handle are rejected before any CUDA call
    rc = _lib.lib.ssw_db_create(C.byref(h), 0, None, 0, 1, 4, 100, None, 0)
    assert rc == 1 and b"null" in _lib.lib.ssw_last_error()
    assert _lib.lib.ssw_set_scan_mode(None, 0) == 1
    assert _lib.lib.ssw_db_destroy(None) == 0


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under seesaw_b200/ may import it."""
    pkg = os.path.join(ROOT, "seesaw_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "seesaw_oracle" not in src and "refstubs" not in src and "import oracle" not in src, fn


def test_every_entry_rejects_null_handles_without_a_device():
    """Argument checks come before any CUDA call: a NULL handle is SSW_ERR_INVALID (1) everywhere."""
    from seesaw_b200 import _lib
    L = _lib.lib
    z = None
    assert L.ssw_scan_topk(z, z, 1, 1, z, z, z, z, z, z) == 1
    assert L.ssw_scan_topk_device(z, z, 1, 1, z, z, z, z, z, z, z) == 1
    assert L.ssw_score_all(z, z, z) == 1
    assert L.ssw_topk_from_scores(z, z, z, 1, z, 0, z, z, z, z) == 1
    assert L.ssw_rescore(z, z, z, z, 0, 0, 0, z, z) == 1
    assert L.ssw_db_set_boxes(z, z, z, z, z, z) == 1
    assert L.ssw_lp_fit(z, z, z, 0, z, z, 1, 1e-5, z, None, None) == 1
    assert L.ssw_lp_destroy(z) == 0 and L.ssw_xchg_destroy(0, z) == 0
    assert L.ssw_knn_build(0, z, 0, 10, 512, 3, 0, 10, z, z) == 1
    assert b"null" in L.ssw_last_error()


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 and as C++ on its own."""
    import subprocess
    h = os.path.join(ROOT, "include", "seesaw_b200.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", h],
                ["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", h]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_c_program_links_against_the_library(tmp_path):
    """A C caller (what a cgo / JNI / ctypes-free maintainer would write) builds and runs against the .so;
    without a GPU it must get SSW_ERR_NO_DEVICE and a message, not a crash."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "seesaw_b200.h"
int main(void) {
  int n = -1;
  if (ssw_device_count(&n) != SSW_OK) return 10;
  float* v = (float*)calloc(4 * 512, sizeof(float));
  int32_t ids[4] = {0, 0, 1, 1};
  ssw_db* db = NULL;
  int rc = ssw_db_create(&db, 0, v, SSW_F32, SSW_F16, 4, 512, ids, 0);
  printf("devices=%d rc=%d msg=%s version=%d\n", n, rc, ssw_last_error(), ssw_version());
  if (n == 0 && rc != SSW_ERR_NO_DEVICE) return 11;
  if (rc == SSW_OK) ssw_db_destroy(db);
  free(v);
  return 0;
}
''')
    exe = tmp_path / "t"
    lib_dir = os.path.join(ROOT, "seesaw_b200")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib_dir, "-lseesaw_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "version=" in r.stdout
