import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_native_built():
    """Build libskbb200.so / liboracle.so when a fresh checkout has no binaries yet (they are
    git-ignored).  nvcc cross-compiles sm_100a without a GPU; on a box without nvcc the
    prebuilt files travel with the snapshot."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "scikit-bio_b200", "libskbb200.so")
    src_dir = os.path.join(ROOT, "scikit-bio_b200", "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "skbio_b200.h"))
    stale = (not os.path.exists(lib)) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in srcs)
    if stale and shutil.which("nvcc"):
        subprocess.check_call(["make", "-s", "-C", src_dir])
    olib = os.path.join(ROOT, "oracle", "liboracle.so")
    osrc = os.path.join(ROOT, "oracle", "oracle_c.c")
    if (not os.path.exists(olib)) or os.path.getmtime(osrc) > os.path.getmtime(olib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _ensure_native_built()


def _load_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)["cases"]


_CASES = None


def golden_cases():
    global _CASES
    if _CASES is None:
        _CASES = _load_cases()
    return _CASES


def case_arrays(c):
    counts = np.array(c["counts"], dtype=c["counts_dtype"])
    parent = np.array(c["tree_parent"], dtype=np.int64)
    length = np.array(c["tree_length"], dtype=np.float64)
    names = np.array(c["tree_names"], dtype=object)
    return counts, parent, length, names


@pytest.fixture(scope="session")
def qiime_tt():
    with open(os.path.join(GOLDEN, "qiime191tt.json")) as f:
        return json.load(f)


def assert_close_rel(got, exp, rtol, what=""):
    """Relative comparison with exact agreement required where the expectation is 0/NaN."""
    got = np.asarray(got, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    assert got.shape == exp.shape, f"{what}: shape {got.shape} != {exp.shape}"
    nan_e = np.isnan(exp)
    assert np.array_equal(np.isnan(got), nan_e), f"{what}: NaN pattern differs"
    inf_e = np.isinf(exp)
    assert np.array_equal(got[inf_e], exp[inf_e]), f"{what}: inf pattern differs"
    ok = ~(nan_e | inf_e)
    g, e = got[ok], exp[ok]
    zero = e == 0.0
    assert np.all(g[zero] == 0.0), f"{what}: expected exact zeros"
    nz = ~zero
    if nz.any():
        rel = np.abs(g[nz] - e[nz]) / np.abs(e[nz])
        assert rel.max() <= rtol, f"{what}: max rel err {rel.max():.3e} > {rtol:g}"
