"""Loader of the UNMODIFIED reference build in oracle/_ref (test / benchmark infrastructure).

`tools/build_ref.sh` compiles /root/reference (scikit-bio) into oracle/_ref/skbio.  This
module makes it importable: it puts oracle/_ref on sys.path and registers stand-ins for the
third-party imports the image lacks, none of which is on the beta-diversity path
(/root/reference/skbio/diversity/_driver.py:231-355):

    array_api_compat -> the copy vendored in SciPy (scipy._external.array_api_compat)
    h5py, biom       -> empty classes (skbio/io/_fileobject.py:12-16 and skbio/table/_base.py:12
                        only need the names for isinstance checks / the io registry)
    natsort, patsy, statsmodels[.*] -> MagicMock

Only tests/, bench.py's reference arm and its cpu_baseline leg may import this module; the
product package (scikit-bio_b200/) never does.
"""

import os
import sys
import types
from unittest import mock

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "skbio", "__init__.py"))


def _install_shims():
    try:
        import array_api_compat  # noqa: F401
    except ImportError:
        import scipy._external.array_api_compat as aac
        import scipy._external.array_api_compat.numpy as aacn

        sys.modules["array_api_compat"] = aac
        sys.modules["array_api_compat.numpy"] = aacn
    try:
        import h5py  # noqa: F401
    except ImportError:
        h5 = types.ModuleType("h5py")
        h5.Group = type("Group", (), {})
        h5.File = type("File", (h5.Group,), {})
        h5.Dataset = type("Dataset", (), {})
        h5.__version__ = "0"
        sys.modules["h5py"] = h5
    try:
        import biom  # noqa: F401
    except ImportError:
        bm = types.ModuleType("biom")
        bm.Table = type("Table", (), {})
        bm.example_table = None
        bm.subsample = lambda *a, **k: None
        bm.__version__ = "0"
        sys.modules["biom"] = bm
    for name in ("natsort", "patsy", "statsmodels", "statsmodels.api", "statsmodels.formula",
                 "statsmodels.formula.api", "statsmodels.stats", "statsmodels.stats.multitest",
                 "statsmodels.regression", "statsmodels.regression.mixed_linear_model",
                 "statsmodels.stats.weightstats"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = mock.MagicMock(name=name)


def load():
    """Import and return the reference's `skbio` package from oracle/_ref."""
    if not available():
        raise ImportError(
            f"reference build not found under {REF_DIR}: run tools/build_ref.sh in the build "
            "container (it needs /root/reference)")
    _install_shims()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import skbio

    if not os.path.abspath(skbio.__file__).startswith(REF_DIR):
        raise ImportError(f"another skbio is already imported from {skbio.__file__}")
    return skbio
