"""Importable alias of the ``scikit-bio_b200/`` package directory.

The product package lives in ``scikit-bio_b200/`` (the name the repository layout asks for),
which is not a valid Python identifier.  This shim makes ``import skbio_b200`` resolve its
sub-modules from that directory, so ``skbio_b200._driver`` is ``scikit-bio_b200/_driver.py``.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "scikit-bio_b200")
__path__.insert(0, _pkg_dir)

with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
del _f
