#!/usr/bin/env bash
# Build the UNMODIFIED reference (scikit-bio, /root/reference) into oracle/_ref/ so that
# bench.py --impl reference, bench.py's cpu_baseline leg and tests/test_reference_build.py
# can call skbio.diversity.beta_diversity itself (SURVEY.md section 8c recipe).
#
#   tools/build_ref.sh [reference_dir]        (default /root/reference)
#
# Nothing is written to the reference directory: the sources are copied to a scratch
# directory, the seven Cython extensions are compiled there (no OpenMP: the beta-diversity
# path has none, setup.py:159-163), and the importable package (python files + the built
# extension modules, without tests) is placed under oracle/_ref/skbio.  oracle/_ref/ is
# git-ignored (build output only, never committed) but travels to the GPU box with the
# snapshot.  Third-party imports the image lacks (biom, h5py, natsort, patsy, statsmodels,
# array_api_compat) are shimmed at import time by oracle/ref_skbio.py.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/skbio" ]; then
    echo "build_ref: $REF/skbio not found (the reference is only present in the build container)" >&2
    exit 2
fi
WORK="$(mktemp -d /tmp/skbio_ref_build.XXXXXX)"
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF/skbio" "$REF/setup.py" "$REF/pyproject.toml" "$REF/README.rst" "$REF/LICENSE.txt" "$WORK/"
# /opt/gcc's -fopenmp cannot link in this image; OpenMP is irrelevant to this path anyway
( cd "$WORK" && DISABLE_OPENMP=1 python setup.py build_ext --inplace > "$WORK/build.log" 2>&1 ) || {
    tail -30 "$WORK/build.log" >&2
    exit 1
}
rm -rf "$OUT"
mkdir -p "$OUT"
# the importable package only: no tests, no generated C, no Cython sources
( cd "$WORK" && find skbio \( -name tests -o -name __pycache__ \) -prune -o -type f \
    \( -name '*.py' -o -name '*.so' \) -print0 | tar --null -T - -cf - | tar -xf - -C "$OUT" )
python - "$OUT" <<'EOF'
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(sys.argv[1]), ".."))
from oracle import ref_skbio
skbio = ref_skbio.load()
from skbio.diversity import _phylogenetic
print("build_ref: skbio", skbio.__version__, "->", os.path.dirname(skbio.__file__))
print("build_ref: native", _phylogenetic.__file__)
EOF
