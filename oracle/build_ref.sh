#!/usr/bin/env bash
# Builds the UNMODIFIED reference (seung-lab/crackle) pybind11 module from the sources where they
# lie under /root/reference into oracle/_ref/ as `_fastcrackle_ref` (renamed via a macro so it can be
# imported beside our own `fastcrackle`).  Test infrastructure only; nothing here is product code.
# Does not run the reference's own build system (setup.py needs pbr, which is absent).
set -euo pipefail
REF="${CRACKLE_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/src/fastcrackle.cpp" ]; then
  echo "reference sources not found at $REF; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
# The reference's pure-Python package and its own test file are staged next to the compiled module (git-ignored like it;
# nothing of the reference enters the history): tests/test_reference_suite.py runs the reference's automated_test.py through
# the reference's Python layer with the B200 library underneath (INTEGRATION.md Option B).
PKG="$OUT/refpkg"
if [ -d "$REF/crackle" ]; then
  rm -rf "$PKG"; mkdir -p "$PKG"
  cp -r "$REF/crackle" "$PKG/crackle"
  cp "$REF/automated_test.py" "$PKG/automated_test.py"
  find "$PKG" -name "__pycache__" -type d -prune -exec rm -rf {} +
fi
EXT="$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")"
TARGET="$OUT/_fastcrackle_ref$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/src/fastcrackle.cpp" ] && [ "${1:-}" != "--force" ]; then
  echo "up to date: $TARGET"; exit 0
fi
WRAP="$(mktemp /tmp/ref_wrap_XXXX.cpp)"
printf '#define fastcrackle _fastcrackle_ref\n#include "fastcrackle.cpp"\n' > "$WRAP"
g++ -std=c++2a -O3 -msse4.2 -mpclmul -shared -fPIC -fvisibility=hidden -pthread \
    $(python3 -m pybind11 --includes) -I"$REF/third_party/fastcrc" -I"$REF/src" \
    "$WRAP" -o "$TARGET"
rm -f "$WRAP"
echo "built $TARGET"
