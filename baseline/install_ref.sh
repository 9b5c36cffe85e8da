#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun).
# The reference's setup.py lists no package data and its gym_sted/routines has no __init__.py, so pip leaves out the
# JSON / weight / fixture files its own code opens at run time; they are copied next to the installed modules.
# Nothing here is product source: baseline/_ref is only read by tests/test_reference_dropin.py and tests/shims.py.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC"/. "$TMP"/                      # /root/reference is read-only; the build writes an egg-info
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" >/dev/null
(cd "$SRC" && find gym_sted -type f \( -name '*.json' -o -name '*.t7' -o -name '*.pbz2' -o -name '*.yml' -o -name '*.yaml' \) -print0 \
    | while IFS= read -r -d '' f; do mkdir -p "$HERE/_ref/$(dirname "$f")"; cp "$f" "$HERE/_ref/$f"; done)
mkdir -p "$HERE/_ref/gym_sted/routines" && cp "$SRC/gym_sted/routines/create.py" "$HERE/_ref/gym_sted/routines/create.py"
rm -rf "$TMP"
echo "reference installed under $HERE/_ref"
