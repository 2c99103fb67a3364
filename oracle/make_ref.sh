#!/bin/sh
# Stage the UNMODIFIED reference (its pure-Python `lib/` and `tools/` trees) into oracle/_ref/ so that it travels to
# the GPU box with the snapshot (oracle/_ref/ is git-ignored, NOT gpurun-ignored; nothing of the reference enters the
# repository history -- it is CC BY-NC-SA).  bench.py's `--impl reference` / `cpu_baseline` legs and
# the "tools run unchanged" tests import it from there; when it is absent they fall back to the oracle port.
#   sh oracle/make_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -d "$SRC/lib/models" ]; then
  echo "make_ref: no reference tree at $SRC (nothing staged)"
  exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/lib" "$SRC/tools" "$SRC/configs" "$HERE/_ref/"
find "$HERE/_ref" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find lib tools configs -type f | sort | xargs sha256sum ) > "$HERE/_ref/SHA256SUMS"
echo "make_ref: staged $(find "$HERE/_ref" -type f | wc -l) files from $SRC into $HERE/_ref"
