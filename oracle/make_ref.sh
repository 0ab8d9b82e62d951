#!/bin/sh
# TEST INFRASTRUCTURE -- not product code.  Stages the UNMODIFIED reference model files next to the oracle so
# that they travel to the GPU box (which has no /root/reference): oracle/_ref/ is git-ignored (the sources
# never enter this repository's history) but NOT gpurun-ignored.  Used only by bench.py's CPU legs
# (`--impl reference` and `cpu_baseline`, kind "reference") and by tests that pin the oracle port.
#   sh oracle/make_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
if [ ! -f "$SRC/Model.py" ]; then
  echo "make_ref: $SRC/Model.py not found (GPU box: the prebuilt oracle/_ref is used as it is)" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
for f in Model.py gnn_transformer.py combination_layer.py; do
  cp -f "$SRC/$f" "$HERE/_ref/$f"
done
( cd "$SRC" && sha256sum Model.py gnn_transformer.py combination_layer.py ) > "$HERE/_ref/SHA256SUMS"
echo "make_ref: staged Model.py gnn_transformer.py combination_layer.py from $SRC into $HERE/_ref"
