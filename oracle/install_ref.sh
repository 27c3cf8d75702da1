#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Makes the unmodified reference (burchim/AVEC) available on the GPU box, which only sees this
# repository: copies the reference's Python tree into baseline/_ref/ (git-ignored, NOT gpurun-ignored, never committed) so that
#   * bench.py --impl reference can time the reference's own code (kind "reference") instead of the restatement,
#   * tests/test_gpu_dropin.py can run the reference's Model.train_step / main.py on the patched encoders.
# The reference has no setup.py / build system (pure Python), so "install" is a copy of nnet/, main.py, functions.py, configs/.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -d "$SRC/nnet" ] || { echo "no reference tree at $SRC"; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/nnet" "$SRC/main.py" "$SRC/functions.py" "$SRC/configs" "$DST/"
[ -f "$SRC/LICENSE" ] && cp "$SRC/LICENSE" "$DST/" || true
find "$DST" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
echo "installed $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
