#!/usr/bin/env bash
# Vendor the UNMODIFIED reference files the parity / baseline harness executes into baseline/_ref/.
#   scripts/vendor_ref.sh [/path/to/reference]        (default /root/reference; run in the build container)
# baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
# travels to the GPU box, where /root/reference does not exist.  Nothing under moma_b200/ imports it: it is used by
#   * scripts/run_ref_loop.py   -- drives helper/loops_moma.py:train_distill_moma unchanged against either the repo's
#                                  modules (MoMA.*, learning.* resolve to this repo) or the reference's own,
#   * tests/test_reference_loop_gpu.py and bench.py's `l2_end_to_end` block.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
[ -d "$REF/MoMA" ] || { echo "vendor_ref: $REF is not a MoMA checkout" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST/helper" "$DST/models"
cp -r "$REF/MoMA" "$REF/learning" "$REF/distiller_zoo" "$DST/"
cp "$REF/helper/__init__.py" "$REF/helper/loops_moma.py" "$REF/helper/util.py" "$DST/helper/"
cp "$REF"/models/__init__.py "$REF"/models/resnet*.py "$REF"/models/mobilenetv2*.py "$REF"/models/shuffle*.py \
   "$REF"/models/ShuffleNet*.py "$REF"/models/vgg.py "$REF"/models/util.py "$DST/models/"
cp -r "$REF/models/efficientnet_pytorch" "$DST/models/"
cp "$REF/train_student_moma.py" "$DST/"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$REF" && find MoMA learning distiller_zoo helper/__init__.py helper/loops_moma.py helper/util.py models train_student_moma.py \
    -type f -name '*.py' 2>/dev/null | sort | while read -r f; do [ -f "$DST/$f" ] && sha256sum "$f"; done ) > "$DST/SHA256SUMS"
echo "vendored $(find "$DST" -name '*.py' | wc -l) files from $REF into $DST (sha256 of every copied file in SHA256SUMS)"
