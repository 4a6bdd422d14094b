#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE — "install" the UNMODIFIED reference where the GPU box can import it.
#
# isl-org/lang-seg has no setup.py / pyproject.toml (it is a scripts-and-modules checkout), so
# `pip install --target baseline/_ref /root/reference` has nothing to install; this script is that step: it copies
# the reference's own Python packages and label files, byte for byte, from /root/reference into baseline/_ref/.
# baseline/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so it travels
# to the GPU box like the built .so. Consumers (they skip / fall back when the directory is absent):
#   * bench.py --impl reference      — times the reference's own LSegModule.evaluate_random on the host cores
#   * tests/test_reference_callers_gpu.py — drives the unmodified callers through the B200 drop-in
# The absent third-party packages (timm, clip, encoding, pytorch_lightning, matplotlib) are provided by
# oracle/ref_standins.py exactly as for oracle/make_golden.py.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/../baseline/_ref"
if [ ! -d "$SRC/modules" ]; then
  echo "make_ref.sh: $SRC is not a lang-seg checkout" >&2
  exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
for d in modules additional_utils data label_files; do
  cp -r "$SRC/$d" "$DST/$d"
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find modules additional_utils data label_files -type f ! -path '*/__pycache__/*' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "reference installed at $DST ($(find "$DST" -type f | wc -l) files)"
