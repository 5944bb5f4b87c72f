#!/bin/bash
# Put the UNMODIFIED reference package where the GPU box can import it: baseline/_ref/ is git-ignored (no reference source
# enters the history) but not gpurun-ignored, so it travels with the snapshot like the built .so does.
# The base contract's `pip install --target baseline/_ref /root/reference` cannot run here: the build backend needs
# setuptools-scm, which is neither installed nor in /opt/wheelhouse, and the reference's dependencies torch-geometric,
# hydra-core and anemoi-utils are absent too (oracle/pyg_shim and oracle/ref_shims stand in for them).
# The package is pure Python, so copying src/anemoi is exactly what an install would put there.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${1:-/root/reference}
[ -d "$SRC/src/anemoi/models" ] || { echo "no reference tree at $SRC"; exit 1; }
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/anemoi"
cp -r "$SRC/src/anemoi" "$ROOT/baseline/_ref/anemoi"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find src/anemoi -type f -name '*.py' | sort | xargs sha256sum ) > "$ROOT/baseline/_ref/SHA256SUMS"
echo "reference copied from $SRC/src/anemoi ($(wc -l < "$ROOT/baseline/_ref/SHA256SUMS") files); import with PYTHONPATH=baseline/_ref:oracle/pyg_shim:oracle/ref_shims" | tee "$ROOT/baseline/_ref/INSTALL.txt"
