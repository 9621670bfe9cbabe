#!/usr/bin/env bash
# Stages the UNMODIFIED reference under baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box).
#   * the package `DominantSparseEigenAD` is pip-installed from a /tmp copy of /root/reference
#     (the checkout is read-only and setup.py wants to write build/);
#   * examples/TFIM/TFIM.py — the model the reference's benchmark-shaped drivers (E0.py:53-67, chiF.py:40-53)
#     use, which setup.py does not package — is copied verbatim next to it as baseline/_ref/ref_examples/TFIM.py.
# Nothing from the reference enters the tracked tree.  Outcome of the install is recorded in DESIGN.md section 3.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
mkdir -p "$HERE/_ref/ref_examples"
cp "$REF/examples/TFIM/TFIM.py" "$HERE/_ref/ref_examples/TFIM.py"
rm -rf "$TMP"
echo "reference staged under $HERE/_ref"
