#!/bin/bash
# Stage the UNMODIFIED reference (its package and its two knn test files) under baseline/_ref/ --
# git-ignored, so never part of this repository's history, but shipped to the GPU box with a
# snapshot -- for tests/test_dropin_gpu.py::test_unmodified_reference_tests_pass_on_the_shim.
# Run in the build container (where /root/reference exists).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
mkdir -p "$ROOT/baseline/_ref/ref_tests"
rm -rf "$ROOT/baseline/_ref/syntropy"
cp -r "$REF/syntropy" "$ROOT/baseline/_ref/syntropy"
cp "$REF/tests/test_knn.py" "$REF/tests/test_knn_temporal.py" "$REF/tests/test_knn_utils.py" "$ROOT/baseline/_ref/ref_tests/"
find "$ROOT/baseline/_ref" -name __pycache__ -prune -exec rm -rf {} +
echo "staged: $(ls "$ROOT/baseline/_ref/ref_tests")"
