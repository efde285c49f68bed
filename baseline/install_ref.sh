#!/bin/sh
# Installs the UNMODIFIED reference (its Python engine + its own CUDA extension, rebuilt for sm_100) under
# baseline/_ref, from a writable copy of /root/reference (the build writes into the source tree), then fixes the
# install layout and pre-generates the prime pickles (baseline/ref_harness.py).  CPU only, ~5 min.
# baseline/_ref is git-ignored (not product source) but travels to the GPU box with gpurun.
set -e
cd "$(dirname "$0")/.."
SRC=${1:-/root/reference}
TMP=$(mktemp -d /tmp/tb200_ref.XXXXXX)
cp -r "$SRC" "$TMP/src"
rm -rf baseline/_ref
CMAKE_ARGS=-DCMAKE_CUDA_ARCHITECTURES=100 python -m pip install --no-index --no-build-isolation --no-deps \
  --find-links /opt/wheelhouse --target baseline/_ref "$TMP/src"
python baseline/ref_harness.py
rm -rf "$TMP"
