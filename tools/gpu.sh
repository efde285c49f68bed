#!/bin/sh
# Rebuild the CUDA library (nvcc cross-compiles here), then run a command on the GPU box.
#   tools/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun "$@"
