#!/bin/bash
# Build the library in-tree, then run a command on a B200 box:  scripts/gpu.sh [--timeout S] [--gpus N] -- '<cmd>'
set -e
cd "$(dirname "$0")/.."
python -m glam_b200.build >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
