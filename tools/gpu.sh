#!/bin/bash
# build libdwb.so here (nvcc cross-compiles), then run a command on a B200 box: tools/gpu.sh [--timeout S] -- 'cmd'
set -e
cd "$(dirname "$0")/.."
python -c "from diffwave_sashimi_b200 import build; build.build()" || exit 1
exec /usr/local/graft/bin/gpurun "$@"
