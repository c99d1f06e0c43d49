#!/bin/bash
# Local wrapper: rebuild the in-tree library if any source is newer, then hand the command to gpurun.
#   bash scripts/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -m probly_search_b200.build > /dev/null
python -c "from oracle import oracle as o; o.build()" > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
