#!/bin/bash
# Build + CPU tests here, then run the given command on the GPU box. Usage: tools/gpu.sh <timeout_s> '<command>' [--gpus N]
set -e
cd "$(dirname "$0")/.."
python -m param_b200.build | tail -2
python -m pytest tests -x -q -m "not gpu" 2>&1 | tail -2
T=$1; shift
CMD=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
