#!/usr/bin/env bash
# compute-sanitizer pass over the single-rank GPU kernel tests (run on a B200 box, e.g. through gpurun):
#   tools/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck
# Only world=1 cases are selected: the multi-rank kernels spin on each other's flags and the sanitizer serialises
# kernel launches, so virtual-cluster tests would dead-lock under it (the barrier watchdog would fire after 30 s).
# Not executed in round 1 (the GPU budget went to benchmarks and numerics); kept as the recipe for round 2.
set -euo pipefail
TOOL=${1:-memcheck}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
mkdir -p gpurun_out
export BYTEPS_FUSED_ENGINE=${BYTEPS_FUSED_ENGINE:-tma}
timeout 1200 compute-sanitizer --tool "$TOOL" --target-processes all --error-exitcode 3 \
  --log-file gpurun_out/sanitizer_"$TOOL".log \
  python -m pytest tests/test_gpu_kernels.py -x -q -k "world1 or (virtual and 1-)" -p no:cacheprovider || true
tail -30 gpurun_out/sanitizer_"$TOOL".log
