#!/usr/bin/env bash
# compute-sanitizer pass over the single-rank GPU kernel tests (run on a B200 box, e.g. through gpurun):
#   tools/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck [all]
# Only world=1 cases are selected: the multi-rank kernels spin on each other's flags and the sanitizer serialises
# kernel launches, so virtual-cluster tests would dead-lock under it (the spin watchdog would trap).
# Default: one case per kernel family (what round 2 ran: 8 passed, 0 errors, 9 s under memcheck,
# profiles/logs/sanitizer_memcheck*.log); `all` selects every world-1 case (several minutes).
set -euo pipefail
TOOL=${1:-memcheck}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
mkdir -p gpurun_out
if [ "${2:-}" = "all" ]; then
  SEL=(tests/test_gpu_kernels.py tests/test_gpu_ring.py tests/test_gpu_compress.py -k "1] and not multigpu")
else
  SEL=("tests/test_gpu_kernels.py::test_inplace_pushpull_virtual[bf16-1]"
       "tests/test_gpu_kernels.py::test_inplace_pushpull_tma_virtual[bf16-1]"
       "tests/test_gpu_kernels.py::test_packed_pushpull_virtual[f32-bf16-True-1]"
       "tests/test_gpu_kernels.py::test_fused_optimizer_virtual[lsu-bf16-adam-1]"
       "tests/test_gpu_kernels.py::test_fused_optimizer_virtual[tma-bf16-sgd-1]"
       "tests/test_gpu_ring.py::test_ring_allreduce_virtual[True-bf16-1]"
       "tests/test_gpu_compress.py::test_gpu_compressor_matches_cpu_reference[kw0-1]"
       "tests/test_gpu_compress.py::test_gpu_compressor_matches_cpu_reference[kw1-1]")
fi
timeout 1200 compute-sanitizer --tool "$TOOL" --error-exitcode 3 \
  --log-file gpurun_out/sanitizer_"$TOOL".log \
  python -m pytest -q -p no:cacheprovider --timeout 600 "${SEL[@]}" || true
tail -30 gpurun_out/sanitizer_"$TOOL".log
