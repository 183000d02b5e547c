#!/usr/bin/env bash
# Run the native-runtime tests under AddressSanitizer (or ThreadSanitizer) in a scratch copy of the tree.
#   tools/sanitize.sh address|thread|undefined [pytest args...]
# The reference only offers ASAN=1 in ps-lite's Makefile; here the whole _core module
# (registry, scheduler, reducer, compressors, transport, server, PS worker) is instrumented and driven by
# the same python tests that run in CI.  Round-1 result: one real bug (heap-use-after-free: the resender
# retransmitted a zero-copy view of a buffer the caller had already released) - fixed, suite clean.
set -euo pipefail
KIND=${1:-address}; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=${SAN_WORKDIR:-/tmp/byteps_b200_san}
rm -rf "$WORK" && mkdir -p "$WORK/repo"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude='byteps_b200/build' --exclude='_core*.so' -cf - .) | (cd "$WORK/repo" && tar xf -)
cd "$WORK/repo"
python - "$KIND" <<'PY'
import sys
kind = sys.argv[1]
p = "byteps_b200/_build.py"
s = open(p).read()
s = s.replace('"-O3", "-std=c++17", "-fPIC"', '"-O1", "-g", "-fsanitize=%s", "-fno-omit-frame-pointer", "-std=c++17", "-fPIC"' % kind)
s = s.replace('objs + ["-fopenmp", "-pthread", "-lrt", "-ldl"]', 'objs + ["-fsanitize=%s", "-fopenmp", "-pthread", "-lrt", "-ldl"]' % kind)
if kind == "thread":
    # libgomp is not instrumented (its barriers are invisible to TSAN): compile the `omp parallel for` loops
    # as plain serial loops, so every remaining report is a race between OUR threads
    s = s.replace('"-fopenmp", ', '"-Wno-unknown-pragmas", ')
    s = s.replace('"-fsanitize=thread", "-fopenmp", "-pthread"', '"-fsanitize=thread", "-pthread"')
open(p, "w").write(s)
PY
python byteps_b200/_build.py core
case "$KIND" in thread) L=tsan;; undefined) L=ubsan;; *) L=asan;; esac
LIB=$(gcc -print-file-name=lib$L.so)
STD=$(gcc -print-file-name=libstdc++.so.6)
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=$WORK/asan
export UBSAN_OPTIONS=print_stacktrace=1:log_path=$WORK/ubsan
export TSAN_OPTIONS=halt_on_error=0:log_path=$WORK/tsan:second_deadlock_stack=1
TESTS=${*:-tests/test_ps.py tests/test_net_features.py tests/test_core_units.py tests/test_ps_api.py}
LD_PRELOAD="$LIB $STD" python -m pytest $TESTS -q -p no:cacheprovider --timeout=900 || true
echo "--- sanitizer reports:"
grep -h "ERROR: AddressSanitizer\|WARNING: ThreadSanitizer\|runtime error" "$WORK"/asan* "$WORK"/tsan* "$WORK"/ubsan* 2>/dev/null | sort | uniq -c || echo none
