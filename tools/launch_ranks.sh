#!/bin/bash
# launch_ranks.sh N script.py [args]: one process per GPU on this host without torchrun (so a profiler that wraps
# THIS script sees every rank as a child process).
N=$1; shift
export MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29650} WORLD_SIZE=$N
pids=()
for ((r = 1; r < N; r++)); do RANK=$r LOCAL_RANK=$r python "$@" & pids+=($!); done
RANK=0 LOCAL_RANK=0 python "$@"
rc=$?
for p in "${pids[@]}"; do wait $p || rc=$?; done
exit $rc
