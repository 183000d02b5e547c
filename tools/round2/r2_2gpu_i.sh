#!/bin/bash
# 2 GPUs: where does a 100 MB CPU-server push_pull spend its time now that pulls are by reference?
O=gpurun_out
mkdir -p $O
rm -rf /tmp/bps_trace; mkdir -p /tmp/bps_trace
(BYTEPS_TRACE_ON=1 BYTEPS_TRACE_START_STEP=6 BYTEPS_TRACE_END_STEP=8 BYTEPS_TRACE_DIR=/tmp/bps_trace BYTEPS_SERVER_PROFILE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 60 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 2>&1 | grep -E "server profile|ps push_pull|rror") > $O/r2i_ps_trace.log
cp /tmp/bps_trace/0/comm.json $O/ps_trace_ref_rank0.json 2>/dev/null
cp /tmp/bps_trace/1/comm.json $O/ps_trace_ref_rank1.json 2>/dev/null
(BYTEPS_SERVER_PROFILE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 60 python -m byteps_b200.launcher.local_cluster -n 1 -s 1 python benchmarks/ps_bench.py --mb 100 2>&1 | grep -E "server profile|ps push_pull|rror") > $O/r2i_ps_1w.log
cat $O/r2i_ps_trace.log $O/r2i_ps_1w.log
