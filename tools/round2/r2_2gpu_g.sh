#!/bin/bash
O=gpurun_out
mkdir -p $O
(BYTEPS_SERVER_PROFILE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 100 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 2>&1 | grep -E "server profile|ps push_pull") > $O/r2g_ps_profile.log
cat $O/r2g_ps_profile.log
