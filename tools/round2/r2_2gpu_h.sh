#!/bin/bash
# 2 GPUs: CPU-server mode with pull-by-reference (colocated IPC) - correctness, then 100 MB timing with it on / off
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_multigpu_ps.py -x -q > $O/r2h_tests.log 2>&1
tail -3 $O/r2h_tests.log
for REF in 1 0; do
  (BYTEPS_PS_PULL_BY_REF=$REF BYTEPS_SERVER_PROFILE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 100 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 --out $O/ps_2gpu_r2_ref$REF.json 2>&1 | grep -E "server profile|ps push_pull|rror") > $O/r2h_ps_ref$REF.log
  cat $O/r2h_ps_ref$REF.log
done
