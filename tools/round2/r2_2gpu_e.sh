#!/bin/bash
O=gpurun_out
mkdir -p $O
for cfg in "1 8 4" "1 1 1"; do
  set -- $cfg
  (BYTEPS_PS_PIPELINE=$1 BYTEPS_IPC_COPY_NUM_THREADS=$2 BYTEPS_SERVER_OMP_THREADS=$3 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 100 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 --out $O/ps_2gpu_r2_copy$2_omp$3.json 2>&1 | tail -1) > $O/r2e_ps_$2_$3.log
  cat $O/r2e_ps_$2_$3.log
done
