#!/bin/bash
O=gpurun_out
mkdir -p $O
N0=$(cat /sys/devices/system/node/node0/cpulist)
N1=$(cat /sys/devices/system/node/node1/cpulist 2>/dev/null)
echo "node0 cpus: $N0  node1 cpus: $N1" > $O/r2f_numa.log
nvidia-smi topo -m 2>/dev/null | head -12 >> $O/r2f_numa.log
for node in "$N0" "$N1"; do
  (BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 100 taskset -c $node python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 2>&1 | tail -1) >> $O/r2f_numa.log
done
cat $O/r2f_numa.log
