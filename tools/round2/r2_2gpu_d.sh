#!/bin/bash
O=gpurun_out
mkdir -p $O
(timeout 300 python -m pytest tests/test_multigpu.py -q -k "cross_barrier_overlaps" 2>&1 | tail -12) > $O/r2d_t_cb.log
# CPU-server pipeline with the timeline on: where do the 16 ms go?
rm -rf /tmp/bps_trace; mkdir -p /tmp/bps_trace
(BYTEPS_TRACE_ON=1 BYTEPS_TRACE_START_STEP=4 BYTEPS_TRACE_END_STEP=6 BYTEPS_TRACE_DIR=/tmp/bps_trace BYTEPS_PS_PIPELINE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 BYTEPS_LOG_LEVEL=WARNING timeout 200 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 2>&1 | tail -4) > $O/r2d_b_ps_trace.log
cp /tmp/bps_trace/0/comm.json $O/ps_trace_rank0.json 2>/dev/null
nproc > $O/r2d_nproc.log; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> $O/r2d_nproc.log
# ncu on the 2-GPU NVLS path: (A) application replay over the whole 2-rank job, (B) one-pass metrics with kernel replay
(MASTER_PORT=29661 timeout 420 ncu --replay-mode application --target-processes all -k regex:pushpull -c 3 \
   --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats \
   --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum \
   --clock-control none --import-source on -o $O/prof_nvls2_app -f bash tools/launch_ranks.sh 2 benchmarks/nvls_profile.py 2>&1 | tail -8) > $O/r2d_ncu_a.log
(export MASTER_ADDR=127.0.0.1 MASTER_PORT=29662 WORLD_SIZE=2
 RANK=1 LOCAL_RANK=1 timeout 120 python benchmarks/nvls_profile.py > $O/r2d_prof_r1.log 2>&1 &
 RANK=0 LOCAL_RANK=0 timeout 120 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum --clock-control none -k regex:pushpull -o $O/prof_nvls2_1pass -f python benchmarks/nvls_profile.py 2>&1 | tail -6
 wait) > $O/r2d_ncu_b.log 2>&1
tail -n 8 $O/r2d_t_cb.log $O/r2d_b_ps_trace.log $O/r2d_nproc.log $O/r2d_ncu_a.log $O/r2d_ncu_b.log; ls -la $O/*.ncu-rep $O/ps_trace_rank0.json
