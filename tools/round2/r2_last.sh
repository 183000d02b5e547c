#!/bin/bash
# last seconds of the round-2 GPU budget: the C API GPU job again (helper library now links misc.cu), then the
# device-timed copy profile of a 1-worker CPU-server push_pull
O=gpurun_out
mkdir -p $O
timeout 40 python -m pytest tests/test_capi.py -q -m gpu -p no:cacheprovider > $O/r2_last_capi.log 2>&1; tail -2 $O/r2_last_capi.log
(BYTEPS_STAGE_PROFILE=1 BYTEPS_SERVER_PROFILE=1 BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 30 python -m byteps_b200.launcher.local_cluster -n 1 -s 1 python benchmarks/ps_bench.py --mb 100 2>&1 | grep -E "stage profile|server profile|ps push_pull|rror") > $O/r2_last_stage_profile.log
cut -c1-1500 $O/r2_last_stage_profile.log
