#!/bin/bash
# Round-2 validation on the 8-GPU lease.  8 GPU-minutes per wall minute: everything has a tight timeout.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
# 1. the default 8-GPU path: every exchange kernel + the whole optimizer, NVLS on (and kernels with NVLS off)
(timeout 420 python -m pytest tests/test_multigpu_nvls.py -q -x -k "(exchange_kernels and (8-1 or 8-0)) or (distributed_optimizer and 8-1) or (priority and 8)" 2>&1 | tail -25) > $O/r2_8_tests.log
# 2. exchange microbenchmarks incl. the ring over the BERT-large / ResNet-50 gradient sets
(timeout 240 $TR --master-port 29701 benchmarks/pushpull_bench.py --quick --skip-api --sizes 16777216,104857600,536870912 --iters 6 --out $O/pushpull_8gpu_r2.json 2>&1 | tail -2 | cut -c1-200) > $O/r2_8_pushpull.log
# 3. training steps: ours vs torch DDP (same graph, fused optimizers)
(BYTEPS_TIMING=1 timeout 200 $TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 --impl ours --verbose > $O/r2_8_resnet_ours.log 2> $O/r2_8_resnet_ours.err; grep -E "bench r0|bps-timing r0" $O/r2_8_resnet_ours.err | head -60 > $O/r2_8_startup.log)
(timeout 200 $TR --master-port 29703 bench.py --gpus 8 --steps 20 --warmup 5 --impl ddp 2>&1 | tail -1) > $O/r2_8_resnet_ddp.log
(timeout 240 $TR --master-port 29704 bench.py --gpus 8 --steps 20 --warmup 5 --impl ours --model bert_large --batch-size 32 2>&1 | tail -1) > $O/r2_8_bert_ours.log
(timeout 240 $TR --master-port 29705 bench.py --gpus 8 --steps 20 --warmup 5 --impl ddp --model bert_large --batch-size 32 2>&1 | tail -1) > $O/r2_8_bert_ddp.log
# 4. ncu on the 8-GPU NVLS data path: application replay over the whole 8-rank job (kernel replay cannot
#    save/restore peer-mapped / multicast memory: "UnknownError"), only rank 0 launches kernels, peers idle
(MASTER_PORT=29706 timeout 330 ncu --replay-mode application --target-processes all -k regex:pushpull -c 4 \
   --section SpeedOfLight --section LaunchStats --section Occupancy \
   --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --import-source on -o $O/prof_nvls8 -f bash tools/launch_ranks.sh 8 benchmarks/nvls_profile.py 2>&1 | tail -6) > $O/r2_8_prof_r0.log
if [ "$R2_EXTRA" = "1" ]; then
# 5. BERT-large in-step variants of the fused exchange (per-bucket LSU / per-bucket TMA / ring persistent)
(BYTEPS_RING=off BYTEPS_FUSED_ENGINE=tma timeout 200 $TR --master-port 29707 bench.py --gpus 8 --steps 20 --warmup 5 --model bert_large --batch-size 32 --no-e2e 2>&1 | tail -1) > $O/r2_8_bert_tma.log
(BYTEPS_RING=off BYTEPS_FUSED_ENGINE=lsu timeout 200 $TR --master-port 29708 bench.py --gpus 8 --steps 20 --warmup 5 --model bert_large --batch-size 32 --no-e2e 2>&1 | tail -1) > $O/r2_8_bert_lsu.log
(BYTEPS_RING=persistent timeout 200 $TR --master-port 29709 bench.py --gpus 8 --steps 20 --warmup 5 --model bert_large --batch-size 32 --no-e2e 2>&1 | tail -1) > $O/r2_8_bert_persistent.log
fi
tail -n 20 $O/r2_8_tests.log
for f in $O/r2_8_resnet_ours.log $O/r2_8_resnet_ddp.log $O/r2_8_bert_ours.log $O/r2_8_bert_ddp.log $O/r2_8_bert_tma.log $O/r2_8_bert_lsu.log $O/r2_8_bert_persistent.log; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("impl","value","ms_per_step","exposed_comm_ms","gpu_launches")}, d["config"].get("ring"), d["config"].get("cuda_graph"), (d.get("e2e") or {}).get("value"), d["clocks"])
except Exception as e:
    print("unparsed:", open(sys.argv[1]).read()[-400:])
PY
done
tail -n 3 $O/r2_8_prof_r0.log; ls -la $O/prof_nvls8.ncu-rep
