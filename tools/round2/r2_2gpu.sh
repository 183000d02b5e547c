#!/bin/bash
# Round-2 validation on a 2-GPU lease: tests first, then benches.  Every piece has its own timeout
# and log under gpurun_out/.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 python -m pytest tests/test_gpu_compress.py tests/test_gpu_ring.py -q 2>&1 | tail -15) > $O/r2_t_1gpu.log
(timeout 900 python -m pytest tests/test_multigpu_nvls.py -q -x 2>&1 | tail -40) > $O/r2_t_nvls2.log
(timeout 600 python -m pytest tests/test_multigpu.py tests/test_multigpu_ps.py -q 2>&1 | tail -25) > $O/r2_t_mgpu2.log
(timeout 200 $TR --master-port 29601 benchmarks/compress_bench.py --out $O/compress_2gpu_r2.json 2>&1 | tail -8) > $O/r2_b_compress2.log
for impl in ours ddp; do
  (timeout 240 $TR --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 --impl $impl 2>&1 | tail -3) > $O/r2_b_resnet2_$impl.log
  (timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 --impl $impl --model bert_large --batch-size 32 2>&1 | tail -3) > $O/r2_b_bert2_$impl.log
done
(timeout 300 $TR --master-port 29604 benchmarks/pushpull_bench.py --quick --skip-api --sizes 16777216,104857600 --iters 10 --out $O/pushpull_2gpu_r2.json 2>&1 | tail -3) > $O/r2_b_pushpull2.log
tail -n 12 $O/r2_t_1gpu.log $O/r2_t_nvls2.log $O/r2_t_mgpu2.log $O/r2_b_compress2.log
for f in $O/r2_b_resnet2_*.log $O/r2_b_bert2_*.log; do echo "== $f"; tail -c 1500 $f; done
echo; tail -c 600 $O/r2_b_pushpull2.log
