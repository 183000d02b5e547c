#!/bin/bash
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 python -m pytest tests/test_gpu_compress.py tests/test_gpu_ring.py tests/test_gpu_native_ops.py -q -s 2>&1 | tail -15) > $O/r2b_t_1gpu.log
(timeout 900 python -m pytest tests/test_multigpu_nvls.py -q -x 2>&1 | tail -40) > $O/r2b_t_nvls2.log
(timeout 400 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -25) > $O/r2b_t_mgpu2.log
(timeout 150 $TR --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --impl ddp --verbose --hang-dump 45 > $O/r2b_ddp.out 2> $O/r2b_ddp.err; tail -3 $O/r2b_ddp.out; grep -E "bench r0|File|line" $O/r2b_ddp.err | tail -40) > $O/r2b_b_ddp.log 2>&1
(timeout 200 $TR --master-port 29612 benchmarks/compress_bench.py --out $O/compress_2gpu_r2.json 2>&1 | tail -6) > $O/r2b_b_compress2.log
(timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/compress_launches.csv python benchmarks/compress_kernels_bench.py --iters 2 2>&1 | tail -6) > $O/r2b_ncu_compress.log
(timeout 300 $TR --master-port 29613 benchmarks/pushpull_bench.py --quick --skip-api --sizes 16777216,104857600 --iters 10 --out $O/pushpull_2gpu_r2.json 2>&1 | tail -2 | cut -c1-300) > $O/r2b_b_pushpull2.log
tail -n 14 $O/r2b_t_1gpu.log $O/r2b_t_nvls2.log $O/r2b_t_mgpu2.log $O/r2b_b_ddp.log $O/r2b_b_compress2.log
