#!/bin/bash
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 400 python -m pytest tests/test_gpu_compress.py tests/test_gpu_ring.py tests/test_gpu_native_ops.py tests/test_gpu_kernels.py tests/test_capi.py -q 2>&1 | tail -15) > $O/r2c_t_1gpu.log
(timeout 600 python -m pytest tests/test_multigpu_nvls.py -q -x 2>&1 | tail -30) > $O/r2c_t_nvls2.log
(timeout 600 python -m pytest tests/test_multigpu.py tests/test_multigpu_ps.py -q 2>&1 | tail -30) > $O/r2c_t_mgpu2.log
(timeout 200 $TR --master-port 29621 benchmarks/compress_bench.py --out $O/compress_2gpu_r2.json 2>&1 | tail -6) > $O/r2c_b_compress2.log
(timeout 150 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --impl ddp 2>&1 | tail -1 | cut -c1-700) > $O/r2c_b_resnet2_ddp.log
(timeout 200 $TR --master-port 29623 bench.py --gpus 2 --steps 20 --warmup 5 --impl ddp --model bert_large --batch-size 32 2>&1 | tail -1 | cut -c1-700) > $O/r2c_b_bert2_ddp.log
# CPU-server mode, 100 MB GPU gradient, 2 workers + 2 servers, shm IPC: pipelined vs not
for pipe in 1 0; do
  (BYTEPS_PS_PIPELINE=$pipe BYTEPS_ENABLE_IPC=1 DMLC_NUM_PORTS=4 timeout 200 python -m byteps_b200.launcher.local_cluster -n 2 -s 2 python benchmarks/ps_bench.py --mb 100 --out $O/ps_2gpu_ipc_pipe$pipe.json 2>&1 | tail -2) > $O/r2c_b_ps_pipe$pipe.log
done
# ncu flow rehearsal on 2 GPUs: rank 0 under ncu, rank 1 plain
(export MASTER_ADDR=127.0.0.1 MASTER_PORT=29624 WORLD_SIZE=2
 RANK=1 LOCAL_RANK=1 timeout 240 python benchmarks/nvls_profile.py > $O/r2c_prof_r1.log 2>&1 &
 RANK=0 LOCAL_RANK=0 timeout 240 ncu --set full --clock-control none --import-source on -k regex:pushpull -o $O/prof_nvls2 -f python benchmarks/nvls_profile.py > $O/r2c_prof_r0.log 2>&1
 wait) 
tail -n 12 $O/r2c_t_1gpu.log $O/r2c_t_nvls2.log $O/r2c_t_mgpu2.log $O/r2c_b_compress2.log $O/r2c_b_resnet2_ddp.log $O/r2c_b_bert2_ddp.log $O/r2c_b_ps_pipe1.log $O/r2c_b_ps_pipe0.log
tail -n 5 $O/r2c_prof_r0.log; ls -la $O/prof_nvls2.ncu-rep
