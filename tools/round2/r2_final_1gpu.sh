#!/bin/bash
# Last GPU call of round 2 (1 GPU): whole GPU suite, compute-sanitizer memcheck over one world-1 case per kernel
# family, PCIe staging probe, smoke + bench.
O=gpurun_out
mkdir -p $O
timeout 170 python -m pytest tests -m gpu -q --timeout 90 -p no:cacheprovider > $O/r2_final_tests.log 2>&1
tail -4 $O/r2_final_tests.log
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file $O/sanitizer_memcheck.log \
  python -m pytest -q -p no:cacheprovider --timeout 70 \
  "tests/test_gpu_kernels.py::test_inplace_pushpull_virtual[bf16-1]" \
  "tests/test_gpu_kernels.py::test_inplace_pushpull_tma_virtual[bf16-1]" \
  "tests/test_gpu_kernels.py::test_packed_pushpull_virtual[f32-bf16-True-1]" \
  "tests/test_gpu_kernels.py::test_fused_optimizer_virtual[lsu-bf16-adam-1]" \
  "tests/test_gpu_kernels.py::test_fused_optimizer_virtual[tma-bf16-sgd-1]" \
  "tests/test_gpu_ring.py::test_ring_allreduce_virtual[True-bf16-1]" \
  "tests/test_gpu_compress.py::test_gpu_compressor_matches_cpu_reference[kw0-1]" \
  "tests/test_gpu_compress.py::test_gpu_compressor_matches_cpu_reference[kw1-1]" \
  > $O/sanitizer_memcheck_pytest.log 2>&1
echo "sanitizer rc=$?" >> $O/sanitizer_memcheck_pytest.log
tail -3 $O/sanitizer_memcheck_pytest.log; tail -5 $O/sanitizer_memcheck.log
timeout 30 python tools/memcpy_probe.py > $O/r2_final_probe.log 2>&1; cat $O/r2_final_probe.log
timeout 60 python __graft_entry__.py smoke > $O/r2_final_smoke.log 2>&1; tail -2 $O/r2_final_smoke.log
timeout 90 python bench.py > $O/r2_final_bench.log 2>&1; tail -1 $O/r2_final_bench.log
