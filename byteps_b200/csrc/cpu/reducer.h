// Host-side elementwise reducer used by the CPU-server mode, the KV server and
// the CPU compressors.
//
// Parity: CpuReducer in /root/reference/byteps/common/cpu_reducer.cc:59-437
// (sum variants for 7 dtypes, OpenMP threads from BYTEPS_OMP_THREAD_PER_GPU,
// AVX+F16C fp16 path).  New here: bf16, AVX-512 kernels chosen at run time, and
// a scale() used for averaging host tensors.
#pragma once
#include <cstddef>
#include <cstdint>

#include "core/types.h"

namespace bps {

class CpuReducer {
 public:
  explicit CpuReducer(int num_threads = 0);  // 0 -> BYTEPS_OMP_THREAD_PER_GPU or 4
  // dst += src
  int sum(void* dst, const void* src, size_t nbytes, int dtype) const;
  // dst = a + b
  int sum(void* dst, const void* a, const void* b, size_t nbytes, int dtype) const;
  // dst += alpha * src
  int sum_scaled(void* dst, const void* src, size_t nbytes, int dtype, float alpha) const;
  // dst = a + alpha * b
  int sum_scaled(void* dst, const void* a, const void* b, size_t nbytes, int dtype, float alpha) const;
  // dst *= alpha (integers: floor-divide by round(1/alpha) when alpha < 1)
  int scale(void* dst, size_t nbytes, int dtype, double alpha) const;
  void copy(void* dst, const void* src, size_t nbytes) const;
  int num_threads() const { return nthreads_; }
  static bool has_avx512();

 private:
  int nthreads_;
};

}  // namespace bps
