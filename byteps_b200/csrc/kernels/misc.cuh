#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace bps {
// dst[0:nbytes) = src (host bytes, snapshotted at launch), stream ordered.
cudaError_t launch_write_blob(void* dst, const void* src, size_t nbytes, cudaStream_t stream);
// overwrite a buffer larger than L2 (benchmark hygiene)
cudaError_t launch_l2_flush(void* buf, size_t nbytes, uint32_t value, cudaStream_t stream);
// x[0..n) *= alpha in place; dtype: 0 f32, 1 bf16, 2 f16 (the 1/size of an averaged push_pull in CPU-server mode,
// applied on the device behind the COPYH2D instead of by a CPU pass over pinned memory)
cudaError_t launch_scale_inplace(void* x, size_t n, int dtype, float alpha, cudaStream_t stream);
}  // namespace bps
