// libbyteps_b200_cuda.so: the CUDA half of the C API (capi/byteps_c_api.h, byteps_push_pull_device).
// libbyteps_b200.so carries no CUDA dependency; it dlopens this library the first time a device
// pointer is handed to it and drives the COPYD2H / COPYH2D stages through the table of
// core/gpu_stage.h (implemented in comm/gpu_stage.cc, which is linked in here).
#include <cuda_runtime_api.h>

#include "core/gpu_stage.h"

namespace bps {
void* gpu_stage_create(int device, int nevents);
void gpu_stage_destroy(void* c);
const BpsGpuStageFns* gpu_stage_fns();
}  // namespace bps

#define BPS_CUDA_API extern "C" __attribute__((visibility("default")))

BPS_CUDA_API const BpsGpuStageFns* byteps_cuda_stage_fns(void) { return bps::gpu_stage_fns(); }
BPS_CUDA_API void* byteps_cuda_stage_create(int device) {
  try {
    return bps::gpu_stage_create(device, 8192);
  } catch (...) {
    return nullptr;
  }
}
BPS_CUDA_API void byteps_cuda_stage_destroy(void* c) { bps::gpu_stage_destroy(c); }
// pinned host staging
BPS_CUDA_API void* byteps_cuda_host_alloc(size_t nbytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, nbytes, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
}
BPS_CUDA_API void byteps_cuda_host_free(void* p) { cudaFreeHost(p); }
// blocking D2H (the init push needs the bytes on the host)
BPS_CUDA_API int byteps_cuda_d2h_sync(int device, void* host, const void* dev, size_t nbytes) {
  cudaSetDevice(device);
  return cudaMemcpy(host, dev, nbytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
BPS_CUDA_API int byteps_cuda_stream_wait_event(void* stream, void* event) {
  return cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0) == cudaSuccess ? 0 : -1;
}
BPS_CUDA_API int byteps_cuda_event_sync(void* event) {
  return cudaEventSynchronize((cudaEvent_t)event) == cudaSuccess ? 0 : -1;
}
BPS_CUDA_API int byteps_cuda_pointer_device(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return a.type == cudaMemoryTypeDevice ? a.device : -1;
}
