// C interface between the CUDA-free runtime (_core) and the CUDA module (_cuda) for the device
// stages of the CPU-server pipeline: COPYD2H before PUSH and COPYH2D after PULL, per partition
// (/root/reference/byteps/common/core_loops.cc:378-443, 650-753).  _cuda implements the table
// (comm/gpu_stage.cc), python hands its address to PSWorker::set_gpu_stage.
#pragma once
#include <stddef.h>

extern "C" {

typedef void (*bps_host_cb)(void* arg);

struct BpsGpuStageFns {
  // make the D2H stream of `ctx` wait for `ready_event` (a cudaEvent_t; may be null)
  void (*wait_ready)(void* ctx, void* ready_event);
  // host <- dev on the D2H stream; returns an event (owned by ctx) that completes with the copy
  void* (*d2h)(void* ctx, void* host, const void* dev, size_t len);
  // 1 when the event has completed
  int (*query)(void* event);
  // dev <- host on the H2D stream (any thread); `cb(arg)` runs on a driver thread when the copy has landed
  // (cb may be null).  Returns 0 on success.
  int (*h2d)(void* ctx, void* dev, const void* host, size_t len, bps_host_cb cb, void* arg);
  // event (owned by ctx) on the H2D stream that covers everything enqueued so far
  void* (*h2d_mark)(void* ctx);
  // page-lock host memory that was mapped by somebody else (a server's shared-memory store) so it can be the
  // source of asynchronous H2D copies.  Returns 0 on success (already registered counts as success).
  int (*host_register)(void* ctx, void* ptr, size_t len);
  // dev[0..n) *= alpha on the H2D stream (dtype codes of core/types.h: F32, F16, BF16); 0 on success
  int (*scale)(void* ctx, void* dev, size_t nbytes, int dtype, double alpha);
};

}  // extern "C"
