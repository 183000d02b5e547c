// CUDA side of core/gpu_stage.h: two side streams and a ring of events per device context.
#include "core/gpu_stage.h"
#include "kernels/misc.cuh"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct StageCtx {
  int device;
  cudaStream_t d2h = nullptr, h2d = nullptr;
  std::vector<cudaEvent_t> events;
  size_t cursor = 0;
  std::mutex mu;
  // BYTEPS_STAGE_PROFILE=1: device-timed start / end of every staged copy (printed when the context is destroyed)
  struct Rec {
    cudaEvent_t a, b;
    size_t len;
    bool h2d;
  };
  bool profile = false;
  std::vector<Rec> recs;

  Rec* begin_rec(cudaStream_t st, size_t len, bool h2d) {
    if (!profile) return nullptr;
    Rec r{nullptr, nullptr, len, h2d};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> g(mu);
    recs.push_back(r);
    return &recs.back();
  }

  cudaEvent_t next_event() {
    std::lock_guard<std::mutex> g(mu);
    cudaEvent_t e = events[cursor % events.size()];
    ++cursor;
    return e;
  }
};

void chk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

void st_wait_ready(void* c, void* ev) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);
  if (ev) cudaStreamWaitEvent(s->d2h, (cudaEvent_t)ev, 0);
}

void* st_d2h(void* c, void* host, const void* dev, size_t len) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);
  cudaEvent_t pb = nullptr;
  if (s->profile) pb = s->begin_rec(s->d2h, len, false)->b;
  cudaMemcpyAsync(host, dev, len, cudaMemcpyDeviceToHost, s->d2h);
  if (pb) cudaEventRecord(pb, s->d2h);
  cudaEvent_t e = s->next_event();
  cudaEventRecord(e, s->d2h);
  return (void*)e;
}

int st_query(void* ev) {
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaSuccess) return 1;
  if (e != cudaErrorNotReady) cudaGetLastError();
  return 0;
}

int st_h2d(void* c, void* dev, const void* host, size_t len, bps_host_cb cb, void* arg) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);       // called from transport / pool threads: they start on device 0
  cudaEvent_t pb = nullptr;
  if (s->profile) pb = s->begin_rec(s->h2d, len, true)->b;
  cudaError_t e = cudaMemcpyAsync(dev, host, len, cudaMemcpyHostToDevice, s->h2d);
  if (pb) cudaEventRecord(pb, s->h2d);
  if (e == cudaSuccess && cb) e = cudaLaunchHostFunc(s->h2d, cb, arg);
  return e == cudaSuccess ? 0 : (int)e;
}

void* st_h2d_mark(void* c) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);
  cudaEvent_t e = s->next_event();
  cudaEventRecord(e, s->h2d);
  return (void*)e;
}

int st_host_register(void* c, void* ptr, size_t len) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);
  cudaError_t e = cudaHostRegister(ptr, len, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return 0;
  }
  if (e != cudaSuccess) cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int st_scale(void* c, void* dev, size_t nbytes, int dtype, double alpha) {
  StageCtx* s = (StageCtx*)c;
  cudaSetDevice(s->device);
  // core dtype codes: F32 = 0, F16 = 2, BF16 = 7 -> kernel codes 0 / 2 / 1
  int code = dtype == 0 ? 0 : dtype == 7 ? 1 : dtype == 2 ? 2 : -1;
  if (code < 0) return -1;
  return bps::launch_scale_inplace(dev, nbytes / (code == 0 ? 4 : 2), code, (float)alpha, s->h2d) == cudaSuccess ? 0 : -1;
}

const BpsGpuStageFns kFns = {st_wait_ready, st_d2h, st_query, st_h2d, st_h2d_mark, st_host_register, st_scale};

}  // namespace

namespace bps {

void* gpu_stage_create(int device, int nevents) {
  StageCtx* s = new StageCtx();
  s->device = device;
  chk(cudaSetDevice(device), "cudaSetDevice");
  chk(cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking), "cudaStreamCreate");
  chk(cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking), "cudaStreamCreate");
  const char* pf = getenv("BYTEPS_STAGE_PROFILE");
  s->profile = pf && *pf && *pf != '0';
  if (s->profile) s->recs.reserve(1 << 16);
  s->events.resize(nevents > 16 ? nevents : 16);
  for (auto& e : s->events) chk(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
  return s;
}

void gpu_stage_destroy(void* c) {
  StageCtx* s = (StageCtx*)c;
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->d2h);
  cudaStreamSynchronize(s->h2d);
  if (s->profile && !s->recs.empty()) {
    // the last (up to) 64 copies: start (ms after the first of them) and duration, per direction, plus the rate
    // while a copy was actually running - tells a slow DMA from an idle stream
    const size_t n = s->recs.size(), from = n > 64 ? n - 64 : 0;
    double busy[2] = {0, 0}, bytes[2] = {0, 0};
    std::string line[2];
    for (size_t i = from; i < n; ++i) {
      const auto& r = s->recs[i];
      float t0 = 0, dt = 0;
      cudaEventElapsedTime(&t0, s->recs[from].a, r.a);
      cudaEventElapsedTime(&dt, r.a, r.b);
      busy[r.h2d] += dt;
      bytes[r.h2d] += (double)r.len;
      char buf[64];
      snprintf(buf, sizeof buf, " %.2f+%.2f", t0, dt);
      line[r.h2d] += buf;
    }
    for (int d = 0; d < 2; ++d)
      fprintf(stderr, "[byteps stage profile] device %d %s: %.1f MB in %.2f ms of copy time (%.1f GB/s while copying); start+duration ms:%s\n",
              s->device, d ? "H2D" : "D2H", bytes[d] / 1e6, busy[d], busy[d] > 0 ? bytes[d] / busy[d] / 1e6 : 0.0,
              line[d].c_str());
    for (auto& r : s->recs) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
  }
  for (auto e : s->events) cudaEventDestroy(e);
  cudaStreamDestroy(s->d2h);
  cudaStreamDestroy(s->h2d);
  delete s;
}

const BpsGpuStageFns* gpu_stage_fns() { return &kFns; }

}  // namespace bps
