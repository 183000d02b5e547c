#include "capi/byteps_c_api.h"
#include "core/gpu_stage.h"

#include <dlfcn.h>

#include <unistd.h>

#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "core/env.h"
#include "core/handle_manager.h"
#include "core/log.h"
#include "core/ps_worker.h"
#include "core/registry.h"
#include "core/trace.h"
#include "net/van.h"
#include "server/server.h"

#define BPS_API extern "C" __attribute__((visibility("default")))

namespace {

using namespace bps;

struct IntAverage {     // integer tensors are averaged with a floor division once the sum is back
  void* data;
  int64_t nbytes;
  int dtype;
};

struct Runtime {
  std::mutex mu;
  bool inited = false;
  int rank = 0, size = 1, local_rank = 0, local_size = 1;
  int num_servers = 0;
  size_t partition_bound = 4096000;
  Registry registry;                       // survives suspend(): names keep their keys
  std::unique_ptr<net::Postoffice> po;
  std::unique_ptr<PSWorker> worker;
  HandleManager solo_handles;              // single-process jobs: push_pull completes at once
  std::map<int, IntAverage> int_avg;       // handle -> pending floor division
  std::map<int, int64_t> bytes_of;         // handle -> bytes (telemetry)
  std::map<std::string, bool> keys_inited;
  // device tensors: CUDA helper library (dlopen'ed on first use), per-device staging contexts, pinned staging
  void* cuda_lib = nullptr;
  const BpsGpuStageFns* gpu_fns = nullptr;
  void* (*cuda_stage_create)(int) = nullptr;
  void* (*cuda_host_alloc)(size_t) = nullptr;
  void (*cuda_host_free)(void*) = nullptr;
  int (*cuda_d2h_sync)(int, void*, const void*, size_t) = nullptr;
  int (*cuda_stream_wait_event)(void*, void*) = nullptr;
  int (*cuda_event_sync)(void*) = nullptr;
  int (*cuda_pointer_device)(const void*) = nullptr;
  std::map<int, void*> gpu_ctx;                         // device -> staging context
  std::map<std::string, std::pair<void*, size_t>> dev_staging;   // tensor -> pinned host buffer
  Telemetry telemetry;
  std::string error;
};

Runtime& rt() {
  static Runtime* r = new Runtime();       // leaked on purpose: may be used during static destruction
  return *r;
}

int fail(const std::string& why) {
  rt().error = why;
  return -1;
}

int start_locked(Runtime& r, int num_workers_override, int num_servers_override, int rank_override) {
  if (r.inited) return 0;
  const int boxes = (int)env_int("DMLC_NUM_WORKER", 1);
  r.local_size = (int)std::max<long long>(1, env_int("BYTEPS_LOCAL_SIZE", 1));
  r.local_rank = (int)env_int("BYTEPS_LOCAL_RANK", 0);
  const int box = (int)env_int("DMLC_WORKER_ID", 0);
  r.size = num_workers_override > 0 ? num_workers_override : boxes * r.local_size;
  r.rank = rank_override >= 0 ? rank_override
                              : (int)env_int("BYTEPS_GLOBAL_RANK", r.local_rank + (long long)box * r.local_size);
  r.num_servers = num_servers_override >= 0 ? num_servers_override : (int)env_int("DMLC_NUM_SERVER", 0);
  // partitions are page-aligned per local device, like the reference (global.cc:129-137)
  const long long page = sysconf(_SC_PAGESIZE) > 0 ? sysconf(_SC_PAGESIZE) : 4096;
  const long long unit = page * r.local_size;
  long long bound = env_int("BYTEPS_PARTITION_BYTES", 4096000);
  r.partition_bound = (size_t)((bound + unit - 1) / unit * unit);
  r.telemetry.configure(env_bool("BYTEPS_TELEMETRY_ON", true), 10.0);
  if (r.num_servers <= 0) {
    if (r.size > 1) return fail("the C API needs servers (DMLC_NUM_SERVER >= 1) for jobs of more than one process");
    r.inited = true;
    return 0;
  }
  net::NetConfig cfg = net::NetConfig::from_env();
  cfg.role = net::Role::kWorker;
  cfg.num_workers = r.size;          // every process is a PS worker node
  cfg.num_servers = r.num_servers;
  cfg.rank_hint = r.rank;
  r.po.reset(new net::Postoffice(cfg));
  PSWorkerConfig wc = PSWorkerConfig::from_env();
  wc.num_pushers = r.size;
  r.worker.reset(new PSWorker(r.po.get(), wc));
  r.po->Start(0, true);
  r.inited = true;
  return 0;
}

int stop_locked(Runtime& r) {
  if (!r.inited) return 0;
  if (r.worker) r.worker->Stop();
  if (r.po) r.po->Finalize(0, true);
  r.worker.reset();
  r.po.reset();
  r.keys_inited.clear();
  r.registry.reset_contexts();
  r.inited = false;
  return 0;
}

void finish_locked(Runtime& r, int handle) {
  auto it = r.int_avg.find(handle);
  if (it != r.int_avg.end()) {
    const IntAverage& a = it->second;
    const int64_t n = a.nbytes / dtype_size(a.dtype);
    const int64_t d = r.size;
    auto floordiv = [d](int64_t v) { return (v >= 0) ? v / d : -((-v + d - 1) / d); };
    switch (a.dtype) {
      case I32: for (int64_t i = 0; i < n; ++i) ((int32_t*)a.data)[i] = (int32_t)floordiv(((int32_t*)a.data)[i]); break;
      case I64: for (int64_t i = 0; i < n; ++i) ((int64_t*)a.data)[i] = floordiv(((int64_t*)a.data)[i]); break;
      case I8: for (int64_t i = 0; i < n; ++i) ((int8_t*)a.data)[i] = (int8_t)floordiv(((int8_t*)a.data)[i]); break;
      case U8: for (int64_t i = 0; i < n; ++i) ((uint8_t*)a.data)[i] = (uint8_t)(((uint8_t*)a.data)[i] / d); break;
      default: break;
    }
    r.int_avg.erase(it);
  }
  auto b = r.bytes_of.find(handle);
  if (b != r.bytes_of.end()) {
    if (r.telemetry.should_record()) r.telemetry.record((size_t)b->second);
    r.bytes_of.erase(b);
  }
}

}  // namespace

BPS_API int byteps_init(void) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  return start_locked(r, -1, -1, -1);
}

BPS_API int byteps_lazy_init(void) { return byteps_init(); }

BPS_API int byteps_shutdown(void) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  return stop_locked(r);
}

BPS_API int byteps_suspend(void) { return byteps_shutdown(); }

BPS_API int byteps_resume(int num_workers, int num_servers, int global_rank) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  if (r.inited) return fail("resume() while running: call suspend() first");
  return start_locked(r, num_workers, num_servers, global_rank);
}

BPS_API int byteps_rank(void) { return rt().rank; }
BPS_API int byteps_size(void) { return rt().size; }
BPS_API int byteps_local_rank(void) { return rt().local_rank; }
BPS_API int byteps_local_size(void) { return rt().local_size; }

BPS_API int byteps_declare_tensor(const char* name) {
  if (!name) return fail("null tensor name");
  return (int)rt().registry.declare(std::string("byteps.") + name);
}

BPS_API int byteps_declare_tensor_kwargs(const char* name, const char* const* keys, const char* const* values, int n) {
  int key = byteps_declare_tensor(name);
  if (key < 0) return key;
  std::unordered_map<std::string, std::string> kw;
  for (int i = 0; i < n; ++i) {
    std::string k = keys[i] ? keys[i] : "";
    if (k.rfind("byteps_", 0) == 0) k = k.substr(7);      // "byteps_compressor_type" -> "compressor_type"
    kw[k] = values[i] ? values[i] : "";
  }
  rt().registry.set_kwargs(std::string("byteps.") + name, kw);
  return key;
}

BPS_API int byteps_push_pull(const char* name, void* data, int64_t nbytes, int dtype, int average, int priority,
                             int version) {
  Runtime& r = rt();
  if (!name || (!data && nbytes > 0) || nbytes < 0) return fail("bad push_pull arguments");
  const int es = dtype_size(dtype);
  if (es == 0 || nbytes % es != 0) return fail("bad dtype / size");
  std::lock_guard<std::mutex> g(r.mu);
  if (!r.inited) return fail("byteps_init() has not been called");
  const std::string full = std::string("byteps.") + name;
  if (!r.worker) {                                  // one process: the sum over one worker is the input
    int h = r.solo_handles.allocate();
    r.solo_handles.mark_done(h, Status::OK());
    r.bytes_of[-h - 2] = nbytes;
    return -h - 2;                                  // solo handles live in their own (negative) id space
  }
  r.registry.declare(full);
  auto ctx = r.registry.context(full);
  r.registry.init_tensor(ctx, (size_t)nbytes, dtype, r.partition_bound, 4096);
  if (ctx->nbytes != (size_t)nbytes || ctx->dtype != dtype)
    return fail("tensor " + full + " changed size or dtype since its first push_pull");
  std::vector<PSWorker::Part> parts;
  for (size_t i = 0; i < ctx->parts.size(); ++i) parts.push_back({ctx->keys[i], ctx->parts[i].offset, ctx->parts[i].len});
  if (!r.keys_inited[full]) {
    // blocking init push per partition: allocates the server-side store and is a barrier across the workers
    for (auto& p : parts) {
      r.worker->InitKey(p.key, (char*)data + p.offset, p.len, dtype, 0);
      if (!ctx->kwargs.empty() && dtype_is_float(dtype)) r.worker->RegisterCompressor(p.key, ctx->kwargs, p.len, dtype);
    }
    r.keys_inited[full] = true;
  }
  const bool fl = dtype_is_float(dtype);
  const double scale = (average && fl) ? 1.0 / r.size : 1.0;
  int h = r.worker->PushPull(full, data, dtype, parts, priority, version, scale, nullptr);
  if (average && !fl) r.int_avg[h] = IntAverage{data, nbytes, dtype};
  r.bytes_of[h] = nbytes;
  return h;
}

namespace {
// libbyteps_b200_cuda.so lives next to this library; BYTEPS_CUDA_LIB overrides the path
bool load_cuda_locked(Runtime& r) {
  if (r.gpu_fns) return true;
  std::string path = env_str("BYTEPS_CUDA_LIB", "");
  if (path.empty()) {
    Dl_info info;
    if (dladdr((void*)&byteps_last_error, &info) && info.dli_fname) {
      path = info.dli_fname;
      size_t slash = path.rfind('/');
      path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/libbyteps_b200_cuda.so";
    } else {
      path = "libbyteps_b200_cuda.so";
    }
  }
  r.cuda_lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!r.cuda_lib) {
    fail(std::string("cannot load the CUDA helper library: ") + dlerror());
    return false;
  }
  auto sym = [&](const char* n) { return dlsym(r.cuda_lib, n); };
  auto fns = (const BpsGpuStageFns* (*)(void))sym("byteps_cuda_stage_fns");
  r.cuda_stage_create = (void* (*)(int))sym("byteps_cuda_stage_create");
  r.cuda_host_alloc = (void* (*)(size_t))sym("byteps_cuda_host_alloc");
  r.cuda_host_free = (void (*)(void*))sym("byteps_cuda_host_free");
  r.cuda_d2h_sync = (int (*)(int, void*, const void*, size_t))sym("byteps_cuda_d2h_sync");
  r.cuda_stream_wait_event = (int (*)(void*, void*))sym("byteps_cuda_stream_wait_event");
  r.cuda_event_sync = (int (*)(void*))sym("byteps_cuda_event_sync");
  r.cuda_pointer_device = (int (*)(const void*))sym("byteps_cuda_pointer_device");
  if (!fns || !r.cuda_stage_create || !r.cuda_host_alloc || !r.cuda_d2h_sync || !r.cuda_stream_wait_event ||
      !r.cuda_event_sync || !r.cuda_pointer_device) {
    fail("the CUDA helper library lacks expected symbols");
    return false;
  }
  r.gpu_fns = fns();
  return true;
}
}  // namespace

BPS_API int byteps_push_pull_device(const char* name, void* dev, int64_t nbytes, int dtype, int average, int priority,
                                    int version, void* ready_event) {
  Runtime& r = rt();
  if (!name || (!dev && nbytes > 0) || nbytes < 0) return fail("bad push_pull arguments");
  const int es = dtype_size(dtype);
  if (es == 0 || nbytes % es != 0) return fail("bad dtype / size");
  std::lock_guard<std::mutex> g(r.mu);
  if (!r.inited) return fail("byteps_init() has not been called");
  if (!r.worker) {                                  // one process: the sum over one worker is the input
    int h = r.solo_handles.allocate();
    r.solo_handles.mark_done(h, Status::OK());
    r.bytes_of[-h - 2] = nbytes;
    return -h - 2;
  }
  if (!dtype_is_float(dtype) && average) return fail("integer averages are host-only (sum on the device, divide later)");
  if (!load_cuda_locked(r)) return -1;
  const int device = r.cuda_pointer_device(dev);
  if (device < 0) return fail("byteps_push_pull_device needs a CUDA device pointer");
  void*& gctx = r.gpu_ctx[device];
  if (!gctx) gctx = r.cuda_stage_create(device);
  if (!gctx) return fail("cannot create the staging streams on the device");
  r.worker->set_gpu_stage(r.gpu_fns);
  const std::string full = std::string("byteps.") + name;
  r.registry.declare(full);
  auto ctx = r.registry.context(full);
  r.registry.init_tensor(ctx, (size_t)nbytes, dtype, r.partition_bound, 4096);
  if (ctx->nbytes != (size_t)nbytes || ctx->dtype != dtype)
    return fail("tensor " + full + " changed size or dtype since its first push_pull");
  auto& stg = r.dev_staging[full];
  if (!stg.first) {
    stg.first = r.cuda_host_alloc((size_t)std::max<int64_t>(nbytes, 16));
    stg.second = (size_t)nbytes;
    if (!stg.first) return fail("cannot allocate pinned host staging");
  }
  std::vector<PSWorker::Part> parts;
  for (size_t i = 0; i < ctx->parts.size(); ++i) parts.push_back({ctx->keys[i], ctx->parts[i].offset, ctx->parts[i].len});
  if (!r.keys_inited[full]) {
    if (ready_event) r.cuda_event_sync(ready_event);
    if (r.cuda_d2h_sync(device, stg.first, dev, (size_t)nbytes) != 0) return fail("cudaMemcpy (init push) failed");
    for (auto& p : parts) {
      r.worker->InitKey(p.key, (char*)stg.first + p.offset, p.len, dtype, 0);
      if (!ctx->kwargs.empty()) r.worker->RegisterCompressor(p.key, ctx->kwargs, p.len, dtype);
    }
    r.keys_inited[full] = true;
  }
  const double scale = average ? 1.0 / r.size : 1.0;
  int h = r.worker->PushPullDevice(full, dev, dev, stg.first, dtype, parts, priority, version, scale, ready_event, gctx);
  r.bytes_of[h] = nbytes;
  return h;
}

BPS_API int byteps_wait_device(int handle, void* stream) {
  Runtime& r = rt();
  if (handle <= -2) return byteps_wait(handle);
  PSWorker* w;
  {
    std::lock_guard<std::mutex> g(r.mu);
    w = r.worker.get();
    if (!w) return fail("no push_pull in flight");
  }
  Status s = w->Wait(handle);
  void* done = w->TakeDoneEvent(handle);
  std::lock_guard<std::mutex> g(r.mu);
  if (!s.ok()) return fail("push_pull failed: " + s.reason);
  finish_locked(r, handle);
  if (done) {
    const int rc = stream == (void*)-1 ? r.cuda_event_sync(done) : r.cuda_stream_wait_event(stream, done);
    if (rc != 0) return fail("waiting for the COPYH2D stage failed");
  }
  return 0;
}

BPS_API int byteps_poll(int handle) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  if (handle <= -2) return r.solo_handles.poll(-handle - 2) ? 1 : 0;
  if (!r.worker) return (int)fail("no push_pull in flight");
  return r.worker->Poll(handle) ? 1 : 0;
}

BPS_API int byteps_wait(int handle) {
  Runtime& r = rt();
  PSWorker* w;
  {
    std::lock_guard<std::mutex> g(r.mu);
    if (handle <= -2) {
      r.solo_handles.wait_and_release(-handle - 2);
      finish_locked(r, handle);
      return 0;
    }
    w = r.worker.get();
    if (!w) return fail("no push_pull in flight");
  }
  Status s = w->Wait(handle);          // not under the lock: other threads keep enqueueing
  std::lock_guard<std::mutex> g(r.mu);
  if (!s.ok()) return fail("push_pull failed: " + s.reason);
  finish_locked(r, handle);
  return 0;
}

BPS_API void* byteps_shm_alloc(const char* name, int64_t nbytes) {
  if (!name || nbytes <= 0) {
    fail("bad shm_alloc arguments");
    return nullptr;
  }
  const std::string full = "BytePS_ShM_" + std::to_string((long)getpid()) + "_" + name;
  void* p = net::ShmRegistry::get().create(full, (size_t)nbytes);
  if (!p) fail("cannot create shared-memory window " + full);
  return p;
}

BPS_API int byteps_shm_free(const char* name) {
  if (!name) return fail("null name");
  net::ShmRegistry::get().release("BytePS_ShM_" + std::to_string((long)getpid()) + "_" + name);
  return 0;
}

BPS_API int byteps_server(void) {
  const std::string role = env_str("DMLC_ROLE", "server");
  if (role != "server" && role != "scheduler") return fail("byteps_server(): DMLC_ROLE must be server or scheduler");
  net::NetConfig cfg = net::NetConfig::from_env();
  cfg.role = role == "server" ? net::Role::kServer : net::Role::kScheduler;
  // DMLC_NUM_WORKER counts worker boxes; every process of a box is a transport-level node
  cfg.num_workers = (int)(env_int("DMLC_NUM_WORKER", 1) * std::max<long long>(1, env_int("BYTEPS_LOCAL_SIZE", 1)));
  if (role == "server" && env_has("DMLC_SERVER_ID")) cfg.rank_hint = (int)env_int("DMLC_SERVER_ID", -1);
  net::Postoffice po(cfg);
  std::unique_ptr<server::SumServer> srv;
  if (role == "server") srv.reset(new server::SumServer(&po, server::ServerConfig::from_env()));
  po.Start(0, !env_bool("BYTEPS_RECOVERY", false));     // a restarted node skips the start barrier
  po.Finalize(0, true);                                 // returns when every node has said goodbye
  if (srv) srv->Stop();
  return 0;
}

BPS_API void byteps_get_pushpull_speed(int64_t* ts_ms, double* mbps) {
  SpeedEntry e = rt().telemetry.get();
  if (ts_ms) *ts_ms = e.ts_ms;
  if (mbps) *mbps = e.mbps;
}

BPS_API const char* byteps_last_error(void) { return rt().error.c_str(); }
