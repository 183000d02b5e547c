// Worker side of the CPU-server mode: the host pipeline
//   [COMPRESS ->] PUSH -> PULL [-> DECOMPRESS]
// per partition, in (priority desc, key asc) order under byte credits.
//
// Parity: the PUSH/PULL/COMPRESS/DECOMPRESS stage loops and FinishOrProceed of
// /root/reference/byteps/common/core_loops.cc:31-137,498-648, InitTensor's
// blocking init push + compressor registration (operations.cc:283-414) and
// key->server placement (global.cc:628-677).  Differences: stage hand-off is
// callback driven (push ack -> pull request -> completion) instead of 1 us
// polling threads; readiness of GPU-staged data is a device-event predicate; the
// GPU stages around it (REDUCE/COPYD2H before, COPYH2D/BROADCAST after) are
// stream-ordered kernels/copies issued by the caller.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "compress/compressor.h"
#include "core/gpu_stage.h"
#include "core/handle_manager.h"
#include "core/ready_table.h"
#include "core/registry.h"
#include "core/scheduler.h"
#include "core/thread_pool.h"
#include "core/trace.h"
#include "cpu/reducer.h"
#include "net/kv_app.h"

namespace bps {

struct PSWorkerConfig {
  std::string hash_fn = "djb2";
  bool mixed_mode = false;
  int mixed_bound = 101;
  uint64_t credit_bytes = 0;          // 0 = unlimited
  size_t min_compress_bytes = 65536;
  int threadpool_size = 4;
  int num_pushers = 0;                // workers that push each key (for mixed-mode hashing)
  static PSWorkerConfig from_env();
};

// int (*)(void* event) -> 1 when the device event has completed
using EventQueryFn = int (*)(void*);

class PSWorker {
 public:
  PSWorker(net::Postoffice* po, const PSWorkerConfig& cfg, int app_id = 0, int customer_id = 0);
  ~PSWorker();
  void Stop();
  void set_event_query(EventQueryFn fn) { event_query_ = fn; }
  void set_timeline(Timeline* t) { timeline_ = t; }
  // NUMA node of the GPU (or of the memory) this worker stages through; sent with init pushes when
  // BYTEPS_NUMA_AWARE=1 so the server places the key's store there (core/numa.h).  Default: BYTEPS_NUMA_NODE or -1.
  void set_numa_node(int node) { numa_node_ = node; }
  int numa_node() const { return numa_node_; }

  // Blocking: announce a key range to its server (global barrier across pushers).
  // `pushers` = how many nodes push this key each round (0 = every worker).
  void InitKey(uint64_t key, const void* data, size_t len, int dtype, int pushers = 0);
  // Blocking: create the worker-side compressor and ship kwargs to the server.
  void RegisterCompressor(uint64_t key, const Kwargs& kw, size_t len, int dtype);
  bool HasCompressor(uint64_t key);
  void SetLearningRate(double lr);

  struct Part {
    uint64_t key;
    size_t offset;
    size_t len;
  };
  // Asynchronous push_pull of host memory `ptr` (in place).  `ready_event`
  // (may be null) must have completed before the data is touched.  When every
  // partition is back, the buffer is scaled by `scale` (1 = sum) and the handle
  // completes.  Returns the handle id.
  // `out` (optional, host memory, same layout as ptr): the result is delivered THERE instead of in `ptr` - per
  // partition, scaled, from a pool thread as soon as its pull completes.  With a colocated server the partition is
  // copied straight out of the server's shared-memory store (pull by reference): `ptr` (a registered window the
  // caller staged the input in) is then only read, and neither the server nor the caller copies the result twice.
  int PushPull(const std::string& name, void* ptr, int dtype, const std::vector<Part>& parts, int priority,
               int version, double scale, void* ready_event, void* out = nullptr);
  // Device tensors, pipelined per partition (reference: COPYD2H / PUSH / PULL / COPYH2D stage loops,
  // core_loops.cc:378-443, 538-618, 650-753): the D2H copy of every partition is issued right here on the
  // context's D2H stream (after `ready_event`), a partition is pushed as soon as ITS copy has landed, and its
  // H2D copy is issued from the pull completion - so D2H, PUSH, PULL and H2D of different partitions overlap
  // and both PCIe directions stay busy.  `host` is the pinned staging buffer, scale is applied on the host per
  // partition.  After Wait(handle), TakeDoneEvent(handle) is the event the consumer stream must wait on.
  void set_gpu_stage(const BpsGpuStageFns* fns) { gpu_ = fns; }
  int PushPullDevice(const std::string& name, const void* dev_in, void* dev_out, void* host, int dtype,
                     const std::vector<Part>& parts, int priority, int version, double scale, void* ready_event,
                     void* gpu_ctx);
  void* TakeDoneEvent(int handle);
  bool Poll(int handle) { return handles_.poll(handle); }
  Status Wait(int handle, int64_t timeout_ms = -1) { return handles_.wait_and_release(handle, timeout_ms); }
  int ServerOf(uint64_t key, size_t len) { return placer_->server_of(key, len); }
  std::vector<uint64_t> ServerLoad() { return placer_->load(); }
  uint64_t bytes_pushed() const { return bytes_pushed_; }

 private:
  struct Job;
  void DispatchLoop();
  void DoPush(const TaskPtr& t);
  void DoPull(const TaskPtr& t);
  void Finish(const TaskPtr& t);
  void DeliverHost(const TaskPtr& t, const void* src);
  std::shared_ptr<Compressor> CompressorOf(uint64_t key);

  net::Postoffice* po_;
  PSWorkerConfig cfg_;
  std::unique_ptr<net::KVWorker> kv_;
  std::unique_ptr<KeyPlacer> placer_;
  std::unique_ptr<ScheduledQueue> push_q_;
  std::unique_ptr<ThreadPool> pool_;
  CpuReducer reducer_;
  HandleManager handles_;
  void Sample(const TaskPtr& t, const char* stage);
  EventQueryFn event_query_ = nullptr;
  const BpsGpuStageFns* gpu_ = nullptr;
  std::mutex done_mu_;
  std::unordered_map<int, void*> done_events_;
  std::unordered_map<void*, bool> registered_;   // server store mappings page-locked for H2D (under done_mu_)
  bool pull_by_ref_ = true;          // BYTEPS_PS_PULL_BY_REF
  int numa_node_ = -1;
  std::string sample_name_;          // BYTEPS_DEBUG_SAMPLE_TENSOR: print first/last element after every stage
  Timeline* timeline_ = nullptr;
  std::thread dispatcher_;
  std::atomic<bool> stop_{false};
  bool eager_pull_ = true;   // BYTEPS_PS_EAGER_PULL=0: wait for the push acknowledgement before pulling
  std::mutex comp_mu_;
  std::unordered_map<uint64_t, std::shared_ptr<Compressor>> compressors_;
  std::unordered_map<uint64_t, std::shared_ptr<std::vector<char>>> comp_bufs_;
  double lr_ = -1.0;                 // last SetLearningRate (applied to compressors registered later too)
  std::atomic<uint64_t> bytes_pushed_{0};
  bool stopped_ = false;
};

}  // namespace bps
