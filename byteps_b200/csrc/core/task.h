// Per-tensor context and per-partition task records.
//
// Parity: BPSContext / TensorTableEntry in /root/reference/byteps/common/common.h:177-264.
// Differences: tasks carry raw (device or host) pointers + a byte range instead
// of a framework Tensor object, and readiness is a generic predicate so the GPU
// path can use a stream-ordered event while the CPU path is always-ready.
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "core/types.h"

namespace bps {

class Compressor;

struct TensorContext {
  std::string name;
  uint32_t declared_key = 0;
  bool initialized = false;
  int dtype = F32;
  size_t nbytes = 0;                 // logical tensor bytes
  size_t aligned_bytes = 0;          // page-rounded staging size
  std::vector<uint64_t> keys;        // one per partition
  std::vector<Partition> parts;      // byte ranges
  void* host_buff = nullptr;         // pinned/shm staging (CPU-server mode)
  std::string shm_name;              // name of the POSIX shm object backing host_buff
  std::unordered_map<std::string, std::string> kwargs;  // compressor configuration
  std::vector<std::shared_ptr<Compressor>> compressors; // one per partition (may be empty)
  std::atomic<int64_t> step_cnt{0};  // completed push_pulls (trace window)
  // trace bookkeeping
  int64_t enqueue_ts_us = 0;
};

struct Task {
  std::shared_ptr<TensorContext> ctx;
  uint64_t key = 0;
  int priority = 0;
  int version = 0;
  int dtype = F32;
  int device = -1;            // -1 = host memory
  void* input = nullptr;      // base pointer of the whole tensor
  void* output = nullptr;     // base pointer of the whole output tensor
  void* host = nullptr;       // base pointer of host staging (may be null)
  void* host_out = nullptr;   // host tasks: where the result goes when it is not the staging buffer itself
  size_t offset = 0;          // byte offset of this partition
  size_t len = 0;             // byte length of this partition
  size_t compressed_len = 0;  // bytes actually on the wire when compressed
  void* compressed = nullptr;
  std::vector<int> stages;    // Stage list
  size_t stage_idx = 0;
  int handle = -1;
  uint32_t total_parts = 1;
  std::shared_ptr<std::atomic<uint32_t>> done_counter;
  std::function<bool()> ready;             // nullptr => always ready
  std::function<void(const Status&)> on_all_done;  // fired by the last partition
  uint64_t seq = 0;           // arrival order (FIFO tiebreak)
  int64_t stage_start_us = 0; // trace
  // device-staged tasks (CPU-server mode with GPU tensors): where the result goes back to, and through what
  void* dev_out = nullptr;    // base pointer of the device output tensor (null: host task)
  void* gpu_ctx = nullptr;    // context of the BpsGpuStageFns table
  double scale = 1.0;         // applied to the partition on the host before COPYH2D
  void* h2d_src = nullptr;    // pull answered by reference: the value in the server's shared-memory store
  void* h2d_region = nullptr; // ... and the mapping it lives in (page-locked once per mapping)
  size_t h2d_region_len = 0;
  int64_t d2h_start_us = 0;   // trace: when the partition's D2H copy was enqueued
  int current_stage() const { return stage_idx < stages.size() ? stages[stage_idx] : -1; }
};

using TaskPtr = std::shared_ptr<Task>;

}  // namespace bps
