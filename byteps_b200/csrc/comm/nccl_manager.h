// Native NCCL manager - the BASELINE arm only.
//
// Parity: NcclManager + the REDUCE/BROADCAST halves of the stage loops
// (/root/reference/byteps/common/nccl_manager.cc:23-181, core_loops.cc:190-376):
// BYTEPS_NCCL_NUM_RINGS communicators with one highest-priority stream each
// (chosen by key % num_rings), unique ids created on the root and exchanged by
// the caller, per-partition ncclReduceScatter (+ncclReduce of the len % nGPU
// tail to the root) then ncclAllGather (+ncclBroadcast of the tail), issued in
// ncclGroupStart/End batches of BYTEPS_NCCL_GROUP_SIZE tasks, one cudaEvent per
// task.  libnccl is resolved at run time with dlopen (the copy torch already
// loaded), so the extension has no link-time NCCL dependency.  Nothing on the
// product path calls into this file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace bps {

class NcclManager {
 public:
  NcclManager(int rank, int world, int device, int num_rings, int group_size);
  ~NcclManager();
  static bool available();
  // 128-byte opaque id (call on the root, once per ring)
  static std::string make_unique_id();
  void init(const std::vector<std::string>& ids);
  int root() const { return world_ - 1; }   // the reference's root is the highest local rank
  // In-place reference-style push_pull.  dtype: bps::DataType code.  Waits for `ready`
  // (may be null) on the ring stream(s), records `done` (may be null) after the last task.
  void push_pull(void* ptr, size_t nbytes, int dtype, uint64_t first_key, size_t partition_bytes, cudaEvent_t ready,
                 cudaEvent_t done);
  size_t tasks_issued() const { return tasks_; }
  cudaStream_t stream(int ring) const { return streams_[ring % streams_.size()]; }
  int num_rings() const { return num_rings_; }

 private:
  int rank_, world_, device_, num_rings_, group_size_;
  std::vector<void*> comms_;
  std::vector<cudaStream_t> streams_;
  std::vector<cudaEvent_t> events_;
  size_t tasks_ = 0;
};

}  // namespace bps
