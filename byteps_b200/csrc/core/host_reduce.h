// Box-local reduction of HOST tensors through shared memory, coordinated by the Unix-datagram signals of
// net/local_signal.h - the path the reference takes when the GPUs of a box cannot reduce among themselves over
// NVLink: every local rank copies its tensor into pinned shared memory, signals the root, the root's CPU reducer
// sums the copies, only the ROOT talks to the servers, and a broadcast signal tells the others to copy the result
// out (/root/reference/byteps/common/core_loops.cc:445-496 PCIE_REDUCE + CpuReducer, :650-753 COPYH2D after
// DO_COPYH2D, communicator.cc:28-276; shared_memory.cc:28-133 for the per-key shm objects).
//
// Here it serves CPU tensors of jobs that run several processes per box in CPU-server mode (and GPU tensors staged
// through the host on machines without peer access): L local ranks cost the servers one push per box instead of L.
//
//   region "BytePS_ShM_box<tag>_<key>":  [slot 0][slot 1]...[slot L-1][window]
//   every rank : contribute(key, src)  copy src -> slot[rank]; non-roots send REDUCE_READY
//   root       : reduce(key)           wait for L-1 REDUCE_READY, broadcast DO_REDUCE; EVERY rank then sums ITS
//                1/L of the slots into the window (the non-roots do that inside collect() and answer
//                PCIE_REDUCE_READY), so the summation runs on L ranks' cores instead of one; returns the window
//                once all shards are in (the caller pushes / pulls the window through PSWorker - it is a registered
//                BytePS_ShM_* object, so a colocated server reads and writes it by reference - or all-reduces it
//                with the other hosts' roots)
//   root       : publish(key, dst)     broadcast DO_BROADCAST, copy window -> dst, wait for L-1 BCAST_READY
//   non-root   : collect(key, dst)     serve DO_REDUCE(key) (sum my shard), wait for DO_BROADCAST(key),
//                                      copy window -> dst, send BCAST_READY
#pragma once
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "core/ready_table.h"
#include "cpu/reducer.h"
#include "net/local_signal.h"

namespace bps {

class HostLocalReduce {
 public:
  // tag: unique per (job, box) - shm object names and socket paths are derived from it
  HostLocalReduce(int local_rank, int local_size, const std::string& tag, int reducer_threads = 0,
                  const std::string& socket_dir = "");
  ~HostLocalReduce();
  bool is_root() const { return comm_->is_root(); }
  int local_rank() const { return rank_; }
  int local_size() const { return size_; }

  // all ranks; returns false if the shared region could not be mapped within timeout_ms
  bool contribute(uint64_t key, const void* src, size_t nbytes, int64_t timeout_ms = 60000);
  // root only: nullptr on timeout.  alpha != 1: every shard is scaled right after it was summed (an average
  // costs no extra pass on the root); the non-roots must pass the same dtype / alpha to collect()
  void* reduce(uint64_t key, size_t nbytes, int dtype, int64_t timeout_ms = -1, double alpha = 1.0);
  bool publish(uint64_t key, void* dst, size_t nbytes, int64_t timeout_ms = -1);
  // non-root only
  bool collect(uint64_t key, void* dst, size_t nbytes, int64_t timeout_ms = -1, int dtype = F32, double alpha = 1.0);
  // window of a key that has been contributed to (root: what reduce() returns)
  void* window(uint64_t key);
  uint64_t signals_received() const;

 private:
  struct Region {
    char* base = nullptr;
    size_t slot = 0;     // bytes per slot (page aligned)
    std::string name;
  };
  Region* region_of(uint64_t key, size_t nbytes, int64_t timeout_ms);
  void follower_loop();
  void sum_my_shard(Region* r, size_t nbytes, int dtype, double alpha);
  bool wait_announced(std::unordered_map<uint64_t, int>& m, uint64_t key, int64_t timeout_ms);

  int rank_, size_;
  std::string tag_;
  std::unique_ptr<LocalComm> comm_;
  std::shared_ptr<ReadyTable> reduce_ready_, shard_ready_, bcast_ready_;
  CpuReducer reducer_;
  std::mutex mu_;
  std::unordered_map<uint64_t, Region> regions_;
  // non-root: DO_BROADCAST keys received from the root but not collected yet
  std::thread follower_;
  bool stop_ = false;
  std::mutex bmu_;
  std::condition_variable bcv_;
  std::unordered_map<uint64_t, int> announced_;      // DO_BROADCAST
  std::unordered_map<uint64_t, int> reduce_asked_;   // DO_REDUCE
};

}  // namespace bps
