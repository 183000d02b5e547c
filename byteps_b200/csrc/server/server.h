// The CPU summation server ("parameter server" of the CPU-server mode).
//
// Parity: /root/reference/byteps/server/{server.cc,server.h,queue.h}.  Same
// per-key protocol - init push (allocate store, reply when every pusher
// arrived) / first push of a round (COPY_FIRST) / later pushes (SUM_RECV) /
// last push (ALL_RECV: publish merged buffer, flush parked pulls) / pulls
// served once per sender per round / async mode sums straight into the store /
// compressed pushes are decompressed before summing and the merged result is
// re-compressed - executed by engine threads fed from priority queues, keys
// load-balanced to threads by accumulated bytes.  Differences: per-key locks
// instead of one global handler mutex, the engine is an object (several
// servers can live in one process for tests), bf16 support.
#pragma once
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <thread>
#include <unordered_map>
#include <vector>

#include "compress/compressor.h"
#include "cpu/reducer.h"
#include "net/kv_app.h"

namespace bps {
namespace server {

enum EngineOp : int { COPY_FIRST = 0, SUM_RECV, ALL_RECV, TERMINATE };

struct EngineMessage {
  uint64_t id = 0;
  int op = TERMINATE;
  uint64_t key = 0;
  int dtype = F32;
  net::SArray<char> src;   // keeps the received payload alive until processed
  size_t len = 0;
  net::KVMeta req;
};

// mutex + condvar queue; with scheduling enabled, a heap ordered by the number
// of pushes already seen for the key (keys closer to completion first), then
// by arrival id (BYTEPS_SERVER_ENABLE_SCHEDULE; reference queue.h:31-105).
class PriorityQueue {
 public:
  explicit PriorityQueue(bool schedule) : schedule_(schedule) {}
  void Push(EngineMessage m);
  void WaitAndPop(EngineMessage* m);
  void ClearCounter(uint64_t key);
  size_t size();

 private:
  bool Before(const EngineMessage& a, const EngineMessage& b);
  bool schedule_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<EngineMessage> heap_;
  std::unordered_map<uint64_t, uint64_t> push_cnt_;
};

struct ServerConfig {
  int engine_threads = 4;       // BYTEPS_SERVER_ENGINE_THREAD
  bool enable_schedule = false; // BYTEPS_SERVER_ENABLE_SCHEDULE
  bool engine_blocking = false; // BYTEPS_SERVER_ENGINE_BLOCKING
  bool sync_mode = true;        // !BYTEPS_ENABLE_ASYNC
  int pushers_per_key = 0;      // 0 -> number of workers
  bool log_keys = false;        // PS_KEY_LOG
  int64_t debug_key = -1;       // BYTEPS_SERVER_DEBUG_KEY (with BYTEPS_SERVER_DEBUG)
  static ServerConfig from_env();
};

class SumServer {
 public:
  SumServer(net::Postoffice* po, const ServerConfig& cfg, int app_id = 0);
  ~SumServer();
  void Stop();
  // statistics
  uint64_t pushes() const { return n_push_; }
  uint64_t pulls() const { return n_pull_; }
  size_t num_keys();

 private:
  struct KeyState {
    std::mutex mu;
    // two buffers: pulls of round r read store[rd] while the pushes of round
    // r+1 are merged into store[wr] (the reference merges in place and relies on
    // timing to keep a late puller from seeing the next round's first copy)
    char* store2[2] = {nullptr, nullptr};
    std::string shm_names[2];                // set when the store lives in shared memory (BYTEPS_ENABLE_IPC=1)
    int rd = 0, wr = 1;
    size_t store_cap = 0;
    size_t len = 0;
    int dtype = F32;
    bool inited = false;
    int pushers = 0;                         // expected pushes per round for THIS key
    std::vector<net::KVMeta> init_reqs;
    std::vector<net::KVMeta> round_reqs;     // pushes seen in the current round
    // pull bookkeeping
    bool push_finished = false;
    std::set<int> seen_sender;
    size_t pull_cnt = 0;
    std::vector<net::KVMeta> parked_pulls;
    // compression
    std::unique_ptr<Compressor> compressor;
    std::vector<net::KVMeta> comp_reqs;
    std::vector<char> comp_out;     // re-compressed merged result
    std::vector<char> decomp;       // scratch for incoming pushes
    const char* merged = nullptr;
    size_t merged_len = 0;
    int tid = -1;
    int numa_node = -1;                      // where the pushers' GPUs hang (majority of the init pushes' hints)
    int64_t t_first_push = 0, t_last_push = 0;   // BYTEPS_SERVER_PROFILE: arrival of the round's first / last push
    // engine thread only: the round's first push, kept (zero-copy) until the second one arrives so that both are
    // merged in ONE pass (store = a + b) instead of copy + read-modify-write
    net::SArray<char> held_first;
    bool holding = false;
  };
  void Handle(const net::KVMeta& req, const net::KVPairs& data, net::KVServer* srv);
  void EngineLoop(int tid);
  KeyState* GetState(uint64_t key);
  int ThreadOf(KeyState* st, size_t len);
  void SendPush(const net::KVMeta& req);
  void SendPull(KeyState* st, uint64_t key, const net::KVMeta& req);
  void Publish(KeyState* st, uint64_t key);   // ALL_RECV; st->mu held
  void Debug(const char* stage, uint64_t key, const void* dst, const void* src, size_t len, int dtype);

  net::Postoffice* po_;
  ServerConfig cfg_;
  int pushers_;
  size_t inline_bytes_ = 16384;   // BYTEPS_SERVER_INLINE_BYTES: keys up to this size are merged on the handler thread
  std::unique_ptr<net::KVServer> kv_;
  CpuReducer reducer_;
  std::mutex map_mu_;
  std::unordered_map<uint64_t, std::unique_ptr<KeyState>> states_;
  std::vector<std::unique_ptr<PriorityQueue>> queues_;
  std::vector<std::thread> threads_;
  std::vector<uint64_t> acc_load_;
  std::mutex load_mu_;
  std::atomic<uint64_t> msg_id_{0};
  std::atomic<uint64_t> n_push_{0}, n_pull_{0};
  // BYTEPS_SERVER_PROFILE=1
  std::atomic<uint64_t> n_push_ref_{0}, push_ref_bytes_{0}, push_payload_bytes_{0}, merge_ns_{0}, merge_bytes_{0};
  // per round of a key above the inline size: first push -> last push (pusher skew), last push -> pulls answered
  std::atomic<uint64_t> rounds_{0}, skew_us_{0}, serve_us_{0}, skew_max_us_{0}, serve_max_us_{0}, parked_at_publish_{0};
  bool profile_ = false;
  std::vector<int> thread_node_;             // NUMA node every engine thread is pinned to (-1: not pinned)
  std::atomic<uint64_t> numa_bound_{0}, numa_refused_{0};
  bool stopped_ = false;
};

}  // namespace server
}  // namespace bps
