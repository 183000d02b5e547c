// Message transport + cluster membership.
//
// Parity map (ps-lite, /root/reference/3rdparty/ps-lite):
//   Van base (src/van.cc)              -> Van: scheduler bootstrap / ADD_NODE id
//                                         assignment, control vs data dispatch,
//                                         barriers, heartbeats, recovery, resend,
//                                         PS_DROP_MSG fault injection, profiling log
//   ZMQVan (src/zmq_van.h)             -> TcpVan: framed messages over plain TCP
//                                         sockets (no ZeroMQ dependency)
//   RDMA IPCTransport (rdma_transport.h:513-712) -> colocated peers exchange the
//                                         payload through POSIX shm, only the meta
//                                         crosses the socket (ShmRegistry)
//   MultiVan (src/multi_van.h)         -> TcpVan with DMLC_NUM_PORTS > 1 stripes
//                                         peers over several listening ports
//   Postoffice (src/postoffice.cc)     -> Postoffice: id scheme (scheduler 1,
//                                         server 8+2r, worker 9+2r), group ids,
//                                         customers, barriers, key ranges, dead nodes
// RDMA / UCX / libfabric vans need NIC hardware + libraries that do not exist
// on a single NVSwitch box; `Van` is the interface a new transport plugs into.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "core/spsc_queue.h"
#include "net/message.h"

namespace bps {
namespace net {

class Postoffice;
class Resender;

struct NetConfig {
  Role role = Role::kWorker;
  int num_workers = 1;
  int num_servers = 1;
  std::string scheduler_host = "127.0.0.1";
  int scheduler_port = 9000;
  std::string node_host = "127.0.0.1";
  int node_port = 0;           // 0 = pick a free port
  int rank_hint = -1;          // stable rank (DMLC_WORKER_ID / BYTEPS_GLOBAL_RANK); -1 = by address order
  int verbose = 0;             // PS_VERBOSE
  int heartbeat_interval_s = 0;  // PS_HEARTBEAT_INTERVAL (0 = off)
  int heartbeat_timeout_s = 0;   // PS_HEARTBEAT_TIMEOUT
  bool resend = false;         // PS_RESEND
  int resend_timeout_ms = 1000;  // PS_RESEND_TIMEOUT
  int drop_msg_pct = 0;        // PS_DROP_MSG
  bool enable_ipc = false;     // BYTEPS_ENABLE_IPC
  std::string profile_path;    // ENABLE_PROFILING + PROFILE_PATH
  bool is_recovery = false;
  int num_lanes = 2;           // DMLC_NUM_PORTS: parallel TCP connections per peer, data striped by key
  std::string van_type = "tcp";  // DMLC_PS_VAN_TYPE: tcp (alias zmq) | shm (one host, no sockets: net/shm_van.h)
  bool local = false;          // DMLC_LOCAL: every node is on this host -> Unix-domain stream sockets
  static NetConfig from_env();
  void resolve_node_host();   // fills node_host when empty (DMLC_INTERFACE, first non-loopback IPv4, 127.0.0.1)
};

// Recycles receive buffers by size: a fresh 4 MB allocation per message costs ~1000 first-touch
// page faults; partitions have a handful of distinct sizes, so a small free list removes them.
class PayloadPool {
 public:
  static PayloadPool& get();
  SArray<char> alloc(size_t n);
  size_t cached_bytes() const { return cached_.load(); }

 private:
  void give_back(char* p, size_t n);
  std::mutex mu_;
  std::unordered_map<size_t, std::vector<char*>> free_;
  std::atomic<size_t> cached_{0};
  size_t cap_ = 512u << 20;
};

// Maps host memory registered as POSIX shm so colocated peers can skip the socket.
class ShmRegistry {
 public:
  struct Region {
    std::string name;
    char* base;
    size_t len;
  };
  static ShmRegistry& get();
  // create-or-open a shm object and map it; the creator owns unlinking
  void* create(const std::string& name, size_t len);
  void* open(const std::string& name, size_t len);
  bool lookup(const void* ptr, size_t len, std::string* name, uint64_t* offset);
  void release(const std::string& name);
  size_t region_len(const std::string& name);   // mapped length (0: not mapped here)
  // unlink BytePS_* objects under /dev/shm whose creating process (the pid in the name) no longer exists - what a
  // killed worker / server leaves behind (staging windows, server stores).  Returns how many were removed.
  static int reap_stale(const std::string& dir = "/dev/shm");

 private:
  std::mutex mu_;
  std::map<std::string, Region> regions_;
  std::map<const char*, std::string> by_base_;     // base address -> name: lookup() is O(log n), not a scan
  std::unordered_set<std::string> owned_;
};

class Van {
 public:
  explicit Van(Postoffice* po);
  virtual ~Van();
  void Start(int customer_id);
  void Stop();
  // returns bytes sent, -1 on failure
  int Send(Message& msg);
  // Zero-copy receive of ONE pull response: the payload of the response to request `timestamp`
  // of (app, customer) is read from the socket straight into dst (ps-lite's RegisterRecvBuffer,
  // kv_app.h:455-566).  One-shot, so a retransmitted duplicate can never scribble over a buffer
  // the application has already handed to the next iteration.
  void ExpectPullResponse(int app_id, int customer_id, int timestamp, char* dst, size_t len);
  bool TakeRecvBuffer(const Meta& meta, char** dst, size_t* len);
  void CancelRecvBuffer(int app_id, int customer_id, int timestamp);
  // retransmission of a message that already carries its signature (resender only)
  int Resend(Message& msg);
  const Node& my_node() const { return my_node_; }
  // true when node `id` shares this host's /dev/shm (same hostname, DMLC_LOCAL, or the shm van): only then may
  // shared-memory names travel instead of payloads (ps-lite's rdma van decides `is_local` the same way)
  virtual bool IsColocated(int id) { (void)id; return false; }
  bool IsReady() const { return ready_.load(); }
  int GetTimestamp() { return timestamp_++; }
  void set_err_handle(std::function<void(int)> h) { err_handle_ = std::move(h); }
  uint64_t direct_recvs() const { return direct_recvs_.load(); }
  uint64_t send_bytes() const { return send_bytes_; }
  uint64_t recv_bytes() const { return recv_bytes_; }

 protected:
  // transport interface ------------------------------------------------------
  virtual int Bind(Node& node, int max_retry) = 0;
  virtual void Connect(const Node& node) = 0;
  virtual int SendMsg(Message& msg) = 0;
  virtual int RecvMsg(Message* msg) = 0;   // blocking; <0 when stopped
  virtual void StopTransport() = 0;
  // Transport threads may hand a DATA message straight to its customer instead of queueing it for the van's
  // receiving thread (one condition-variable hop less per message).  Only when nothing else has to see the message
  // first: no resender, no fault injection, no profiling, cluster ready.  Returns false when the caller must queue it.
  bool TryDirectData(Message* msg, int nbytes);

  Postoffice* po_;
  Node scheduler_;
  Node my_node_;
  bool is_scheduler_ = false;
  std::function<void(int)> err_handle_;

 private:
  void Receiving();
  void Heartbeat();
  void ProcessAddNode(Message* msg, Meta* nodes, Meta* recovery_nodes);
  void ProcessBarrier(Message* msg, bool instance);
  void ProcessHeartbeat(Message* msg);
  void ProcessData(Message* msg);
  void ProcessTerminate();
  void UpdateLocalID(Message* msg, std::unordered_set<int>* deadnodes, Meta* nodes, Meta* recovery);
  void ProfileEvent(const Message& msg, bool send);

  std::atomic<bool> ready_{false};
  std::atomic<int> registered_{0};   // scheduler: nodes that have sent ADD_NODE (start-up diagnostics)
  bool direct_dispatch_ = true;   // BYTEPS_VAN_DIRECT_DISPATCH
  std::atomic<bool> direct_ok_{false};
  std::atomic<int> timestamp_{0};
  std::atomic<uint64_t> send_bytes_{0}, recv_bytes_{0};
  std::atomic<uint64_t> direct_recvs_{0};
  struct RecvSlot {
    char* dst;
    size_t len;
  };
  std::mutex recv_slots_mu_;
  std::unordered_map<uint64_t, RecvSlot> recv_slots_;   // (app, customer, timestamp) -> destination
  std::thread receiver_, heartbeat_;
  std::atomic<bool> stopping_{false};
  int num_servers_ = 0, num_workers_ = 0;
  std::vector<int> barrier_count_;
  std::unordered_map<std::string, int> connected_nodes_;  // "host:port" -> id
  std::unordered_map<int, int> shared_node_mapping_;
  Resender* resender_ = nullptr;
  std::mutex start_mu_;
  int init_stage_ = 0;
  FILE* profile_ = nullptr;
  std::mutex profile_mu_;
  Meta add_node_nodes_;
  Meta add_node_recovery_;
  friend class Resender;
};

class TcpVan : public Van {
 public:
  explicit TcpVan(Postoffice* po) : Van(po) {}
  ~TcpVan() override;

 protected:
  int Bind(Node& node, int max_retry) override;
  void Connect(const Node& node) override;
  int SendMsg(Message& msg) override;
  int RecvMsg(Message* msg) override;
  void StopTransport() override;
  bool IsColocated(int id) override;

 private:
  // One peer = `lanes` TCP connections (the reference's MultiVan opens DMLC_NUM_PORTS vans per
  // node, multi_van.h:59-285; here it is one van with several lanes).  Lane 0 carries control
  // traffic; data messages go to lane key % lanes, so per-key order is kept while different keys
  // use different sockets (kernel send queues and reader threads run in parallel).
  struct Lane {
    int fd = -1;
    std::mutex mu;
  };
  struct Sender {
    std::vector<std::unique_ptr<Lane>> lanes;
    std::string addr;
    bool colocated = false;
  };
  void AcceptLoop();
  void ReadLoop(int fd);
  bool ipc_send_strip(Message& msg, Sender* s);
  void ipc_recv_attach(Message* msg);

  int listen_fd_ = -1;
  bool local_ = false;   // DMLC_LOCAL=1: Unix-domain stream sockets instead of TCP
  int BindLocal(Node& node, int max_retry);
  std::thread acceptor_;
  std::vector<std::thread> readers_;
  std::vector<int> reader_fds_;
  std::mutex readers_mu_;
  std::mutex senders_mu_;
  std::unordered_map<int, std::shared_ptr<Sender>> senders_;
  std::mutex q_mu_;
  std::condition_variable q_cv_;
  std::queue<Message> recv_q_;
  std::atomic<bool> closed_{false};
};

// --------------------------------------------------------------------------------
class Customer;

class Postoffice {
 public:
  explicit Postoffice(const NetConfig& cfg);
  ~Postoffice();
  // connect to the cluster; blocks until the scheduler has assigned ids (and,
  // if do_barrier, until every node has started)
  void Start(int customer_id, bool do_barrier = true);
  void Finalize(int customer_id, bool do_barrier = true);
  void AddCustomer(Customer* c);
  void RemoveCustomer(Customer* c);
  Customer* GetCustomer(int app_id, int customer_id, int timeout_s = 0);
  // find the customer and Accept() the message atomically w.r.t. RemoveCustomer
  bool Deliver(int app_id, int customer_id, const Message& msg, int timeout_s = 0);
  void Barrier(int customer_id, int node_group);
  void ManageBarrier(int customer_id);   // called by the van on barrier release
  Van* van() { return van_.get(); }
  const NetConfig& cfg() const { return cfg_; }

  // id arithmetic
  static int WorkerRankToID(int rank) { return rank * 2 + 9; }
  static int ServerRankToID(int rank) { return rank * 2 + 8; }
  static int IDtoRank(int id) { return std::max((id - 8) / 2, 0); }
  int num_workers() const { return cfg_.num_workers; }
  int num_servers() const { return cfg_.num_servers; }
  int my_rank() const { return IDtoRank(van_->my_node().id); }
  bool is_worker() const { return cfg_.role == Role::kWorker; }
  bool is_server() const { return cfg_.role == Role::kServer; }
  bool is_scheduler() const { return cfg_.role == Role::kScheduler; }
  const std::vector<int>& GetNodeIDs(int node_group) const;
  // uniform key ranges over uint64 (one per server)
  const std::vector<std::pair<uint64_t, uint64_t>>& GetServerKeyRanges();
  // heartbeats
  void UpdateHeartbeat(int node_id, time_t t);
  std::vector<int> GetDeadNodes(int timeout_s);
  bool is_recovery() const { return cfg_.is_recovery; }
  int verbose() const { return cfg_.verbose; }

 private:
  void InitNodeIDs();
  NetConfig cfg_;
  std::unique_ptr<Van> van_;
  std::mutex mu_;
  std::unordered_map<int, std::unordered_map<int, Customer*>> customers_;
  std::unordered_map<int, std::vector<int>> node_ids_;
  std::mutex barrier_mu_;
  std::condition_variable barrier_cv_;
  std::unordered_map<int, bool> barrier_done_;
  std::vector<std::pair<uint64_t, uint64_t>> key_ranges_;
  std::mutex hb_mu_;
  std::unordered_map<int, time_t> heartbeats_;
  time_t start_time_ = 0;
  std::mutex start_mu_;
  int init_stage_ = 0;
};

// --------------------------------------------------------------------------------
// Per-app receive thread + request tracker (ps-lite Customer, src/customer.cc:20-82)
class Customer {
 public:
  using RecvHandle = std::function<void(const Message&)>;
  Customer(int app_id, int customer_id, RecvHandle h, Postoffice* po);
  ~Customer();
  int app_id() const { return app_id_; }
  int customer_id() const { return customer_id_; }
  int NewRequest(int recver);            // returns a timestamp expecting 1 response per node in `recver`
  void WaitRequest(int timestamp);
  int NumResponse(int timestamp);
  void AddResponse(int timestamp, int num = 1);
  void Accept(const Message& m);         // called by the van

 private:
  void Receiving();
  int app_id_, customer_id_;
  RecvHandle handle_;
  Postoffice* po_;
  std::mutex q_mu_;
  std::condition_variable q_cv_;
  std::queue<Message> q_;
  bool stop_ = false;
  // DMLC_LOCKLESS_QUEUE=1: busy-polled ring instead of mutex + condvar (core/spsc_queue.h)
  std::unique_ptr<SpscQueue<Message>> ring_;
  std::atomic<bool> ring_stop_{false};
  std::mutex tracker_mu_;
  std::condition_variable tracker_cv_;
  std::vector<std::pair<int, int>> tracker_;
  std::thread thread_;
};

}  // namespace net
}  // namespace bps
