#include "net/van.h"

#include <net/if.h>

#include <arpa/inet.h>
#include <dirent.h>
#include <fcntl.h>
#include <ifaddrs.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstring>
#include <chrono>
#include <cstddef>
#include <cstdlib>
#include <ctime>
#include <random>
#include <thread>

#include "core/env.h"
#include "core/log.h"
#include "net/resender.h"
#include "net/shm_van.h"

namespace bps {
namespace net {

#define VLOG(po, lvl) \
  if ((po)->verbose() >= (lvl)) BPS_LOG(INFO)

// ================================================================ config
static std::string interface_ipv4(const std::string& nic) {
  std::string out;
  ifaddrs* ifs = nullptr;
  if (getifaddrs(&ifs) != 0) return out;
  for (ifaddrs* i = ifs; i; i = i->ifa_next) {
    if (!i->ifa_addr || i->ifa_addr->sa_family != AF_INET || nic != i->ifa_name) continue;
    char buf[INET_ADDRSTRLEN];
    if (inet_ntop(AF_INET, &reinterpret_cast<sockaddr_in*>(i->ifa_addr)->sin_addr, buf, sizeof(buf))) out = buf;
    break;
  }
  freeifaddrs(ifs);
  if (out.empty()) BPS_LOG(WARNING) << "DMLC_INTERFACE=" << nic << " has no IPv4 address";
  return out;
}

// ps-lite's GetIP(): the first IPv4 address of an interface that is up and not loopback.
static std::string first_external_ipv4() {
  std::string out;
  ifaddrs* ifs = nullptr;
  if (getifaddrs(&ifs) != 0) return out;
  for (ifaddrs* i = ifs; i; i = i->ifa_next) {
    if (!i->ifa_addr || i->ifa_addr->sa_family != AF_INET) continue;
    if ((i->ifa_flags & IFF_LOOPBACK) || !(i->ifa_flags & IFF_UP)) continue;
    char buf[INET_ADDRSTRLEN];
    if (inet_ntop(AF_INET, &reinterpret_cast<sockaddr_in*>(i->ifa_addr)->sin_addr, buf, sizeof(buf))) {
      out = buf;
      break;
    }
  }
  freeifaddrs(ifs);
  return out;
}

// What this node advertises to the scheduler: DMLC_NODE_HOST, else DMLC_INTERFACE's address, else - like
// ps-lite (van.cc:520-562) - the first non-loopback IPv4; a job whose scheduler is on the loopback
// interface is a one-host job and keeps 127.0.0.1.
void NetConfig::resolve_node_host() {
  if (!node_host.empty()) return;
  std::string nic = env_str("DMLC_INTERFACE", "");
  if (!nic.empty()) node_host = interface_ipv4(nic);
  if (!node_host.empty()) return;
  const bool local_job = scheduler_host == "localhost" || scheduler_host.rfind("127.", 0) == 0;
  if (!local_job) node_host = first_external_ipv4();
  if (node_host.empty()) node_host = "127.0.0.1";
}

NetConfig NetConfig::from_env() {
  NetConfig c;
  std::string role = env_str("DMLC_ROLE", "worker");
  c.role = role == "server" ? Role::kServer : role == "scheduler" ? Role::kScheduler : Role::kWorker;
  c.num_workers = (int)env_int("DMLC_NUM_WORKER", 1);
  c.num_servers = (int)env_int("DMLC_NUM_SERVER", 1);
  c.scheduler_host = env_str("DMLC_PS_ROOT_URI", "127.0.0.1");
  c.scheduler_port = (int)env_int("DMLC_PS_ROOT_PORT", 9000);
  c.node_host = env_str("DMLC_NODE_HOST", "");
  c.resolve_node_host();
  c.node_port = (int)env_int("DMLC_PORT", env_int("PORT", 0));
  if (env_bool("DMLC_ENABLE_RDMA", false) || env_bool("DMLC_ENABLE_UCX", false))
    BPS_LOG(WARNING) << "DMLC_ENABLE_RDMA / DMLC_ENABLE_UCX: no verbs/UCX transport in this build; using the TCP van"
                     << " (DMLC_NUM_PORTS lanes, BYTEPS_ENABLE_IPC for colocated peers)";
  if (c.role == Role::kWorker && env_has("DMLC_WORKER_ID")) c.rank_hint = (int)env_int("DMLC_WORKER_ID", -1);
  if (env_has("DMLC_RANK")) c.rank_hint = (int)env_int("DMLC_RANK", -1);
  c.verbose = (int)env_int("PS_VERBOSE", 0);
  c.heartbeat_interval_s = (int)env_int("PS_HEARTBEAT_INTERVAL", 0);
  c.heartbeat_timeout_s = (int)env_int("PS_HEARTBEAT_TIMEOUT", 0);
  c.resend = env_bool("PS_RESEND", false);
  c.resend_timeout_ms = (int)env_int("PS_RESEND_TIMEOUT", 1000);
  c.drop_msg_pct = (int)env_int("PS_DROP_MSG", 0);
  c.enable_ipc = env_bool("BYTEPS_ENABLE_IPC", false);
  // connections per peer: 2 on small hosts (push and pull streams overlap; more only adds threads), up to 8 where
  // there are cores to run the extra reader threads (one lane saturates at 3-4 GB/s of loopback / NIC copy work)
  const long long cores = (long long)std::thread::hardware_concurrency();
  const long long dflt_lanes = std::min<long long>(8, std::max<long long>(2, cores / 16));
  c.num_lanes = (int)std::min<long long>(16, std::max<long long>(1, env_int("DMLC_NUM_PORTS", dflt_lanes)));
  c.local = env_bool("DMLC_LOCAL", false);
  c.van_type = env_str("DMLC_PS_VAN_TYPE", "tcp");
  if (env_bool("ENABLE_PROFILING", false)) c.profile_path = env_str("PROFILE_PATH", "./van_profile.log");
  return c;
}

// ================================================================ shm registry
PayloadPool& PayloadPool::get() {
  static PayloadPool* p = new PayloadPool();   // leaked on purpose: buffers may outlive static destruction
  return *p;
}

SArray<char> PayloadPool::alloc(size_t n) {
  if (n < (64u << 10)) return SArray<char>(n);             // small messages: plain new/delete
  char* buf = nullptr;
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.find(n);
    if (it != free_.end() && !it->second.empty()) {
      buf = it->second.back();
      it->second.pop_back();
      cached_ -= n;
    }
  }
  if (!buf) buf = new char[n];
  return SArray<char>(buf, n, [this, n](char* p) { give_back(p, n); });
}

void PayloadPool::give_back(char* p, size_t n) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (cached_ + n <= cap_) {
      free_[n].push_back(p);
      cached_ += n;
      return;
    }
  }
  delete[] p;
}

ShmRegistry& ShmRegistry::get() {
  static ShmRegistry r;
  return r;
}

static void* map_shm(const std::string& name, size_t len, bool create) {
  int fd = shm_open(("/" + name).c_str(), create ? (O_CREAT | O_RDWR) : O_RDWR, 0666);
  if (fd < 0) return nullptr;
  if (create && ftruncate(fd, (off_t)len) != 0) {
    close(fd);
    return nullptr;
  }
  if (!create) {
    struct stat st;
    if (fstat(fd, &st) == 0 && (size_t)st.st_size < len) {
      close(fd);
      return nullptr;
    }
  }
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return p == MAP_FAILED ? nullptr : p;
}

void* ShmRegistry::create(const std::string& name, size_t len) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(name);
  if (it != regions_.end() && it->second.len >= len) return it->second.base;
  void* p = map_shm(name, len, true);
  BPS_CHECK(p != nullptr) << "shm create failed for " << name << " (" << len << " bytes)";
  // grown: the smaller mapping stays mapped (zero-copy views of in-flight messages may still point into it), it is
  // only forgotten
  if (it != regions_.end()) by_base_.erase(it->second.base);
  regions_[name] = Region{name, (char*)p, len};
  by_base_[(const char*)p] = name;
  owned_.insert(name);
  return p;
}

void* ShmRegistry::open(const std::string& name, size_t len) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(name);
  if (it != regions_.end() && it->second.len >= len) return it->second.base;
  void* p = map_shm(name, len, false);
  if (!p) return nullptr;
  if (it != regions_.end()) by_base_.erase(it->second.base);   // see create(): not unmapped on purpose
  regions_[name] = Region{name, (char*)p, len};
  by_base_[(const char*)p] = name;
  return p;
}

bool ShmRegistry::lookup(const void* ptr, size_t len, std::string* name, uint64_t* offset) {
  std::lock_guard<std::mutex> g(mu_);
  const char* p = (const char*)ptr;
  auto it = by_base_.upper_bound(p);        // first region that starts after p; the candidate is the one before
  if (it == by_base_.begin()) return false;
  --it;
  const Region& r = regions_[it->second];
  if (p >= r.base && p + len <= r.base + r.len) {
    *name = r.name;
    *offset = (uint64_t)(p - r.base);
    return true;
  }
  return false;
}

size_t ShmRegistry::region_len(const std::string& name) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(name);
  return it == regions_.end() ? 0 : it->second.len;
}

void ShmRegistry::release(const std::string& name) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(name);
  if (it == regions_.end()) return;
  by_base_.erase(it->second.base);
  munmap(it->second.base, it->second.len);
  regions_.erase(it);
  if (owned_.erase(name)) shm_unlink(("/" + name).c_str());
}

int ShmRegistry::reap_stale(const std::string& dir) {
  // names carry the creator's pid right after the prefix: BytePS_ShM_<pid>_..., BytePS_SrvStore_<pid>_...
  static const char* prefixes[] = {"BytePS_ShM_", "BytePS_SrvStore_"};
  int removed = 0;
  DIR* d = opendir(dir.c_str());
  if (!d) return 0;
  while (dirent* e = readdir(d)) {
    for (const char* pre : prefixes) {
      const size_t n = strlen(pre);
      if (strncmp(e->d_name, pre, n) != 0 || !isdigit((unsigned char)e->d_name[n])) continue;
      const long pid = strtol(e->d_name + n, nullptr, 10);
      if (pid <= 1 || pid == (long)getpid()) continue;
      if (kill((pid_t)pid, 0) == 0 || errno != ESRCH) continue;     // alive (or not ours to judge)
      if (unlink((dir + "/" + e->d_name).c_str()) == 0) ++removed;
    }
  }
  closedir(d);
  return removed;
}

// ================================================================ resender
bool Resender::AddIncoming(const Message& msg) {
  if (msg.meta.control.cmd == Control::TERMINATE) return false;
  if (msg.meta.control.cmd == Control::ACK) {
    std::lock_guard<std::mutex> g(mu_);
    send_buff_.erase(msg.meta.control.msg_sig);
    return true;
  }
  // a joining node has no id yet (its ADD_NODE carries sender = kEmpty): it cannot be ACKed
  if (msg.meta.sender == kEmpty || msg.meta.msg_sig == 0) return false;
  uint64_t sig = msg.meta.msg_sig;
  bool dup;
  {
    std::lock_guard<std::mutex> g(mu_);
    dup = !acked_.insert(sig).second;
  }
  // always (re-)acknowledge: the previous ACK may have been lost
  Message ack;
  ack.meta.recver = msg.meta.sender;
  ack.meta.sender = msg.meta.recver;
  ack.meta.control.cmd = Control::ACK;
  ack.meta.control.msg_sig = sig;
  van_->Send(ack);
  if (dup) BPS_LOG(DEBUG) << "resender: duplicated message dropped";
  return dup;
}

void Resender::Monitoring() {
  while (!exit_) {
    std::this_thread::sleep_for(std::chrono::milliseconds(std::max(1, timeout_ms_ / 4)));
    std::vector<Message> resend;
    std::vector<uint64_t> give_up;
    int64_t now = Now();
    {
      std::lock_guard<std::mutex> g(mu_);
      for (auto& kv : send_buff_) {
        Entry& e = kv.second;
        if (e.send + (int64_t)timeout_ms_ * (1 + e.num_retry) < now) {
          if (e.num_retry >= max_retry_) {
            // the peer is gone (e.g. it already finalized): report, do not take the process down
            BPS_LOG(ERROR) << "message to node " << e.msg.meta.recver << " was not ACKed after " << max_retry_
                           << " retries; giving up";
            give_up.push_back(kv.first);
            continue;
          }
          resend.push_back(e.msg);
          ++e.num_retry;
        }
      }
      for (uint64_t sig : give_up) send_buff_.erase(sig);
    }
    for (auto& m : resend) {
      if (exit_) break;
      van_->Resend(m);
    }
  }
}

// ================================================================ Van
Van::Van(Postoffice* po) : po_(po) {}
Van::~Van() {}

void Van::Start(int customer_id) {
  std::unique_lock<std::mutex> lk(start_mu_);
  const NetConfig& c = po_->cfg();
  if (init_stage_ == 0) {
    scheduler_.hostname = c.scheduler_host;
    scheduler_.port = c.scheduler_port;
    scheduler_.role = Role::kScheduler;
    scheduler_.id = kScheduler;
    is_scheduler_ = c.role == Role::kScheduler;
    if (is_scheduler_) {
      my_node_ = scheduler_;
    } else {
      my_node_.hostname = c.node_host;
      my_node_.role = c.role;
      my_node_.port = c.node_port;
      my_node_.id = kEmpty;
      my_node_.customer_id = customer_id;
      my_node_.aux_id = c.rank_hint;
      my_node_.is_recovery = c.is_recovery;
    }
    barrier_count_.assign(8, 0);
    direct_dispatch_ = env_bool("BYTEPS_VAN_DIRECT_DISPATCH", true);
    my_node_.port = Bind(my_node_, is_scheduler_ ? 0 : 40);
    BPS_CHECK_NE(my_node_.port, -1) << "bind failed for " << my_node_.debug();
    VLOG(po_, 1) << "bind to " << my_node_.debug();
    Connect(scheduler_);
    if (!c.profile_path.empty()) profile_ = fopen(c.profile_path.c_str(), "a");
    if (c.resend) resender_ = new Resender(c.resend_timeout_ms, 10, this);
    direct_ok_.store(direct_dispatch_ && !resender_ && !profile_ && c.drop_msg_pct == 0 && c.verbose < 3,
                     std::memory_order_release);
    receiver_ = std::thread([this] { Receiving(); });
    init_stage_ = 1;
  }
  lk.unlock();
  if (!is_scheduler_) {
    // tell the scheduler about me
    Message msg;
    Node cnode = my_node_;
    cnode.customer_id = customer_id;
    msg.meta.recver = kScheduler;
    msg.meta.control.cmd = Control::ADD_NODE;
    msg.meta.control.node.push_back(cnode);
    msg.meta.timestamp = timestamp_++;
    Send(msg);
  }
  {
    // Every role blocks here until the scheduler has seen all DMLC_NUM_WORKER x BYTEPS_LOCAL_SIZE worker processes
    // and DMLC_NUM_SERVER servers.  A scheduler / server started without BYTEPS_LOCAL_SIZE (a worker-side
    // variable in the reference) counts differently from multi-GPU worker boxes and would wait forever in
    // silence: say what is missing, periodically.
    const auto t0 = std::chrono::steady_clock::now();
    const long long warn_s = std::max<long long>(1, env_int("BYTEPS_START_WARN_S", 30));
    long long warned = 0;
    while (!ready_.load()) {
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
      const long long waited =
          std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
      if (waited / warn_s > warned) {
        warned = waited / warn_s;
        if (is_scheduler_)
          BPS_LOG(WARNING) << "scheduler: " << registered_.load() << " of " << po_->num_workers() + po_->num_servers()
                           << " nodes registered after " << waited << " s (expecting " << po_->num_workers()
                           << " worker processes = DMLC_NUM_WORKER x BYTEPS_LOCAL_SIZE and " << po_->num_servers()
                           << " servers); the scheduler and the servers need the same BYTEPS_LOCAL_SIZE as the workers";
        else
          BPS_LOG(WARNING) << my_node_.debug() << ": not admitted by the scheduler at " << scheduler_.hostname << ":"
                           << scheduler_.port << " after " << waited << " s (it waits for " << po_->num_workers()
                           << " worker processes and " << po_->num_servers() << " servers)";
      }
    }
  }
  lk.lock();
  if (init_stage_ == 1) {
    if (!is_scheduler_ && c.heartbeat_interval_s > 0) heartbeat_ = std::thread([this] { Heartbeat(); });
    init_stage_ = 2;
  }
}

void Van::Stop() {
  if (po_->verbose() >= 3) BPS_LOG(INFO) << my_node_.debug() << " Stop() enter";
  VLOG(po_, 1) << my_node_.debug() << " stopping: draining the resender";
  if (resender_ && !resender_->Drain(std::max<int64_t>(5000, 4 * (int64_t)resender_->timeout_ms())))
    BPS_LOG(WARNING) << my_node_.debug() << ": stopping with unacknowledged messages";
  VLOG(po_, 1) << my_node_.debug() << " stopping: terminate self";
  stopping_ = true;
  direct_ok_.store(false, std::memory_order_release);
  // unblock the receiving thread with a TERMINATE addressed to myself
  Message exit;
  exit.meta.control.cmd = Control::TERMINATE;
  exit.meta.recver = my_node_.id;
  exit.meta.customer_id = 0;
  int sent = SendMsg(exit);
  if (po_->verbose() >= 3) BPS_LOG(INFO) << my_node_.debug() << " terminate-self send rc=" << sent;
  if (receiver_.joinable()) receiver_.join();
  if (heartbeat_.joinable()) heartbeat_.join();
  delete resender_;
  resender_ = nullptr;
  StopTransport();
  if (profile_) {
    fclose(profile_);
    profile_ = nullptr;
  }
  ready_ = false;
  init_stage_ = 0;
  connected_nodes_.clear();
  shared_node_mapping_.clear();
  num_servers_ = num_workers_ = 0;
  add_node_nodes_ = Meta();
  add_node_recovery_ = Meta();
  stopping_ = false;
}

int Van::Send(Message& msg) {
  if (msg.meta.sender == kEmpty) msg.meta.sender = my_node_.id;
  const bool reliable = resender_ && msg.meta.sender != kEmpty;
  // every transmission gets its OWN signature: callers reuse Message objects (the scheduler's
  // barrier release goes to every node from one object, responses copy request metas), and a
  // signature shared by two messages makes the second one untracked here and a "duplicate" there
  if (reliable && msg.meta.control.cmd != Control::ACK) {
    static std::atomic<uint64_t> serial{1};
    msg.meta.msg_sig = Resender::Signature(msg.meta) ^ (serial.fetch_add(1) * 0x9E3779B97F4A7C15ull) ^
                       ((uint64_t)std::random_device{}() << 17);
    if (msg.meta.msg_sig == 0) msg.meta.msg_sig = 1;
  }
  int n = SendMsg(msg);
  if (n < 0) {
    if (err_handle_) err_handle_(msg.meta.recver);
    if (!stopping_) BPS_LOG(WARNING) << my_node_.debug() << ": send to node " << msg.meta.recver << " failed";
    return -1;
  }
  send_bytes_ += n;
  if (reliable) resender_->AddOutgoing(msg);
  if (profile_) ProfileEvent(msg, true);
  if (po_->verbose() >= 2) BPS_LOG(INFO) << my_node_.debug() << " sent " << n << "B to " << msg.meta.recver;
  return n;
}

static uint64_t recv_slot_key(int app_id, int customer_id, int timestamp) {
  return ((uint64_t)(uint32_t)app_id << 48) ^ ((uint64_t)(uint32_t)customer_id << 32) ^ (uint64_t)(uint32_t)timestamp;
}

void Van::ExpectPullResponse(int app_id, int customer_id, int timestamp, char* dst, size_t len) {
  std::lock_guard<std::mutex> g(recv_slots_mu_);
  recv_slots_[recv_slot_key(app_id, customer_id, timestamp)] = RecvSlot{dst, len};
}

void Van::CancelRecvBuffer(int app_id, int customer_id, int timestamp) {
  std::lock_guard<std::mutex> g(recv_slots_mu_);
  recv_slots_.erase(recv_slot_key(app_id, customer_id, timestamp));
}

bool Van::TakeRecvBuffer(const Meta& meta, char** dst, size_t* len) {
  if (meta.request || !meta.pull || !meta.control.empty() || meta.simple_app) return false;
  std::lock_guard<std::mutex> g(recv_slots_mu_);
  auto it = recv_slots_.find(recv_slot_key(meta.app_id, meta.customer_id, meta.timestamp));
  if (it == recv_slots_.end()) return false;
  *dst = it->second.dst;
  *len = it->second.len;
  recv_slots_.erase(it);     // one-shot
  ++direct_recvs_;
  return true;
}

int Van::Resend(Message& msg) {
  int n = SendMsg(msg);
  if (n < 0) {
    if (!stopping_) BPS_LOG(DEBUG) << my_node_.debug() << ": resend to node " << msg.meta.recver << " failed";
    return -1;
  }
  send_bytes_ += n;
  if (po_->verbose() >= 2) BPS_LOG(INFO) << my_node_.debug() << " re-sent " << n << "B to " << msg.meta.recver;
  return n;
}

void Van::ProfileEvent(const Message& msg, bool send) {
  if (msg.meta.control.cmd != Control::EMPTY || (!msg.meta.push && !msg.meta.pull)) return;
  auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch())
                .count();
  std::lock_guard<std::mutex> g(profile_mu_);
  if (!profile_) return;
  fprintf(profile_, "%llu\t%s_van_%s_%s\t%lld\n", (unsigned long long)msg.meta.key,
          po_->is_server() ? "server" : "worker", send ? "send" : "recv", msg.meta.push ? "push" : "pull",
          (long long)us);
}

void Van::Receiving() {
  std::unordered_set<int> dead_set;
  std::mt19937 rng((unsigned)time(nullptr) + (unsigned)my_node_.port);
  const int drop = po_->cfg().drop_msg_pct;
  while (true) {
    Message msg;
    int n = RecvMsg(&msg);
    if (n < 0) break;
    // fault injection: randomly drop after the cluster is up (PS_DROP_MSG)
    if (drop > 0 && ready_.load() && msg.meta.control.cmd != Control::TERMINATE &&
        (int)(rng() % 100) < drop) {
      BPS_LOG(DEBUG) << "drop message (fault injection)";
      continue;
    }
    recv_bytes_ += n;
    if (po_->verbose() >= 3)
      BPS_LOG(INFO) << my_node_.debug() << " recv cmd=" << (int)msg.meta.control.cmd << " from " << msg.meta.sender
                    << " ts=" << msg.meta.timestamp;
    if (resender_ && resender_->AddIncoming(msg)) continue;
    if (profile_) ProfileEvent(msg, false);
    auto cmd = msg.meta.control.cmd;
    if (cmd == Control::TERMINATE) {
      ProcessTerminate();
      break;
    } else if (cmd == Control::ADD_NODE) {
      ProcessAddNode(&msg, &add_node_nodes_, &add_node_recovery_);
    } else if (cmd == Control::BARRIER) {
      ProcessBarrier(&msg, false);
    } else if (cmd == Control::INSTANCE_BARRIER) {
      ProcessBarrier(&msg, true);
    } else if (cmd == Control::HEARTBEAT) {
      ProcessHeartbeat(&msg);
    } else if (cmd == Control::ACK) {
      // consumed by the resender when enabled
    } else {
      ProcessData(&msg);
    }
  }
}

void Van::ProcessTerminate() { VLOG(po_, 1) << my_node_.debug() << " is stopped"; }


void Van::UpdateLocalID(Message* msg, std::unordered_set<int>* dead, Meta* nodes, Meta* recovery) {
  auto& ctrl = msg->meta.control;
  size_t num_nodes = (size_t)(po_->num_servers() + po_->num_workers());
  if (msg->meta.sender == kEmpty) {
    BPS_CHECK(is_scheduler_);
    BPS_CHECK_EQ(ctrl.node.size(), (size_t)1);
    if (nodes->control.node.size() < num_nodes) {
      nodes->control.node.push_back(ctrl.node[0]);
      registered_.store((int)nodes->control.node.size());
    } else {
      // a node died and restarted: hand it the id of a dead node of the same role
      BPS_CHECK(ready_.load());
      for (size_t i = 0; i + 1 < nodes->control.node.size(); ++i) {
        const Node& node = nodes->control.node[i];
        if (dead->count(node.id) && node.role == ctrl.node[0].role) {
          Node& rn = ctrl.node[0];
          rn.id = node.id;
          rn.is_recovery = true;
          VLOG(po_, 1) << "replace dead node " << node.debug() << " by " << rn.debug();
          nodes->control.node[i] = rn;
          recovery->control.node.push_back(rn);
          break;
        }
      }
    }
  }
  for (const Node& node : ctrl.node) {
    if (my_node_.hostname == node.hostname && my_node_.port == node.port) {
      if (my_node_.id == kEmpty) my_node_ = node;
    }
  }
}

void Van::ProcessAddNode(Message* msg, Meta* nodes, Meta* recovery) {
  std::unordered_set<int> dead_set;
  if (is_scheduler_ && ready_.load()) {
    auto dead = po_->GetDeadNodes(po_->cfg().heartbeat_timeout_s);
    dead_set.insert(dead.begin(), dead.end());
  }
  UpdateLocalID(msg, &dead_set, nodes, recovery);
  if (is_scheduler_) {
    recovery->control.cmd = Control::ADD_NODE;
    time_t t = time(nullptr);
    size_t num_nodes = (size_t)(po_->num_servers() + po_->num_workers());
    if (nodes->control.node.size() == num_nodes && !ready_.load()) {
      auto& v = nodes->control.node;
      // order: explicit BYTEPS_ORDERED_HOSTS, else by (host, port)
      std::string ordered = env_str("BYTEPS_ORDERED_HOSTS", "");
      if (!ordered.empty()) {
        std::unordered_map<std::string, size_t> pos;
        size_t i = 0, p;
        std::string s = ordered;
        while (true) {
          p = s.find(',');
          std::string h = s.substr(0, p);
          pos[h.substr(0, h.find(':'))] = i++;
          if (p == std::string::npos) break;
          s.erase(0, p + 1);
        }
        std::stable_sort(v.begin(), v.end(), [&pos](const Node& a, const Node& b) {
          return pos[a.hostname] < pos[b.hostname];
        });
      } else {
        std::sort(v.begin(), v.end(), [](const Node& a, const Node& b) {
          if (a.hostname != b.hostname) return a.hostname < b.hostname;
          return a.port < b.port;
        });
      }
      bool preferred = false;
      for (auto& n : v) preferred |= (n.aux_id != -1);
      if (preferred) {
        std::unordered_set<int> sr, wr;
        for (auto& n : v) {
          auto& set = n.role == Role::kServer ? sr : wr;
          if (n.aux_id < 0 || !set.insert(n.aux_id).second) {
            preferred = false;   // incomplete / duplicate hints: fall back to address order
            break;
          }
        }
        if (preferred) {
          for (int i = 0; i < po_->num_servers(); ++i) preferred &= sr.count(i) > 0;
          for (int i = 0; i < po_->num_workers(); ++i) preferred &= wr.count(i) > 0;
        }
      }
      for (auto& n : v) {
        std::string hp = n.hostname + ":" + std::to_string(n.port);
        int id = n.role == Role::kServer ? Postoffice::ServerRankToID(preferred ? n.aux_id : num_servers_)
                                         : Postoffice::WorkerRankToID(preferred ? n.aux_id : num_workers_);
        if (!connected_nodes_.count(hp)) {
          n.id = id;
          Connect(n);
          po_->UpdateHeartbeat(n.id, t);
          connected_nodes_[hp] = id;
          VLOG(po_, 1) << "assign id=" << id << " to " << n.debug();
        } else {
          shared_node_mapping_[id] = connected_nodes_[hp];
          n.id = connected_nodes_[hp];
        }
        if (n.role == Role::kServer) ++num_servers_;
        else ++num_workers_;
      }
      v.push_back(my_node_);
      nodes->control.cmd = Control::ADD_NODE;
      Message back;
      back.meta = *nodes;
      for (int r : po_->GetNodeIDs(kWorkerGroup + kServerGroup)) {
        if (!shared_node_mapping_.count(r)) {
          back.meta.recver = r;
          back.meta.sender = my_node_.id;
          back.meta.timestamp = timestamp_++;
          back.meta.msg_sig = 0;
          Send(back);
        }
      }
      VLOG(po_, 1) << "the scheduler is connected to " << num_workers_ << " workers and " << num_servers_
                   << " servers";
      ready_ = true;
    } else if (!recovery->control.node.empty()) {
      auto dead = po_->GetDeadNodes(po_->cfg().heartbeat_timeout_s);
      std::unordered_set<int> ds(dead.begin(), dead.end());
      BPS_CHECK_EQ(recovery->control.node.size(), (size_t)1);
      Connect(recovery->control.node[0]);
      po_->UpdateHeartbeat(recovery->control.node[0].id, t);
      for (int r : po_->GetNodeIDs(kWorkerGroup + kServerGroup)) {
        if (r != recovery->control.node[0].id && ds.count(r)) continue;  // never talk to dead nodes
        Message back;
        back.meta = (r == recovery->control.node[0].id) ? *nodes : *recovery;
        back.meta.control.cmd = Control::ADD_NODE;
        back.meta.recver = r;
        back.meta.sender = my_node_.id;
        back.meta.timestamp = timestamp_++;
        back.meta.msg_sig = 0;
        Send(back);
      }
      recovery->control.node.clear();
    }
  } else {
    // worker/server: connect to the peers I talk to
    for (const Node& node : msg->meta.control.node) {
      std::string hp = node.hostname + ":" + std::to_string(node.port);
      if (!connected_nodes_.count(hp)) {
        // same-role nodes never exchange data (ps-lite skips those links too)
        if (node.role != my_node_.role || node.id == my_node_.id) Connect(node);
        connected_nodes_[hp] = node.id;
      } else if (node.is_recovery) {
        if (node.role != my_node_.role) Connect(node);   // re-connect to a restarted peer
      }
      if (!node.is_recovery && node.role == Role::kServer) ++num_servers_;
      if (!node.is_recovery && node.role == Role::kWorker) ++num_workers_;
    }
    VLOG(po_, 1) << my_node_.debug() << " is connected to others";
    ready_ = true;
  }
}

void Van::ProcessBarrier(Message* msg, bool instance) {
  auto& ctrl = msg->meta.control;
  if (msg->meta.request) {
    int group = ctrl.barrier_group;
    if ((size_t)group >= barrier_count_.size()) barrier_count_.resize(group + 1, 0);
    ++barrier_count_[group];
    VLOG(po_, 1) << "barrier count for group " << group << ": " << barrier_count_[group];
    if (barrier_count_[group] == (int)po_->GetNodeIDs(group).size()) {
      barrier_count_[group] = 0;
      Message res;
      res.meta.request = false;
      res.meta.app_id = msg->meta.app_id;
      res.meta.customer_id = msg->meta.customer_id;
      res.meta.control.cmd = instance ? Control::INSTANCE_BARRIER : Control::BARRIER;
      for (int r : po_->GetNodeIDs(group)) {
        if (!shared_node_mapping_.count(r)) {
          res.meta.recver = r;
          res.meta.sender = my_node_.id;
          res.meta.timestamp = timestamp_++;
          res.meta.msg_sig = 0;
          Send(res);
        }
      }
    }
  } else {
    po_->ManageBarrier(msg->meta.customer_id);
  }
}

void Van::ProcessHeartbeat(Message* msg) {
  time_t t = time(nullptr);
  for (auto& node : msg->meta.control.node) {
    po_->UpdateHeartbeat(node.id, t);
    if (is_scheduler_) {
      Message back;
      back.meta.control.cmd = Control::HEARTBEAT;
      back.meta.recver = node.id;
      back.meta.timestamp = timestamp_++;
      Send(back);   // echo so the node knows the scheduler is alive
    }
  }
  if (!is_scheduler_) po_->UpdateHeartbeat(kScheduler, t);
}

void Van::Heartbeat() {
  const int interval = po_->cfg().heartbeat_interval_s;
  int slept_ms = 0;
  while (interval > 0 && ready_.load() && !stopping_) {
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    slept_ms += 50;
    if (slept_ms < interval * 1000) continue;
    slept_ms = 0;
    Message msg;
    msg.meta.recver = kScheduler;
    msg.meta.control.cmd = Control::HEARTBEAT;
    msg.meta.control.node.push_back(my_node_);
    msg.meta.timestamp = timestamp_++;
    Send(msg);
  }
}

bool Van::TryDirectData(Message* msg, int nbytes) {
  // one atomic flag, computed in Start() once resender / profiling are known and cleared first thing in Stop():
  // transport threads never look at those (non-atomic) members themselves
  if (!direct_ok_.load(std::memory_order_acquire) || !ready_.load()) return false;
  const Meta& m = msg->meta;
  if (!m.control.empty() || !(m.push || m.pull) || m.simple_app) return false;
  recv_bytes_ += nbytes;
  ProcessData(msg);
  return true;
}

void Van::ProcessData(Message* msg) {
  int app_id = msg->meta.app_id;
  int customer_id = po_->is_worker() ? msg->meta.customer_id : app_id;
  // lookup and hand-over happen under the post office's lock, so a customer that is being destroyed
  // (RemoveCustomer takes the same lock first) can never be entered afterwards (ThreadSanitizer: ~Customer vs Accept)
  if (!po_->Deliver(app_id, customer_id, *msg, 5))
    BPS_LOG(WARNING) << "no customer for app " << app_id << " customer " << customer_id << "; message dropped";
}

// ================================================================ TcpVan
static bool write_all(int fd, const iovec* iov_in, int cnt) {
  std::vector<iovec> iov(iov_in, iov_in + cnt);
  size_t idx = 0;
  while (idx < iov.size()) {
    int n = (int)std::min<size_t>(iov.size() - idx, 64);
    ssize_t w = writev(fd, &iov[idx], n);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    size_t left = (size_t)w;
    while (left > 0 && idx < iov.size()) {
      if (left >= iov[idx].iov_len) {
        left -= iov[idx].iov_len;
        ++idx;
      } else {
        iov[idx].iov_base = (char*)iov[idx].iov_base + left;
        iov[idx].iov_len -= left;
        left = 0;
      }
    }
    while (idx < iov.size() && iov[idx].iov_len == 0) ++idx;
  }
  return true;
}

static bool read_all(int fd, void* buf, size_t len) {
  char* p = (char*)buf;
  while (len > 0) {
    ssize_t r = read(fd, p, len);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    len -= (size_t)r;
  }
  return true;
}

static void tune_socket(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  int buf = 4 << 20;
  setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
  setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
}

TcpVan::~TcpVan() { StopTransport(); }

// DMLC_LOCAL=1 (ps-lite zmq_van.h:109-115,180-181 switches tcp://host:port to ipc:///tmp/<port>): every node of
// the job is on this host, so the stream sockets are Unix-domain ones.  The "port" stays the node's identity in
// the node table; the address is the Linux abstract name "\0byteps_van_<port>" (nothing to unlink afterwards).
static socklen_t unix_addr(sockaddr_un* a, int port) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  int n = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "byteps_van_%d", port);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

int TcpVan::BindLocal(Node& node, int max_retry) {
  listen_fd_ = socket(AF_UNIX, SOCK_STREAM, 0);
  if (listen_fd_ < 0) return -1;
  std::mt19937 rng((unsigned)time(nullptr) ^ ((unsigned)getpid() << 8));
  int port = node.port > 0 ? node.port : 10000 + (int)(rng() % 40000);
  for (int i = 0; i <= max_retry; ++i) {
    sockaddr_un a;
    socklen_t len = unix_addr(&a, port);
    if (bind(listen_fd_, (sockaddr*)&a, len) == 0) {
      if (listen(listen_fd_, 256) != 0) return -1;
      closed_ = false;
      acceptor_ = std::thread([this] { AcceptLoop(); });
      return port;
    }
    port = 10000 + (int)(rng() % 40000);
  }
  close(listen_fd_);
  listen_fd_ = -1;
  return -1;
}

int TcpVan::Bind(Node& node, int max_retry) {
  local_ = po_->cfg().local;
  if (local_) return BindLocal(node, max_retry);
  listen_fd_ = socket(AF_INET, SOCK_STREAM, 0);
  if (listen_fd_ < 0) return -1;
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  int port = node.port;
  std::mt19937 rng((unsigned)time(nullptr) ^ (unsigned)getpid());
  for (int i = 0; i <= max_retry; ++i) {
    sockaddr_in a;
    memset(&a, 0, sizeof(a));
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_ANY);
    a.sin_port = htons((uint16_t)port);
    if (bind(listen_fd_, (sockaddr*)&a, sizeof(a)) == 0) {
      if (port == 0) {
        socklen_t len = sizeof(a);
        getsockname(listen_fd_, (sockaddr*)&a, &len);
        port = ntohs(a.sin_port);
        if (port == po_->cfg().scheduler_port + 1 && i < max_retry) {
          // scheduler port + 1 is where the workers' torch.distributed rendezvous listens (common/__init__.py);
          // the kernel likes to hand out the neighbour of a port it has just assigned - give it back
          close(listen_fd_);
          listen_fd_ = socket(AF_INET, SOCK_STREAM, 0);
          if (listen_fd_ < 0) return -1;
          setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
          port = 10000 + (int)(rng() % 40000);
          if (port == po_->cfg().scheduler_port + 1) ++port;
          continue;
        }
      }
      if (listen(listen_fd_, 256) != 0) return -1;
      closed_ = false;
      acceptor_ = std::thread([this] { AcceptLoop(); });
      return port;
    }
    if (i == max_retry) break;
    port = 10000 + (int)(rng() % 40000);
    if (port == po_->cfg().scheduler_port + 1) ++port;
  }
  close(listen_fd_);
  listen_fd_ = -1;
  return -1;
}

void TcpVan::AcceptLoop() {
  while (!closed_) {
    pollfd p{listen_fd_, POLLIN, 0};
    int r = poll(&p, 1, 100);
    if (r <= 0) continue;
    int fd = accept(listen_fd_, nullptr, nullptr);
    if (fd < 0) continue;
    tune_socket(fd);
    std::lock_guard<std::mutex> g(readers_mu_);
    reader_fds_.push_back(fd);
    readers_.emplace_back([this, fd] { ReadLoop(fd); });
  }
}

struct FrameHeader {
  uint32_t magic;
  uint32_t meta_len;
  uint32_t ndata;
  uint32_t pad;
};
static constexpr uint32_t kMagic = 0x62707332;  // "bps2"

void TcpVan::ReadLoop(int fd) {
  while (!closed_) {
    FrameHeader h;
    if (!read_all(fd, &h, sizeof(h))) break;
    if (h.magic != kMagic || h.ndata > 64 || h.meta_len > (64u << 20)) {
      BPS_LOG(ERROR) << "corrupt frame; closing connection";
      break;
    }
    std::vector<uint64_t> lens(h.ndata);
    if (h.ndata && !read_all(fd, lens.data(), h.ndata * 8)) break;
    std::string meta(h.meta_len, '\0');
    if (!read_all(fd, &meta[0], h.meta_len)) break;
    Message msg;
    if (!meta_unpack(meta.data(), meta.size(), &msg.meta)) {
      BPS_LOG(ERROR) << "cannot decode meta; closing connection";
      break;
    }
    bool ok = true;
    for (uint32_t i = 0; i < h.ndata; ++i) {
      SArray<char> a;
      char* direct = nullptr;
      size_t direct_len = 0;
      if (i == 0 && h.ndata == 1 && TakeRecvBuffer(msg.meta, &direct, &direct_len) && direct_len == lens[i]) {
        a = SArray<char>(direct, (size_t)lens[i]);          // the requester's own memory, not owned
      } else {
        a = PayloadPool::get().alloc((size_t)lens[i]);
      }
      if (lens[i] && !read_all(fd, a.data(), (size_t)lens[i])) {
        ok = false;
        break;
      }
      a.src_dev = msg.meta.src_dev; a.src_id = msg.meta.src_id;
      a.dst_dev = msg.meta.dst_dev; a.dst_id = msg.meta.dst_id;
      msg.data.push_back(a);
    }
    if (!ok) break;
    ipc_recv_attach(&msg);
    if (TryDirectData(&msg, (int)std::min<size_t>(msg.data_bytes() + 64, 0x7fffffff))) continue;
    {
      std::lock_guard<std::mutex> g(q_mu_);
      recv_q_.push(std::move(msg));
    }
    q_cv_.notify_one();
  }
  // Deregister BEFORE closing: once closed the number can be handed to any new socket of this
  // process (several vans may live in one process), and a later StopTransport() shutting down
  // a stale entry would kill that unrelated connection.
  std::lock_guard<std::mutex> g(readers_mu_);
  reader_fds_.erase(std::remove(reader_fds_.begin(), reader_fds_.end(), fd), reader_fds_.end());
  close(fd);
}

void TcpVan::ipc_recv_attach(Message* msg) {
  // colocated sender left the payload in shared memory: build a zero-copy view
  Meta& m = msg->meta;
  if (m.shm_name.empty() || !m.push || !m.request || !msg->data.empty()) return;
  void* base = ShmRegistry::get().open(m.shm_name, (size_t)(m.shm_offset + m.shm_len));
  if (!base) {
    BPS_LOG(ERROR) << "cannot map shm " << m.shm_name;
    return;
  }
  msg->data.push_back(SArray<char>((char*)base + m.shm_offset, (size_t)m.shm_len, false));
}

void TcpVan::Connect(const Node& node) {
  BPS_CHECK_NE(node.id, kEmpty);
  BPS_CHECK_NE(node.port, 0);
  int id = node.id;
  {
    std::lock_guard<std::mutex> g(senders_mu_);
    auto it = senders_.find(id);
    if (it != senders_.end()) {
      for (auto& l : it->second->lanes)
        if (l->fd >= 0) close(l->fd);
      senders_.erase(it);
    }
  }
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  std::string port = std::to_string(node.port);
  sockaddr_un ua;
  socklen_t ualen = 0;
  if (local_) {
    ualen = unix_addr(&ua, node.port);
  } else if (getaddrinfo(node.hostname.c_str(), port.c_str(), &hints, &res) != 0 || !res) {
    BPS_LOG(ERROR) << "cannot resolve " << node.hostname;
    return;
  }
  auto s = std::make_shared<Sender>();
  // the scheduler only ever sees control traffic: one lane is enough
  const int lanes = (node.role == Role::kScheduler || po_->cfg().role == Role::kScheduler)
                        ? 1 : std::max(1, po_->cfg().num_lanes);
  auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(env_int("BYTEPS_CONNECT_TIMEOUT_S", 60));
  for (int l = 0; l < lanes; ++l) {
    int fd = -1;
    while (std::chrono::steady_clock::now() < deadline && !closed_) {
      fd = socket(local_ ? AF_UNIX : AF_INET, SOCK_STREAM, 0);
      if (fd < 0) break;
      if (local_ ? connect(fd, (sockaddr*)&ua, ualen) == 0 : connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      close(fd);
      fd = -1;
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    if (fd < 0) {
      BPS_LOG(ERROR) << "cannot connect to " << node.debug();
      for (auto& ln : s->lanes) close(ln->fd);
      if (res) freeaddrinfo(res);
      return;
    }
    tune_socket(fd);
    auto lane = std::make_unique<Lane>();
    lane->fd = fd;
    s->lanes.push_back(std::move(lane));
  }
  if (res) freeaddrinfo(res);
  s->addr = (local_ ? std::string("unix:") : node.hostname + ":") + port;
  s->colocated = local_ || (node.hostname == my_node_.hostname) || node.hostname == "127.0.0.1" ||
                 node.hostname == "localhost";
  std::lock_guard<std::mutex> g(senders_mu_);
  senders_[id] = s;
}

bool TcpVan::IsColocated(int id) {
  std::lock_guard<std::mutex> g(senders_mu_);
  auto it = senders_.find(id);
  return it != senders_.end() && it->second->colocated;
}

bool TcpVan::ipc_send_strip(Message& msg, Sender* s) {
  if (!po_->cfg().enable_ipc || !s->colocated || msg.data.size() != 1) return false;
  Meta& m = msg.meta;
  if (!(m.push && m.request)) return false;
  std::string name;
  uint64_t off;
  if (!ShmRegistry::get().lookup(msg.data[0].data(), msg.data[0].size(), &name, &off)) return false;
  m.shm_name = name;
  m.shm_offset = off;
  m.shm_len = msg.data[0].size();
  return true;
}

int TcpVan::SendMsg(Message& msg) {
  int id = msg.meta.recver;
  BPS_CHECK_NE(id, kEmpty);
  std::shared_ptr<Sender> s;
  {
    std::lock_guard<std::mutex> g(senders_mu_);
    auto it = senders_.find(id);
    if (it == senders_.end()) {
      if (!closed_) BPS_LOG(WARNING) << "there is no socket to node " << id;
      return -1;
    }
    s = it->second;
  }
  Message stripped;
  const Message* out = &msg;
  if (ipc_send_strip(msg, s.get())) {
    stripped.meta = msg.meta;   // payload stays in shm; only the meta travels
    out = &stripped;
  }
  std::string meta = meta_pack(out->meta);
  FrameHeader h{kMagic, (uint32_t)meta.size(), (uint32_t)out->data.size(), 0};
  std::vector<uint64_t> lens;
  for (auto& d : out->data) lens.push_back(d.size());
  std::vector<iovec> iov;
  iov.push_back({&h, sizeof(h)});
  if (!lens.empty()) iov.push_back({lens.data(), lens.size() * 8});
  iov.push_back({&meta[0], meta.size()});
  size_t total = sizeof(h) + lens.size() * 8 + meta.size();
  for (auto& d : out->data) {
    if (d.size()) iov.push_back({d.data(), d.size()});
    total += d.size();
  }
  const bool is_data = msg.meta.control.empty() && (msg.meta.push || msg.meta.pull) && !msg.meta.simple_app;
  Lane* lane = s->lanes[is_data ? msg.meta.key % s->lanes.size() : 0].get();
  std::lock_guard<std::mutex> g(lane->mu);
  if (lane->fd < 0 || !write_all(lane->fd, iov.data(), (int)iov.size())) return -1;
  return (int)std::min<size_t>(total, 0x7fffffff);
}

int TcpVan::RecvMsg(Message* msg) {
  std::unique_lock<std::mutex> lk(q_mu_);
  q_cv_.wait(lk, [this] { return !recv_q_.empty() || closed_.load(); });
  if (recv_q_.empty()) return -1;
  *msg = std::move(recv_q_.front());
  recv_q_.pop();
  size_t n = msg->data_bytes() + 64;
  return (int)std::min<size_t>(n, 0x7fffffff);
}

void TcpVan::StopTransport() {
  if (closed_.exchange(true)) {
    // already closed (or never bound)
  }
  q_cv_.notify_all();
  if (acceptor_.joinable()) acceptor_.join();
  if (listen_fd_ >= 0) {
    close(listen_fd_);
    listen_fd_ = -1;
  }
  {
    std::lock_guard<std::mutex> g(senders_mu_);
    for (auto& kv : senders_) {
      for (auto& ln : kv.second->lanes) {
        std::lock_guard<std::mutex> g2(ln->mu);
        if (ln->fd < 0) continue;
        shutdown(ln->fd, SHUT_RDWR);
        close(ln->fd);
        ln->fd = -1;
      }
    }
    senders_.clear();
  }
  std::vector<std::thread> rs;
  {
    std::lock_guard<std::mutex> g(readers_mu_);
    for (int fd : reader_fds_) shutdown(fd, SHUT_RDWR);
    rs.swap(readers_);
    reader_fds_.clear();
  }
  for (auto& t : rs)
    if (t.joinable()) t.join();
  std::lock_guard<std::mutex> g(q_mu_);
  while (!recv_q_.empty()) recv_q_.pop();
}

// ================================================================ Postoffice
// Transport factory (ps-lite: Van::Create + DMLC_PS_VAN_TYPE / DMLC_ENABLE_RDMA, van.cc:82-112).  "tcp" (alias
// "zmq": the default socket transport) and "shm" (one host, no sockets) exist in this build; the verbs / UCX /
// libfabric names fail loudly instead of silently degrading.
static Van* create_van(Postoffice* po) {
  const std::string& type = po->cfg().van_type;
  if (type == "tcp" || type == "zmq" || type == "0" || type.empty()) return new TcpVan(po);
  if (type == "shm") return new ShmVan(po);
  BPS_LOG_FATAL << "DMLC_PS_VAN_TYPE=" << type << ": transport not built (available: tcp, shm; rdma/ucx/fabric need "
                << "libraries that are not part of this build)";
  return nullptr;
}

Postoffice::Postoffice(const NetConfig& cfg) : cfg_(cfg) {
  if (cfg_.van_type == "shm") cfg_.enable_ipc = true;   // every peer is colocated: registered windows go by reference
  van_.reset(create_van(this));
  InitNodeIDs();
}

Postoffice::~Postoffice() {}

void Postoffice::InitNodeIDs() {
  node_ids_.clear();
  for (int i = 0; i < cfg_.num_workers; ++i) {
    int id = WorkerRankToID(i);
    for (int g : {id, kWorkerGroup, kWorkerGroup + kServerGroup, kWorkerGroup + kScheduler,
                  kWorkerGroup + kServerGroup + kScheduler})
      node_ids_[g].push_back(id);
  }
  for (int i = 0; i < cfg_.num_servers; ++i) {
    int id = ServerRankToID(i);
    for (int g : {id, kServerGroup, kWorkerGroup + kServerGroup, kServerGroup + kScheduler,
                  kWorkerGroup + kServerGroup + kScheduler})
      node_ids_[g].push_back(id);
  }
  for (int g : {kScheduler, kScheduler + kServerGroup + kWorkerGroup, kScheduler + kWorkerGroup,
                kScheduler + kServerGroup})
    node_ids_[g].push_back(kScheduler);
}

const std::vector<int>& Postoffice::GetNodeIDs(int node_group) const {
  auto it = node_ids_.find(node_group);
  BPS_CHECK(it != node_ids_.end()) << "node group " << node_group << " does not exist";
  return it->second;
}

void Postoffice::Start(int customer_id, bool do_barrier) {
  {
    std::lock_guard<std::mutex> g(start_mu_);
    if (init_stage_ == 0) {
      start_time_ = time(nullptr);
      init_stage_ = 1;
    }
  }
  van_->Start(customer_id);
  if (do_barrier) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler);
}

void Postoffice::Finalize(int customer_id, bool do_barrier) {
  if (do_barrier) Barrier(customer_id, kWorkerGroup + kServerGroup + kScheduler);
  if (customer_id == 0) {
    van_->Stop();
    std::lock_guard<std::mutex> g(start_mu_);
    init_stage_ = 0;
    barrier_done_.clear();
  }
}

void Postoffice::AddCustomer(Customer* c) {
  std::lock_guard<std::mutex> g(mu_);
  auto& m = customers_[c->app_id()];
  BPS_CHECK(!m.count(c->customer_id())) << "customer " << c->customer_id() << " already exists";
  m[c->customer_id()] = c;
  std::lock_guard<std::mutex> g2(barrier_mu_);
  barrier_done_[c->customer_id()] = false;
}

void Postoffice::RemoveCustomer(Customer* c) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = customers_.find(c->app_id());
  if (it != customers_.end()) it->second.erase(c->customer_id());
}

bool Postoffice::Deliver(int app_id, int customer_id, const Message& msg, int timeout_s) {
  for (int i = 0; i <= timeout_s * 1000; ++i) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = customers_.find(app_id);
      if (it != customers_.end()) {
        auto jt = it->second.find(customer_id);
        if (jt != it->second.end()) {
          jt->second->Accept(msg);
          return true;
        }
      }
    }
    if (i < timeout_s * 1000) std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return false;
}

Customer* Postoffice::GetCustomer(int app_id, int customer_id, int timeout_s) {
  for (int i = 0; i <= timeout_s * 1000; ++i) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = customers_.find(app_id);
      if (it != customers_.end()) {
        auto jt = it->second.find(customer_id);
        if (jt != it->second.end()) return jt->second;
      }
    }
    if (i < timeout_s * 1000) std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  return nullptr;
}

void Postoffice::Barrier(int customer_id, int node_group) {
  if (GetNodeIDs(node_group).size() <= 1) return;
  Role r = van_->my_node().role;
  if (r == Role::kScheduler) BPS_CHECK(node_group & kScheduler);
  else if (r == Role::kWorker) BPS_CHECK(node_group & kWorkerGroup);
  else BPS_CHECK(node_group & kServerGroup);
  std::unique_lock<std::mutex> lk(barrier_mu_);
  barrier_done_[customer_id] = false;
  Message req;
  req.meta.recver = kScheduler;
  req.meta.request = true;
  req.meta.control.cmd = Control::BARRIER;
  req.meta.app_id = 0;
  req.meta.customer_id = customer_id;
  req.meta.control.barrier_group = node_group;
  req.meta.timestamp = van_->GetTimestamp();
  lk.unlock();
  van_->Send(req);
  lk.lock();
  barrier_cv_.wait(lk, [this, customer_id] { return barrier_done_[customer_id]; });
}

void Postoffice::ManageBarrier(int customer_id) {
  {
    std::lock_guard<std::mutex> g(barrier_mu_);
    barrier_done_[customer_id] = true;
  }
  barrier_cv_.notify_all();
}

const std::vector<std::pair<uint64_t, uint64_t>>& Postoffice::GetServerKeyRanges() {
  std::lock_guard<std::mutex> g(mu_);
  if (key_ranges_.empty()) {
    const uint64_t kMax = ~0ull;
    for (int i = 0; i < cfg_.num_servers; ++i)
      key_ranges_.emplace_back(kMax / cfg_.num_servers * i, kMax / cfg_.num_servers * (i + 1));
  }
  return key_ranges_;
}

void Postoffice::UpdateHeartbeat(int node_id, time_t t) {
  std::lock_guard<std::mutex> g(hb_mu_);
  heartbeats_[node_id] = t;
}

std::vector<int> Postoffice::GetDeadNodes(int timeout_s) {
  std::vector<int> dead;
  if (!van_->IsReady() || timeout_s == 0) return dead;
  time_t now = time(nullptr);
  const auto& nodes = is_scheduler() ? GetNodeIDs(kWorkerGroup + kServerGroup) : GetNodeIDs(kScheduler);
  std::lock_guard<std::mutex> g(hb_mu_);
  for (int r : nodes) {
    auto it = heartbeats_.find(r);
    if ((it == heartbeats_.end() || it->second + timeout_s < now) && start_time_ + timeout_s < now)
      dead.push_back(r);
  }
  return dead;
}

// ================================================================ Customer
Customer::Customer(int app_id, int customer_id, RecvHandle h, Postoffice* po)
    : app_id_(app_id), customer_id_(customer_id), handle_(std::move(h)), po_(po) {
  if (env_bool("DMLC_LOCKLESS_QUEUE", false)) ring_ = std::make_unique<SpscQueue<Message>>(8192);
  po_->AddCustomer(this);
  thread_ = std::thread([this] { Receiving(); });
}

Customer::~Customer() {
  po_->RemoveCustomer(this);
  {
    std::lock_guard<std::mutex> g(q_mu_);
    stop_ = true;
  }
  ring_stop_.store(true, std::memory_order_release);
  q_cv_.notify_all();
  if (thread_.joinable()) thread_.join();
}

int Customer::NewRequest(int recver) {
  std::lock_guard<std::mutex> g(tracker_mu_);
  int num = (int)po_->GetNodeIDs(recver).size();
  tracker_.emplace_back(num, 0);
  return (int)tracker_.size() - 1;
}

void Customer::WaitRequest(int ts) {
  std::unique_lock<std::mutex> lk(tracker_mu_);
  tracker_cv_.wait(lk, [this, ts] { return tracker_[ts].first == tracker_[ts].second; });
}

int Customer::NumResponse(int ts) {
  std::lock_guard<std::mutex> g(tracker_mu_);
  return tracker_[ts].second;
}

void Customer::AddResponse(int ts, int num) {
  {
    std::lock_guard<std::mutex> g(tracker_mu_);
    tracker_[ts].second += num;
  }
  tracker_cv_.notify_all();
}

void Customer::Accept(const Message& m) {
  if (ring_) {
    ring_->push(m);
    return;
  }
  {
    std::lock_guard<std::mutex> g(q_mu_);
    q_.push(m);
  }
  q_cv_.notify_one();
}

void Customer::Receiving() {
  while (true) {
    Message m;
    if (ring_) {
      if (!ring_->wait_pop(&m, ring_stop_)) break;
    } else {
      std::unique_lock<std::mutex> lk(q_mu_);
      q_cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
      if (stop_ && q_.empty()) break;
      m = std::move(q_.front());
      q_.pop();
    }
    handle_(m);
    if (!m.meta.request) {
      std::lock_guard<std::mutex> g(tracker_mu_);
      if (m.meta.timestamp >= 0 && (size_t)m.meta.timestamp < tracker_.size()) ++tracker_[m.meta.timestamp].second;
    }
    tracker_cv_.notify_all();
  }
}

}  // namespace net
}  // namespace bps
