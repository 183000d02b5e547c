#include "net/shm_van.h"

#include <dirent.h>
#include <fcntl.h>
#include <linux/futex.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <ctime>
#include <random>
#include <thread>

#include "core/env.h"
#include "core/log.h"

namespace bps {
namespace net {

namespace {

constexpr uint32_t kQueueMagic = 0x62707351;   // "bpsQ"
constexpr uint32_t kArenaMagic = 0x62707341;   // "bpsA"
constexpr size_t kHeaderBytes = 4096;          // queue / arena header page
constexpr size_t kAlign = 64;

enum BlobKind : uint32_t {
  kInline = 0,      // bytes inside the slot
  kArena = 1,       // bytes in the sender's arena towards me; `advance` returns the space
  kRegistered = 2,  // bytes in a window the sender registered (ShmRegistry): zero-copy view
  kOneOff = 3,      // a dedicated segment for an oversized payload; the receiver unlinks it
};

struct Blob {
  uint32_t kind;
  uint32_t pad;
  uint64_t len;
  uint64_t off;
  uint64_t advance;       // kArena: value of the arena's consumed counter once this blob has been copied out
  char name[56];          // kArena / kRegistered / kOneOff: shm object name
  uint64_t object_len;    // kArena: size of the arena object (so the receiver can map it)
};
static_assert(sizeof(Blob) == 96, "blob descriptor layout");

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

long futex(std::atomic<uint32_t>* addr, int op, uint32_t val, const timespec* ts) {
  return syscall(SYS_futex, reinterpret_cast<uint32_t*>(addr), op, val, ts, nullptr, 0);
}

std::string queue_name(int port) { return "/bps_shmvan_" + std::to_string(port); }

// ThreadSanitizer keys synchronisation on virtual addresses.  When two vans live in ONE process (the in-process test
// clusters) the same slot is mapped twice, so the release store of the sender and the acquire load of the receiver
// hit different addresses and the happens-before edge they establish is invisible to it.  Tell it explicitly.
#if defined(__SANITIZE_THREAD__)
extern "C" void __tsan_acquire(void* addr);
extern "C" void __tsan_release(void* addr);
char tsan_token;
inline void hb_release() { __tsan_release(&tsan_token); }
inline void hb_acquire() { __tsan_acquire(&tsan_token); }
#else
inline void hb_release() {}
inline void hb_acquire() {}
#endif

}  // namespace

struct ShmVan::QueueHeader {
  uint32_t magic;
  uint32_t owner_pid;
  uint32_t slots;
  uint32_t slot_bytes;
  alignas(64) std::atomic<uint64_t> tail;        // producers claim positions here
  alignas(64) std::atomic<uint32_t> doorbell;    // bumped after every publication
  std::atomic<uint32_t> sleeping;                // the consumer is (about to be) parked on the doorbell
};

struct ShmVan::Slot {
  std::atomic<uint64_t> seq;     // == position: free for that position; == position + 1: published
  uint32_t nblobs;               // blob 0 is the packed meta
  uint32_t sender_port;
  Blob blobs[kMaxBlobs];
  char bytes[kSlotBytes - 16 - sizeof(Blob) * kMaxBlobs];
};

struct ShmVan::ArenaHeader {
  uint32_t magic;
  uint32_t pad;
  uint64_t size;                               // bytes of the data area (after the header page)
  alignas(64) std::atomic<uint64_t> consumed;  // bytes the receiver has released (monotonic)
};

static_assert(sizeof(ShmVan::Slot) == ShmVan::kSlotBytes, "slot layout");

ShmVan::~ShmVan() { StopTransport(); }

bool ShmVan::MapObject(const std::string& name, size_t len, bool create, Mapping* out) {
  int fd = shm_open(name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) return false;
  if (create && ftruncate(fd, (off_t)len) != 0) {
    close(fd);
    shm_unlink(name.c_str());
    return false;
  }
  if (!create) {
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < len) {   // the creator has not sized it yet
      close(fd);
      return false;
    }
  }
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    if (create) shm_unlink(name.c_str());
    return false;
  }
  out->name = name;
  out->base = static_cast<char*>(p);
  out->len = len;
  out->owner = create;
  return true;
}

void ShmVan::Unmap(Mapping* m, bool unlink) {
  if (m->base) munmap(m->base, m->len);
  if (unlink && !m->name.empty()) shm_unlink(m->name.c_str());
  *m = Mapping();
}

// Objects of processes that died without StopTransport() (kill -9, a crash) would stay in /dev/shm for ever:
// sweep them once per process.  Arena / one-off names carry the creator's pid, queues record it in their header.
static void collect_garbage() {
  DIR* d = opendir("/dev/shm");
  if (!d) return;
  auto pid_dead = [](long pid) { return pid > 0 && kill((pid_t)pid, 0) != 0 && errno == ESRCH; };
  while (dirent* e = readdir(d)) {
    const std::string f = e->d_name;
    if (f.rfind("bps_shmvan_", 0) != 0) continue;
    const std::string name = "/" + f;
    long pid = -1;
    unsigned a, b, n;
    if (sscanf(f.c_str(), "bps_shmvan_big_%ld_%u", &pid, &n) == 2 ||
        sscanf(f.c_str(), "bps_shmvan_%u_to_%u_%ld_%u", &a, &b, &pid, &n) == 4) {
      if (pid_dead(pid)) shm_unlink(name.c_str());
      continue;
    }
    int fd = shm_open(name.c_str(), O_RDONLY, 0600);      // a queue: the owner is in the header
    if (fd < 0) continue;
    uint32_t head[2] = {0, 0};
    const bool got = read(fd, head, sizeof(head)) == (ssize_t)sizeof(head);
    close(fd);
    if (got && head[0] == kQueueMagic && pid_dead((long)head[1])) shm_unlink(name.c_str());
  }
  closedir(d);
}

int ShmVan::Bind(Node& node, int max_retry) {
  static std::once_flag swept;
  std::call_once(swept, collect_garbage);
  arena_bytes_ = (size_t)std::max<long long>(1, env_int("BYTEPS_SHMVAN_ARENA_MB", 32)) << 20;
  const size_t len = kHeaderBytes + (size_t)kSlots * kSlotBytes;
  std::mt19937 rng((unsigned)time(nullptr) ^ ((unsigned)getpid() << 10) ^ (unsigned)(uintptr_t)this);
  const bool fixed = node.port > 0;
  int port = fixed ? node.port : 10000 + (int)(rng() % 40000);
  for (int i = 0; i <= std::max(max_retry, 8); ++i) {
    const std::string name = queue_name(port);
    if (MapObject(name, len, true, &my_queue_)) {
      auto* q = reinterpret_cast<QueueHeader*>(my_queue_.base);
      q->owner_pid = (uint32_t)getpid();
      q->slots = kSlots;
      q->slot_bytes = kSlotBytes;
      new (&q->tail) std::atomic<uint64_t>(0);
      new (&q->doorbell) std::atomic<uint32_t>(0);
      new (&q->sleeping) std::atomic<uint32_t>(0);
      auto* slots = reinterpret_cast<Slot*>(my_queue_.base + kHeaderBytes);
      for (uint32_t s = 0; s < kSlots; ++s) new (&slots[s].seq) std::atomic<uint64_t>(s);
      std::atomic_thread_fence(std::memory_order_seq_cst);
      q->magic = kQueueMagic;      // last: a peer that sees the magic sees an initialised queue
      head_ = 0;
      closed_ = false;
      return port;
    }
    // the name exists: a live van (pick another identity) or the leftover of a crashed job (reclaim it)
    Mapping old;
    if (MapObject(name, kHeaderBytes, false, &old)) {
      auto* q = reinterpret_cast<QueueHeader*>(old.base);
      const uint32_t pid = q->owner_pid;
      const bool dead = q->magic == kQueueMagic && pid != 0 && kill((pid_t)pid, 0) != 0 && errno == ESRCH;
      Unmap(&old, dead);
      if (dead) continue;          // retry the same port
    }
    if (fixed) {
      BPS_LOG(ERROR) << "shm van: " << name << " is owned by a running process";
      return -1;
    }
    port = 10000 + (int)(rng() % 40000);
  }
  return -1;
}

void ShmVan::Connect(const Node& node) {
  BPS_CHECK_NE(node.id, kEmpty);
  BPS_CHECK_NE(node.port, 0);
  auto p = std::make_shared<Peer>();
  p->port = node.port;
  const size_t qlen = kHeaderBytes + (size_t)kSlots * kSlotBytes;
  auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(env_int("BYTEPS_CONNECT_TIMEOUT_S", 60));
  bool ok = false;
  while (std::chrono::steady_clock::now() < deadline && !closed_) {
    if (MapObject(queue_name(node.port), qlen, false, &p->queue)) {
      if (reinterpret_cast<QueueHeader*>(p->queue.base)->magic == kQueueMagic) {
        ok = true;
        break;
      }
      Unmap(&p->queue, false);      // created but not initialised yet
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
  if (!ok) {
    BPS_LOG(ERROR) << "shm van: cannot attach to the queue of " << node.debug();
    return;
  }
  static std::atomic<uint32_t> nonce{0};
  const std::string aname = "/bps_shmvan_" + std::to_string(my_node_.port) + "_to_" + std::to_string(node.port) +
                            "_" + std::to_string((uint32_t)getpid()) + "_" + std::to_string(nonce++);
  if (!MapObject(aname, kHeaderBytes + arena_bytes_, true, &p->arena)) {
    BPS_LOG(ERROR) << "shm van: cannot create arena " << aname;
    Unmap(&p->queue, false);
    return;
  }
  auto* a = reinterpret_cast<ArenaHeader*>(p->arena.base);
  a->size = arena_bytes_;
  new (&a->consumed) std::atomic<uint64_t>(0);
  a->magic = kArenaMagic;
  std::shared_ptr<Peer> old;
  {
    std::lock_guard<std::mutex> g(peers_mu_);
    auto it = peers_.find(node.id);
    if (it != peers_.end()) old = it->second;
    peers_[node.id] = p;
  }
  if (old) {          // reconnect (recovery): the previous arena is no longer written
    std::lock_guard<std::mutex> g(old->mu);
    Unmap(&old->arena, true);
    Unmap(&old->queue, false);
  }
}

// Ring allocation of n bytes in my arena towards p (caller holds p->mu).  Blocks while the receiver has not
// released enough space; returns nullptr when the van is closing.
char* ShmVan::ArenaAlloc(Peer* p, size_t n, uint64_t* off) {
  auto* a = reinterpret_cast<ArenaHeader*>(p->arena.base);
  const uint64_t size = a->size;
  n = (n + kAlign - 1) / kAlign * kAlign;
  uint64_t head = p->arena_head;
  if (head % size + n > size) head += size - head % size;     // never straddle the end
  int spins = 0;
  while (head + n - a->consumed.load(std::memory_order_acquire) > size) {
    if (closed_) return nullptr;
    if (++spins < 2000) cpu_relax();
    else std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  *off = head % size;
  p->arena_head = head + n;
  return p->arena.base + kHeaderBytes + *off;
}

void ShmVan::Ring(QueueHeader* q) {
  q->doorbell.fetch_add(1, std::memory_order_seq_cst);
  if (q->sleeping.load(std::memory_order_seq_cst)) futex(&q->doorbell, FUTEX_WAKE, 1, nullptr);
}

int ShmVan::SendMsg(Message& msg) {
  const int id = msg.meta.recver;
  BPS_CHECK_NE(id, kEmpty);
  std::shared_ptr<Peer> p;
  {
    std::lock_guard<std::mutex> g(peers_mu_);
    auto it = peers_.find(id);
    if (it == peers_.end()) {
      if (!closed_) BPS_LOG(WARNING) << "shm van: not connected to node " << id;
      return -1;
    }
    p = it->second;
  }
  BPS_CHECK_LT(msg.data.size(), (size_t)kMaxBlobs) << "too many data blobs in one message";
  // A push request whose payload sits in a registered window travels by reference (the worker keeps the window
  // untouched until the matching pull has returned - same contract as the TCP van's colocated IPC path).
  Meta& m = msg.meta;
  bool by_ref = false;
  std::string ref_name;
  uint64_t ref_off = 0;
  if (po_->cfg().enable_ipc && m.push && m.request && msg.data.size() == 1 &&
      ShmRegistry::get().lookup(msg.data[0].data(), msg.data[0].size(), &ref_name, &ref_off)) {
    m.shm_name = ref_name;
    m.shm_offset = ref_off;
    m.shm_len = msg.data[0].size();
    by_ref = true;
  }
  const std::string meta = meta_pack(m);
  size_t total = meta.size();

  std::lock_guard<std::mutex> g(p->mu);
  if (!p->queue.base) return -1;
  auto* q = reinterpret_cast<QueueHeader*>(p->queue.base);
  auto* slots = reinterpret_cast<Slot*>(p->queue.base + kHeaderBytes);
  // ---- claim a slot (bounded MPMC ring, one CAS per message)
  uint64_t pos = q->tail.load(std::memory_order_relaxed);
  Slot* s = nullptr;
  int spins = 0;
  while (true) {
    s = &slots[pos & (kSlots - 1)];
    const uint64_t seq = s->seq.load(std::memory_order_acquire);
    const int64_t dif = (int64_t)(seq - pos);
    if (dif == 0) {
      if (q->tail.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) break;
    } else if (dif < 0) {          // ring full: the receiver is behind
      if (closed_) return -1;
      if (++spins < 2000) cpu_relax();
      else std::this_thread::sleep_for(std::chrono::microseconds(50));
      pos = q->tail.load(std::memory_order_relaxed);
    } else {
      pos = q->tail.load(std::memory_order_relaxed);
    }
  }
  // ---- fill it
  size_t inl = 0;      // bytes of s->bytes in use
  auto place = [&](Blob* b, const char* src, size_t len) -> bool {
    memset(b, 0, sizeof(*b));
    b->len = len;
    if (inl + len <= sizeof(s->bytes)) {
      b->kind = kInline;
      b->off = inl;
      if (len) memcpy(s->bytes + inl, src, len);
      inl += (len + 7) & ~size_t(7);
      return true;
    }
    auto* a = reinterpret_cast<ArenaHeader*>(p->arena.base);
    if (len <= a->size / 2) {
      uint64_t off;
      char* dst = ArenaAlloc(p.get(), len, &off);
      if (!dst) return false;
      memcpy(dst, src, len);
      b->kind = kArena;
      b->off = off;
      b->advance = p->arena_head;
      b->object_len = p->arena.len;
      snprintf(b->name, sizeof(b->name), "%s", p->arena.name.c_str());
      return true;
    }
    // oversized: a segment of its own, unlinked by the receiver
    static std::atomic<uint64_t> big{0};
    const std::string name = "/bps_shmvan_big_" + std::to_string((uint32_t)getpid()) + "_" + std::to_string(big++);
    Mapping seg;
    if (!MapObject(name, len, true, &seg)) return false;
    memcpy(seg.base, src, len);
    Unmap(&seg, false);
    b->kind = kOneOff;
    snprintf(b->name, sizeof(b->name), "%s", name.c_str());
    return true;
  };
  bool ok = place(&s->blobs[0], meta.data(), meta.size());
  uint32_t nb = 1;
  for (auto& d : msg.data) {
    if (!ok) break;
    Blob* b = &s->blobs[nb++];
    if (by_ref) {
      memset(b, 0, sizeof(*b));
      b->kind = kRegistered;      // described by meta.shm_*; nothing to copy
      b->len = d.size();
    } else {
      ok = place(b, d.data(), d.size());
    }
    total += d.size();
  }
  s->nblobs = ok ? nb : 0;        // 0: the receiver skips a slot that could not be filled
  s->sender_port = (uint32_t)my_node_.port;
  hb_release();
  s->seq.store(pos + 1, std::memory_order_release);
  Ring(q);
  return ok ? (int)std::min<size_t>(total, 0x7fffffff) : -1;
}

ShmVan::Mapping* ShmVan::PeerArena(const std::string& name, size_t len) {
  std::lock_guard<std::mutex> g(arenas_mu_);
  auto it = peer_arenas_.find(name);
  if (it != peer_arenas_.end()) return &it->second;
  Mapping m;
  if (!MapObject(name, len, false, &m)) return nullptr;
  return &peer_arenas_.emplace(name, m).first->second;
}

int ShmVan::RecvMsg(Message* msg) {
  auto* q = reinterpret_cast<QueueHeader*>(my_queue_.base);
  auto* slots = reinterpret_cast<Slot*>(my_queue_.base + kHeaderBytes);
  while (true) {
    if (!my_queue_.base) return -1;
    Slot* s = &slots[head_ & (kSlots - 1)];
    // ---- wait for the slot at head_ to be published: poll briefly, then park on the doorbell
    int spins = 0;
    while (s->seq.load(std::memory_order_acquire) != head_ + 1) {
      if (closed_) return -1;
      if (++spins < 4000) {
        cpu_relax();
        continue;
      }
      const uint32_t bell = q->doorbell.load(std::memory_order_seq_cst);
      q->sleeping.store(1, std::memory_order_seq_cst);
      if (s->seq.load(std::memory_order_acquire) != head_ + 1 && !closed_) {
        timespec ts{0, 100 * 1000 * 1000};
        futex(&q->doorbell, FUTEX_WAIT, bell, &ts);
      }
      q->sleeping.store(0, std::memory_order_seq_cst);
      spins = 0;
    }
    hb_acquire();
    // ---- decode
    bool ok = s->nblobs >= 1 && s->nblobs <= kMaxBlobs;
    Message out;
    size_t total = 0;
    for (uint32_t i = 0; ok && i < s->nblobs; ++i) {
      const Blob& b = s->blobs[i];
      const char* src = nullptr;
      Mapping oneoff;
      Mapping* arena = nullptr;
      if (b.kind == kInline) {
        ok = b.off + b.len <= sizeof(s->bytes);
        src = s->bytes + b.off;
      } else if (b.kind == kArena) {
        arena = PeerArena(std::string(b.name, strnlen(b.name, sizeof(b.name))), (size_t)b.object_len);
        ok = arena != nullptr && kHeaderBytes + b.off + b.len <= arena->len;
        if (ok) src = arena->base + kHeaderBytes + b.off;
      } else if (b.kind == kOneOff) {
        ok = MapObject(std::string(b.name, strnlen(b.name, sizeof(b.name))), (size_t)b.len, false, &oneoff);
        if (ok) src = oneoff.base;
      } else if (b.kind != kRegistered) {
        ok = false;
      }
      if (!ok) break;
      if (i == 0) {
        ok = b.kind != kRegistered && meta_unpack(src, (size_t)b.len, &out.meta);
      } else if (b.kind == kRegistered) {
        // zero-copy view of the sender's registered window (meta.shm_* names it)
        void* base = ShmRegistry::get().open(out.meta.shm_name, (size_t)(out.meta.shm_offset + out.meta.shm_len));
        ok = base != nullptr && out.meta.shm_len == b.len;
        if (ok) out.data.push_back(SArray<char>((char*)base + out.meta.shm_offset, (size_t)b.len, false));
      } else {
        SArray<char> a;
        char* direct = nullptr;
        size_t direct_len = 0;
        if (i == 1 && s->nblobs == 2 && TakeRecvBuffer(out.meta, &direct, &direct_len) && direct_len == b.len) {
          a = SArray<char>(direct, (size_t)b.len);      // the requester's own buffer (ExpectPullResponse)
        } else {
          a = PayloadPool::get().alloc((size_t)b.len);
        }
        if (b.len) memcpy(a.data(), src, (size_t)b.len);
        a.src_dev = out.meta.src_dev; a.src_id = out.meta.src_id;
        a.dst_dev = out.meta.dst_dev; a.dst_id = out.meta.dst_id;
        out.data.push_back(a);
      }
      if (arena)      // give the space back to the sender
        reinterpret_cast<ArenaHeader*>(arena->base)->consumed.store(b.advance, std::memory_order_release);
      if (oneoff.base) Unmap(&oneoff, true);
      total += (size_t)b.len;
    }
    const bool skipped = s->nblobs == 0;
    s->seq.store(head_ + kSlots, std::memory_order_release);      // free the slot for the next lap
    ++head_;
    if (skipped) continue;
    if (!ok) {
      BPS_LOG(ERROR) << "shm van: dropping an undecodable message";
      continue;
    }
    *msg = std::move(out);
    return (int)std::min<size_t>(total + 64, 0x7fffffff);
  }
}

void ShmVan::StopTransport() {
  closed_ = true;
  if (my_queue_.base) {
    auto* q = reinterpret_cast<QueueHeader*>(my_queue_.base);
    q->doorbell.fetch_add(1);
    futex(&q->doorbell, FUTEX_WAKE, 8, nullptr);
  }
  std::unordered_map<int, std::shared_ptr<Peer>> peers;
  {
    std::lock_guard<std::mutex> g(peers_mu_);
    peers.swap(peers_);
  }
  for (auto& kv : peers) {
    std::lock_guard<std::mutex> g(kv.second->mu);
    Unmap(&kv.second->arena, true);
    Unmap(&kv.second->queue, false);
  }
  {
    std::lock_guard<std::mutex> g(arenas_mu_);
    for (auto& kv : peer_arenas_) Unmap(&kv.second, false);
    peer_arenas_.clear();
  }
  // Van::Stop() joins the receiving thread before it calls this, so nobody polls the queue any more
  Unmap(&my_queue_, true);
}

}  // namespace net
}  // namespace bps
