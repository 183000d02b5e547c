// Wire structures of the KV transport.
//
// Parity: ps-lite's Node / Control / Meta / Message
// (/root/reference/3rdparty/ps-lite/include/ps/internal/message.h:74-329) and
// SArray (/root/reference/3rdparty/ps-lite/include/ps/sarray.h:46-352).  The
// encoding is our own compact little-endian layout (see meta_pack/meta_unpack)
// rather than ps-lite's RawMeta struct dump.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace bps {
namespace net {

// node-group bitmask used for barriers and broadcast ids
constexpr int kScheduler = 1;
constexpr int kServerGroup = 2;
constexpr int kWorkerGroup = 4;
constexpr int kEmpty = -1;

enum class Role : int { kServer = 0, kWorker = 1, kScheduler = 2 };

enum DeviceType : int { DEV_UNK = 0, DEV_CPU = 1, DEV_GPU = 2 };

// Reference-counted byte range; zero-copy over external memory when built
// with a no-op deleter.  src/dst device tags travel with the data like in the
// GPU-aware ps-lite revision.
template <typename T>
class SArray {
 public:
  SArray() = default;
  explicit SArray(size_t n) { resize(n); }
  SArray(T* data, size_t n, bool take_ownership = false) { reset(data, n, take_ownership); }
  template <typename Deleter>
  SArray(T* data, size_t n, Deleter d) : size_(n), ptr_(data, d) {}
  void reset(T* data, size_t n, bool take_ownership = false) {
    size_ = n;
    if (take_ownership) ptr_.reset(data, [](T* p) { delete[] p; });
    else ptr_.reset(data, [](T*) {});
  }
  void resize(size_t n) {
    T* p = new T[n ? n : 1];
    size_ = n;
    ptr_.reset(p, [](T* q) { delete[] q; });
  }
  void copy_from(const T* src, size_t n) {
    resize(n);
    if (n) memcpy(ptr_.get(), src, n * sizeof(T));
  }
  SArray<T> segment(size_t begin, size_t end) const {
    SArray<T> r;
    r.size_ = end - begin;
    r.ptr_ = std::shared_ptr<T>(ptr_, ptr_.get() + begin);
    r.src_dev = src_dev; r.src_id = src_id; r.dst_dev = dst_dev; r.dst_id = dst_id;
    return r;
  }
  T* data() const { return ptr_.get(); }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  T& operator[](size_t i) const { return ptr_.get()[i]; }
  int src_dev = DEV_UNK, src_id = -1, dst_dev = DEV_UNK, dst_id = -1;

 private:
  size_t size_ = 0;
  std::shared_ptr<T> ptr_;
};

struct Node {
  Role role = Role::kWorker;
  int id = kEmpty;
  int customer_id = 0;
  std::string hostname;
  int port = 0;
  bool is_recovery = false;
  int aux_id = -1;   // rank hint supplied by the node (stable ordering)
  std::string debug() const {
    std::ostringstream os;
    os << (role == Role::kServer ? "server" : role == Role::kWorker ? "worker" : "scheduler") << "[" << id << "]@"
       << hostname << ":" << port;
    return os.str();
  }
};

struct Control {
  enum Command : int { EMPTY = 0, TERMINATE, ADD_NODE, BARRIER, ACK, HEARTBEAT, INSTANCE_BARRIER };
  Command cmd = EMPTY;
  std::vector<Node> node;
  int barrier_group = 0;
  uint64_t msg_sig = 0;
  bool empty() const { return cmd == EMPTY; }
};

struct Meta {
  int head = kEmpty;
  int app_id = kEmpty;
  int customer_id = kEmpty;
  int timestamp = kEmpty;
  int sender = kEmpty;
  int recver = kEmpty;
  bool request = false;
  bool push = false;
  bool pull = false;
  bool simple_app = false;
  std::string body;
  Control control;
  uint64_t key = 0;
  int cmd = 0;            // user command (request type x dtype pairing)
  uint64_t val_len = 0;
  uint64_t msg_sig = 0;   // resender signature
  // colocated IPC: payload lives in a POSIX shm object instead of the socket
  std::string shm_name;
  uint64_t shm_offset = 0;
  uint64_t shm_len = 0;
  int src_dev = DEV_UNK, src_id = -1, dst_dev = DEV_UNK, dst_id = -1;
};

struct Message {
  Meta meta;
  std::vector<SArray<char>> data;
  void add_data(const SArray<char>& a) { data.push_back(a); }
  size_t data_bytes() const {
    size_t n = 0;
    for (auto& d : data) n += d.size();
    return n;
  }
};

// little-endian byte codec
std::string meta_pack(const Meta& m);
bool meta_unpack(const char* buf, size_t len, Meta* m);

}  // namespace net
}  // namespace bps
