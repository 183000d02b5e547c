// Reliable delivery: ACK + timeout retransmit + duplicate suppression.
// Parity: /root/reference/3rdparty/ps-lite/src/resender.h:15-141 (PS_RESEND,
// PS_RESEND_TIMEOUT, at most 10 retries).
#pragma once
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "core/log.h"
#include "net/message.h"

namespace bps {
namespace net {

class Van;

class Resender {
 public:
  Resender(int timeout_ms, int max_retry, Van* van) : timeout_ms_(timeout_ms), max_retry_(max_retry), van_(van) {
    monitor_ = std::thread([this] { Monitoring(); });
  }
  ~Resender() {
    exit_ = true;
    if (monitor_.joinable()) monitor_.join();
  }

  static uint64_t Signature(const Meta& m) {
    // (app, sender, recver, timestamp, request, push, pull) -> 64 bits
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) {
      h ^= v;
      h *= 1099511628211ull;
    };
    mix((uint64_t)(uint32_t)m.app_id);
    mix((uint64_t)(uint32_t)m.sender);
    mix((uint64_t)(uint32_t)m.recver);
    mix((uint64_t)(uint32_t)m.timestamp);
    mix((uint64_t)((m.request ? 1 : 0) | (m.push ? 2 : 0) | (m.pull ? 4 : 0) | (m.simple_app ? 8 : 0)));
    mix((uint64_t)m.control.cmd);
    return h ? h : 1;
  }

  // remember an outgoing message until it is ACKed
  void AddOutgoing(const Message& msg) {
    if (msg.meta.control.cmd == Control::ACK) return;
    std::lock_guard<std::mutex> g(mu_);
    uint64_t sig = msg.meta.msg_sig;
    if (send_buff_.count(sig)) return;
    Entry e;
    // Deep copy of the payload: outgoing data is usually a zero-copy view of the caller's buffer, and the
    // caller may release it as soon as the RESPONSE arrives - while the request's ACK can still be lost
    // and a retransmission pending (found by AddressSanitizer: heap-use-after-free in writev).
    e.msg.meta = msg.meta;
    for (const auto& d : msg.data) {
      SArray<char> c;
      c.copy_from(d.data(), d.size());
      c.src_dev = d.src_dev; c.src_id = d.src_id; c.dst_dev = d.dst_dev; c.dst_id = d.dst_id;
      e.msg.data.push_back(c);
    }
    e.send = Now();
    e.num_retry = 0;
    send_buff_[sig] = e;
  }

  // returns true if the message must be dropped (ACK consumed or duplicate)
  bool AddIncoming(const Message& msg);

  // Block until every outgoing message was ACKed (or given up), at most max_ms.  A node must not
  // tear its van down while a peer may still need a retransmission from it - e.g. the scheduler's
  // barrier release that fault injection dropped at the receiver.
  bool Drain(int64_t max_ms) {
    const int64_t deadline = Now() + max_ms;
    while (Now() < deadline) {
      {
        std::lock_guard<std::mutex> g(mu_);
        if (send_buff_.empty()) return true;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
    return false;
  }
  int timeout_ms() const { return timeout_ms_; }

 private:
  struct Entry {
    Message msg;
    int64_t send;
    int num_retry;
  };
  static int64_t Now() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
  }
  void Monitoring();

  int timeout_ms_, max_retry_;
  Van* van_;
  std::mutex mu_;
  std::unordered_map<uint64_t, Entry> send_buff_;
  std::unordered_set<uint64_t> acked_;
  std::atomic<bool> exit_{false};
  std::thread monitor_;
};

}  // namespace net
}  // namespace bps
