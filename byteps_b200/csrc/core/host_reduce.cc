#include "core/host_reduce.h"

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>

#include "core/log.h"
#include "net/van.h"

namespace bps {

static size_t page_up(size_t n) { return (n + 4095) / 4096 * 4096; }

HostLocalReduce::HostLocalReduce(int local_rank, int local_size, const std::string& tag, int reducer_threads,
                                 const std::string& socket_dir)
    : rank_(local_rank), size_(local_size), tag_(tag), reducer_(reducer_threads) {
  BPS_CHECK_GT(local_size, 0);
  std::vector<int> members;
  for (int r = 0; r < local_size; ++r) members.push_back(r);
  comm_.reset(new LocalComm(local_rank, members, socket_dir, "hr" + tag, /*start_listening=*/false));
  // the root has to hear from every OTHER local rank
  reduce_ready_ = std::make_shared<ReadyTable>(local_size - 1, "HOST_REDUCE");
  shard_ready_ = std::make_shared<ReadyTable>(local_size - 1, "HOST_SHARD");
  bcast_ready_ = std::make_shared<ReadyTable>(local_size - 1, "HOST_BCAST");
  if (comm_->is_root()) {
    comm_->set_tables(reduce_ready_.get(), shard_ready_.get(), bcast_ready_.get(), nullptr);
    comm_->start();       // datagrams that arrived meanwhile are queued in the socket: none is lost
  } else {
    follower_ = std::thread([this] { follower_loop(); });
  }
}

HostLocalReduce::~HostLocalReduce() {
  {
    std::lock_guard<std::mutex> g(bmu_);
    stop_ = true;
  }
  bcv_.notify_all();
  if (follower_.joinable()) follower_.join();
  comm_.reset();
  std::lock_guard<std::mutex> g(mu_);
  for (auto& kv : regions_) net::ShmRegistry::get().release(kv.second.name);
}

uint64_t HostLocalReduce::signals_received() const { return comm_->received(); }

void HostLocalReduce::follower_loop() {
  while (true) {
    {
      std::lock_guard<std::mutex> g(bmu_);
      if (stop_) return;
    }
    LocalMsg m;
    if (!comm_->recv_from_root(&m, 200)) continue;      // 200 ms receive timeout: notices stop_
    if (m.signal != SIG_DO_BROADCAST && m.signal != SIG_DO_REDUCE) continue;
    {
      std::lock_guard<std::mutex> g(bmu_);
      ++(m.signal == SIG_DO_BROADCAST ? announced_ : reduce_asked_)[m.key];
    }
    bcv_.notify_all();
  }
}

HostLocalReduce::Region* HostLocalReduce::region_of(uint64_t key, size_t nbytes, int64_t timeout_ms) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(key);
  if (it != regions_.end()) {
    BPS_CHECK_GE(it->second.slot, nbytes) << "host reduce: key " << key << " was sized " << it->second.slot
                                          << " bytes per rank, now " << nbytes;
    return &it->second;
  }
  Region r;
  r.slot = page_up(nbytes ? nbytes : 1);
  r.name = "BytePS_ShM_box" + tag_ + "_" + std::to_string(key);
  const size_t total = r.slot * (size_t)(size_ + 1);
  if (comm_->is_root()) {
    r.base = (char*)net::ShmRegistry::get().create(r.name, total);
  } else {
    // the root creates the object the first time it sees the key: wait for it (and for its final size)
    const auto t0 = std::chrono::steady_clock::now();
    while (!(r.base = (char*)net::ShmRegistry::get().open(r.name, total))) {
      if (timeout_ms >= 0 &&
          std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_ms)
        return nullptr;
      usleep(500);
    }
  }
  if (!r.base) return nullptr;
  auto ins = regions_.emplace(key, r);
  return &ins.first->second;
}

bool HostLocalReduce::contribute(uint64_t key, const void* src, size_t nbytes, int64_t timeout_ms) {
  Region* r = region_of(key, nbytes, timeout_ms);
  if (!r) return false;
  reducer_.copy(r->base + (size_t)rank_ * r->slot, src, nbytes);
  if (!comm_->is_root()) return comm_->send_to_root(SIG_REDUCE_READY, key);
  return true;
}

void* HostLocalReduce::window(uint64_t key) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = regions_.find(key);
  return it == regions_.end() ? nullptr : it->second.base + (size_t)size_ * it->second.slot;
}

// window[shard of this rank] = alpha * sum over the slots; shards are cut at 4 KB so they never share a line
void HostLocalReduce::sum_my_shard(Region* r, size_t nbytes, int dtype, double alpha) {
  const size_t per = ((nbytes + (size_t)size_ - 1) / (size_t)size_ + 4095) / 4096 * 4096;
  const size_t lo = std::min(nbytes, per * (size_t)rank_), hi = std::min(nbytes, lo + per);
  if (hi <= lo) return;
  const size_t n = hi - lo;
  char* win = r->base + (size_t)size_ * r->slot + lo;
  if (size_ == 1) {
    reducer_.copy(win, r->base + lo, n);
  } else {
    reducer_.sum(win, r->base + lo, r->base + r->slot + lo, n, dtype);      // slot0 + slot1 in one pass
    for (int s = 2; s < size_; ++s) reducer_.sum(win, r->base + (size_t)s * r->slot + lo, n, dtype);
  }
  if (alpha != 1.0) reducer_.scale(win, n, dtype, alpha);
}

void* HostLocalReduce::reduce(uint64_t key, size_t nbytes, int dtype, int64_t timeout_ms, double alpha) {
  BPS_CHECK(comm_->is_root()) << "reduce() is the root's stage";
  Region* r = region_of(key, nbytes, timeout_ms);
  if (!r) return nullptr;
  if (size_ > 1) {
    if (!reduce_ready_->wait_ready(key, timeout_ms)) return nullptr;
    reduce_ready_->clear_ready_count(key);
    // every copy is in its slot: all ranks sum their shard of the window at once
    if (!comm_->broadcast(SIG_DO_REDUCE, key)) return nullptr;
  }
  sum_my_shard(r, nbytes, dtype, alpha);
  if (size_ > 1) {
    if (!shard_ready_->wait_ready(key, timeout_ms)) return nullptr;
    shard_ready_->clear_ready_count(key);
  }
  return r->base + (size_t)size_ * r->slot;
}

bool HostLocalReduce::publish(uint64_t key, void* dst, size_t nbytes, int64_t timeout_ms) {
  BPS_CHECK(comm_->is_root()) << "publish() is the root's stage";
  void* win = window(key);
  if (!win) return false;
  bool ok = size_ == 1 || comm_->broadcast(SIG_DO_BROADCAST, key);
  if (dst && dst != win) reducer_.copy(dst, win, nbytes);
  if (size_ > 1) {
    // the window is rewritten by the next round's reduce(): every follower must have copied it out first
    if (!bcast_ready_->wait_ready(key, timeout_ms)) return false;
    bcast_ready_->clear_ready_count(key);
  }
  return ok;
}

bool HostLocalReduce::wait_announced(std::unordered_map<uint64_t, int>& m, uint64_t key, int64_t timeout_ms) {
  std::unique_lock<std::mutex> lk(bmu_);
  auto have = [&] {
    auto it = m.find(key);
    return stop_ || (it != m.end() && it->second > 0);
  };
  if (timeout_ms < 0) bcv_.wait(lk, have);
  else if (!bcv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), have)) return false;
  if (stop_) return false;
  if (--m[key] == 0) m.erase(key);
  return true;
}

bool HostLocalReduce::collect(uint64_t key, void* dst, size_t nbytes, int64_t timeout_ms, int dtype, double alpha) {
  BPS_CHECK(!comm_->is_root()) << "collect() is a follower's stage";
  // 1. the root has seen every copy: sum my shard of the window and say so
  if (!wait_announced(reduce_asked_, key, timeout_ms)) return false;
  Region* r = region_of(key, nbytes, timeout_ms);
  if (!r) return false;
  sum_my_shard(r, nbytes, dtype, alpha);
  if (!comm_->send_to_root(SIG_PCIE_REDUCE_READY, key)) return false;
  // 2. the window holds the final result (after the server round trip / the exchange between hosts)
  if (!wait_announced(announced_, key, timeout_ms)) return false;
  reducer_.copy(dst, r->base + (size_t)size_ * r->slot, nbytes);
  return comm_->send_to_root(SIG_BCAST_READY, key);
}

}  // namespace bps
