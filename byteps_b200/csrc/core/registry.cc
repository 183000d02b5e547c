#include "core/registry.h"

#include <cstdlib>

#include <functional>

#include "core/log.h"

namespace bps {

uint32_t Registry::declare(const std::string& name) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = by_name_.find(name);
  if (it != by_name_.end()) return it->second->declared_key;
  BPS_CHECK_LT(order_.size(), (size_t)65536) << "at most 65536 tensors can be declared";
  auto ctx = std::make_shared<TensorContext>();
  ctx->name = name;
  ctx->declared_key = static_cast<uint32_t>(order_.size());
  order_.push_back(name);
  by_name_[name] = ctx;
  BPS_LOG(DEBUG) << "declared tensor " << name << " -> key " << ctx->declared_key;
  return ctx->declared_key;
}

bool Registry::is_declared(const std::string& name) const {
  std::lock_guard<std::mutex> g(mu_);
  return by_name_.count(name) != 0;
}

std::shared_ptr<TensorContext> Registry::context(const std::string& name) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = by_name_.find(name);
  return it == by_name_.end() ? nullptr : it->second;
}

std::shared_ptr<TensorContext> Registry::context_by_key(uint32_t k) {
  std::lock_guard<std::mutex> g(mu_);
  if (k >= order_.size()) return nullptr;
  return by_name_[order_[k]];
}

bool Registry::init_tensor(const std::shared_ptr<TensorContext>& ctx, size_t nbytes, int dtype, size_t bound,
                           size_t page) {
  std::lock_guard<std::mutex> g(mu_);
  if (ctx->initialized && ctx->nbytes == nbytes && ctx->dtype == dtype) return false;
  ctx->nbytes = nbytes;
  ctx->dtype = dtype;
  ctx->aligned_bytes = round_up(nbytes ? nbytes : 1, page ? page : 4096);
  ctx->parts = partition_bytes(nbytes, bound);
  BPS_CHECK_LE(ctx->parts.size(), (size_t)65536) << "tensor " << ctx->name << " has too many partitions";
  ctx->keys.clear();
  for (size_t i = 0; i < ctx->parts.size(); ++i) ctx->keys.push_back(make_key(ctx->declared_key, (uint32_t)i));
  ctx->compressors.clear();
  ctx->initialized = true;
  return true;
}

std::vector<std::string> Registry::declared_names() const {
  std::lock_guard<std::mutex> g(mu_);
  return order_;
}

size_t Registry::size() const {
  std::lock_guard<std::mutex> g(mu_);
  return order_.size();
}

void Registry::reset_contexts() {
  std::lock_guard<std::mutex> g(mu_);
  for (auto& kv : by_name_) {
    auto& c = kv.second;
    c->initialized = false;
    c->keys.clear();
    c->parts.clear();
    c->compressors.clear();
    c->host_buff = nullptr;
    c->shm_name.clear();
    c->step_cnt = 0;
  }
}

void Registry::set_kwargs(const std::string& name, const std::unordered_map<std::string, std::string>& kw) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = by_name_.find(name);
  if (it != by_name_.end()) it->second->kwargs = kw;
}

uint64_t hash_naive(uint64_t key) { return ((key >> 16) + (key % 65536)) * 9973; }
uint64_t hash_builtin(uint64_t key) {
  // BYTEPS_BUILT_IN_HASH_COEF multiplies the hash (global.cc:169-172,603) to spread keys differently
  static const uint64_t coef = [] {
    const char* e = getenv("BYTEPS_BUILT_IN_HASH_COEF");
    long v = e ? atol(e) : 1;
    return (uint64_t)(v > 0 ? v : 1);
  }();
  return std::hash<std::string>()(std::to_string(key)) * coef;
}
uint64_t hash_djb2(uint64_t key) {
  std::string s = std::to_string(key);
  uint64_t h = 5381;
  for (unsigned char c : s) h = ((h << 5) + h) + c;
  return h;
}
uint64_t hash_sdbm(uint64_t key) {
  std::string s = std::to_string(key);
  uint64_t h = 0;
  for (unsigned char c : s) h = c + (h << 6) + (h << 16) - h;
  return h;
}

KeyPlacer::KeyPlacer(const std::string& fn, int num_servers, int num_workers, bool mixed_mode, int mixed_bound)
    : fn_(fn),
      num_servers_(num_servers > 0 ? num_servers : 1),
      num_workers_(num_workers),
      mixed_mode_(mixed_mode),
      mixed_bound_(mixed_bound),
      load_(num_servers_, 0) {}

// Mixed mode: some servers are colocated with workers; bias placement so
// network load balances (global.cc:566-596).
int KeyPlacer::mixed(uint64_t key) const {
  int noncol = num_servers_ - num_workers_;
  int col = num_workers_;
  BPS_CHECK_GE(mixed_bound_, num_servers_);
  BPS_CHECK_GT(noncol, 0) << "mixed mode needs non-colocated servers";
  double ratio = (2.0 * noncol * (num_workers_ - 1)) /
                 ((double)num_workers_ * (num_workers_ + noncol) - 2.0 * noncol);
  BPS_CHECK_LE(ratio, 1.0) << "mixed mode: more non-colocated servers than workers";
  double threshold = ratio * mixed_bound_;
  uint64_t r = hash_djb2(key) % mixed_bound_;
  if ((double)r < threshold) return (int)(hash_djb2(r) % noncol);
  return noncol + (int)(hash_djb2(r) % col);
}

int KeyPlacer::server_of(uint64_t key, size_t len) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = memo_.find(key);
  if (it != memo_.end()) return it->second;
  int s = 0;
  if (fn_ == "naive") s = hash_naive(key) % num_servers_;
  else if (fn_ == "built_in") s = hash_builtin(key) % num_servers_;
  else if (fn_ == "djb2") s = hash_djb2(key) % num_servers_;
  else if (fn_ == "sdbm") s = hash_sdbm(key) % num_servers_;
  else if (fn_ == "mixed") {
    BPS_CHECK(mixed_mode_) << "BYTEPS_KEY_HASH_FN=mixed also needs BYTEPS_ENABLE_MIXED_MODE";
    s = mixed(key);
  } else {
    BPS_LOG_FATAL << "unsupported BYTEPS_KEY_HASH_FN '" << fn_ << "' (naive|built_in|djb2|sdbm|mixed)";
  }
  memo_[key] = s;
  load_[s] += len;
  return s;
}

}  // namespace bps
