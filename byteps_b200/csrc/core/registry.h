// Tensor name registry, key allocation and key->server placement.
//
// Parity: declare/re-declare + name->context table in
// /root/reference/byteps/common/global.cc:412-436 and the key hashing family
// (naive / built_in / djb2 / sdbm / mixed) in global.cc:566-677.  Ours is an
// explicit object (no process-wide statics) so suspend/resume simply rebuilds
// the Engine while the Registry's declaration order survives.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "core/task.h"

namespace bps {

class Registry {
 public:
  // Returns the declared key; idempotent for a known name.
  uint32_t declare(const std::string& name);
  bool is_declared(const std::string& name) const;
  std::shared_ptr<TensorContext> context(const std::string& name);  // nullptr if unknown
  std::shared_ptr<TensorContext> context_by_key(uint32_t declared_key);
  // Fill partitioning for a tensor the first time its size is known.
  // Returns false if it was already initialised with the same size.
  bool init_tensor(const std::shared_ptr<TensorContext>& ctx, size_t nbytes, int dtype, size_t partition_bound,
                   size_t page);
  std::vector<std::string> declared_names() const;  // in declaration order
  size_t size() const;
  // Drop per-context state but keep names+order so keys stay stable
  // (suspend/resume, /root/reference/byteps/common/global.cc:431-436).
  void reset_contexts();
  void set_kwargs(const std::string& name, const std::unordered_map<std::string, std::string>& kw);

 private:
  mutable std::mutex mu_;
  std::vector<std::string> order_;
  std::unordered_map<std::string, std::shared_ptr<TensorContext>> by_name_;
};

// ---- key -> server placement -------------------------------------------------
uint64_t hash_naive(uint64_t key);
uint64_t hash_builtin(uint64_t key);
uint64_t hash_djb2(uint64_t key);
uint64_t hash_sdbm(uint64_t key);

class KeyPlacer {
 public:
  KeyPlacer(const std::string& fn, int num_servers, int num_workers, bool mixed_mode, int mixed_bound);
  int server_of(uint64_t key, size_t len);  // memoised; accumulates per-server load
  const std::vector<uint64_t>& load() const { return load_; }
  int num_servers() const { return num_servers_; }

 private:
  int mixed(uint64_t key) const;
  std::string fn_;
  int num_servers_, num_workers_;
  bool mixed_mode_;
  int mixed_bound_;
  std::mutex mu_;
  std::unordered_map<uint64_t, int> memo_;
  std::vector<uint64_t> load_;
};

}  // namespace bps
