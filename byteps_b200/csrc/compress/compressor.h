// Host-side gradient compressors (reference implementations + CPU-server mode).
//
// Parity: /root/reference/byteps/common/compressor/{compressor.h,compressor_registry.cc,
// utils.h,error_feedback.cc,momentum.cc} and impl/{onebit,topk,randomk,dithering,
// vanilla_error_feedback,nesterov_momentum}.cc.  Same algorithms, same kwargs
// names ("compressor_type", "compressor_k", "compressor_onebit_scaling",
// "ef_type", "momentum_type", "momentum_mu", "seed", "dithering_partition",
// "dithering_normalize") and the same xorshift128+ random stream, so results
// can be checked against an independent numpy model exactly like the
// reference's tests do.  Differences: explicit dst buffers (no hidden _buf
// aliasing rules), bf16 support, 32-bit indices/bit-words for every dtype
// (the reference's same-width index overflows for fp16 tensors > 65536
// elements), and the learning rate for error feedback is set through an API
// instead of an mmap'd "lr.s" file (the file is still honoured if present).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "core/types.h"

namespace bps {

using Kwargs = std::unordered_map<std::string, std::string>;

std::string kwargs_serialize(const Kwargs& kw);
Kwargs kwargs_deserialize(const std::string& s);

// xorshift128+ with the reference's seeding convention (state = {seed, seed}).
class XorShift128Plus {
 public:
  XorShift128Plus() {
    std::random_device rd;
    a_ = rd();
    b_ = rd();
  }
  void set_seed(uint64_t seed) { a_ = b_ = seed; }
  uint64_t next() {
    uint64_t t = a_;
    const uint64_t s = b_;
    a_ = s;
    t ^= t << 23;
    t ^= t >> 17;
    t ^= s ^ (s >> 26);
    b_ = t;
    return t + s;
  }
  uint64_t randint(uint64_t low, uint64_t high) { return next() % (high - low) + low; }
  double rand() { return double(next()) / double(kMax); }
  bool bernoulli(double p) { return double(next()) < p * double(kMax); }

 private:
  static constexpr uint64_t kMax = std::numeric_limits<uint64_t>::max();
  uint64_t a_, b_;
};

// MSB-first bit stream over 32-bit words.
class BitWriter {
 public:
  explicit BitWriter(uint32_t* d) : d_(d) {}
  void put(bool x) {
    acc_ = (acc_ << 1) | (x ? 1u : 0u);
    if (++used_ == 32) {
      d_[blocks_++] = acc_;
      used_ = 0;
      acc_ = 0;
    }
  }
  void flush() {
    if (used_ > 0) d_[blocks_] = acc_ << (32 - used_);
  }
  size_t bits() const { return blocks_ * 32 + used_; }
  size_t blocks() const { return (bits() + 31) / 32; }

 private:
  uint32_t* d_;
  uint32_t acc_ = 0;
  size_t used_ = 0, blocks_ = 0;
};

class BitReader {
 public:
  explicit BitReader(const uint32_t* d) : d_(d) {}
  bool get() {
    if (used_ == 0) {
      acc_ = d_[blocks_++];
      used_ = 32;
    }
    return (acc_ >> --used_) & 1u;
  }
  size_t bits() const { return blocks_ * 32 - used_; }

 private:
  const uint32_t* d_;
  uint32_t acc_ = 0;
  size_t used_ = 0, blocks_ = 0;
};

void elias_delta_encode(BitWriter& w, unsigned long x);
unsigned long elias_delta_decode(BitReader& r);
uint32_t round_next_pow2(uint32_t v);

class Compressor {
 public:
  Compressor(size_t nbytes, int dtype) : nbytes_(nbytes), dtype_(dtype) {}
  virtual ~Compressor() = default;
  // Upper bound of the compressed size for this tensor.
  virtual size_t max_compressed_bytes() const = 0;
  // grad may be modified (decorators add error/momentum in place).  Returns bytes written to dst.
  virtual size_t compress(void* grad, void* dst) = 0;
  // dst receives nbytes() of dtype(); src and dst must not alias.
  virtual void decompress(const void* src, size_t csize, void* dst) = 0;
  // error = corrected - decompress(compressed), fused.
  virtual void fast_update_error(void* error, const void* corrected, const void* compressed, size_t csize);
  virtual void set_lr(double) {}
  virtual const char* name() const = 0;
  size_t nbytes() const { return nbytes_; }
  int dtype() const { return dtype_; }
  size_t numel() const { return nbytes_ / dtype_size(dtype_); }

 protected:
  size_t nbytes_;
  int dtype_;
};

using CompressorCtor = std::function<std::unique_ptr<Compressor>(const Kwargs&, size_t, int, bool)>;

class CompressorRegistry {
 public:
  static void add(const std::string& name, CompressorCtor c);
  // Peels decorators in the reference's order: momentum -> error feedback ->
  // compressor; the server side skips momentum.  nullptr when kwargs name no compressor.
  static std::unique_ptr<Compressor> create(const Kwargs& kw, size_t nbytes, int dtype, bool server_side = false);
  static std::vector<std::string> names();
};

}  // namespace bps
