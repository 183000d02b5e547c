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
  // Same draw, compared on its 53 most significant bits (one signed int -> double conversion instead of the
  // unsigned 64-bit sequence): differs from bernoulli() only when the draw lies within 2^-53 of the threshold.
  bool bernoulli53(double p) { return double(int64_t(next() >> 11)) < p * 9007199254740992.0; }

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
  // the `nb` low bits of v, most significant first (same stream as nb calls of put)
  void put_bits(uint64_t v, int nb) {
    while (nb > 0) {
      const int room = 32 - (int)used_;
      const int take = nb < room ? nb : room;
      const uint32_t chunk = (uint32_t)((v >> (nb - take)) & ((take == 32) ? 0xffffffffull : ((1ull << take) - 1)));
      acc_ = (take == 32) ? chunk : ((acc_ << take) | chunk);
      used_ += take;
      nb -= take;
      if (used_ == 32) {
        d_[blocks_++] = acc_;
        used_ = 0;
        acc_ = 0;
      }
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
  // the next nb bits as a number, first bit most significant (nb <= 64)
  uint64_t get_bits(int nb) {
    uint64_t v = 0;
    while (nb > 0) {
      if (used_ == 0) {
        acc_ = d_[blocks_++];
        used_ = 32;
      }
      const int take = nb < (int)used_ ? nb : (int)used_;
      const uint32_t chunk = (take == 32) ? acc_ : ((acc_ >> (used_ - take)) & ((1u << take) - 1));
      v = (take == 64) ? chunk : ((v << take) | chunk);
      used_ -= take;
      nb -= take;
    }
    return v;
  }
  size_t bits() const { return blocks_ * 32 - used_; }

 private:
  const uint32_t* d_;
  uint32_t acc_ = 0;
  size_t used_ = 0, blocks_ = 0;
};

// Elias-delta code of x >= 1: floor(log2(len)) zeros, len = 1 + floor(log2 x) in binary, then x without its leading 1.
inline void elias_delta_encode(BitWriter& w, unsigned long x) {
  const int len = 64 - __builtin_clzl(x);
  const int lol = 31 - __builtin_clz((unsigned)len);
  w.put_bits((uint64_t)len, 2 * lol + 1);      // len < 2^(lol+1): the padding is exactly the lol leading zeros
  if (len > 1) w.put_bits(x & ((len - 1 >= 64) ? ~0ull : ((1ull << (len - 1)) - 1)), len - 1);
}

inline unsigned long elias_delta_decode(BitReader& r) {
  int lol = 0;
  while (!r.get()) ++lol;
  const int len = (int)((1ull << lol) | r.get_bits(lol));
  if (len <= 1) return 1;
  return (unsigned long)((len - 1 >= 64 ? 0ull : (1ull << (len - 1))) | r.get_bits(len - 1));
}

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
  // dst += decompress(src): what a server does with every push after the first of a round.  The result equals
  // decompressing into a scratch buffer and adding it element-wise; sparse payloads touch only their k entries.
  virtual void decompress_add(const void* src, size_t csize, void* dst);
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
