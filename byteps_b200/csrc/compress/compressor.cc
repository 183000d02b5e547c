#include "compress/compressor.h"

// x86 SIMD fast paths (fp32) are optional: the generic typed loops below them do the same work (aarch64 hosts and
// x86 without AVX2 build and run; BPS_NO_X86_SIMD forces the generic build)
#if (defined(__x86_64__) || defined(__i386__)) && defined(__AVX2__) && !defined(BPS_NO_X86_SIMD)
#define BPS_X86_SIMD 1
#include <immintrin.h>
#else
#define BPS_X86_SIMD 0
#endif

#include <array>
#include <type_traits>

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>

#include "core/env.h"
#include "core/log.h"
#include "cpu/half.h"
#include "cpu/reducer.h"

namespace bps {

// ---------------------------------------------------------------- utilities
std::string kwargs_serialize(const Kwargs& kw) {
  std::map<std::string, std::string> sorted(kw.begin(), kw.end());
  std::ostringstream os;
  os << sorted.size();
  for (auto& p : sorted) os << " " << p.first << " " << p.second;
  return os.str();
}

Kwargs kwargs_deserialize(const std::string& s) {
  Kwargs kw;
  std::istringstream is(s);
  size_t n = 0;
  is >> n;
  for (size_t i = 0; i < n; ++i) {
    std::string k, v;
    is >> k >> v;
    kw[k] = v;
  }
  return kw;
}

static inline int ilog2(unsigned long x) { return 63 - __builtin_clzl(x); }

uint32_t round_next_pow2(uint32_t v) {
  v -= 1;
  v |= v >> 1;
  v |= v >> 2;
  v |= v >> 4;
  v |= v >> 8;
  v |= v >> 16;
  return v + 1;
}

template <typename T>
static T kw_get(const Kwargs& kw, const std::string& k, bool optional, T dflt = T()) {
  auto it = kw.find(k);
  if (it == kw.end()) {
    if (!optional) BPS_LOG_FATAL << "hyper-parameter '" << k << "' is required";
    return dflt;
  }
  std::istringstream ss(it->second);
  T v{};
  if (std::is_same<T, bool>::value) {
    std::string s = it->second;
    for (auto& c : s) c = tolower(c);
    return (T)(s == "true" || s == "1" || s == "yes");
  }
  ss >> v;
  return v;
}

// ---------------------------------------------------------------- typed access
struct TF32 {
  using S = float;
  static float ld(const S* p, size_t i) { return p[i]; }
  static void st(S* p, size_t i, float v) { p[i] = v; }
};
struct TF64 {
  using S = double;
  static double ld(const S* p, size_t i) { return p[i]; }
  static void st(S* p, size_t i, double v) { p[i] = v; }
};
struct TF16 {
  using S = uint16_t;
  static float ld(const S* p, size_t i) { return f16_to_f32(p[i]); }
  static void st(S* p, size_t i, float v) { p[i] = f32_to_f16(v); }
};
struct TBF16 {
  using S = uint16_t;
  static float ld(const S* p, size_t i) { return bf16_to_f32(p[i]); }
  static void st(S* p, size_t i, float v) { p[i] = f32_to_bf16(v); }
};

#define BPS_DISPATCH_FLOAT(dtype, FN, ...)                                      \
  switch (dtype) {                                                              \
    case F32: FN<TF32>(__VA_ARGS__); break;                                     \
    case F64: FN<TF64>(__VA_ARGS__); break;                                     \
    case F16: FN<TF16>(__VA_ARGS__); break;                                     \
    case BF16: FN<TBF16>(__VA_ARGS__); break;                                   \
    default: BPS_LOG_FATAL << "compressors need a floating dtype, got " << dtype_name(dtype); \
  }

void Compressor::decompress_add(const void* src, size_t csize, void* dst) {
  static thread_local std::vector<char> tmp;
  tmp.resize(nbytes_);
  decompress(src, csize, tmp.data());
  CpuReducer r(1);
  r.sum(dst, tmp.data(), nbytes_, dtype_);
}

void Compressor::fast_update_error(void* error, const void* corrected, const void* compressed, size_t csize) {
  // generic (unfused) fallback: error = corrected - D(compressed)
  std::vector<char> tmp(nbytes_);
  decompress(compressed, csize, tmp.data());
  CpuReducer r(1);
  r.sum_scaled(error, corrected, tmp.data(), nbytes_, dtype_, -1.0f);
}

// ---------------------------------------------------------------- onebit
// payload: [ceil(n/32) uint32 words, 1 = negative, MSB first][float scale]
class OnebitCompressor : public Compressor {
 public:
  OnebitCompressor(size_t nbytes, int dtype, bool scaled) : Compressor(nbytes, dtype), scaled_(scaled) {}
  const char* name() const override { return "onebit"; }
  size_t max_compressed_bytes() const override { return ((numel() + 31) / 32) * 4 + 4; }

  // No OpenMP in here: partitions are compressed concurrently by the worker's thread pool / the server's engine
  // threads, and a parallel region per 4 MB partition inside those threads oversubscribes the cores (measured:
  // 0.2 GB/s with 8 spinning OpenMP threads vs 2 GB/s with one, benchmarks/cpu_compress_bench.py).  The fp32 case is
  // hand-vectorised instead.
  template <typename A>
  void do_compress(const void* src_, uint32_t* dst, size_t n, size_t* out) {
    const typename A::S* src = (const typename A::S*)src_;
    const size_t chunks = (n + 31) / 32;
    float scale = 1.0f;
    if (BPS_X86_SIMD && std::is_same<A, TF32>::value) {
      compress_f32((const float*)src_, dst, n, &scale);
    } else {
      if (scaled_) {
        double sum = 0.0;
        for (size_t i = 0; i < n; ++i) sum += std::fabs((double)A::ld(src, i));
        scale = (float)(sum / (double)n);
      }
      for (size_t c = 0; c < chunks; ++c) {
        uint32_t x = 0;
        for (size_t j = 0; j < 32; ++j) {
          size_t i = c * 32 + j;
          x <<= 1;
          if (i < n) x |= (A::ld(src, i) < 0) ? 1u : 0u;
        }
        dst[c] = x;
      }
    }
    memcpy(&dst[chunks], &scale, 4);
    *out = chunks * 4 + 4;
  }

  static uint32_t bit_reverse32(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
  }

  // AVX2: |x| summed in four double lanes, signs taken with a compare (so -0.0 and NaN count as non-negative, like
  // `x < 0`) and packed 32 per word, element 0 in the most significant bit.
  #if BPS_X86_SIMD
  void compress_f32(const float* src, uint32_t* dst, size_t n, float* scale) {
    const size_t full = n / 32;
    const __m256 zero = _mm256_setzero_ps();
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    __m256d acc0 = _mm256_setzero_pd(), acc1 = _mm256_setzero_pd();
    for (size_t c = 0; c < full; ++c) {
      uint32_t lsb_first = 0;
      for (int g = 0; g < 4; ++g) {
        __m256 v = _mm256_loadu_ps(src + c * 32 + g * 8);
        lsb_first |= (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(v, zero, _CMP_LT_OQ)) << (8 * g);
        if (scaled_) {
          __m256 a = _mm256_and_ps(v, absmask);
          acc0 = _mm256_add_pd(acc0, _mm256_cvtps_pd(_mm256_castps256_ps128(a)));
          acc1 = _mm256_add_pd(acc1, _mm256_cvtps_pd(_mm256_extractf128_ps(a, 1)));
        }
      }
      dst[c] = bit_reverse32(lsb_first);
    }
    double lanes[4];
    _mm256_storeu_pd(lanes, _mm256_add_pd(acc0, acc1));
    double sum = (lanes[0] + lanes[1]) + (lanes[2] + lanes[3]);
    if (full * 32 < n) {          // ragged tail: at most 31 elements
      uint32_t x = 0;
      for (size_t j = 0; j < 32; ++j) {
        size_t i = full * 32 + j;
        x <<= 1;
        if (i < n) {
          x |= (src[i] < 0) ? 1u : 0u;
          sum += std::fabs((double)src[i]);
        }
      }
      dst[full] = x;
    }
    if (scaled_) *scale = (float)(sum / (double)n);
  }
#else
  void compress_f32(const float*, uint32_t*, size_t, float*) {}
#endif
  size_t compress(void* grad, void* dst) override {
    size_t out = 0;
    BPS_DISPATCH_FLOAT(dtype_, do_compress, grad, (uint32_t*)dst, numel(), &out);
    return out;
  }

  // mode 0: dst = sign*scale ; mode 1: error = corrected - sign*scale
  template <typename A>
  void do_expand(const uint32_t* src, size_t csize, void* dst_, const void* corr_, int mode) {
    typename A::S* dst = (typename A::S*)dst_;
    const typename A::S* corr = (const typename A::S*)corr_;
    const size_t n = numel();
    BPS_CHECK_GE(csize, (size_t)4) << "onebit payload without its scale";
    const size_t chunks = (csize - 4) / 4;
    float scale;
    memcpy(&scale, &src[chunks], 4);
    size_t c0 = 0;
    if (BPS_X86_SIMD && std::is_same<A, TF32>::value) c0 = expand_f32(src, std::min(chunks, n / 32), scale, (float*)dst_, (const float*)corr_, mode);
    for (size_t c = c0; c < chunks; ++c) {
      uint32_t x = src[c];
      for (size_t j = 0; j < 32; ++j) {
        size_t i = c * 32 + j;
        if (i >= n) break;
        float v = ((x >> (31 - j)) & 1u) ? -scale : scale;
        if (mode == 0) A::st(dst, i, v);
        else if (mode == 1) A::st(dst, i, A::ld(corr, i) - v);
        else {                                   // v as the tensor dtype would store it, then the dtype's own sum
          typename A::S q;
          A::st(&q, 0, v);
          A::st(dst, i, A::ld(dst, i) + A::ld(&q, 0));
        }
      }
    }
  }

  // AVX2 expansion of `full` complete words; returns how many words it handled.
  #if BPS_X86_SIMD
  static size_t expand_f32(const uint32_t* src, size_t full, float scale, float* dst, const float* corr, int mode) {
    const __m256 pos = _mm256_set1_ps(scale), neg = _mm256_set1_ps(-scale);
    const __m256i lane_bit = _mm256_setr_epi32(1, 2, 4, 8, 16, 32, 64, 128);
    for (size_t c = 0; c < full; ++c) {
      const uint32_t lsb_first = bit_reverse32(src[c]);      // bit j = element j
      for (int g = 0; g < 4; ++g) {
        __m256i byte = _mm256_set1_epi32((int)((lsb_first >> (8 * g)) & 0xffu));
        __m256i hit = _mm256_cmpeq_epi32(_mm256_and_si256(byte, lane_bit), lane_bit);
        __m256 v = _mm256_blendv_ps(pos, neg, _mm256_castsi256_ps(hit));
        float* d = dst + c * 32 + g * 8;
        if (mode == 0) _mm256_storeu_ps(d, v);
        else if (mode == 1) _mm256_storeu_ps(d, _mm256_sub_ps(_mm256_loadu_ps(corr + c * 32 + g * 8), v));
        else _mm256_storeu_ps(d, _mm256_add_ps(_mm256_loadu_ps(d), v));
      }
    }
    return full;
  }
#else
  static size_t expand_f32(const uint32_t*, size_t, float, float*, const float*, int) { return 0; }
#endif
  void decompress(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)src, csize, dst, nullptr, 0);
  }
  void fast_update_error(void* error, const void* corrected, const void* compressed, size_t csize) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)compressed, csize, error, corrected, 1);
  }
  void decompress_add(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)src, csize, dst, nullptr, 2);
  }

 private:
  bool scaled_;
};

// ---------------------------------------------------------------- sparse (index,value) pairs
// payload: k records of {uint32 index; value} where value is stored in the
// tensor dtype, padded so each record is 8 bytes (16 bytes for fp64).
template <typename A>
struct PairRec;
template <>
struct PairRec<TF32> {
  uint32_t idx;
  float val;
};
template <>
struct PairRec<TF64> {
  uint64_t idx;
  double val;
};
template <>
struct PairRec<TF16> {
  uint32_t idx;
  uint16_t val;
  uint16_t pad;
};
template <>
struct PairRec<TBF16> {
  uint32_t idx;
  uint16_t val;
  uint16_t pad;
};

static size_t pair_bytes(int dtype) { return dtype == F64 ? 16 : 8; }

class SparseBase : public Compressor {
 public:
  SparseBase(size_t nbytes, int dtype, unsigned k) : Compressor(nbytes, dtype), k_(k) {
    BPS_CHECK_GT(k, 0u);
    // an absolute k can exceed a small tensor (a bias next to a conv weight): keep everything then
    const size_t n = numel();
    if (n > 0 && (size_t)k_ > n) k_ = (unsigned)n;
  }
  size_t max_compressed_bytes() const override { return (size_t)k_ * pair_bytes(dtype_); }

  template <typename A>
  void do_scatter(const void* src_, size_t csize, void* dst_, const void* corr_, int mode) {
    using R = PairRec<A>;
    typename A::S* dst = (typename A::S*)dst_;
    const R* recs = (const R*)src_;
    const size_t cnt = csize / sizeof(R);
    const size_t n = numel();
    if (mode == 0) memset(dst, 0, nbytes_);
    else if (dst_ != corr_) memcpy(dst, corr_, nbytes_);
    for (size_t i = 0; i < cnt; ++i) {
      size_t ix = (size_t)recs[i].idx;
      if (ix >= n) continue;
      if (mode == 0) dst[ix] = recs[i].val;   // later duplicates overwrite, as in the reference
      else A::st(dst, ix, 0.0f);
    }
  }
  void decompress(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_scatter, src, csize, dst, nullptr, 0);
  }
  void fast_update_error(void* error, const void* corrected, const void* compressed, size_t csize) override {
    BPS_DISPATCH_FLOAT(dtype_, do_scatter, compressed, csize, error, corrected, 1);
  }
  // O(k): only the transmitted entries are touched.  A payload may name an index twice (random-k); decompress() lets
  // the LAST record win, so the records are visited back to front and an index is added once.
  template <typename A>
  void do_scatter_add(const void* src_, size_t csize, void* dst_) {
    using R = PairRec<A>;
    typename A::S* dst = (typename A::S*)dst_;
    const R* recs = (const R*)src_;
    const size_t cnt = csize / sizeof(R);
    const size_t n = numel();
    static thread_local std::vector<uint64_t> seen;
    if (seen.size() < (n + 63) / 64) seen.resize((n + 63) / 64, 0);
    for (size_t i = cnt; i-- > 0;) {
      const size_t ix = (size_t)recs[i].idx;
      if (ix >= n) continue;
      uint64_t& w = seen[ix >> 6];
      const uint64_t bit = 1ull << (ix & 63);
      if (w & bit) continue;
      w |= bit;
      A::st(dst, ix, A::ld(dst, ix) + A::ld(&recs[i].val, 0));
    }
    for (size_t i = 0; i < cnt; ++i) {
      const size_t ix = (size_t)recs[i].idx;
      if (ix < n) seen[ix >> 6] &= ~(1ull << (ix & 63));
    }
  }
  void decompress_add(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_scatter_add, src, csize, dst);
  }

 protected:
  unsigned k_;
};

class TopkCompressor : public SparseBase {
 public:
  using SparseBase::SparseBase;
  const char* name() const override { return "topk"; }
  template <typename A>
  void do_compress(const void* src_, void* dst_, size_t n, size_t* out) {
    using R = PairRec<A>;
    const typename A::S* src = (const typename A::S*)src_;
    R* heap = (R*)dst_;
    BPS_CHECK_LE((size_t)k_, n) << "topk: k larger than the tensor";
    // min-heap on |value| of the k best seen so far
    auto cmp = [](const R& a, const R& b) { return std::fabs((double)A::ld(&a.val, 0)) > std::fabs((double)A::ld(&b.val, 0)); };
    size_t size = 0;
    if (BPS_X86_SIMD && std::is_same<A, TF32>::value) {
      topk_f32((const float*)src_, (PairRec<TF32>*)dst_, n);
      *out = (size_t)k_ * sizeof(R);
      return;
    }
    for (size_t i = 0; i < n; ++i) {
      if (i < k_) {
        R r{};
        r.idx = (decltype(r.idx))i;
        r.val = src[i];
        heap[size++] = r;
        std::push_heap(heap, heap + size, cmp);
      } else if (std::fabs((double)A::ld(src, i)) > std::fabs((double)A::ld(&heap[0].val, 0))) {
        std::pop_heap(heap, heap + size, cmp);
        R r{};
        r.idx = (decltype(r.idx))i;
        r.val = src[i];
        heap[size - 1] = r;
        std::push_heap(heap, heap + size, cmp);
      }
    }
    *out = (size_t)k_ * sizeof(R);
  }
  // Same algorithm and visiting order as the generic path (a min-heap of the k largest magnitudes seen so far, a
  // newcomer must be strictly larger than the minimum), so the selected set is the same; but once the heap is full,
  // eight candidates at a time are tested against the current threshold with one AVX2 compare - almost all blocks
  // are rejected without touching the heap - and a replacement is one sift-down instead of pop_heap + push_heap.
  // Large tensors, fp32: estimate a threshold slightly BELOW the k-th largest magnitude from 16 K samples, collect
  // everything above it in one vectorised pass (a few k candidates) and select the exact top k among those.  The
  // result is exact whenever at least k candidates were found (then every element of the true top k is a candidate);
  // otherwise - or when the candidate buffer overflows (many equal values) - the caller falls back to the heap.
  // Ties on the k-th magnitude go to the lower index.
  #if BPS_X86_SIMD
  bool topk_select_f32(const float* src, PairRec<TF32>* out, size_t n) {
    using R = PairRec<TF32>;
    const size_t k = k_, m = 16384;
    if (n < 8 * m) return false;
    const double hits = (double)k * (double)m / (double)n;       // expected samples inside the true top k
    if (hits < 24.0) return false;
    const size_t rank = (size_t)(hits + 5.0 * std::sqrt(hits) + 10.0);
    if (rank >= m / 2) return false;                              // k is a large fraction: the heap is fine
    static thread_local std::vector<float> sample;
    static thread_local std::vector<R> cand;
    sample.resize(m);
    const size_t stride = n / m;
    for (size_t i = 0; i < m; ++i) {
      const size_t j = i * stride + (size_t)(((uint32_t)i * 2654435761u) >> 8) % stride;
      sample[i] = std::fabs(src[j]);
    }
    std::nth_element(sample.begin(), sample.begin() + rank, sample.end(), std::greater<float>());
    const float lo = sample[rank];
    if (!(lo > 0.0f)) return false;
    const size_t cap = 4 * (size_t)((double)rank * (double)n / (double)m) + 64;
    cand.resize(cap + 8);
    size_t cnt = 0;
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    const __m256 thr = _mm256_set1_ps(lo);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      int mask = _mm256_movemask_ps(_mm256_cmp_ps(_mm256_and_ps(_mm256_loadu_ps(src + i), absmask), thr, _CMP_GT_OQ));
      while (mask) {
        const int lane = __builtin_ctz(mask);
        mask &= mask - 1;
        cand[cnt].idx = (uint32_t)(i + lane);
        cand[cnt].val = src[i + lane];
        ++cnt;
      }
      if (cnt > cap) return false;
    }
    for (; i < n; ++i)
      if (std::fabs(src[i]) > lo) {
        cand[cnt].idx = (uint32_t)i;
        cand[cnt].val = src[i];
        ++cnt;
      }
    if (cnt < k || cnt > cap) return false;
    auto better = [](const R& a, const R& b) {
      const float x = std::fabs(a.val), y = std::fabs(b.val);
      return x > y || (x == y && a.idx < b.idx);
    };
    if (cnt > k) std::nth_element(cand.begin(), cand.begin() + k, cand.begin() + cnt, better);
    memcpy(out, cand.data(), k * sizeof(R));
    return true;
  }
#else
  bool topk_select_f32(const float*, PairRec<TF32>*, size_t) { return false; }
#endif

  #if BPS_X86_SIMD
  void topk_f32(const float* src, PairRec<TF32>* heap, size_t n) {
    using R = PairRec<TF32>;
    if (topk_select_f32(src, heap, n)) return;
    auto cmp = [](const R& a, const R& b) { return std::fabs(a.val) > std::fabs(b.val); };
    size_t size = 0;
    size_t i = 0;
    for (; i < k_ && i < n; ++i) {
      R r{};
      r.idx = (decltype(r.idx))i;
      r.val = src[i];
      heap[size++] = r;
      std::push_heap(heap, heap + size, cmp);
    }
    auto offer = [&](size_t j) {
      const float key = std::fabs(src[j]);
      if (!(key > std::fabs(heap[0].val))) return;
      // replace the minimum and sift the newcomer down: one pass instead of pop_heap + push_heap
      size_t hole = 0;
      for (;;) {
        size_t child = 2 * hole + 1;
        if (child >= size) break;
        if (child + 1 < size && std::fabs(heap[child + 1].val) < std::fabs(heap[child].val)) ++child;
        if (!(std::fabs(heap[child].val) < key)) break;
        heap[hole] = heap[child];
        hole = child;
      }
      heap[hole].idx = (decltype(heap[hole].idx))j;
      heap[hole].val = src[j];
    };
    const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    for (; i + 8 <= n; i += 8) {
      __m256 a = _mm256_and_ps(_mm256_loadu_ps(src + i), absmask);
      __m256 thr = _mm256_set1_ps(std::fabs(heap[0].val));
      int m = _mm256_movemask_ps(_mm256_cmp_ps(a, thr, _CMP_GT_OQ));
      while (m) {                  // in index order, against the threshold as it rises
        int lane = __builtin_ctz(m);
        m &= m - 1;
        offer(i + lane);
      }
    }
    for (; i < n; ++i) offer(i);
  }
#else
  void topk_f32(const float*, PairRec<TF32>*, size_t) {}
#endif
  size_t compress(void* grad, void* dst) override {
    size_t out = 0;
    BPS_DISPATCH_FLOAT(dtype_, do_compress, grad, dst, numel(), &out);
    return out;
  }
};

class RandomkCompressor : public SparseBase {
 public:
  RandomkCompressor(size_t nbytes, int dtype, unsigned k, unsigned seed) : SparseBase(nbytes, dtype, k) {
    if (seed) rng_.set_seed(seed);
  }
  const char* name() const override { return "randomk"; }
  template <typename A>
  void do_compress(const void* src_, void* dst_, size_t n, size_t* out) {
    using R = PairRec<A>;
    const typename A::S* src = (const typename A::S*)src_;
    R* recs = (R*)dst_;
    for (size_t i = 0; i < k_; ++i) {
      uint64_t ix = rng_.randint(0, n);
      R r{};
      r.idx = (decltype(r.idx))ix;
      r.val = src[ix];
      recs[i] = r;
    }
    *out = (size_t)k_ * sizeof(R);
  }
  size_t compress(void* grad, void* dst) override {
    size_t out = 0;
    BPS_DISPATCH_FLOAT(dtype_, do_compress, grad, dst, numel(), &out);
    return out;
  }

 private:
  XorShift128Plus rng_;
};

// ---------------------------------------------------------------- dithering
// payload: [Elias-delta bitstream words][uint32 bit count][float scale]
class DitheringCompressor : public Compressor {
 public:
  enum Partition { LINEAR = 0, NATURAL = 1 };
  enum Normalize { MAX = 0, L2 = 1 };
  DitheringCompressor(size_t nbytes, int dtype, unsigned s, unsigned seed, int ptype, int ntype)
      : Compressor(nbytes, dtype), s_(s), ptype_(ptype), ntype_(ntype) {
    BPS_CHECK_GT(s, 0u);
    if (seed) rng_.set_seed(seed);
  }
  const char* name() const override { return "dithering"; }
  // worst case per element: two Elias-delta codes (<= 2*(2*6+32) bits) + sign; budget 16 bytes/elem.
  size_t max_compressed_bytes() const override { return numel() * 16 + 16; }

  // fp32, linear levels: the per-element arithmetic (double-precision normalisation, floor, fraction) does not depend
  // on the random stream, so it runs first, four lanes at a time, into scratch arrays; the sequential part that is
  // left - one xorshift draw per element in index order, then the bit writer - no longer waits on a division.
  // Same operations in the same precision as the generic loop: the payload is bit-identical.
  #if BPS_X86_SIMD
  void compress_linear_f32(const float* src, uint32_t* dst, size_t n, size_t* out) {
    double scale = 0.0;
    const __m256 absmask8 = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
    size_t i = 0;
    if (ntype_ == MAX) {
      __m256 mx = _mm256_setzero_ps();
      for (; i + 8 <= n; i += 8) mx = _mm256_max_ps(_mm256_and_ps(_mm256_loadu_ps(src + i), absmask8), mx);   // NaN: keep mx
      float lanes[8];
      _mm256_storeu_ps(lanes, mx);
      for (float v : lanes) scale = std::max(scale, (double)v);
      for (; i < n; ++i) scale = std::max(scale, (double)std::fabs(src[i]));
    } else {
      for (i = 0; i < n; ++i) {      // sequential on purpose: the rounding of the running sum is part of the contract
        double v = src[i];
        scale += v * v;
      }
      scale = std::sqrt(scale);
    }
    BitWriter w(dst);
    if (scale > 0) {
      static thread_local std::vector<int32_t> floor_q;
      static thread_local std::vector<float> frac;
      floor_q.resize(n + 4);
      frac.resize(n + 4);
      const __m128 absmask4 = _mm_castsi128_ps(_mm_set1_epi32(0x7fffffff));
      const __m256d vscale = _mm256_set1_pd(scale), vs = _mm256_set1_pd((double)s_);
      for (i = 0; i + 4 <= n; i += 4) {
        __m128 ax = _mm_and_ps(_mm_loadu_ps(src + i), absmask4);
        __m128 nrm = _mm256_cvtpd_ps(_mm256_mul_pd(_mm256_div_pd(_mm256_cvtps_pd(ax), vscale), vs));
        __m128 fl = _mm_floor_ps(nrm);
        _mm_storeu_si128((__m128i*)(floor_q.data() + i), _mm_cvttps_epi32(fl));
        _mm_storeu_ps(frac.data() + i, _mm_sub_ps(nrm, fl));
      }
      for (; i < n; ++i) {
        float nrm = (float)((std::fabs(src[i]) / scale) * s_);
        float fl = std::floor(nrm);
        floor_q[i] = (int32_t)fl;
        frac[i] = nrm - fl;
      }
      // Elias-delta codes of 1..255 as (bits, length): gap + sign + level of one element go out in ONE put_bits
      static const std::array<std::pair<uint32_t, uint8_t>, 256> small = [] {
        std::array<std::pair<uint32_t, uint8_t>, 256> t{};
        for (unsigned x = 1; x < 256; ++x) {
          const int len = 32 - __builtin_clz(x), lol = 31 - __builtin_clz((unsigned)len);
          t[x] = {((uint32_t)len << (len - 1)) | (x & ((1u << (len - 1)) - 1)), (uint8_t)(2 * lol + len)};
        }
        return t;
      }();
      size_t last = (size_t)-1;
      for (i = 0; i < n; ++i) {
        const unsigned q = (unsigned)floor_q[i] + (rng_.bernoulli53(frac[i]) ? 1u : 0u);
        if (!q) continue;
        const size_t gap = i - last;
        last = i;
        const uint64_t sign = std::signbit(src[i]) ? 1u : 0u;
        if (gap < 256 && q < 256) {
          const auto& a = small[gap];
          const auto& b = small[q];
          w.put_bits((((uint64_t)a.first << 1 | sign) << b.second) | b.first, a.second + 1 + b.second);
        } else {
          elias_delta_encode(w, gap);
          w.put(sign != 0);
          elias_delta_encode(w, q);
        }
      }
    }
    w.flush();
    size_t blocks = w.blocks();
    dst[blocks] = (uint32_t)w.bits();
    float fs = (float)scale;
    memcpy(&dst[blocks + 1], &fs, 4);
    *out = blocks * 4 + 8;
  }
#else
  void compress_linear_f32(const float*, uint32_t*, size_t, size_t*) {}
#endif

  template <typename A>
  void do_compress(const void* src_, uint32_t* dst, size_t n, size_t* out) {
    const typename A::S* src = (const typename A::S*)src_;
    if (BPS_X86_SIMD && std::is_same<A, TF32>::value && ptype_ == LINEAR) {
      compress_linear_f32((const float*)src_, dst, n, out);
      return;
    }
    double scale = 0.0;
    if (ntype_ == MAX) {
      for (size_t i = 0; i < n; ++i) scale = std::max(scale, (double)std::fabs(A::ld(src, i)));
    } else {
      for (size_t i = 0; i < n; ++i) {
        double v = A::ld(src, i);
        scale += v * v;
      }
      scale = std::sqrt(scale);
    }
    BitWriter w(dst);
    size_t last = (size_t)-1;
    if (scale > 0) {
      const unsigned level = 1u << (s_ - 1);
      for (size_t i = 0; i < n; ++i) {
        float x = (float)A::ld(src, i);
        float ax = std::fabs(x);
        unsigned q;
        if (ptype_ == LINEAR) {
          float normalized = (float)((ax / scale) * s_);
          float fl = std::floor(normalized);
          q = (unsigned)fl + (rng_.bernoulli(normalized - fl) ? 1u : 0u);
        } else {
          double normalized = (ax / scale) * level;
          unsigned fl = round_next_pow2((uint32_t)std::ceil(normalized)) >> 1;
          unsigned length = fl ? fl : 1;
          double p = (normalized - fl) / length;
          q = fl + length * (rng_.bernoulli(p) ? 1u : 0u);
        }
        if (q) {
          size_t diff = i - last;
          last = i;
          elias_delta_encode(w, diff);
          w.put(std::signbit(x));
          elias_delta_encode(w, q);
        }
      }
    }
    w.flush();
    size_t blocks = w.blocks();
    dst[blocks] = (uint32_t)w.bits();
    float fs = (float)scale;
    memcpy(&dst[blocks + 1], &fs, 4);
    *out = blocks * 4 + 8;
  }
  size_t compress(void* grad, void* dst) override {
    size_t out = 0;
    BPS_DISPATCH_FLOAT(dtype_, do_compress, grad, (uint32_t*)dst, numel(), &out);
    return out;
  }

  template <typename A>
  void do_expand(const uint32_t* src, size_t csize, void* dst_, const void* corr_, int mode) {
    typename A::S* dst = (typename A::S*)dst_;
    // mode 0: dst = D(src); mode 1: dst = corr - D(src); mode 2: dst += D(src) (untransmitted entries are zeros)
    if (csize < 8) {              // not a dithering payload (empty response): nothing was transmitted
      if (mode == 0) memset(dst, 0, nbytes_);
      else if (mode == 1 && dst_ != corr_) memcpy(dst, corr_, nbytes_);
      return;
    }
    const size_t blocks = (csize - 8) / 4;
    const uint32_t bits = src[blocks];
    float scale;
    memcpy(&scale, &src[blocks + 1], 4);
    if (mode == 0) memset(dst, 0, nbytes_);
    else if (mode == 1 && dst_ != corr_) memcpy(dst, corr_, nbytes_);
    unsigned s = (ptype_ == NATURAL) ? (1u << (s_ - 1)) : s_;
    BitReader r(src);
    size_t last = (size_t)-1;
    const size_t n = numel();
    while (r.bits() < bits) {
      size_t diff = elias_delta_decode(r);
      size_t i = last + diff;
      last = i;
      int sb = r.get();
      unsigned q = (unsigned)elias_delta_decode(r);
      if (i >= n) break;
      float num = q * scale / s;
      float v = (1 - (sb << 1)) * num;
      if (mode == 0) A::st(dst, i, v);
      else if (mode == 1) A::st(dst, i, (float)A::ld(dst, i) - v);
      else {                                       // v as the tensor dtype stores it, summed in the dtype's own way
        typename A::S qv;
        A::st(&qv, 0, v);
        A::st(dst, i, A::ld(dst, i) + A::ld(&qv, 0));
      }
    }
  }
  void decompress(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)src, csize, dst, nullptr, 0);
  }
  void fast_update_error(void* error, const void* corrected, const void* compressed, size_t csize) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)compressed, csize, error, corrected, 1);
  }
  void decompress_add(const void* src, size_t csize, void* dst) override {
    BPS_DISPATCH_FLOAT(dtype_, do_expand, (const uint32_t*)src, csize, dst, nullptr, 2);
  }

 private:
  unsigned s_;
  int ptype_, ntype_;
  XorShift128Plus rng_;
};

// ---------------------------------------------------------------- decorators
// g += (lr_prev/lr_cur) * e ; c = C(g) ; e = g - D(c)
class VanillaErrorFeedback : public Compressor {
 public:
  VanillaErrorFeedback(size_t nbytes, int dtype, std::unique_ptr<Compressor> inner)
      : Compressor(nbytes, dtype), inner_(std::move(inner)), error_(nbytes, 0), reducer_(1) {
    std::string f = env_str("BYTEPS_LR_FILE", "lr.s");
    fd_ = open(f.c_str(), O_RDONLY);
    if (fd_ >= 0) {
      void* p = mmap(nullptr, 8, PROT_READ, MAP_SHARED, fd_, 0);
      if (p != MAP_FAILED) {
        mm_ = p;
        pre_lr_ = cur_lr_ = *reinterpret_cast<double*>(mm_);
      }
    }
  }
  ~VanillaErrorFeedback() override {
    if (mm_) munmap(mm_, 8);
    if (fd_ >= 0) close(fd_);
  }
  const char* name() const override { return "vanilla_ef"; }
  size_t max_compressed_bytes() const override { return inner_->max_compressed_bytes(); }
  void set_lr(double lr) override {
    api_lr_ = lr;
    inner_->set_lr(lr);
  }
  size_t compress(void* grad, void* dst) override {
    if (api_lr_ > 0) cur_lr_ = api_lr_;
    else if (mm_) cur_lr_ = *reinterpret_cast<double*>(mm_);
    double ratio = (cur_lr_ > 0 && pre_lr_ > 0) ? pre_lr_ / cur_lr_ : 1.0;
    // The corrected gradient g + ratio * e is built IN the error buffer, compressed from there, and the error is
    // then updated in place (e = corrected - D(c)): sparse compressors only clear their k entries instead of copying
    // the partition, and `grad` is left untouched.
    reducer_.sum_scaled(error_.data(), grad, error_.data(), nbytes_, dtype_, (float)ratio);
    pre_lr_ = cur_lr_;
    size_t cs = inner_->compress(error_.data(), dst);
    inner_->fast_update_error(error_.data(), error_.data(), dst, cs);
    return cs;
  }
  void decompress(const void* src, size_t csize, void* dst) override { inner_->decompress(src, csize, dst); }
  void decompress_add(const void* src, size_t csize, void* dst) override { inner_->decompress_add(src, csize, dst); }
  const void* error() const { return error_.data(); }

 private:
  std::unique_ptr<Compressor> inner_;
  std::vector<char> error_;
  CpuReducer reducer_;
  int fd_ = -1;
  void* mm_ = nullptr;
  double pre_lr_ = 1.0, cur_lr_ = 1.0, api_lr_ = -1.0;
};

// m = mu*m + g ; g += mu*m
class NesterovMomentum : public Compressor {
 public:
  NesterovMomentum(size_t nbytes, int dtype, std::unique_ptr<Compressor> inner, float mu)
      : Compressor(nbytes, dtype), inner_(std::move(inner)), mom_(nbytes, 0), mu_(mu), reducer_(1) {}
  const char* name() const override { return "nesterov_momentum"; }
  size_t max_compressed_bytes() const override { return inner_->max_compressed_bytes(); }
  void set_lr(double lr) override { inner_->set_lr(lr); }
  size_t compress(void* grad, void* dst) override {
    if (dtype_ == F32) {
      // both updates in one pass over the partition (same expressions as the two reducer calls below)
      float* __restrict m = (float*)mom_.data();
      float* __restrict g = (float*)grad;
      const float mu = mu_;
      const size_t n = numel();
      for (size_t i = 0; i < n; ++i) {
        const float mi = g[i] + mu * m[i];
        m[i] = mi;
        g[i] = g[i] + mu * mi;
      }
    } else {
      reducer_.sum_scaled(mom_.data(), grad, mom_.data(), nbytes_, dtype_, mu_);
      reducer_.sum_scaled(grad, mom_.data(), nbytes_, dtype_, mu_);
    }
    return inner_->compress(grad, dst);
  }
  void decompress(const void* src, size_t csize, void* dst) override { inner_->decompress(src, csize, dst); }
  void decompress_add(const void* src, size_t csize, void* dst) override { inner_->decompress_add(src, csize, dst); }

 private:
  std::unique_ptr<Compressor> inner_;
  std::vector<char> mom_;
  float mu_;
  CpuReducer reducer_;
};

// ---------------------------------------------------------------- registry
static std::mutex& reg_mu() {
  static std::mutex m;
  return m;
}
static std::map<std::string, CompressorCtor>& reg_map() {
  static std::map<std::string, CompressorCtor> m;
  return m;
}

void CompressorRegistry::add(const std::string& name, CompressorCtor c) {
  std::lock_guard<std::mutex> g(reg_mu());
  reg_map()[name] = std::move(c);
}

static unsigned resolve_k(const Kwargs& kw, size_t nbytes, int dtype) {
  float factor = kw_get<float>(kw, "compressor_k", false);
  BPS_CHECK_GT(factor, 0.0f) << "compressor_k must be positive";
  unsigned k;
  if (factor < 1) {
    k = (unsigned)(factor * (nbytes / dtype_size(dtype)));
    if (k == 0) k = 1;
  } else {
    k = (unsigned)factor;
  }
  return k;
}

static void register_builtin() {
  static std::once_flag once;
  std::call_once(once, [] {
    auto& m = reg_map();
    m["onebit_compressor_type"] = [](const Kwargs& kw, size_t n, int d, bool) {
      bool scaled = kw_get<bool>(kw, "compressor_onebit_scaling", true, false);
      return std::unique_ptr<Compressor>(new OnebitCompressor(n, d, scaled));
    };
    m["topk_compressor_type"] = [](const Kwargs& kw, size_t n, int d, bool) {
      return std::unique_ptr<Compressor>(new TopkCompressor(n, d, resolve_k(kw, n, d)));
    };
    m["randomk_compressor_type"] = [](const Kwargs& kw, size_t n, int d, bool) {
      unsigned seed = kw_get<unsigned>(kw, "seed", true, 0);
      return std::unique_ptr<Compressor>(new RandomkCompressor(n, d, resolve_k(kw, n, d), seed));
    };
    m["dithering_compressor_type"] = [](const Kwargs& kw, size_t n, int d, bool) {
      unsigned k = kw_get<unsigned>(kw, "compressor_k", false);
      unsigned seed = kw_get<unsigned>(kw, "seed", true, 0);
      int pt = kw_get<int>(kw, "dithering_partition", true, 0);
      int nt = kw_get<int>(kw, "dithering_normalize", true, 0);
      return std::unique_ptr<Compressor>(new DitheringCompressor(n, d, k, seed, pt, nt));
    };
    m["vanilla_ef_type"] = [](const Kwargs& kw, size_t n, int d, bool server) {
      Kwargs c = kw;
      c.erase("ef_type");
      c.erase("momentum_type");
      auto inner = CompressorRegistry::create(c, n, d, server);
      BPS_CHECK(inner != nullptr) << "error feedback needs a compressor_type";
      return std::unique_ptr<Compressor>(new VanillaErrorFeedback(n, d, std::move(inner)));
    };
    m["nesterov_momentum_type"] = [](const Kwargs& kw, size_t n, int d, bool server) {
      Kwargs c = kw;
      c.erase("momentum_type");
      auto inner = CompressorRegistry::create(c, n, d, server);
      BPS_CHECK(inner != nullptr) << "momentum needs a compressor_type";
      float mu = kw_get<float>(kw, "momentum_mu", false);
      return std::unique_ptr<Compressor>(new NesterovMomentum(n, d, std::move(inner), mu));
    };
  });
}

std::unique_ptr<Compressor> CompressorRegistry::create(const Kwargs& kw, size_t nbytes, int dtype, bool server_side) {
  register_builtin();
  static const char* worker_order[] = {"momentum_type", "ef_type", "compressor_type"};
  static const char* server_order[] = {"ef_type", "compressor_type"};
  const char** order = server_side ? server_order : worker_order;
  int cnt = server_side ? 2 : 3;
  for (int i = 0; i < cnt; ++i) {
    auto it = kw.find(order[i]);
    if (it == kw.end()) continue;
    std::string key = it->second + "_" + order[i];
    CompressorCtor ctor;
    {
      std::lock_guard<std::mutex> g(reg_mu());
      auto f = reg_map().find(key);
      if (f == reg_map().end()) BPS_LOG_FATAL << "no compressor registered under name: " << key;
      ctor = f->second;
    }
    return ctor(kw, nbytes, dtype, server_side);
  }
  return nullptr;
}

std::vector<std::string> CompressorRegistry::names() {
  register_builtin();
  std::lock_guard<std::mutex> g(reg_mu());
  std::vector<std::string> out;
  for (auto& kv : reg_map()) out.push_back(kv.first);
  return out;
}

}  // namespace bps
