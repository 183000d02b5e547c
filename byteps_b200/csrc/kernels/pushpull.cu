// Fused push-pull kernels for sm_100a (see pushpull.cuh for the contract).
//
// Geometry.  The flat wire range is cut into UNITS of 8 elements (16 B for
// 16-bit wire types, 32 B for fp32).  Rank r owns the contiguous shard of
// units [r*per, (r+1)*per).  Inside a shard, tiles of (blockDim * UNROLL) units
// are dealt round-robin to the CTAs; the same (shard-local tile -> blockIdx)
// map is used on every rank, so the CTA that packs a tile on rank A is the
// peer-barrier partner of the CTA that reduces it on rank B.  That is what lets
// every cross-rank dependency be expressed with per-CTA flag barriers and no
// grid-wide sync.
#include "kernels/pushpull.cuh"

#include "kernels/common.cuh"

namespace bps {

namespace {

constexpr int kPeerChunk = 8;  // peers handled per statically-unrolled pass

template <class T>
struct UnitOf {
  static constexpr int kVecs = 8 / T::kPerVec;        // 16-byte vectors per unit
  static constexpr int kBytes = 8 * T::kBytes;        // bytes per unit
};

template <class T>
__device__ __forceinline__ void unit_to_float(const Vec16* v, float* f) {
#pragma unroll
  for (int k = 0; k < UnitOf<T>::kVecs; ++k) T::unpack(v[k], f + k * T::kPerVec);
}
template <class T>
__device__ __forceinline__ void float_to_unit(const float* f, Vec16* v) {
#pragma unroll
  for (int k = 0; k < UnitOf<T>::kVecs; ++k) v[k] = T::pack(f + k * T::kPerVec);
}

// ---------------------------------------------------------------- segment cursor
struct SegCursor {
  const SegDesc* segs;
  int nsegs;
  int i;
  __device__ __forceinline__ void init(const SegDesc* s, int n, int64_t e) {
    segs = s;
    nsegs = n;
    int lo = 0, hi = n - 1;
    while (lo < hi) {  // last seg with start <= e
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&segs[mid].start) <= e) lo = mid;
      else hi = mid - 1;
    }
    i = lo;
  }
  __device__ __forceinline__ void seek(int64_t e) {
    while (i + 1 < nsegs && __ldg(&segs[i + 1].start) <= e) ++i;
  }
};

// gather 8 elements starting at flat element e from the user tensors (zero padded)
template <class U>
__device__ __forceinline__ void gather_unit(SegCursor& c, int64_t e, float* f) {
  c.seek(e);
  const int64_t start = __ldg(&c.segs[c.i].start);
  const int64_t n = __ldg(&c.segs[c.i].n);
  const char* src = (const char*)__ldg((const unsigned long long*)&c.segs[c.i].src);
  const int64_t local = e - start;
  if (local + 8 <= n && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    Vec16 v[UnitOf<U>::kVecs];
#pragma unroll
    for (int k = 0; k < UnitOf<U>::kVecs; ++k) v[k] = ld_stream16(src + local * U::kBytes + k * 16);
    unit_to_float<U>(v, f);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (local + k < n && local + k >= 0) ? U::load1(src, local + k) : 0.f;
  }
}

template <class U>
__device__ __forceinline__ void scatter_unit(SegCursor& c, int64_t e, const float* f) {
  c.seek(e);
  const int64_t start = __ldg(&c.segs[c.i].start);
  const int64_t n = __ldg(&c.segs[c.i].n);
  char* dst = (char*)__ldg((const unsigned long long*)&c.segs[c.i].dst);
  const int64_t local = e - start;
  if (local + 8 <= n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    Vec16 v[UnitOf<U>::kVecs];
    float_to_unit<U>(f, v);
#pragma unroll
    for (int k = 0; k < UnitOf<U>::kVecs; ++k) st_stream16(dst + local * U::kBytes + k * 16, v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (local + k < n && local + k >= 0) U::store1(dst, local + k, f[k]);
  }
}

// ---------------------------------------------------------------- reduce over peers
template <class W, int UNROLL>
__device__ __forceinline__ void reduce_units_p2p(const PeerView& pv, size_t off, const size_t (&idx)[UNROLL],
                                                 const bool (&valid)[UNROLL], int rot, float (&acc)[UNROLL][8]) {
  constexpr int KV = UnitOf<W>::kVecs;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[u][k] = 0.f;
  for (int base = 0; base < pv.world; base += kPeerChunk) {
    Vec16 v[UNROLL][kPeerChunk][KV];
#pragma unroll
    for (int j = 0; j < kPeerChunk; ++j) {
      if (base + j < pv.world) {
        int p = base + j + rot;
        if (p >= pv.world) p -= pv.world;
        const char* src = pv.data[p] + off;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (valid[u]) {
#pragma unroll
            for (int k = 0; k < KV; ++k) v[u][j][k] = ld_stream16(src + idx[u] * UnitOf<W>::kBytes + k * 16);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kPeerChunk; ++j) {
      if (base + j < pv.world) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (valid[u]) {
            float f[8];
            unit_to_float<W>(v[u][j], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[u][k] += f[k];
          }
        }
      }
    }
  }
}

template <class W>
__device__ __forceinline__ void reduce_unit_nvls(const PeerView& pv, size_t off, size_t unit, float* acc) {
  Vec16 v[UnitOf<W>::kVecs];
#pragma unroll
  for (int k = 0; k < UnitOf<W>::kVecs; ++k) v[k] = W::mm_reduce(pv.mc_data + off + unit * UnitOf<W>::kBytes + k * 16);
  unit_to_float<W>(v, acc);
}

// write one unit into the window of every rank
template <class P>
__device__ __forceinline__ void sink_peers(const PeerView& pv, size_t off, size_t unit, const float* f, bool nvls) {
  Vec16 v[UnitOf<P>::kVecs];
  float_to_unit<P>(f, v);
  if (nvls) {
#pragma unroll
    for (int k = 0; k < UnitOf<P>::kVecs; ++k) mm_st16(pv.mc_data + off + unit * UnitOf<P>::kBytes + k * 16, v[k]);
  } else {
    for (int base = 0; base < pv.world; base += kPeerChunk) {
#pragma unroll
      for (int j = 0; j < kPeerChunk; ++j) {
        if (base + j < pv.world) {
          int p = base + j + pv.rank;
          if (p >= pv.world) p -= pv.world;
          char* dst = pv.data[p] + off + unit * UnitOf<P>::kBytes;
#pragma unroll
          for (int k = 0; k < UnitOf<P>::kVecs; ++k) st_stream16(dst + k * 16, v[k]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- epilogues
struct EpiScale {
  float scale;
  __device__ __forceinline__ void operator()(float* g, size_t) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] *= scale;
  }
};

__device__ __forceinline__ void ld8(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// g (sum of gradients) -> new weights, updating the fp32 master shard in place.
struct EpiSGD {
  float* master;
  float* mom;
  size_t shard_begin;  // units
  float scale;
  OptHParams hp;
  __device__ __forceinline__ void operator()(float* g, size_t unit) const {
    const size_t li = (unit - shard_begin) * 8;
    float w[8], m[8];
    ld8(master + li, w);
    const bool has_mom = hp.momentum != 0.f;
    if (has_mom) ld8(mom + li, m);
    const float gs = scale * hp.grad_scale;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gk = g[k] * gs + hp.weight_decay * w[k];
      if (has_mom) {
        float b = hp.first_step ? gk : hp.momentum * m[k] + (1.f - hp.dampening) * gk;
        m[k] = b;
        gk = hp.nesterov ? gk + hp.momentum * b : b;
      }
      w[k] -= hp.lr * gk;
      g[k] = w[k];
    }
    st8(master + li, w);
    if (has_mom) st8(mom + li, m);
  }
};

struct EpiAdam {
  float* master;
  float* m1;
  float* m2;
  size_t shard_begin;
  float scale;
  OptHParams hp;
  __device__ __forceinline__ void operator()(float* g, size_t unit) const {
    const size_t li = (unit - shard_begin) * 8;
    float w[8], m[8], v[8];
    ld8(master + li, w);
    ld8(m1 + li, m);
    ld8(m2 + li, v);
    const float gs = scale * hp.grad_scale;
    const float inv_c1 = 1.f / hp.bias_c1;
    const float inv_c2 = 1.f / hp.bias_c2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gk = g[k] * gs;
      if (hp.adamw) w[k] *= (1.f - hp.lr * hp.weight_decay);
      else gk += hp.weight_decay * w[k];
      m[k] = hp.beta1 * m[k] + (1.f - hp.beta1) * gk;
      v[k] = hp.beta2 * v[k] + (1.f - hp.beta2) * gk * gk;
      const float denom = sqrtf(v[k] * inv_c2) + hp.eps;
      w[k] -= hp.lr * (m[k] * inv_c1) / denom;
      g[k] = w[k];
    }
    st8(master + li, w);
    st8(m1 + li, m);
    st8(m2 + li, v);
  }
};

// ---------------------------------------------------------------- phases
// Visit the units of shard r that belong to this CTA; fn(unit_index).
template <int UNROLL, class F>
__device__ __forceinline__ void for_owned_tiles(size_t s0, size_t s1, F&& fn) {
  const size_t tile = (size_t)blockDim.x * UNROLL;
  for (size_t t = s0 + (size_t)blockIdx.x * tile; t < s1; t += (size_t)gridDim.x * tile) fn(t);
}

template <class U, class W, int UNROLL>
__device__ __forceinline__ void pack_phase(const PeerView& pv, const SegDesc* segs, int nsegs, size_t stage_off,
                                           size_t total_units) {
  if (nsegs <= 0) return;
  SegCursor cur;
  bool inited = false;
  char* stage = pv.data[pv.rank] + stage_off;
  for (int r = 0; r < pv.world; ++r) {
    size_t s0, s1;
    shard_units(total_units, pv.world, r, &s0, &s1);
    for_owned_tiles<UNROLL>(s0, s1, [&](size_t t) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        size_t unit = t + (size_t)u * blockDim.x + threadIdx.x;
        if (unit < s1) {
          int64_t e = (int64_t)unit * 8;
          if (!inited) {
            cur.init(segs, nsegs, e);
            inited = true;
          }
          float f[8];
          gather_unit<U>(cur, e, f);
          Vec16 v[UnitOf<W>::kVecs];
          float_to_unit<W>(f, v);
#pragma unroll
          for (int k = 0; k < UnitOf<W>::kVecs; ++k) st_stream16(stage + unit * UnitOf<W>::kBytes + k * 16, v[k]);
        }
      }
    });
  }
}

template <class U, class W, int UNROLL>
__device__ __forceinline__ void unpack_phase(const PeerView& pv, const SegDesc* segs, int nsegs, size_t stage_off,
                                             size_t total_units) {
  if (nsegs <= 0) return;
  SegCursor cur;
  bool inited = false;
  const char* stage = pv.data[pv.rank] + stage_off;
  for (int r = 0; r < pv.world; ++r) {
    size_t s0, s1;
    shard_units(total_units, pv.world, r, &s0, &s1);
    for_owned_tiles<UNROLL>(s0, s1, [&](size_t t) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        size_t unit = t + (size_t)u * blockDim.x + threadIdx.x;
        if (unit < s1) {
          int64_t e = (int64_t)unit * 8;
          if (!inited) {
            cur.init(segs, nsegs, e);
            inited = true;
          }
          Vec16 v[UnitOf<W>::kVecs];
#pragma unroll
          for (int k = 0; k < UnitOf<W>::kVecs; ++k) v[k] = ld_stream16(stage + unit * UnitOf<W>::kBytes + k * 16);
          float f[8];
          unit_to_float<W>(v, f);
          scatter_unit<U>(cur, e, f);
        }
      }
    });
  }
}

// reduce the units of [s0,s1) owned by this CTA; epi(acc, unit); sink(acc, unit)
template <class W, int UNROLL, class Epi, class Sink>
__device__ __forceinline__ void reduce_phase(const PeerView& pv, size_t off, size_t s0, size_t s1, bool nvls, int rot,
                                             Epi& epi, Sink&& sink) {
  for_owned_tiles<UNROLL>(s0, s1, [&](size_t t) {
    if (nvls) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        size_t unit = t + (size_t)u * blockDim.x + threadIdx.x;
        if (unit < s1) {
          float acc[8];
          reduce_unit_nvls<W>(pv, off, unit, acc);
          epi(acc, unit);
          sink(acc, unit);
        }
      }
    } else {
      size_t idx[UNROLL];
      bool valid[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        idx[u] = t + (size_t)u * blockDim.x + threadIdx.x;
        valid[u] = idx[u] < s1;
      }
      float acc[UNROLL][8];
      reduce_units_p2p<W, UNROLL>(pv, off, idx, valid, rot, acc);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (valid[u]) {
          epi(acc[u], idx[u]);
          sink(acc[u], idx[u]);
        }
      }
    }
  });
}

template <class W>
struct UnrollOf {
  static constexpr int value = (W::kBytes == 4) ? 1 : 2;
};

// ---------------------------------------------------------------- kernels
// mode: 0 = reduce-scatter + all-gather, 1 = reduce-scatter only, 2 = all-gather only
template <class W>
__global__ void __launch_bounds__(512) pushpull_inplace_kernel(PeerView pv, size_t off, size_t total_units, float scale,
                                                               int mode, int nvls, int channel) {
  constexpr int UNROLL = UnrollOf<W>::value;
  size_t s0, s1;
  shard_units(total_units, pv.world, pv.rank, &s0, &s1);
  barrier_peers(pv, channel);
  if (mode == 2) {
    const char* mine = pv.data[pv.rank] + off;
    for_owned_tiles<UNROLL>(s0, s1, [&](size_t t) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        size_t unit = t + (size_t)u * blockDim.x + threadIdx.x;
        if (unit < s1) {
          Vec16 v[UnitOf<W>::kVecs];
#pragma unroll
          for (int k = 0; k < UnitOf<W>::kVecs; ++k) v[k] = ld_stream16(mine + unit * UnitOf<W>::kBytes + k * 16);
          float f[8];
          unit_to_float<W>(v, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] *= scale;
          sink_peers<W>(pv, off, unit, f, nvls != 0);
        }
      }
    });
  } else {
    EpiScale epi{scale};
    if (mode == 0) {
      reduce_phase<W, UNROLL>(pv, off, s0, s1, nvls != 0, pv.rank + 1 >= pv.world ? 0 : pv.rank + 1, epi,
                              [&](const float* f, size_t unit) { sink_peers<W>(pv, off, unit, f, nvls != 0); });
    } else {
      char* mine = pv.data[pv.rank] + off;
      reduce_phase<W, UNROLL>(pv, off, s0, s1, nvls != 0, pv.rank + 1 >= pv.world ? 0 : pv.rank + 1, epi,
                              [&](const float* f, size_t unit) {
                                Vec16 v[UnitOf<W>::kVecs];
                                float_to_unit<W>(f, v);
#pragma unroll
                                for (int k = 0; k < UnitOf<W>::kVecs; ++k)
                                  st_stream16(mine + unit * UnitOf<W>::kBytes + k * 16, v[k]);
                              });
    }
  }
  barrier_peers(pv, channel);
}

template <class U, class W>
__global__ void __launch_bounds__(512) pushpull_packed_kernel(PeerView pv, const SegDesc* segs, int nsegs,
                                                              size_t stage_off, size_t total_units, float scale,
                                                              int nvls, int one_shot, int end_barrier, int channel) {
  constexpr int UNROLL = UnrollOf<W>::value;
  pack_phase<U, W, UNROLL>(pv, segs, nsegs, stage_off, total_units);
  barrier_peers(pv, channel);
  EpiScale epi{scale};
  if (one_shot) {
    // every rank reduces every unit (fixed peer order -> bit-identical results
    // on all ranks) and scatters straight into the user tensors.
    SegCursor cur;
    bool inited = false;
    for (int r = 0; r < pv.world; ++r) {
      size_t s0, s1;
      shard_units(total_units, pv.world, r, &s0, &s1);
      reduce_phase<W, UNROLL>(pv, stage_off, s0, s1, nvls != 0, 0, epi, [&](const float* f, size_t unit) {
        int64_t e = (int64_t)unit * 8;
        if (!inited) {
          cur.init(segs, nsegs, e);
          inited = true;
        }
        // round through the wire dtype so one-shot and two-shot agree bit for bit
        Vec16 v[UnitOf<W>::kVecs];
        float_to_unit<W>(f, v);
        float g[8];
        unit_to_float<W>(v, g);
        scatter_unit<U>(cur, e, g);
      });
    }
    if (end_barrier) barrier_peers(pv, channel);
  } else {
    size_t s0, s1;
    shard_units(total_units, pv.world, pv.rank, &s0, &s1);
    reduce_phase<W, UNROLL>(pv, stage_off, s0, s1, nvls != 0, pv.rank + 1 >= pv.world ? 0 : pv.rank + 1, epi,
                            [&](const float* f, size_t unit) { sink_peers<W>(pv, stage_off, unit, f, nvls != 0); });
    barrier_peers(pv, channel);
    unpack_phase<U, W, UNROLL>(pv, segs, nsegs, stage_off, total_units);
  }
}

template <class G, class W, class P, class Epi>
__global__ void __launch_bounds__(512) pushpull_fused_opt_kernel(PeerView pv, const SegDesc* segs, int nsegs,
                                                                 size_t stage_off, size_t param_off,
                                                                 size_t total_units, Epi epi_proto,
                                                                 const OptHParams* hp, int nvls, int channel) {
  constexpr int UNROLL = 1;
  pack_phase<G, W, UNROLL>(pv, segs, nsegs, stage_off, total_units);
  barrier_peers(pv, channel);
  size_t s0, s1;
  shard_units(total_units, pv.world, pv.rank, &s0, &s1);
  Epi epi = epi_proto;
  epi.shard_begin = s0;
  epi.hp = *hp;
  reduce_phase<W, UNROLL>(pv, stage_off, s0, s1, nvls != 0, pv.rank + 1 >= pv.world ? 0 : pv.rank + 1, epi,
                          [&](const float* f, size_t unit) { sink_peers<P>(pv, param_off, unit, f, nvls != 0); });
  barrier_peers(pv, channel);
}

__global__ void barrier_only_kernel(PeerView pv, int channel) { barrier_peers(pv, channel); }

// ---------------------------------------------------------------- dispatch helpers
inline size_t units_of(size_t nelem) { return (nelem + 7) / 8; }

template <class F>
cudaError_t dispatch_wire(int wire, F&& f) {
  switch (wire) {
    case WIRE_F32: return f(TagF32{});
    case WIRE_BF16: return f(TagBF16{});
    case WIRE_F16: return f(TagF16{});
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_pushpull_inplace(const PeerView& pv, int wire, size_t off, size_t nelem, float scale,
                                    const LaunchCfg& cfg, cudaStream_t stream) {
  if (cfg.blocks < 1 || cfg.blocks > kMaxBlocks || (off & 15)) return cudaErrorInvalidValue;
  return dispatch_wire(wire, [&](auto w) {
    using W = decltype(w);
    pushpull_inplace_kernel<W><<<cfg.blocks, cfg.threads, 0, stream>>>(pv, off, units_of(nelem), scale, 0,
                                                                       cfg.use_nvls, cfg.channel);
    return cudaGetLastError();
  });
}

cudaError_t launch_reduce_scatter(const PeerView& pv, int wire, size_t off, size_t nelem, const LaunchCfg& cfg,
                                  cudaStream_t stream) {
  if (cfg.blocks < 1 || cfg.blocks > kMaxBlocks || (off & 15)) return cudaErrorInvalidValue;
  return dispatch_wire(wire, [&](auto w) {
    using W = decltype(w);
    pushpull_inplace_kernel<W><<<cfg.blocks, cfg.threads, 0, stream>>>(pv, off, units_of(nelem), 1.0f, 1,
                                                                       cfg.use_nvls, cfg.channel);
    return cudaGetLastError();
  });
}

cudaError_t launch_all_gather(const PeerView& pv, int wire, size_t off, size_t nelem, float scale,
                              const LaunchCfg& cfg, cudaStream_t stream) {
  if (cfg.blocks < 1 || cfg.blocks > kMaxBlocks || (off & 15)) return cudaErrorInvalidValue;
  return dispatch_wire(wire, [&](auto w) {
    using W = decltype(w);
    pushpull_inplace_kernel<W><<<cfg.blocks, cfg.threads, 0, stream>>>(pv, off, units_of(nelem), scale, 2,
                                                                       cfg.use_nvls, cfg.channel);
    return cudaGetLastError();
  });
}

cudaError_t launch_pushpull_packed(const PeerView& pv, int user_dtype, int wire, const SegDesc* segs, int nsegs,
                                   size_t stage_off, size_t total_elems, float scale, const LaunchCfg& cfg,
                                   cudaStream_t stream) {
  if (cfg.blocks < 1 || cfg.blocks > kMaxBlocks || (stage_off & 15) || (total_elems & 7)) return cudaErrorInvalidValue;
  const size_t units = total_elems / 8;
#define BPS_PACKED(U, W)                                                                                             \
  pushpull_packed_kernel<U, W><<<cfg.blocks, cfg.threads, 0, stream>>>(pv, segs, nsegs, stage_off, units, scale,     \
                                                                       cfg.use_nvls, cfg.one_shot, cfg.end_barrier,  \
                                                                       cfg.channel);                                 \
  return cudaGetLastError();
  if (user_dtype == WIRE_F32 && wire == WIRE_F32) { BPS_PACKED(TagF32, TagF32) }
  if (user_dtype == WIRE_F32 && wire == WIRE_BF16) { BPS_PACKED(TagF32, TagBF16) }
  if (user_dtype == WIRE_F32 && wire == WIRE_F16) { BPS_PACKED(TagF32, TagF16) }
  if (user_dtype == WIRE_BF16 && wire == WIRE_BF16) { BPS_PACKED(TagBF16, TagBF16) }
  if (user_dtype == WIRE_F16 && wire == WIRE_F16) { BPS_PACKED(TagF16, TagF16) }
  if (user_dtype == WIRE_BF16 && wire == WIRE_F32) { BPS_PACKED(TagBF16, TagF32) }
  if (user_dtype == WIRE_F16 && wire == WIRE_F32) { BPS_PACKED(TagF16, TagF32) }
#undef BPS_PACKED
  return cudaErrorInvalidValue;
}

cudaError_t launch_pushpull_fused_opt(const PeerView& pv, int grad_dtype, int wire, int param_dtype, int opt_kind,
                                      const SegDesc* segs, int nsegs, size_t stage_off, size_t param_off,
                                      size_t total_elems, float scale, float* master, float* state0, float* state1,
                                      const OptHParams* hp, const LaunchCfg& cfg, cudaStream_t stream) {
  if (cfg.blocks < 1 || cfg.blocks > kMaxBlocks || (stage_off & 15) || (param_off & 15) || (total_elems & 7))
    return cudaErrorInvalidValue;
  const size_t units = total_elems / 8;
#define BPS_FUSED(G, W, P)                                                                                           \
  if (opt_kind == OPT_SGD) {                                                                                         \
    EpiSGD e{master, state0, 0, scale, OptHParams{}};                                                                \
    pushpull_fused_opt_kernel<G, W, P, EpiSGD><<<cfg.blocks, cfg.threads, 0, stream>>>(                              \
        pv, segs, nsegs, stage_off, param_off, units, e, hp, cfg.use_nvls, cfg.channel);                             \
  } else if (opt_kind == OPT_ADAM) {                                                                                 \
    EpiAdam e{master, state0, state1, 0, scale, OptHParams{}};                                                       \
    pushpull_fused_opt_kernel<G, W, P, EpiAdam><<<cfg.blocks, cfg.threads, 0, stream>>>(                             \
        pv, segs, nsegs, stage_off, param_off, units, e, hp, cfg.use_nvls, cfg.channel);                             \
  } else {                                                                                                           \
    return cudaErrorInvalidValue;                                                                                    \
  }                                                                                                                  \
  return cudaGetLastError();
  if (grad_dtype == WIRE_BF16 && wire == WIRE_BF16 && param_dtype == WIRE_BF16) { BPS_FUSED(TagBF16, TagBF16, TagBF16) }
  if (grad_dtype == WIRE_F32 && wire == WIRE_F32 && param_dtype == WIRE_F32) { BPS_FUSED(TagF32, TagF32, TagF32) }
  if (grad_dtype == WIRE_F32 && wire == WIRE_BF16 && param_dtype == WIRE_F32) { BPS_FUSED(TagF32, TagBF16, TagF32) }
  if (grad_dtype == WIRE_F16 && wire == WIRE_F16 && param_dtype == WIRE_F16) { BPS_FUSED(TagF16, TagF16, TagF16) }
  if (grad_dtype == WIRE_BF16 && wire == WIRE_F32 && param_dtype == WIRE_BF16) { BPS_FUSED(TagBF16, TagF32, TagBF16) }
#undef BPS_FUSED
  return cudaErrorInvalidValue;
}

cudaError_t launch_barrier(const PeerView& pv, int blocks, int channel, cudaStream_t stream) {
  if (blocks < 1 || blocks > kMaxBlocks) return cudaErrorInvalidValue;
  barrier_only_kernel<<<blocks, 32, 0, stream>>>(pv, channel);
  return cudaGetLastError();
}

}  // namespace bps
