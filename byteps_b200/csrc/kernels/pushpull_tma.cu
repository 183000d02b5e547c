// TMA-fed variant of the in-place push-pull (P2P reduce-scatter + all-gather).
//
// The LSU kernel in pushpull.cu keeps 16 x 16-byte peer loads in flight per
// thread in registers.  Here the copy engine of the SM does the NVLink traffic:
// one elected producer thread issues `cp.async.bulk` (1-D TMA, SASS UBLKCP)
// copies of every peer's slice of the tile into a multi-stage shared-memory
// ring, completion is counted on mbarriers, consumer warps only add up rows of
// shared memory, and the reduced tile leaves through bulk stores to every
// peer.  In flight per CTA: STAGES x world x TILE bytes, independent of the
// register file; consumers never stall on NVLink latency.
#include "kernels/common.cuh"
#include "kernels/pushpull.cuh"

namespace bps {

namespace {

constexpr int kTileUnits = 256;                 // 16-byte units per peer per tile (4 KiB)
constexpr int kTileBytes = kTileUnits * 16;
constexpr int kConsumers = kTileUnits;          // one unit per consumer thread
constexpr int kThreadsTma = kConsumers + 32;    // + producer warp
constexpr int kMaxStages = 4;

struct TmaSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
};

template <class W>
__global__ void __launch_bounds__(kThreadsTma) pushpull_inplace_tma_kernel(PeerView pv, size_t off,
                                                                           size_t total_groups, float scale,
                                                                           int stages, int channel) {
  constexpr int E = W::kPerVec;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TmaSmem* sm = reinterpret_cast<TmaSmem*>(smem_raw);
  unsigned char* ring = smem_raw + 128;                       // [stage][world + 1][kTileBytes]
  const int world = pv.world;
  const size_t stage_bytes = (size_t)(world + 1) * kTileBytes;
  const int warp = threadIdx.x >> 5;
  const bool is_producer = warp == (kConsumers >> 5);

  size_t b, e;
  shard_units(total_groups, world, pv.rank, &b, &e);
  const size_t s0 = b * (8 / E), s1 = e * (8 / E);           // my shard in 16-byte units

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], 1);
    }
    mbar_fence_init();
  }
  barrier_peers(pv, channel);    // includes __syncthreads: barriers initialised, every peer's input is ready
  fence_proxy_async();           // order the async-proxy reads below after the acquire above

  const size_t first = s0 + (size_t)blockIdx.x * kTileUnits;
  const size_t stride = (size_t)gridDim.x * kTileUnits;
  const int rot = pv.rank + 1 >= world ? 0 : pv.rank + 1;

  if (is_producer) {
    if ((threadIdx.x & 31) == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (size_t t = first; t < s1; t += stride) {
        const uint32_t units = (uint32_t)((s1 - t) < (size_t)kTileUnits ? (s1 - t) : (size_t)kTileUnits);
        const uint32_t bytes = units * 16;
        mbar_wait(&sm->empty[s], phase ^ 1);                 // slot free (first pass falls through)
        mbar_arrive_expect_tx(&sm->full[s], bytes * world);
        unsigned char* dst = ring + (size_t)s * stage_bytes;
        for (int j = 0; j < world; ++j) {
          int p = j + rot;
          if (p >= world) p -= world;
          bulk_g2s(dst + (size_t)p * kTileBytes, pv.data[p] + off + t * 16, bytes, &sm->full[s]);
        }
        if (++s == stages) {
          s = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    int s = 0;
    uint32_t phase = 0;
    for (size_t t = first; t < s1; t += stride) {
      const uint32_t units = (uint32_t)((s1 - t) < (size_t)kTileUnits ? (s1 - t) : (size_t)kTileUnits);
      unsigned char* buf = ring + (size_t)s * stage_bytes;
      unsigned char* outp = buf + (size_t)world * kTileBytes;
      mbar_wait(&sm->full[s], phase);                        // all peers' slices have landed
      if (threadIdx.x < units) {
        float acc[E];
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] = 0.f;
        for (int p = 0; p < world; ++p) {                   // fixed order: bit-reproducible
          float f[E];
          W::unpack(lds16(buf + (size_t)p * kTileBytes + threadIdx.x * 16), f);
#pragma unroll
          for (int k = 0; k < E; ++k) acc[k] += f[k];
        }
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] *= scale;
        sts16(outp + threadIdx.x * 16, W::pack(acc));
      }
      fence_proxy_async_smem();                              // generic-proxy writes -> visible to the bulk store
      named_bar_sync(1, kConsumers);
      if (threadIdx.x == 0) {
        const uint32_t bytes = units * 16;
        for (int j = 0; j < world; ++j) {
          int p = j + pv.rank;
          if (p >= world) p -= world;
          bulk_s2g(pv.data[p] + off + t * 16, outp, bytes);
        }
        bulk_commit();
        bulk_wait_read<0>();                                 // the out tile has been read: slot reusable
        mbar_arrive(&sm->empty[s]);
      }
      if (++s == stages) {
        s = 0;
        phase ^= 1;
      }
    }
    if (threadIdx.x == 0) {
      bulk_wait<0>();                                        // every bulk store of this CTA is complete
      fence_proxy_async();
    }
  }
  barrier_peers(pv, channel);
}

}  // namespace

cudaError_t launch_pushpull_inplace_tma(const PeerView& pv, int wire, size_t off, size_t nelem, float scale,
                                        int blocks, int stages, int channel, cudaStream_t stream) {
  if (blocks < 1 || blocks > kMaxBlocks || (off & 15) || stages < 1 || stages > kMaxStages)
    return cudaErrorInvalidValue;
  const size_t smem = 128 + (size_t)stages * (pv.world + 1) * kTileBytes;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const size_t groups = (nelem + 7) / 8;
  cudaError_t err;
#define BPS_TMA(W)                                                                                             \
  err = cudaFuncSetAttribute(pushpull_inplace_tma_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             (int)smem);                                                                       \
  if (err != cudaSuccess) return err;                                                                          \
  pushpull_inplace_tma_kernel<W><<<blocks, kThreadsTma, smem, stream>>>(pv, off, groups, scale, stages, channel); \
  return cudaGetLastError();
  switch (wire) {
    case WIRE_F32: BPS_TMA(TagF32)
    case WIRE_BF16: BPS_TMA(TagBF16)
    case WIRE_F16: BPS_TMA(TagF16)
  }
#undef BPS_TMA
  return cudaErrorInvalidValue;
}

}  // namespace bps
