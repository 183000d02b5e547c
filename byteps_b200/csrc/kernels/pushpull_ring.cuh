// Host-visible API of the descriptor-ring exchange kernel (pushpull_ring.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels/peer_view.h"
#include "kernels/pushpull.cuh"

namespace bps {

enum RingKind : int { RING_ALLREDUCE = 0, RING_SGD = 1, RING_ADAM = 2 };

// One bucket / partition.  The table lives in device memory; entry i must describe the same
// window, slot and priority on every rank (pointers are rank-local).
struct RingDesc {
  uint64_t grad_off;    // byte offset of the gradient window in the symmetric arena (wire dtype)
  uint64_t param_off;   // byte offset of the parameter window (fused kinds)
  uint64_t groups;      // element count / 8
  uint64_t bytes;       // wire bytes (credit accounting)
  float* master;        // fp32 master shard / optimizer state of THIS rank (fused kinds)
  float* state0;
  float* state1;
  const OptHParams* hp;
  float scale;
  int32_t kind;         // RingKind
  int32_t wire;         // WireDType
  int32_t priority;     // higher goes first when scheduling is on
  uint32_t slot;        // signal slot, < kRingSlots, unique inside a launch
  uint32_t pad[3];
};
static_assert(sizeof(RingDesc) == 96, "RingDesc layout is part of the python binding");

constexpr int kRingMarkBatch = 32;
struct RingSlotList {
  int n;
  uint32_t slot[kRingMarkBatch];
};

// One launch consumes descs[0..n).  blocks = worker CTAs (one more CTA is added for the scheduler
// when sched != 0).  self_mark: the launch marks its own descriptors ready at entry (everything
// was produced earlier in stream order); otherwise ring_mark must be issued by the producers.
// credit_bytes: scheduling window (0 = unlimited), BYTEPS_SCHEDULING_CREDIT x partition bytes.
// All descriptors of a launch share one (wire dtype, kind) class; RingDesc::wire/kind are checked
// by the host wrapper.
cudaError_t launch_pushpull_ring(const PeerView& pv, int wire, int kind, const RingDesc* descs, int n, int blocks,
                                 int use_nvls, int sched, int self_mark, unsigned long long credit_bytes,
                                 cudaStream_t stream, int solo = 0);

// Publish "my gradients for these slots are complete" to every rank, ordered after everything
// already enqueued on `stream`.
cudaError_t launch_ring_mark(const PeerView& pv, const uint32_t* slots, int n, cudaStream_t stream);

// Write the device's globaltimer into RingState::stamps[idx] (end-of-backward marker for the
// exposed-communication measurement; works inside CUDA graphs).
cudaError_t launch_ring_stamp(const PeerView& pv, int idx, cudaStream_t stream);

// Force-load every kernel of the ring module (see the lazy-loading note in pushpull_ring.cu).
cudaError_t ring_preload();

}  // namespace bps
