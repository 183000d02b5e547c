// Host-safe definitions shared by the kernels and the symmetric-memory setup.
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace bps {

constexpr int kMaxRanks = 16;        // signal-pad stride (8 GPUs per box today)
constexpr int kMaxBlocks = 512;      // max CTAs that take part in a cross-rank barrier
constexpr int kSigBytes = kMaxBlocks * kMaxRanks * 4 * 2;  // two independent channels

// What a kernel needs to talk to its peers.  All pointers are valid in THIS
// process' address space (peer buffers are IPC/VMM-mapped at setup time).
struct PeerView {
  char* data[kMaxRanks];      // symmetric data region of every rank
  uint32_t* sig[kMaxRanks];   // signal pad of every rank: [channel][block][src_rank]
  char* mc_data;              // multicast alias of the data region (nullptr: no NVLS)
  uint32_t* epoch;            // local, private: [channel][block] barrier generation
  int rank;
  int world;
};

// ---- descriptor ring (pushpull_ring.cu) ------------------------------------------------------
// Peer-visible flags live in the signal pad after the two barrier channels; every value is a
// monotonically increasing per-slot generation, so nothing is ever reset and the state
// survives CUDA-graph replays.
constexpr int kRingSlots = 1024;
constexpr size_t kRingReadyOff = kSigBytes;                                        // [slot][src rank] u32
constexpr size_t kRingDoneOff = kRingReadyOff + (size_t)kRingSlots * kMaxRanks * 4;  // [slot][src rank] u32
constexpr size_t kRingOrderOff = kRingDoneOff + (size_t)kRingSlots * kMaxRanks * 4;  // [position] u64, root -> all
constexpr size_t kRingPadEnd = kRingOrderOff + (size_t)kRingSlots * 8;
constexpr int kRingStamps = 16;

// Private (not peer-mapped) per-rank state; allocated right behind the barrier epochs.
struct RingState {
  uint32_t expected[kRingSlots];   // generation at which each slot is consumed next (starts at 1)
  uint32_t marked[kRingSlots];     // generation this rank last published for the slot
  uint32_t arrive[kRingSlots];     // CTAs that finished the slot in the current launch
  uint32_t order_pos[kRingSlots];  // trace: position at which the slot was processed in the last launch
  unsigned long long t_start[kRingSlots];   // trace: globaltimer (ns) when the first CTA picked the slot up
  unsigned long long t_end[kRingSlots];     // trace: ... when the last CTA finished it
  unsigned long long stamps[kRingStamps];   // user stamps (ring_stamp): e.g. end of backward
  unsigned long long done_bytes;   // bytes completed in the current launch (credit accounting on the root)
  unsigned long long spin_limit;   // watchdog, in SM clocks (0 = built-in default)
  uint32_t launch_id;              // completed ring launches
  uint32_t exit_count;             // CTAs that left the current launch
};

__host__ __device__ inline RingState* ring_state_of(uint32_t* epoch) {
  return reinterpret_cast<RingState*>(epoch + 2 * kMaxBlocks);
}
constexpr size_t kPrivateStateBytes = 2 * kMaxBlocks * sizeof(uint32_t) + sizeof(RingState);

}  // namespace bps
