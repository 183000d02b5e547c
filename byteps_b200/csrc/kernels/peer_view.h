// Host-safe definitions shared by the kernels and the symmetric-memory setup.
#pragma once
#include <stdint.h>

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

}  // namespace bps
