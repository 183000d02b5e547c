// Symmetric, peer-mapped device memory for one-process-per-GPU jobs.
//
// Replaces the reference's intra-node plumbing - NCCL communicators
// (/root/reference/byteps/common/nccl_manager.cc:74-165), UDS READY/DO_* signals
// (communicator.cc:128-276) and the cudaHostRegister'ed POSIX shm staging
// (shared_memory.cc:28-82) - with: one device allocation per rank that every
// peer maps into its own address space (CUDA VMM handles passed as file
// descriptors over a Unix socket, or legacy cudaIpc handles), an optional
// NVLS multicast alias of the same memory, and a signal pad of flags that the
// kernels use to synchronise directly over NVLink.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "kernels/peer_view.h"

namespace bps {

class SymmMem {
 public:
  // mode: "auto" | "vmm" | "ipc" | "local".  token: job-unique string used for
  // the Unix socket names.
  SymmMem(int rank, int world, int device, size_t data_bytes, const std::string& mode, const std::string& token);
  ~SymmMem();
  SymmMem(const SymmMem&) = delete;

  const std::string& mode() const { return mode_; }
  // Opaque per-rank blob to all-gather (ipc: 64-byte handle, vmm: socket name).
  std::string export_info() const;
  // Map every peer.  infos[r] is rank r's export_info().
  void import_peers(const std::vector<std::string>& infos);

  // NVLS multicast (vmm mode only).  Call order on every rank:
  //   mc_supported() -> [rank 0] mc_create() -> all: mc_join(root_info) -> job barrier -> mc_bind()
  bool mc_supported() const { return mc_supported_; }
  std::string mc_create();
  void mc_join(const std::string& root_info);
  void mc_bind();
  bool has_multicast() const { return mc_va_ != 0; }

  PeerView view() const;
  void* local_ptr() const { return (void*)local_; }
  void* peer_ptr(int r) const { return (void*)peers_[r]; }
  void* mc_ptr() const { return (void*)mc_va_; }
  size_t data_bytes() const { return data_bytes_; }
  size_t alloc_bytes() const { return alloc_bytes_; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  int device() const { return device_; }
  void close_server();

 private:
  void alloc_vmm();
  void alloc_plain();
  void start_fd_server();
  int fetch_fd(const std::string& sock_name, char what);

  int rank_, world_, device_;
  size_t data_bytes_, alloc_bytes_ = 0;
  std::string mode_, token_, sock_name_;
  uintptr_t local_ = 0;
  std::vector<uintptr_t> peers_;
  std::vector<unsigned long long> peer_handles_;  // imported CUmemGenericAllocationHandle
  unsigned long long mem_handle_ = 0;
  int mem_fd_ = -1;
  // multicast
  bool mc_supported_ = false;
  size_t mc_gran_ = 0;
  unsigned long long mc_handle_ = 0;
  int mc_fd_ = -1;
  uintptr_t mc_va_ = 0;
  // fd server
  int listen_fd_ = -1;
  std::thread server_;
  std::atomic<bool> stop_{false};
  uint32_t* epoch_ = nullptr;
  std::string ipc_handle_;
};

// number of bytes reserved at the end of every allocation for flags
constexpr size_t kSignalPadBytes = 1 << 18;
static_assert(kRingPadEnd <= kSignalPadBytes, "signal pad too small");

}  // namespace bps
