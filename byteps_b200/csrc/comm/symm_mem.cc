#include "comm/symm_mem.h"

#include <cuda.h>
#include <cuda_runtime_api.h>
#include <errno.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <stdexcept>

#include "core/env.h"
#include "core/log.h"
#include "core/types.h"

namespace bps {

namespace {

#define RT_CHECK(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));        \
  } while (0)

// Driver entry points are fetched through the runtime so the extension carries
// no link-time dependency on libcuda (it must import on a CPU-only box).
struct Drv {
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = 0;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = 0;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = 0;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = 0;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = 0;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = 0;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = 0;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = 0;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = 0;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = 0;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = 0;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = 0;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = 0;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = 0;
  CUresult (*DeviceGet)(CUdevice*, int) = 0;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = 0;
  CUresult (*GetErrorString)(CUresult, const char**) = 0;
  bool ok = false;
};

template <class F>
bool load_sym(const char* name, F* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(fn);
  return true;
}

Drv& drv() {
  static Drv d = [] {
    Drv x;
    bool ok = true;
    ok &= load_sym("cuMemGetAllocationGranularity", &x.MemGetAllocationGranularity);
    ok &= load_sym("cuMemCreate", &x.MemCreate);
    ok &= load_sym("cuMemRelease", &x.MemRelease);
    ok &= load_sym("cuMemAddressReserve", &x.MemAddressReserve);
    ok &= load_sym("cuMemAddressFree", &x.MemAddressFree);
    ok &= load_sym("cuMemMap", &x.MemMap);
    ok &= load_sym("cuMemUnmap", &x.MemUnmap);
    ok &= load_sym("cuMemSetAccess", &x.MemSetAccess);
    ok &= load_sym("cuMemExportToShareableHandle", &x.MemExportToShareableHandle);
    ok &= load_sym("cuMemImportFromShareableHandle", &x.MemImportFromShareableHandle);
    ok &= load_sym("cuDeviceGet", &x.DeviceGet);
    ok &= load_sym("cuDeviceGetAttribute", &x.DeviceGetAttribute);
    ok &= load_sym("cuGetErrorString", &x.GetErrorString);
    // multicast is optional
    load_sym("cuMulticastCreate", &x.MulticastCreate);
    load_sym("cuMulticastAddDevice", &x.MulticastAddDevice);
    load_sym("cuMulticastBindMem", &x.MulticastBindMem);
    load_sym("cuMulticastGetGranularity", &x.MulticastGetGranularity);
    x.ok = ok;
    return x;
  }();
  return d;
}

void drv_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = "?";
  if (drv().GetErrorString) drv().GetErrorString(r, &msg);
  throw std::runtime_error(std::string(what) + " failed: " + (msg ? msg : "?") + " (" + std::to_string((int)r) + ")");
}
#define DRV_CHECK(call) drv_check(drv().call, #call)

sockaddr_un abstract_addr(const std::string& name, socklen_t* len) {
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem entry to clean up
  size_t n = name.size() < sizeof(a.sun_path) - 2 ? name.size() : sizeof(a.sun_path) - 2;
  memcpy(a.sun_path + 1, name.data(), n);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}

bool send_fd(int sock, int fd) {
  char byte = 'F';
  iovec iov{&byte, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == 1;
}

int recv_fd(int sock) {
  char byte = 0;
  iovec iov{&byte, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, 0) != 1) return -1;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd = -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

}  // namespace

SymmMem::SymmMem(int rank, int world, int device, size_t data_bytes, const std::string& mode, const std::string& token)
    : rank_(rank), world_(world), device_(device), data_bytes_(round_up(data_bytes, 4096)), mode_(mode), token_(token) {
  if (world < 1 || world > kMaxRanks) throw std::runtime_error("SymmMem: world must be in [1," + std::to_string(kMaxRanks) + "]");
  RT_CHECK(cudaSetDevice(device_));
  RT_CHECK(cudaFree(0));
  peers_.assign(world_, 0);
  peer_handles_.assign(world_, 0);
  if (mode_ == "auto") mode_ = (world_ == 1) ? "local" : (drv().ok ? "vmm" : "ipc");
  if (mode_ == "vmm") {
    try {
      alloc_vmm();
    } catch (const std::exception& e) {
      if (mode != "auto") throw;
      BPS_LOG(WARNING) << "VMM symmetric allocation failed (" << e.what() << "); falling back to cudaIpc";
      mode_ = "ipc";
    }
  }
  if (mode_ == "ipc" || mode_ == "local") alloc_plain();
  else if (mode_ != "vmm") throw std::runtime_error("SymmMem: unknown mode " + mode_);
  peers_[rank_] = local_;
  RT_CHECK(cudaMemset((void*)local_, 0, alloc_bytes_));
  // private state: barrier generations + descriptor-ring bookkeeping (peer_view.h::RingState)
  RT_CHECK(cudaMalloc((void**)&epoch_, kPrivateStateBytes));
  RT_CHECK(cudaMemset(epoch_, 0, kPrivateStateBytes));
  {
    std::vector<uint32_t> ones(kRingSlots, 1u);
    RingState* rs = ring_state_of(epoch_);
    RT_CHECK(cudaMemcpy(rs->expected, ones.data(), sizeof(uint32_t) * kRingSlots, cudaMemcpyHostToDevice));
    // BYTEPS_SPIN_TIMEOUT_MS: how long a kernel waits for a peer before it traps (a crashed rank or a
    // mismatched launch order becomes a CUDA error instead of a hung GPU)
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_);
    if (khz <= 0) khz = 1900000;
    unsigned long long limit = (unsigned long long)env_int("BYTEPS_SPIN_TIMEOUT_MS", 30000) * (unsigned long long)khz;
    RT_CHECK(cudaMemcpy(&rs->spin_limit, &limit, sizeof(limit), cudaMemcpyHostToDevice));
  }
  RT_CHECK(cudaDeviceSynchronize());
}

void SymmMem::alloc_vmm() {
  if (!drv().ok) throw std::runtime_error("CUDA VMM driver entry points unavailable");
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  DRV_CHECK(MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  // multicast capability decides the size granularity
  CUdevice dev;
  DRV_CHECK(DeviceGet(&dev, device_));
  int mc = 0;
  if (world_ > 1 && drv().MulticastCreate && !env_bool("BYTEPS_DISABLE_NVLS", false))
    drv().DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  size_t need = data_bytes_ + kSignalPadBytes;
  if (mc) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = world_;
    mp.size = round_up(need, gran);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (drv().MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg) {
      mc_gran_ = mg;
      mc_supported_ = true;
      if (mg > gran) gran = mg;
    }
  }
  alloc_bytes_ = round_up(need, gran);
  CUmemGenericAllocationHandle h;
  DRV_CHECK(MemCreate(&h, alloc_bytes_, &prop, 0));
  mem_handle_ = h;
  CUdeviceptr va = 0;
  DRV_CHECK(MemAddressReserve(&va, alloc_bytes_, gran, 0, 0));
  DRV_CHECK(MemMap(va, alloc_bytes_, 0, h, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DRV_CHECK(MemSetAccess(va, alloc_bytes_, &acc, 1));
  local_ = (uintptr_t)va;
  if (world_ > 1) {
    int fd = -1;
    DRV_CHECK(MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    mem_fd_ = fd;
    sock_name_ = "bps-symm-" + token_ + "-" + std::to_string(rank_);
    start_fd_server();
  }
}

void SymmMem::alloc_plain() {
  alloc_bytes_ = round_up(data_bytes_ + kSignalPadBytes, 2 << 20);
  void* p = nullptr;
  RT_CHECK(cudaMalloc(&p, alloc_bytes_));
  local_ = (uintptr_t)p;
  if (mode_ == "ipc") {
    cudaIpcMemHandle_t h;
    RT_CHECK(cudaIpcGetMemHandle(&h, p));
    ipc_handle_.assign(reinterpret_cast<const char*>(&h), sizeof(h));
  }
}

void SymmMem::start_fd_server() {
  listen_fd_ = socket(AF_UNIX, SOCK_STREAM, 0);
  if (listen_fd_ < 0) throw std::runtime_error("SymmMem: socket() failed");
  socklen_t len;
  sockaddr_un a = abstract_addr(sock_name_, &len);
  if (bind(listen_fd_, (sockaddr*)&a, len) != 0 || listen(listen_fd_, 64) != 0) {
    close(listen_fd_);
    listen_fd_ = -1;
    throw std::runtime_error(std::string("SymmMem: bind/listen failed: ") + strerror(errno));
  }
  server_ = std::thread([this] {
    while (!stop_.load()) {
      pollfd pfd{listen_fd_, POLLIN, 0};
      int pr = poll(&pfd, 1, 100);
      if (pr <= 0) continue;
      int c = accept(listen_fd_, nullptr, nullptr);
      if (c < 0) continue;
      char what = 0;
      if (recv(c, &what, 1, 0) == 1) {
        int fd = (what == 'M') ? mem_fd_ : (what == 'C' ? mc_fd_ : -1);
        if (fd >= 0) send_fd(c, fd);
      }
      close(c);
    }
  });
}

int SymmMem::fetch_fd(const std::string& sock_name, char what) {
  auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(env_int("BYTEPS_SYMM_TIMEOUT_S", 60));
  while (std::chrono::steady_clock::now() < deadline) {
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0) break;
    socklen_t len;
    sockaddr_un a = abstract_addr(sock_name, &len);
    if (connect(s, (sockaddr*)&a, len) == 0) {
      int fd = -1;
      if (send(s, &what, 1, 0) == 1) fd = recv_fd(s);
      close(s);
      if (fd >= 0) return fd;
    } else {
      close(s);
    }
    usleep(2000);
  }
  throw std::runtime_error("SymmMem: could not fetch fd '" + std::string(1, what) + "' from " + sock_name);
}

std::string SymmMem::export_info() const {
  if (mode_ == "vmm") return sock_name_;
  if (mode_ == "ipc") return ipc_handle_;
  return "";
}

void SymmMem::import_peers(const std::vector<std::string>& infos) {
  if (world_ == 1 || mode_ == "local") return;
  if ((int)infos.size() != world_) throw std::runtime_error("import_peers: need one info per rank");
  RT_CHECK(cudaSetDevice(device_));
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    if (mode_ == "vmm") {
      int fd = fetch_fd(infos[r], 'M');
      CUmemGenericAllocationHandle h;
      DRV_CHECK(MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      close(fd);
      CUdeviceptr va = 0;
      DRV_CHECK(MemAddressReserve(&va, alloc_bytes_, 2 << 20, 0, 0));
      DRV_CHECK(MemMap(va, alloc_bytes_, 0, h, 0));
      CUmemAccessDesc acc;
      memset(&acc, 0, sizeof(acc));
      acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      acc.location.id = device_;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      DRV_CHECK(MemSetAccess(va, alloc_bytes_, &acc, 1));
      peers_[r] = (uintptr_t)va;
      peer_handles_[r] = h;
    } else {
      if (infos[r].size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("bad ipc handle size");
      cudaIpcMemHandle_t h;
      memcpy(&h, infos[r].data(), sizeof(h));
      void* p = nullptr;
      RT_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      peers_[r] = (uintptr_t)p;
    }
  }
}

std::string SymmMem::mc_create() {
  if (!mc_supported_ || mode_ != "vmm") return "";
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = world_;
  mp.size = alloc_bytes_;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  CUresult r = drv().MulticastCreate(&h, &mp);
  if (r != CUDA_SUCCESS) {
    BPS_LOG(WARNING) << "cuMulticastCreate failed (" << (int)r << "); NVLS disabled";
    return "";
  }
  mc_handle_ = h;
  int fd = -1;
  r = drv().MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    BPS_LOG(WARNING) << "multicast export failed (" << (int)r << "); NVLS disabled";
    mc_handle_ = 0;
    return "";
  }
  mc_fd_ = fd;
  return sock_name_;
}

void SymmMem::mc_join(const std::string& root_info) {
  if (root_info.empty()) throw std::runtime_error("mc_join: empty root info");
  RT_CHECK(cudaSetDevice(device_));
  if (!mc_handle_) {
    int fd = fetch_fd(root_info, 'C');
    CUmemGenericAllocationHandle h;
    DRV_CHECK(MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    close(fd);
    mc_handle_ = h;
  }
  CUdevice dev;
  DRV_CHECK(DeviceGet(&dev, device_));
  DRV_CHECK(MulticastAddDevice(mc_handle_, dev));
}

void SymmMem::mc_bind() {
  RT_CHECK(cudaSetDevice(device_));
  DRV_CHECK(MulticastBindMem(mc_handle_, 0, mem_handle_, 0, alloc_bytes_, 0));
  CUdeviceptr va = 0;
  DRV_CHECK(MemAddressReserve(&va, alloc_bytes_, mc_gran_, 0, 0));
  DRV_CHECK(MemMap(va, alloc_bytes_, 0, mc_handle_, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DRV_CHECK(MemSetAccess(va, alloc_bytes_, &acc, 1));
  mc_va_ = (uintptr_t)va;
}

PeerView SymmMem::view() const {
  PeerView pv;
  memset(&pv, 0, sizeof(pv));
  for (int r = 0; r < world_; ++r) {
    pv.data[r] = (char*)peers_[r];
    pv.sig[r] = (uint32_t*)(peers_[r] + alloc_bytes_ - kSignalPadBytes);
  }
  pv.mc_data = (char*)mc_va_;
  pv.epoch = epoch_;
  pv.rank = rank_;
  pv.world = world_;
  return pv;
}

void SymmMem::close_server() {
  stop_.store(true);
  if (server_.joinable()) server_.join();
  if (listen_fd_ >= 0) {
    close(listen_fd_);
    listen_fd_ = -1;
  }
  if (mem_fd_ >= 0) {
    close(mem_fd_);
    mem_fd_ = -1;
  }
  if (mc_fd_ >= 0) {
    close(mc_fd_);
    mc_fd_ = -1;
  }
}

SymmMem::~SymmMem() {
  close_server();
  // Unmapping peer memory while peers may still be running kernels is unsafe;
  // callers barrier before destruction.  Errors during teardown are ignored.
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  if (mode_ == "vmm" && drv().ok) {
    if (mc_va_) {
      drv().MemUnmap((CUdeviceptr)mc_va_, alloc_bytes_);
      drv().MemAddressFree((CUdeviceptr)mc_va_, alloc_bytes_);
    }
    for (int r = 0; r < world_; ++r) {
      if (r == rank_ || !peers_[r]) continue;
      drv().MemUnmap((CUdeviceptr)peers_[r], alloc_bytes_);
      drv().MemAddressFree((CUdeviceptr)peers_[r], alloc_bytes_);
      if (peer_handles_[r]) drv().MemRelease(peer_handles_[r]);
    }
    if (mc_handle_) drv().MemRelease(mc_handle_);
    if (local_) {
      drv().MemUnmap((CUdeviceptr)local_, alloc_bytes_);
      drv().MemAddressFree((CUdeviceptr)local_, alloc_bytes_);
    }
    if (mem_handle_) drv().MemRelease(mem_handle_);
  } else {
    if (mode_ == "ipc") {
      for (int r = 0; r < world_; ++r)
        if (r != rank_ && peers_[r]) cudaIpcCloseMemHandle((void*)peers_[r]);
    }
    if (local_) cudaFree((void*)local_);
  }
  if (epoch_) cudaFree(epoch_);
}

}  // namespace bps
