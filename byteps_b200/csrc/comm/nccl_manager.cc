#include "comm/nccl_manager.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <cstring>
#include <stdexcept>

#include "core/types.h"

namespace bps {

namespace {

struct NcclUniqueId {
  char internal[128];
};
using ncclComm_t = void*;

struct Api {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Api& api() {
  static Api a = [] {
    Api x;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return x;
    auto sym = [h](const char* n) { return dlsym(h, n); };
    x.GetUniqueId = (decltype(x.GetUniqueId))sym("ncclGetUniqueId");
    x.CommInitRank = (decltype(x.CommInitRank))sym("ncclCommInitRank");
    x.CommDestroy = (decltype(x.CommDestroy))sym("ncclCommDestroy");
    x.ReduceScatter = (decltype(x.ReduceScatter))sym("ncclReduceScatter");
    x.AllGather = (decltype(x.AllGather))sym("ncclAllGather");
    x.Reduce = (decltype(x.Reduce))sym("ncclReduce");
    x.Broadcast = (decltype(x.Broadcast))sym("ncclBroadcast");
    x.GroupStart = (decltype(x.GroupStart))sym("ncclGroupStart");
    x.GroupEnd = (decltype(x.GroupEnd))sym("ncclGroupEnd");
    x.GetErrorString = (decltype(x.GetErrorString))sym("ncclGetErrorString");
    x.ok = x.GetUniqueId && x.CommInitRank && x.CommDestroy && x.ReduceScatter && x.AllGather && x.Reduce &&
           x.Broadcast && x.GroupStart && x.GroupEnd;
    return x;
  }();
  return a;
}

void nccl_check(int r, const char* what) {
  if (r == 0) return;
  const char* msg = api().GetErrorString ? api().GetErrorString(r) : "?";
  throw std::runtime_error(std::string(what) + " failed: " + msg);
}
void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ncclDataType_t values
int nccl_dtype(int d) {
  switch (d) {
    case I8: return 0;
    case U8: return 1;
    case I32: return 2;
    case I64: return 4;
    case F16: return 6;
    case F32: return 7;
    case F64: return 8;
    case BF16: return 9;
    default: throw std::runtime_error("unsupported dtype for NCCL");
  }
}
constexpr int kNcclSum = 0;

}  // namespace

bool NcclManager::available() { return api().ok; }

std::string NcclManager::make_unique_id() {
  if (!api().ok) throw std::runtime_error("libnccl.so.2 not available");
  NcclUniqueId id;
  nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
  return std::string(id.internal, sizeof(id.internal));
}

NcclManager::NcclManager(int rank, int world, int device, int num_rings, int group_size)
    : rank_(rank), world_(world), device_(device), num_rings_(num_rings < 1 ? 1 : num_rings),
      group_size_(group_size < 1 ? 1 : group_size) {}

void NcclManager::init(const std::vector<std::string>& ids) {
  if (!api().ok) throw std::runtime_error("libnccl.so.2 not available");
  if ((int)ids.size() != num_rings_) throw std::runtime_error("need one unique id per ring");
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  int lo, hi;
  cuda_check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "priority range");
  for (int r = 0; r < num_rings_; ++r) {
    NcclUniqueId id;
    if (ids[r].size() != sizeof(id.internal)) throw std::runtime_error("bad unique id size");
    memcpy(id.internal, ids[r].data(), sizeof(id.internal));
    ncclComm_t c = nullptr;
    nccl_check(api().CommInitRank(&c, world_, id, rank_), "ncclCommInitRank");
    comms_.push_back(c);
    cudaStream_t s;
    cuda_check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi), "stream");   // highest priority
    streams_.push_back(s);
  }
}

NcclManager::~NcclManager() {
  cudaSetDevice(device_);
  for (auto s : streams_) {
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
  }
  for (auto e : events_) cudaEventDestroy(e);
  for (auto c : comms_)
    if (c && api().ok) api().CommDestroy(c);
}

void NcclManager::push_pull(void* ptr, size_t nbytes, int dtype, uint64_t first_key, size_t part_bytes,
                            cudaEvent_t ready, cudaEvent_t done) {
  const int es = dtype_size(dtype);
  const int nd = nccl_dtype(dtype);
  // bound rounded up to world*page like the reference (global.cc:142)
  const size_t page = 4096 * (size_t)world_;
  const size_t bound = round_up(part_bytes ? part_bytes : nbytes, page);
  auto parts = ::bps::partition_bytes(nbytes, bound);
  if (ready)
    for (auto s : streams_) cuda_check(cudaStreamWaitEvent(s, ready, 0), "wait ready");
  char* base = static_cast<char*>(ptr);
  const int N = world_, root = world_ - 1;
  for (size_t g0 = 0; g0 < parts.size(); g0 += group_size_) {
    const size_t g1 = g0 + group_size_ < parts.size() ? g0 + group_size_ : parts.size();
    for (int phase = 0; phase < 2; ++phase) {   // REDUCE group, then BROADCAST group
      if (N > 1) nccl_check(api().GroupStart(), "ncclGroupStart");
      for (size_t i = g0; i < g1; ++i) {
        const uint64_t key = first_key + i;
        const int ring = (int)(key % num_rings_);
        char* p = base + parts[i].offset;
        const size_t elems = parts[i].len / es;
        const size_t per = elems / N, tail = elems % N;
        if (N == 1) continue;   // one GPU: the reference skips PostNcclCalls as well
        if (phase == 0) {
          if (per) nccl_check(api().ReduceScatter(p, p + (size_t)rank_ * per * es, per, nd, kNcclSum, comms_[ring], streams_[ring]), "ncclReduceScatter");
          if (tail) nccl_check(api().Reduce(p + per * N * es, p + per * N * es, tail, nd, kNcclSum, root, comms_[ring], streams_[ring]), "ncclReduce");
        } else {
          if (per) nccl_check(api().AllGather(p + (size_t)rank_ * per * es, p, per, nd, comms_[ring], streams_[ring]), "ncclAllGather");
          if (tail) nccl_check(api().Broadcast(p + per * N * es, p + per * N * es, tail, nd, root, comms_[ring], streams_[ring]), "ncclBroadcast");
        }
        ++tasks_;
      }
      if (N > 1) nccl_check(api().GroupEnd(), "ncclGroupEnd");
    }
  }
  if (done) {
    // every ring stream must have finished: chain them into ring 0
    for (size_t r = 1; r < streams_.size(); ++r) {
      if (events_.size() < r) {
        cudaEvent_t e;
        cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
        events_.push_back(e);
      }
      cuda_check(cudaEventRecord(events_[r - 1], streams_[r]), "record");
      cuda_check(cudaStreamWaitEvent(streams_[0], events_[r - 1], 0), "wait");
    }
    cuda_check(cudaEventRecord(done, streams_[0]), "record done");
  }
}

}  // namespace bps
