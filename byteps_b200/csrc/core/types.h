// Common value types for the byteps_b200 runtime.
//
// Parity: /root/reference/byteps/common/common.h:59-118 (DataType, QueueType,
// StatusType), common.cc:98-144 (command pairing, dtype lengths), with bf16
// added as a first-class dtype (the reference has none).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace bps {

// First seven values keep the reference/mshadow numbering so that wire
// commands stay interpretable; BF16 is new.
enum DataType : int {
  F32 = 0,
  F64 = 1,
  F16 = 2,
  U8 = 3,
  I32 = 4,
  I8 = 5,
  I64 = 6,
  BF16 = 7,
  DTYPE_COUNT = 8,
};

inline int dtype_size(int d) {
  switch (d) {
    case U8: case I8: return 1;
    case F16: case BF16: return 2;
    case F32: case I32: return 4;
    case F64: case I64: return 8;
    default: return 0;
  }
}

inline const char* dtype_name(int d) {
  static const char* n[] = {"float32", "float64", "float16", "uint8", "int32", "int8", "int64", "bfloat16"};
  return (d >= 0 && d < DTYPE_COUNT) ? n[d] : "invalid";
}

inline bool dtype_is_float(int d) { return d == F32 || d == F64 || d == F16 || d == BF16; }

// Pipeline stages a partition can flow through.  Names match the reference's
// QueueType so timeline traces remain comparable; how each stage is executed
// differs (GPU stages are stream-ordered kernels, not NCCL calls).
enum Stage : int {
  COORDINATE_REDUCE = 0,
  REDUCE,
  COPYD2H,
  PCIE_REDUCE,
  COORDINATE_PUSH,
  COMPRESS,
  PUSH,
  PULL,
  DECOMPRESS,
  COPYH2D,
  COORDINATE_BROADCAST,
  BROADCAST,
  STAGE_COUNT
};

inline const char* stage_name(int s) {
  static const char* n[] = {"COORDINATE_REDUCE", "REDUCE", "COPYD2H", "PCIE_REDUCE", "COORDINATE_PUSH",
                            "COMPRESS", "PUSH", "PULL", "DECOMPRESS", "COPYH2D", "COORDINATE_BROADCAST",
                            "BROADCAST"};
  return (s >= 0 && s < STAGE_COUNT) ? n[s] : "?";
}

enum StatusCode : int { ST_OK = 0, ST_UNKNOWN, ST_PRECONDITION, ST_ABORTED, ST_INVALID_ARGUMENT, ST_IN_PROGRESS };

struct Status {
  StatusCode code = ST_OK;
  std::string reason;
  static Status OK() { return Status{}; }
  static Status InProgress() { return Status{ST_IN_PROGRESS, ""}; }
  static Status Error(StatusCode c, std::string r) { return Status{c, std::move(r)}; }
  bool ok() const { return code == ST_OK; }
  bool in_progress() const { return code == ST_IN_PROGRESS; }
};

enum RequestType : int { kDefaultPushPull = 0, kRowSparsePushPull = 1, kCompressedPushPull = 2 };

// Cantor pairing of (request type, dtype) -> one int carried in the wire header.
inline int command_encode(int req, int dtype) { return ((req + dtype) * (req + dtype + 1)) / 2 + dtype; }
inline void command_decode(int cmd, int* req, int* dtype) {
  int w = 0;
  while ((w + 1) * (w + 2) / 2 <= cmd) ++w;
  int t = w * (w + 1) / 2;
  *dtype = cmd - t;
  *req = w - *dtype;
}

inline size_t round_up(size_t v, size_t m) { return m ? ((v + m - 1) / m) * m : v; }

// Compressed payloads are padded so that trailing metadata words stay aligned.
inline size_t align_payload(size_t size, int dtype) {
  size_t m = static_cast<size_t>(dtype_size(dtype)) * dtype_size(dtype) * 8;
  return round_up(size, m ? m : 8);
}

// key layout: (declared_key << 16) | partition_index
inline uint64_t make_key(uint32_t declared, uint32_t part) { return (static_cast<uint64_t>(declared) << 16) | (part & 0xffffu); }
inline uint32_t key_declared(uint64_t k) { return static_cast<uint32_t>(k >> 16); }
inline uint32_t key_part(uint64_t k) { return static_cast<uint32_t>(k & 0xffffu); }

struct Partition {
  size_t offset;
  size_t len;
};

// Split [0,size) into chunks of at most `bound` bytes (bound is pre-aligned by
// the caller to local_size*page so shards stay vector-aligned).
inline std::vector<Partition> partition_bytes(size_t size, size_t bound) {
  std::vector<Partition> out;
  if (bound == 0) bound = size ? size : 1;
  size_t off = 0;
  while (off < size) {
    size_t len = (size - off < bound) ? (size - off) : bound;
    out.push_back({off, len});
    off += len;
  }
  if (out.empty()) out.push_back({0, 0});
  return out;
}

}  // namespace bps
