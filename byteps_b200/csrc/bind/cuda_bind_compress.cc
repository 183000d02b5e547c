// pybind11 surface of the fused GPU compressors (kernels/compress_fused.cu).
#include <cuda_runtime_api.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bind/cuda_bind_ext.h"
#include "kernels/compress_fused.cuh"

namespace py = pybind11;
using namespace bps;

namespace {
void chk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// xorshift128+ state transition as a 128x128 matrix over GF(2) (row r = the input bits that feed
// output bit r), and its powers T^(2^j): what the device needs to jump ahead in the CPU
// compressor's random stream (compress/compressor.cc; reference utils.h:74-113).
struct BitMat {
  uint64_t row[128][2];
};

void xs_step(uint64_t& a, uint64_t& b) {
  uint64_t t = a;
  const uint64_t s = b;
  a = s;
  t ^= t << 23;
  t ^= t >> 17;
  t ^= s ^ (s >> 26);
  b = t;
}

BitMat transition() {
  BitMat m;
  memset(&m, 0, sizeof(m));
  for (int c = 0; c < 128; ++c) {
    uint64_t a = c < 64 ? (1ull << c) : 0, b = c >= 64 ? (1ull << (c - 64)) : 0;
    xs_step(a, b);
    for (int r = 0; r < 128; ++r) {
      const bool bit = r < 64 ? ((a >> r) & 1) : ((b >> (r - 64)) & 1);
      if (bit) m.row[r][c >> 6] |= 1ull << (c & 63);
    }
  }
  return m;
}

BitMat square(const BitMat& m) {
  BitMat o;
  memset(&o, 0, sizeof(o));
  for (int r = 0; r < 128; ++r) {
    uint64_t x0 = 0, x1 = 0;
    for (int c = 0; c < 128; ++c)
      if ((m.row[r][c >> 6] >> (c & 63)) & 1) {
        x0 ^= m.row[c][0];
        x1 ^= m.row[c][1];
      }
    o.row[r][0] = x0;
    o.row[r][1] = x1;
  }
  return o;
}
}  // namespace

void bind_cuda_compress(py::module_& m) {
  m.attr("FUSED_MAX_BLOCKS") = kFusedMaxBlocks;
  m.attr("TOPK_SCRATCH_BYTES") = kTopkScratchBytes;

  m.def("xorshift_jump_table", []() {
    std::string out;
    BitMat t = transition();
    for (int j = 0; j < 32; ++j) {
      out.append(reinterpret_cast<const char*>(&t), sizeof(t));
      t = square(t);
    }
    return py::bytes(out);
  }, "T^(2^j), j < 32, of the xorshift128+ transition: 32 x 128 rows x 2 uint64");

  m.def("onebit_pre", [](uintptr_t g, int dtype, uintptr_t mom, float mu, uintptr_t err, float ratio, uintptr_t p_out,
                         size_t n, uintptr_t words, bool use_scale, uintptr_t parts, uintptr_t counter, uintptr_t s) {
    chk(launch_onebit_pre((const void*)g, dtype, (float*)mom, mu, (const float*)err, ratio, (float*)p_out, n,
                          (uint32_t*)words, use_scale ? 1 : 0, (float*)parts, (uint32_t*)counter, (cudaStream_t)s),
        "onebit_pre");
  });
  m.def("onebit_sum", [](uintptr_t slots, size_t slot_bytes, int world, int me, size_t n, uintptr_t err, uintptr_t err2,
                         uintptr_t c2_out, uintptr_t out, int dtype, float mult, bool use_scale2, uintptr_t parts,
                         uintptr_t counter, uintptr_t scale2_out, uintptr_t s) {
    chk(launch_onebit_sum((const void*)slots, slot_bytes, world, me, n, (float*)err, (const float*)err2,
                          (float*)c2_out, (void*)out, dtype, mult, use_scale2 ? 1 : 0, (float*)parts,
                          (uint32_t*)counter, (float*)scale2_out, (cudaStream_t)s),
        "onebit_sum");
  });
  m.def("onebit_out", [](uintptr_t c2, size_t n, uintptr_t scale2, uintptr_t err2, uintptr_t out, int dtype, float mult,
                         uintptr_t s) {
    chk(launch_onebit_out((const float*)c2, n, (const float*)scale2, (float*)err2, (void*)out, dtype, mult,
                          (cudaStream_t)s),
        "onebit_out");
  });
  m.def("payload_push", [](const PeerView& pv, size_t win_off, size_t slot_bytes, size_t bytes, int blocks, int channel,
                           uintptr_t s) {
    chk(launch_payload_push(pv, win_off, slot_bytes, bytes, blocks, channel, (cudaStream_t)s), "payload_push");
  });
  m.def("topk_pre", [](uintptr_t g, int dtype, uintptr_t mom, float mu, uintptr_t err, float ratio, uintptr_t p_out,
                       size_t n, uint32_t k, uintptr_t scratch, uintptr_t s) {
    chk(launch_topk_pre((const void*)g, dtype, (float*)mom, mu, (const float*)err, ratio, (float*)p_out, n, k,
                        (void*)scratch, (cudaStream_t)s),
        "topk_pre");
  });
  m.def("topk_finish", [](uintptr_t x, size_t n, uint32_t k, int first_level, uintptr_t pairs, bool zero_kept,
                          uintptr_t scratch, uintptr_t s) {
    chk(launch_topk_finish((float*)x, n, k, first_level, (uint32_t*)pairs, zero_kept ? 1 : 0, (void*)scratch,
                           (cudaStream_t)s),
        "topk_finish");
  });
  m.def("sparse_add_pairs", [](uintptr_t pairs, uint32_t k, size_t n, uintptr_t dst, uintptr_t s) {
    chk(launch_sparse_add_pairs((const uint32_t*)pairs, k, n, (float*)dst, (cudaStream_t)s), "sparse_add_pairs");
  });
  m.def("scatter_pairs", [](uintptr_t pairs, uint32_t k, size_t n, uintptr_t out, int dtype, float mult, uintptr_t s) {
    chk(launch_scatter_pairs((const uint32_t*)pairs, k, n, (void*)out, dtype, mult, (cudaStream_t)s), "scatter_pairs");
  });
  m.def("cast_scale4", [](uintptr_t in, size_t n, uintptr_t out, int dtype, float mult, uintptr_t s) {
    chk(launch_cast_scale4((const float*)in, n, (void*)out, dtype, mult, (cudaStream_t)s), "cast_scale4");
  });
  m.def("randomk_draw", [](uintptr_t state, uintptr_t jump, uint32_t k, size_t n, uintptr_t idx, uintptr_t s) {
    chk(launch_randomk_draw((uint64_t*)state, (const uint64_t*)jump, k, n, (uint32_t*)idx, (cudaStream_t)s),
        "randomk_draw");
  });
  m.def("randomk_pre", [](uintptr_t g, int dtype, uintptr_t mom, float mu, uintptr_t err, float ratio, size_t n,
                          uintptr_t idx, uint32_t k, uintptr_t vals, uintptr_t s) {
    chk(launch_randomk_pre((const void*)g, dtype, (float*)mom, mu, (float*)err, ratio, n, (const uint32_t*)idx, k,
                           (float*)vals, (cudaStream_t)s),
        "randomk_pre");
  });
  m.def("dither_sum_slots", [](uintptr_t slots, size_t slot_bytes, int world, size_t n, int s_levels, int partition,
                               uintptr_t sum, uintptr_t s) {
    chk(launch_dither_sum_slots((const void*)slots, slot_bytes, world, n, s_levels, partition, (float*)sum,
                                (cudaStream_t)s),
        "dither_sum_slots");
  });
  m.def("dense_sum_slots", [](uintptr_t slots, size_t slot_bytes, int world, uint32_t k, uintptr_t out, uintptr_t s) {
    chk(launch_dense_sum_slots((const void*)slots, slot_bytes, world, k, (float*)out, (cudaStream_t)s),
        "dense_sum_slots");
  });
}
