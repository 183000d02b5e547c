// pybind11 surface of the descriptor-ring exchange (kernels/pushpull_ring.cu).
#include <cuda_runtime_api.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "bind/cuda_bind_ext.h"
#include "kernels/pushpull_ring.cuh"

namespace py = pybind11;
using namespace bps;

namespace {
void chk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

void bind_cuda_ring(py::module_& m) {
  m.attr("RING_SLOTS") = kRingSlots;
  m.attr("RING_DESC_BYTES") = (int)sizeof(RingDesc);
  m.attr("RING_ALLREDUCE") = (int)RING_ALLREDUCE;
  m.attr("RING_SGD") = (int)RING_SGD;
  m.attr("RING_ADAM") = (int)RING_ADAM;
  m.attr("RING_STAMPS") = kRingStamps;

  m.def(
      "pushpull_ring",
      [](const PeerView& pv, int wire, int kind, uintptr_t descs, int n, int blocks, bool nvls, bool sched,
         bool self_mark, unsigned long long credit_bytes, uintptr_t stream, bool solo) {
        chk(launch_pushpull_ring(pv, wire, kind, (const RingDesc*)descs, n, blocks, nvls ? 1 : 0, sched ? 1 : 0,
                                 self_mark ? 1 : 0, credit_bytes, (cudaStream_t)stream, solo ? 1 : 0),
            "pushpull_ring");
      },
      py::arg("view"), py::arg("wire"), py::arg("kind"), py::arg("descs"), py::arg("n"), py::arg("blocks"),
      py::arg("nvls") = false, py::arg("sched") = false, py::arg("self_mark") = true, py::arg("credit_bytes") = 0,
      py::arg("stream") = 0, py::arg("solo") = false,
      "One launch consumes a device table of n RingDesc entries (same wire dtype and kind)");

  m.def("ring_preload", []() { chk(ring_preload(), "ring_preload"); });

  m.def(
      "ring_mark",
      [](const PeerView& pv, const std::vector<uint32_t>& slots, uintptr_t stream) {
        chk(launch_ring_mark(pv, slots.data(), (int)slots.size(), (cudaStream_t)stream), "ring_mark");
      },
      py::arg("view"), py::arg("slots"), py::arg("stream") = 0);

  m.def(
      "ring_stamp",
      [](const PeerView& pv, int idx, uintptr_t stream) {
        chk(launch_ring_stamp(pv, idx, (cudaStream_t)stream), "ring_stamp");
      },
      py::arg("view"), py::arg("idx"), py::arg("stream") = 0);

  // Trace of the last launches: per slot (order position, first pick-up, last finish) in
  // globaltimer nanoseconds, plus the user stamps.  Synchronous copy (call after the step).
  m.def(
      "ring_trace",
      [](const PeerView& pv, const std::vector<uint32_t>& slots) {
        RingState* rs = ring_state_of(pv.epoch);
        std::vector<unsigned long long> t0(kRingSlots), t1(kRingSlots), stamps(kRingStamps);
        std::vector<uint32_t> pos(kRingSlots);
        chk(cudaMemcpy(t0.data(), rs->t_start, sizeof(unsigned long long) * kRingSlots, cudaMemcpyDeviceToHost),
            "ring_trace");
        chk(cudaMemcpy(t1.data(), rs->t_end, sizeof(unsigned long long) * kRingSlots, cudaMemcpyDeviceToHost),
            "ring_trace");
        chk(cudaMemcpy(pos.data(), rs->order_pos, sizeof(uint32_t) * kRingSlots, cudaMemcpyDeviceToHost),
            "ring_trace");
        chk(cudaMemcpy(stamps.data(), rs->stamps, sizeof(unsigned long long) * kRingStamps, cudaMemcpyDeviceToHost),
            "ring_trace");
        py::list out;
        for (uint32_t s : slots) {
          if (s >= (uint32_t)kRingSlots) throw std::runtime_error("ring_trace: bad slot");
          out.append(py::make_tuple(pos[s], t0[s], t1[s]));
        }
        py::list st;
        for (auto v : stamps) st.append(v);
        return py::make_tuple(out, st);
      },
      py::arg("view"), py::arg("slots"));
}
