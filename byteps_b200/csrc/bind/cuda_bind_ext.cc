#include "bind/cuda_bind_ext.h"

#include <cuda_runtime_api.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>

#include "kernels/misc.cuh"

namespace py = pybind11;
using namespace bps;

static void chk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

extern "C" int bps_event_query(void* ev) {
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaSuccess) return 1;
  if (e != cudaErrorNotReady) cudaGetLastError();
  return 0;
}

void bind_cuda_ext(py::module_& m) {
  // address of a C function the CUDA-free runtime can call to poll a device event
  m.def("event_query_fn", []() { return (uintptr_t)&bps_event_query; });
  m.def(
      "write_blob",
      [](uintptr_t dst, const py::bytes& data, uintptr_t stream) {
        std::string s(data);
        chk(launch_write_blob((void*)dst, s.data(), s.size(), (cudaStream_t)stream), "write_blob");
      },
      py::arg("dst"), py::arg("data"), py::arg("stream") = 0);
  m.def(
      "l2_flush",
      [](uintptr_t buf, size_t nbytes, uint32_t value, uintptr_t stream) {
        chk(launch_l2_flush((void*)buf, nbytes, value, (cudaStream_t)stream), "l2_flush");
      },
      py::arg("buf"), py::arg("nbytes"), py::arg("value") = 0, py::arg("stream") = 0);
}
