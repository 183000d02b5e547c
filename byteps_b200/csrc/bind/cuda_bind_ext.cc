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

void bind_cuda_ext(py::module_& m) {
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
