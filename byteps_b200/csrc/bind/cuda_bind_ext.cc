#include "bind/cuda_bind_ext.h"

namespace py = pybind11;

void bind_cuda_ext(py::module_& m) {}
