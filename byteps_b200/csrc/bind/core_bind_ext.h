// Second half of the _core bindings (network transport, KV apps, server,
// host pipeline engine) lives in its own translation unit to keep compile
// times down.
#pragma once
#include <pybind11/pybind11.h>

void bind_core_ext(pybind11::module_& m);
