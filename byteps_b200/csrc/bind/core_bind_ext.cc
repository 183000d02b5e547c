#include "bind/core_bind_ext.h"

namespace py = pybind11;

void bind_net(py::module_& m);     // net/bind_net.cc
void bind_server(py::module_& m);  // server/bind_server.cc
void bind_engine(py::module_& m);  // core/bind_engine.cc

void bind_core_ext(py::module_& m) {
#ifdef BPS_WITH_NET
  bind_net(m);
  bind_server(m);
  bind_engine(m);
#endif
}
