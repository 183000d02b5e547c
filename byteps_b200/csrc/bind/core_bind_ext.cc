// Bindings for the KV transport, the summation server and the PS worker pipeline.
#include "bind/core_bind_ext.h"

#include <pybind11/functional.h>
#include <pybind11/stl.h>

#include "core/env.h"
#include "core/host_reduce.h"
#include "core/numa.h"
#include "core/ps_worker.h"
#include "net/kv_app.h"
#include "net/local_signal.h"
#include "net/van.h"
#include "server/server.h"

namespace py = pybind11;
using namespace bps;
using namespace bps::net;

namespace {

NetConfig make_cfg(const std::string& role, int num_workers, int num_servers, const std::string& sched_host,
                   int sched_port, const std::string& node_host, int rank_hint, py::dict extra) {
  NetConfig c = NetConfig::from_env();
  c.role = role == "server" ? Role::kServer : role == "scheduler" ? Role::kScheduler : Role::kWorker;
  c.num_workers = num_workers;
  c.num_servers = num_servers;
  c.scheduler_host = sched_host;
  c.scheduler_port = sched_port;
  // an explicit host wins; '' keeps what the environment says (DMLC_NODE_HOST, DMLC_INTERFACE, or the first
  // non-loopback address when the scheduler is on another host)
  c.node_host = !node_host.empty() ? node_host : env_str("DMLC_NODE_HOST", "");
  c.resolve_node_host();
  c.rank_hint = rank_hint;
  c.node_port = 0;
  for (auto item : extra) {
    std::string k = py::str(item.first);
    if (k == "verbose") c.verbose = item.second.cast<int>();
    else if (k == "heartbeat_interval_s") c.heartbeat_interval_s = item.second.cast<int>();
    else if (k == "heartbeat_timeout_s") c.heartbeat_timeout_s = item.second.cast<int>();
    else if (k == "resend") c.resend = item.second.cast<bool>();
    else if (k == "resend_timeout_ms") c.resend_timeout_ms = item.second.cast<int>();
    else if (k == "drop_msg_pct") c.drop_msg_pct = item.second.cast<int>();
    else if (k == "enable_ipc") c.enable_ipc = item.second.cast<bool>();
    else if (k == "profile_path") c.profile_path = item.second.cast<std::string>();
    else if (k == "is_recovery") c.is_recovery = item.second.cast<bool>();
    else if (k == "node_port") c.node_port = item.second.cast<int>();
    else if (k == "num_lanes") c.num_lanes = std::max(1, std::min(16, item.second.cast<int>()));
    else if (k == "local") c.local = item.second.cast<bool>();
    else if (k == "van_type") c.van_type = item.second.cast<std::string>();
    else throw std::runtime_error("unknown NetConfig field " + k);
  }
  return c;
}

}  // namespace

void bind_core_ext(py::module_& m) {
  m.attr("GROUP_SCHEDULER") = kScheduler;
  m.attr("GROUP_SERVER") = kServerGroup;
  m.attr("GROUP_WORKER") = kWorkerGroup;
  m.attr("GROUP_ALL") = kScheduler + kServerGroup + kWorkerGroup;

  m.def("resolve_node_host", [](const std::string& sched_host, const std::string& node_host) {
    // the address a node with these settings would advertise to the scheduler (tests, doctor)
    return make_cfg("worker", 1, 1, sched_host, 9000, node_host, -1, py::dict()).node_host;
  });

  m.def("meta_roundtrip", [](int head, const std::string& body, uint64_t key, int cmd) {
    Meta a;
    a.head = head;
    a.body = body;
    a.key = key;
    a.cmd = cmd;
    a.push = true;
    a.request = true;
    a.control.cmd = Control::ADD_NODE;
    Node n;
    n.hostname = "h";
    n.port = 7;
    n.id = 9;
    a.control.node.push_back(n);
    std::string s = meta_pack(a);
    Meta b;
    bool ok = meta_unpack(s.data(), s.size(), &b);
    return py::make_tuple(ok, b.head, b.body, b.key, b.cmd, b.push, b.request, (int)b.control.cmd,
                          b.control.node.size() ? b.control.node[0].port : -1, s.size());
  });

  m.def("meta_pack_sample", [](int nodes, const std::string& body) {
    Meta a;
    a.body = body;
    a.key = 0x1234567890ull;
    a.push = true;
    a.control.cmd = nodes ? Control::ADD_NODE : Control::EMPTY;
    for (int i = 0; i < nodes; ++i) {
      Node n;
      n.hostname = "host" + std::to_string(i);
      n.port = 1000 + i;
      n.id = 8 + i;
      a.control.node.push_back(n);
    }
    return py::bytes(meta_pack(a));
  });
  // decoder robustness (fuzzing): arbitrary bytes must be rejected or decoded, never read out of bounds
  m.def("meta_unpack_bytes", [](const py::bytes& b) {
    std::string s = b;
    Meta out;
    bool ok = meta_unpack(s.data(), s.size(), &out);
    return py::make_tuple(ok, (int)out.control.node.size(), out.body.size());
  });

  py::class_<Postoffice, std::shared_ptr<Postoffice>>(m, "Postoffice")
      .def(py::init([](const std::string& role, int nw, int ns, const std::string& sh, int sp, const std::string& nh,
                       int rank_hint, py::dict extra) {
             return std::make_shared<Postoffice>(make_cfg(role, nw, ns, sh, sp, nh, rank_hint, extra));
           }),
           py::arg("role"), py::arg("num_workers"), py::arg("num_servers"), py::arg("scheduler_host") = "127.0.0.1",
           py::arg("scheduler_port") = 9000, py::arg("node_host") = "127.0.0.1", py::arg("rank_hint") = -1,
           py::arg("extra") = py::dict())
      .def("start", [](Postoffice& p, int cid, bool barrier) {
             py::gil_scoped_release r;
             p.Start(cid, barrier);
           }, py::arg("customer_id") = 0, py::arg("barrier") = true)
      .def("finalize", [](Postoffice& p, int cid, bool barrier) {
             py::gil_scoped_release r;
             p.Finalize(cid, barrier);
           }, py::arg("customer_id") = 0, py::arg("barrier") = true)
      .def("barrier", [](Postoffice& p, int cid, int group) {
             py::gil_scoped_release r;
             p.Barrier(cid, group);
           }, py::arg("customer_id") = 0, py::arg("group") = kWorkerGroup)
      .def("my_rank", &Postoffice::my_rank)
      .def("my_id", [](Postoffice& p) { return p.van()->my_node().id; })
      .def("my_port", [](Postoffice& p) { return p.van()->my_node().port; })
      .def("num_workers", &Postoffice::num_workers)
      .def("num_servers", &Postoffice::num_servers)
      .def("server_key_ranges", &Postoffice::GetServerKeyRanges)
      .def("dead_nodes", &Postoffice::GetDeadNodes)
      .def("send_bytes", [](Postoffice& p) { return p.van()->send_bytes(); })
      .def("recv_bytes", [](Postoffice& p) { return p.van()->recv_bytes(); })
      .def("direct_recvs", [](Postoffice& p) { return p.van()->direct_recvs(); },
           "pull responses whose payload was read from the socket straight into the destination buffer")
      .def_static("worker_rank_to_id", &Postoffice::WorkerRankToID)
      .def_static("server_rank_to_id", &Postoffice::ServerRankToID)
      .def_static("id_to_rank", &Postoffice::IDtoRank);

  // ---- raw KV apps (transport tests / benchmarks)
  py::class_<KVWorker>(m, "KVWorker")
      .def(py::init([](std::shared_ptr<Postoffice> po, int app, int cid) { return new KVWorker(app, cid, po.get()); }),
           py::arg("postoffice"), py::arg("app_id") = 0, py::arg("customer_id") = 0, py::keep_alive<1, 2>())
      .def("push", [](KVWorker& w, int server, uint64_t key, uintptr_t ptr, size_t len, int cmd) {
             py::gil_scoped_release r;
             SArray<char> v((char*)ptr, len, false);
             int ts = w.ZPush(server, key, v, cmd);
             w.Wait(ts);
           }, py::arg("server"), py::arg("key"), py::arg("ptr"), py::arg("len"), py::arg("cmd") = 0)
      .def("pull", [](KVWorker& w, int server, uint64_t key, uintptr_t ptr, size_t len, int cmd) {
             py::gil_scoped_release r;
             int ts_early = -1;     // asks ZPull to keep the received length for pulled_len
             int ts = w.ZPull(server, key, (char*)ptr, len, cmd, nullptr, &ts_early);
             w.Wait(ts);
             return w.pulled_len(ts);
           }, py::arg("server"), py::arg("key"), py::arg("ptr"), py::arg("len"), py::arg("cmd") = 0)
      .def("request", [](KVWorker& w, int head, const std::string& body, int recv_id) {
        py::gil_scoped_release r;
        int ts = w.Request(head, body, recv_id);
        w.Wait(ts);
      });

  // echo-style KV server used by transport tests: stores the last push per key, answers pulls with it
  struct EchoServer {
    std::unique_ptr<KVServer> kv;
    std::mutex mu;
    std::unordered_map<uint64_t, std::string> store;
    std::vector<std::string> simple_bodies;
  };
  py::class_<EchoServer>(m, "EchoKVServer")
      .def(py::init([](std::shared_ptr<Postoffice> po, int app) {
             auto* e = new EchoServer();
             e->kv.reset(new KVServer(app, po.get()));
             e->kv->set_kv_request_handle([e](const KVMeta& req, const KVPairs& d, KVServer* s) {
               KVPairs res;
               if (req.push) {
                 std::lock_guard<std::mutex> g(e->mu);
                 e->store[req.key].assign(d.vals.data(), d.vals.size());
               } else {
                 std::lock_guard<std::mutex> g(e->mu);
                 auto& v = e->store[req.key];
                 res.key = req.key;
                 res.vals.copy_from(v.data(), v.size());
               }
               s->Response(req, res);
             });
             e->kv->set_request_handle([e](const Message& msg, SimpleApp* app) {
               {
                 std::lock_guard<std::mutex> g(e->mu);
                 e->simple_bodies.push_back(msg.meta.body);
               }
               app->Response(msg, "ack:" + msg.meta.body);
             });
             return e;
           }),
           py::arg("postoffice"), py::arg("app_id") = 0, py::keep_alive<1, 2>())
      .def("num_keys", [](EchoServer& e) {
        std::lock_guard<std::mutex> g(e.mu);
        return e.store.size();
      })
      .def("simple_bodies", [](EchoServer& e) {
        std::lock_guard<std::mutex> g(e.mu);
        return e.simple_bodies;
      })
      .def("stop", [](EchoServer& e) { e.kv.reset(); });

  // ---- the summation server
  py::class_<server::SumServer>(m, "SumServer")
      .def(py::init([](std::shared_ptr<Postoffice> po, int threads, py::object schedule, py::object blocking,
                       py::object sync, int pushers, bool log_keys, int64_t debug_key) {
             // None = keep what the environment says (BYTEPS_SERVER_ENABLE_SCHEDULE, BYTEPS_SERVER_ENGINE_BLOCKING,
             // BYTEPS_ENABLE_ASYNC); an explicit bool overrides it
             server::ServerConfig c = server::ServerConfig::from_env();
             if (threads > 0) c.engine_threads = threads;
             if (!schedule.is_none()) c.enable_schedule = schedule.cast<bool>();
             if (!blocking.is_none()) c.engine_blocking = blocking.cast<bool>();
             if (!sync.is_none()) c.sync_mode = sync.cast<bool>();
             if (pushers > 0) c.pushers_per_key = pushers;
             c.log_keys = c.log_keys || log_keys;
             if (debug_key >= 0) c.debug_key = debug_key;
             return new server::SumServer(po.get(), c);
           }),
           py::arg("postoffice"), py::arg("engine_threads") = 0, py::arg("enable_schedule") = py::none(),
           py::arg("engine_blocking") = py::none(), py::arg("sync_mode") = py::none(), py::arg("pushers_per_key") = 0,
           py::arg("log_keys") = false, py::arg("debug_key") = -1, py::keep_alive<1, 2>())
      .def("stop", [](server::SumServer& s) {
        py::gil_scoped_release r;
        s.Stop();
      })
      .def("pushes", &server::SumServer::pushes)
      .def("pulls", &server::SumServer::pulls)
      .def("num_keys", &server::SumServer::num_keys);

  // ---- PS worker pipeline
  py::class_<PSWorker>(m, "PSWorker")
      .def(py::init([](std::shared_ptr<Postoffice> po, const std::string& hash_fn, uint64_t credit_bytes,
                       size_t min_compress, int pool, int num_pushers) {
             PSWorkerConfig c = PSWorkerConfig::from_env();
             if (!hash_fn.empty()) c.hash_fn = hash_fn;
             if (credit_bytes) c.credit_bytes = credit_bytes;
             if (min_compress != (size_t)-1) c.min_compress_bytes = min_compress;
             if (pool > 0) c.threadpool_size = pool;
             c.num_pushers = num_pushers;
             return new PSWorker(po.get(), c);
           }),
           py::arg("postoffice"), py::arg("hash_fn") = "", py::arg("credit_bytes") = 0,
           py::arg("min_compress_bytes") = (size_t)-1, py::arg("threadpool_size") = 0, py::arg("num_pushers") = 0,
           py::keep_alive<1, 2>())
      .def("stop", [](PSWorker& w) {
        py::gil_scoped_release r;
        w.Stop();
      })
      .def("set_event_query", [](PSWorker& w, uintptr_t fn) { w.set_event_query((EventQueryFn)fn); })
      .def("set_timeline", [](PSWorker& w, std::shared_ptr<Timeline> t) { w.set_timeline(t.get()); },
           py::keep_alive<1, 2>())
      .def("init_key", [](PSWorker& w, uint64_t key, uintptr_t ptr, size_t len, int dtype, int pushers) {
        py::gil_scoped_release r;
        w.InitKey(key, (const void*)ptr, len, dtype, pushers);
      }, py::arg("key"), py::arg("ptr"), py::arg("len"), py::arg("dtype"), py::arg("pushers") = 0)
      .def("register_compressor", [](PSWorker& w, uint64_t key, const Kwargs& kw, size_t len, int dtype) {
        py::gil_scoped_release r;
        w.RegisterCompressor(key, kw, len, dtype);
      })
      .def("has_compressor", &PSWorker::HasCompressor)
      .def("set_learning_rate", &PSWorker::SetLearningRate)
      .def("push_pull",
           [](PSWorker& w, const std::string& name, uintptr_t ptr, int dtype,
              const std::vector<std::tuple<uint64_t, size_t, size_t>>& parts, int priority, int version, double scale,
              uintptr_t ready_event, uintptr_t out) {
             std::vector<PSWorker::Part> ps;
             for (auto& t : parts) ps.push_back({std::get<0>(t), std::get<1>(t), std::get<2>(t)});
             return w.PushPull(name, (void*)ptr, dtype, ps, priority, version, scale, (void*)ready_event, (void*)out);
           },
           py::arg("name"), py::arg("ptr"), py::arg("dtype"), py::arg("parts"), py::arg("priority") = 0,
           py::arg("version") = 0, py::arg("scale") = 1.0, py::arg("ready_event") = 0, py::arg("out") = 0)
      .def("set_numa_node", &PSWorker::set_numa_node)
      .def("numa_node", &PSWorker::numa_node)
      .def("set_gpu_stage", [](PSWorker& w, uintptr_t fns) { w.set_gpu_stage((const BpsGpuStageFns*)fns); })
      .def("push_pull_device",
           [](PSWorker& w, const std::string& name, uintptr_t dev_in, uintptr_t dev_out, uintptr_t host, int dtype,
              const std::vector<std::tuple<uint64_t, size_t, size_t>>& parts, int priority, int version,
              double scale, uintptr_t ready_event, uintptr_t gpu_ctx) {
             std::vector<PSWorker::Part> ps;
             for (auto& p : parts) ps.push_back({std::get<0>(p), std::get<1>(p), std::get<2>(p)});
             return w.PushPullDevice(name, (const void*)dev_in, (void*)dev_out, (void*)host, dtype, ps, priority,
                                     version, scale, (void*)ready_event, (void*)gpu_ctx);
           },
           py::arg("name"), py::arg("dev_in"), py::arg("dev_out"), py::arg("host"), py::arg("dtype"), py::arg("parts"),
           py::arg("priority") = 0, py::arg("version") = 0, py::arg("scale") = 1.0, py::arg("ready_event") = 0,
           py::arg("gpu_ctx") = 0)
      .def("take_done_event", [](PSWorker& w, int h) { return (uintptr_t)w.TakeDoneEvent(h); })
      .def("poll", &PSWorker::Poll)
      .def("wait", [](PSWorker& w, int h, int64_t timeout_ms) {
             Status s;
             {
               py::gil_scoped_release r;
               s = w.Wait(h, timeout_ms);
             }
             if (s.code != ST_OK && s.code != ST_IN_PROGRESS) throw std::runtime_error("push_pull failed: " + s.reason);
             return s.code == ST_OK;
           }, py::arg("handle"), py::arg("timeout_ms") = -1)
      .def("server_of", &PSWorker::ServerOf)
      .def("server_load", &PSWorker::ServerLoad)
      .def("bytes_pushed", &PSWorker::bytes_pushed);

  // ---- intra-box UDS signalling
  for (int i = 0; i < SIG_COUNT; ++i) {
    static const char* names[] = {"SIG_REDUCE_READY", "SIG_PCIE_REDUCE_READY", "SIG_BCAST_READY", "SIG_PUSH_READY",
                                  "SIG_DO_REDUCE", "SIG_DO_BROADCAST", "SIG_DO_GROUP", "SIG_DO_COPYH2D"};
    m.attr(names[i]) = i;
  }
  py::class_<LocalComm>(m, "LocalComm")
      .def(py::init<int, const std::vector<int>&, const std::string&, const std::string&>(), py::arg("local_rank"),
           py::arg("members"), py::arg("dir") = "", py::arg("suffix") = "bps")
      .def_property_readonly("rank", &LocalComm::rank)
      .def_property_readonly("root", &LocalComm::root)
      .def("is_root", &LocalComm::is_root)
      .def("send_to_root", [](LocalComm& c, int sig, uint64_t key) {
        py::gil_scoped_release r;
        return c.send_to_root(sig, key);
      })
      .def("broadcast", [](LocalComm& c, int sig, uint64_t key) {
        py::gil_scoped_release r;
        return c.broadcast(sig, key);
      })
      .def("recv_from_root", [](LocalComm& c, int timeout_ms) -> py::object {
        LocalMsg msg;
        bool ok;
        {
          py::gil_scoped_release r;
          ok = c.recv_from_root(&msg, timeout_ms);
        }
        if (!ok) return py::none();
        return py::make_tuple(msg.src, msg.signal, msg.key);
      }, py::arg("timeout_ms") = 5000)
      .def("set_tables", [](LocalComm& c, std::shared_ptr<ReadyTable> a, std::shared_ptr<ReadyTable> b,
                            std::shared_ptr<ReadyTable> d, std::shared_ptr<ReadyTable> e) {
        c.set_tables(a.get(), b.get(), d.get(), e.get());
      }, py::arg("reduce") = nullptr, py::arg("pcie") = nullptr, py::arg("bcast") = nullptr, py::arg("push") = nullptr,
         py::keep_alive<1, 2>(), py::keep_alive<1, 3>(), py::keep_alive<1, 4>(), py::keep_alive<1, 5>())
      .def("received", &LocalComm::received);

  // ---- box-local reduction of host tensors (core/host_reduce.h)
  py::class_<HostLocalReduce>(m, "HostLocalReduce")
      .def(py::init<int, int, const std::string&, int, const std::string&>(), py::arg("local_rank"), py::arg("local_size"),
           py::arg("tag"), py::arg("reducer_threads") = 0, py::arg("socket_dir") = "")
      .def("is_root", &HostLocalReduce::is_root)
      .def_property_readonly("local_rank", &HostLocalReduce::local_rank)
      .def_property_readonly("local_size", &HostLocalReduce::local_size)
      .def("contribute", [](HostLocalReduce& h, uint64_t key, uintptr_t src, size_t nbytes, int64_t timeout_ms) {
             py::gil_scoped_release r;
             return h.contribute(key, (const void*)src, nbytes, timeout_ms);
           }, py::arg("key"), py::arg("src"), py::arg("nbytes"), py::arg("timeout_ms") = 60000)
      .def("reduce", [](HostLocalReduce& h, uint64_t key, size_t nbytes, int dtype, int64_t timeout_ms, double alpha) {
             py::gil_scoped_release r;
             return (uintptr_t)h.reduce(key, nbytes, dtype, timeout_ms, alpha);
           }, py::arg("key"), py::arg("nbytes"), py::arg("dtype"), py::arg("timeout_ms") = -1, py::arg("alpha") = 1.0)
      .def("publish", [](HostLocalReduce& h, uint64_t key, uintptr_t dst, size_t nbytes, int64_t timeout_ms) {
             py::gil_scoped_release r;
             return h.publish(key, (void*)dst, nbytes, timeout_ms);
           }, py::arg("key"), py::arg("dst"), py::arg("nbytes"), py::arg("timeout_ms") = -1)
      .def("collect", [](HostLocalReduce& h, uint64_t key, uintptr_t dst, size_t nbytes, int64_t timeout_ms, int dtype,
                         double alpha) {
             py::gil_scoped_release r;
             return h.collect(key, (void*)dst, nbytes, timeout_ms, dtype, alpha);
           }, py::arg("key"), py::arg("dst"), py::arg("nbytes"), py::arg("timeout_ms") = -1, py::arg("dtype") = (int)F32,
           py::arg("alpha") = 1.0)
      .def("window", [](HostLocalReduce& h, uint64_t key) { return (uintptr_t)h.window(key); })
      .def("signals_received", &HostLocalReduce::signals_received);

  // ---- shm registry (colocated IPC + pinned staging buffers)
  // Host-memory implementation of the device-stage table (core/gpu_stage.h): "device" pointers are plain host
  // memory, copies are memcpy, events are always complete.  It lets the CPU test-suite drive
  // PSWorker::PushPullDevice - the per-partition D2H / PUSH / PULL / H2D pipeline and the pull-by-reference
  // path - without a GPU, and records what the worker asked for.
  struct HostStage {
    std::atomic<uint64_t> d2h{0}, h2d{0}, h2d_bytes{0}, registered{0}, scaled{0};
    std::atomic<bool> refuse_register{false};
    std::mutex mu;
    std::vector<std::pair<uintptr_t, size_t>> h2d_sources;
  };
  static HostStage hs;
  static int hs_event = 0;
  static const BpsGpuStageFns hs_fns = {
      [](void*, void*) {},
      [](void*, void* host, const void* dev, size_t len) -> void* {
        memcpy(host, dev, len);
        hs.d2h++;
        return &hs_event;
      },
      [](void*) -> int { return 1; },
      [](void*, void* dev, const void* host, size_t len, bps_host_cb cb, void* arg) -> int {
        memcpy(dev, host, len);
        hs.h2d++;
        hs.h2d_bytes += len;
        {
          std::lock_guard<std::mutex> g(hs.mu);
          hs.h2d_sources.emplace_back((uintptr_t)host, len);
        }
        if (cb) cb(arg);
        return 0;
      },
      [](void*) -> void* { return &hs_event; },
      [](void*, void*, size_t) -> int {
        hs.registered++;
        return hs.refuse_register ? 1 : 0;
      },
      [](void*, void* dev, size_t nbytes, int dtype, double alpha) -> int {
        static CpuReducer r;
        r.scale(dev, nbytes, dtype, alpha);
        hs.scaled++;
        return 0;
      },
  };
  m.def("host_stage_fns", [] { return (uintptr_t)&hs_fns; });
  m.def("host_stage_reset", [](bool refuse_register) {
    hs.d2h = hs.h2d = hs.h2d_bytes = hs.registered = hs.scaled = 0;
    hs.refuse_register = refuse_register;
    std::lock_guard<std::mutex> g(hs.mu);
    hs.h2d_sources.clear();
  }, py::arg("refuse_register") = false);
  m.def("host_stage_stats", [] {
    py::dict d;
    d["d2h"] = (uint64_t)hs.d2h;
    d["h2d"] = (uint64_t)hs.h2d;
    d["h2d_bytes"] = (uint64_t)hs.h2d_bytes;
    d["registered"] = (uint64_t)hs.registered;
    d["scaled"] = (uint64_t)hs.scaled;
    std::lock_guard<std::mutex> g(hs.mu);
    d["h2d_sources"] = hs.h2d_sources;
    return d;
  });
  m.def("ipc_stats", [] {
    auto& st = bps::net::IpcStats::get();
    py::dict d;
    d["ref_responses"] = (uint64_t)st.ref_responses;
    d["ref_bytes"] = (uint64_t)st.ref_bytes;
    d["shm_responses"] = (uint64_t)st.shm_responses;
    d["payload_responses"] = (uint64_t)st.payload_responses;
    return d;
  });

  // NUMA placement helpers (core/numa.h)
  m.def("numa_num_nodes", &bps::numa_num_nodes);
  m.def("numa_node_of_pci", &bps::numa_node_of_pci);
  m.def("numa_cpus_of_node", &bps::numa_cpus_of_node);
  m.def("parse_cpu_list", &bps::parse_cpu_list);
  m.def("numa_bind_memory", [](uintptr_t p, size_t len, int node) { return bps::numa_bind_memory((void*)p, len, node); });
  m.def("numa_node_of_addr", [](uintptr_t p) { return bps::numa_node_of_addr((const void*)p); });
  m.def("numa_pin_thread_to_node", &bps::numa_pin_thread_to_node);
  m.def("numa_aware", &bps::numa_aware);
  m.def("numa_prefer_node_for_process", &bps::numa_prefer_node_for_process);
  m.def("numa_pack_head", &bps::numa_pack_head);
  m.def("numa_head_pushers", &bps::numa_head_pushers);
  m.def("numa_head_node", &bps::numa_head_node);
  m.def("shm_create", [](const std::string& name, size_t len) { return (uintptr_t)ShmRegistry::get().create(name, len); });
  m.def("shm_open", [](const std::string& name, size_t len) { return (uintptr_t)ShmRegistry::get().open(name, len); });
  m.def("shm_release", [](const std::string& name) { ShmRegistry::get().release(name); });
  m.def("shm_reap_stale", [](const std::string& dir) { return ShmRegistry::reap_stale(dir); }, py::arg("dir") = "/dev/shm");
  m.def("shm_lookup", [](uintptr_t p, size_t len) -> py::object {
    std::string name;
    uint64_t off = 0;
    if (!ShmRegistry::get().lookup((const void*)p, len, &name, &off)) return py::none();
    return py::make_tuple(name, off);
  });
}
