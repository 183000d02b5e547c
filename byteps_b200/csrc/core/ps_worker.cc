#include "core/ps_worker.h"

#include "core/numa.h"

#include <cstring>

#include "core/env.h"
#include "core/log.h"
#include "cpu/half.h"

namespace bps {

PSWorkerConfig PSWorkerConfig::from_env() {
  PSWorkerConfig c;
  c.hash_fn = env_str("BYTEPS_KEY_HASH_FN", "djb2");
  c.mixed_mode = env_bool("BYTEPS_ENABLE_MIXED_MODE", false);
  c.mixed_bound = (int)env_int("BYTEPS_MIXED_MODE_BOUND", 101);
  long long credit = env_int("BYTEPS_SCHEDULING_CREDIT", 0);
  long long part = env_int("BYTEPS_PARTITION_BYTES", 4096000);
  c.credit_bytes = credit > 0 ? (uint64_t)(credit * part) : 0;
  c.min_compress_bytes = (size_t)env_int("BYTEPS_MIN_COMPRESS_BYTES", 65536);
  c.threadpool_size = (int)env_int("BYTEPS_THREADPOOL_SIZE", 4);
  return c;
}

PSWorker::PSWorker(net::Postoffice* po, const PSWorkerConfig& cfg, int app_id, int customer_id)
    : po_(po), cfg_(cfg), reducer_(0) {
  kv_.reset(new net::KVWorker(app_id, customer_id, po));
  int pushers = cfg.num_pushers > 0 ? cfg.num_pushers : po->num_workers();
  placer_.reset(new KeyPlacer(cfg.hash_fn, po->num_servers(), pushers, cfg.mixed_mode, cfg.mixed_bound));
  // PUSH is the scheduled stage: priority order + byte credits, like the reference's
  // scheduled queue (there it is the REDUCE queue on the signal root).
  push_q_.reset(new ScheduledQueue(PUSH, true, cfg.credit_bytes));
  pool_.reset(new ThreadPool((size_t)std::max(1, cfg.threadpool_size)));
  eager_pull_ = env_int("BYTEPS_PS_EAGER_PULL", 1) != 0;
  sample_name_ = env_str("BYTEPS_DEBUG_SAMPLE_TENSOR", "");
  pull_by_ref_ = env_bool("BYTEPS_PS_PULL_BY_REF", true);
  numa_node_ = (int)env_int("BYTEPS_NUMA_NODE", -1);
  dispatcher_ = std::thread([this] { DispatchLoop(); });
}

PSWorker::~PSWorker() { Stop(); }

void PSWorker::Stop() {
  if (stopped_) return;
  stopped_ = true;
  stop_ = true;
  push_q_->stop();
  if (dispatcher_.joinable()) dispatcher_.join();
  pool_->shutdown();
  kv_.reset();
}

void PSWorker::InitKey(uint64_t key, const void* data, size_t len, int dtype, int pushers) {
  int server = placer_->server_of(key, len);
  net::SArray<char> vals((char*)data, len, false);
  int cmd = command_encode(kDefaultPushPull, dtype);
  // low 16 bits: pushers of this key; above: where my GPU hangs, so the server can place the store next to it
  const int head = numa_pack_head(pushers, numa_aware() && numa_num_nodes() > 1 ? numa_node_ : -1);
  kv_->Wait(kv_->ZPush(server, key, vals, cmd, nullptr, head));
}

std::shared_ptr<Compressor> PSWorker::CompressorOf(uint64_t key) {
  std::lock_guard<std::mutex> g(comp_mu_);
  auto it = compressors_.find(key);
  return it == compressors_.end() ? nullptr : it->second;
}

bool PSWorker::HasCompressor(uint64_t key) { return CompressorOf(key) != nullptr; }

void PSWorker::RegisterCompressor(uint64_t key, const Kwargs& kw, size_t len, int dtype) {
  if (len < cfg_.min_compress_bytes) return;   // small tensors are not worth compressing
  std::shared_ptr<Compressor> c(CompressorRegistry::create(kw, len, dtype, false).release());
  if (!c) return;
  if (lr_ > 0) c->set_lr(lr_);     // the rate announced before this tensor's first push_pull
  {
    std::lock_guard<std::mutex> g(comp_mu_);
    compressors_[key] = c;
    comp_bufs_[key] = std::make_shared<std::vector<char>>(c->max_compressed_bytes() + 64);
  }
  std::string content = kwargs_serialize(kw);
  net::SArray<char> vals;
  vals.copy_from(content.data(), content.size());
  int server = placer_->server_of(key, len);
  int cmd = command_encode(kCompressedPushPull, dtype);
  kv_->Wait(kv_->ZPush(server, key, vals, cmd));
}

void PSWorker::SetLearningRate(double lr) {
  std::lock_guard<std::mutex> g(comp_mu_);
  lr_ = lr;
  for (auto& kv : compressors_) kv.second->set_lr(lr);
}

int PSWorker::PushPull(const std::string& name, void* ptr, int dtype, const std::vector<Part>& parts, int priority,
                       int version, double scale, void* ready_event, void* out) {
  if (out == ptr) out = nullptr;
  int h = handles_.allocate();
  if (parts.empty()) {
    handles_.mark_done(h, Status::OK());
    return h;
  }
  auto counter = std::make_shared<std::atomic<uint32_t>>(0);
  auto ctx = std::make_shared<TensorContext>();
  ctx->name = name;
  ctx->enqueue_ts_us = now_us();
  size_t total = 0;
  for (auto& p : parts) total = std::max(total, p.offset + p.len);
  EventQueryFn eq = event_query_;
  const int es = dtype_size(dtype);
  for (auto& p : parts) {
    auto t = std::make_shared<Task>();
    t->ctx = ctx;
    t->key = p.key;
    t->priority = priority;
    t->version = version;
    t->dtype = dtype;
    t->input = t->output = t->host = ptr;
    t->host_out = out;
    t->scale = scale;           // applied per partition when the result goes to `out`
    t->offset = p.offset;
    t->len = p.len;
    t->handle = h;
    t->total_parts = (uint32_t)parts.size();
    t->done_counter = counter;
    if (ready_event && eq) t->ready = [eq, ready_event]() { return eq(ready_event) != 0; };
    CpuReducer* red = &reducer_;
    HandleManager* hm = &handles_;
    Timeline* tl = timeline_;
    t->on_all_done = [=](const Status& s) {
      if (s.ok() && scale != 1.0 && !out) red->scale(ptr, (total / es) * es, dtype, scale);
      if (tl && tl->enabled()) tl->record(name, "", ~0ull, ctx->enqueue_ts_us, now_us() - ctx->enqueue_ts_us);
      hm->mark_done(h, s);
    };
    push_q_->add(t);
  }
  return h;
}

// BYTEPS_DEBUG_SAMPLE_TENSOR=<substring of the tensor name>: first and last element of the partition after
// every stage, like the reference (core_loops.cc:37-67; it selects by key, names are what users know here).
void PSWorker::Sample(const TaskPtr& t, const char* stage) {
  if (sample_name_.empty() || !t->ctx || t->ctx->name.find(sample_name_) == std::string::npos) return;
  // after the pull a host task with a separate output holds its result there, not in the staging window
  const bool delivered = t->host_out && !t->dev_out && strcmp(stage, "PULL") == 0;
  const char* base = (const char*)(delivered ? t->host_out : t->host) + t->offset;
  const int es = dtype_size(t->dtype);
  if (!base || es <= 0 || t->len < (size_t)es) return;
  auto val = [&](const char* p) -> double {
    switch (t->dtype) {
      case F32: return *(const float*)p;
      case F64: return *(const double*)p;
      case F16: return f16_to_f32(*(const uint16_t*)p);
      case BF16: return bf16_to_f32(*(const uint16_t*)p);
      case I32: return *(const int32_t*)p;
      case I64: return (double)*(const int64_t*)p;
      case I8: return *(const int8_t*)p;
      default: return *(const uint8_t*)p;
    }
  };
  BPS_LOG(WARNING) << "sample " << t->ctx->name << " key=" << t->key << " stage=" << stage << " first=" << val(base)
                   << " last=" << val(base + (t->len / es - 1) * es) << " len=" << t->len;
}

namespace {
struct H2dTrace {
  Timeline* tl;
  std::string name;
  uint64_t key;
  int64_t t0;
};
void h2d_landed(void* arg) {     // runs on a CUDA driver thread: no CUDA calls here
  H2dTrace* h = (H2dTrace*)arg;
  h->tl->record(h->name, stage_name(COPYH2D), h->key, h->t0, now_us() - h->t0);
  delete h;
}
}  // namespace

int PSWorker::PushPullDevice(const std::string& name, const void* dev_in, void* dev_out, void* host, int dtype,
                             const std::vector<Part>& parts, int priority, int version, double scale,
                             void* ready_event, void* gpu_ctx) {
  BPS_CHECK(gpu_ != nullptr) << "PushPullDevice without a GPU stage table (set_gpu_stage)";
  int h = handles_.allocate();
  if (parts.empty()) {
    handles_.mark_done(h, Status::OK());
    return h;
  }
  auto counter = std::make_shared<std::atomic<uint32_t>>(0);
  auto ctx = std::make_shared<TensorContext>();
  ctx->name = name;
  ctx->enqueue_ts_us = now_us();
  const BpsGpuStageFns* gpu = gpu_;
  gpu->wait_ready(gpu_ctx, ready_event);
  HandleManager* hm = &handles_;
  Timeline* tl = timeline_;
  for (auto& p : parts) {
    auto t = std::make_shared<Task>();
    t->ctx = ctx;
    t->key = p.key;
    t->priority = priority;
    t->version = version;
    t->dtype = dtype;
    t->input = (void*)dev_in;
    t->output = dev_out;
    t->dev_out = dev_out;
    t->gpu_ctx = gpu_ctx;
    t->scale = scale;
    t->host = host;
    t->offset = p.offset;
    t->len = p.len;
    t->handle = h;
    t->total_parts = (uint32_t)parts.size();
    t->done_counter = counter;
    // COPYD2H of THIS partition; the push waits for this event only
    t->d2h_start_us = now_us();
    void* ev = gpu->d2h(gpu_ctx, (char*)host + p.offset, (const char*)dev_in + p.offset, p.len);
    t->ready = [gpu, ev]() { return gpu->query(ev) != 0; };
    t->on_all_done = [=](const Status& s) {
      // every partition's H2D copy has been enqueued: one event on that stream covers them all
      void* done = gpu->h2d_mark(gpu_ctx);
      {
        std::lock_guard<std::mutex> g(done_mu_);
        done_events_[h] = done;
      }
      if (tl && tl->enabled()) tl->record(name, "", ~0ull, ctx->enqueue_ts_us, now_us() - ctx->enqueue_ts_us);
      hm->mark_done(h, s);
    };
    push_q_->add(t);
  }
  return h;
}

void* PSWorker::TakeDoneEvent(int handle) {
  std::lock_guard<std::mutex> g(done_mu_);
  auto it = done_events_.find(handle);
  if (it == done_events_.end()) return nullptr;
  void* e = it->second;
  done_events_.erase(it);
  return e;
}

void PSWorker::DispatchLoop() {
  while (!stop_) {
    TaskPtr t = push_q_->wait_get(2000);
    if (!t) continue;
    t->stage_start_us = now_us();
    if (t->dev_out) {
      if (timeline_ && timeline_->enabled())
        timeline_->record(t->ctx->name, stage_name(COPYD2H), t->key, t->d2h_start_us, now_us() - t->d2h_start_us);
      Sample(t, "COPYD2H");
    }
    auto comp = CompressorOf(t->key);
    if (comp) {
      // COMPRESS stage on the pool, then PUSH
      pool_->enqueue([this, t, comp] {
        std::shared_ptr<std::vector<char>> buf;
        {
          std::lock_guard<std::mutex> g(comp_mu_);
          buf = comp_bufs_[t->key];
        }
        int64_t t0 = now_us();
        t->compressed_len = comp->compress((char*)t->host + t->offset, buf->data());
        t->compressed = buf->data();
        if (timeline_ && timeline_->enabled())
          timeline_->record(t->ctx->name, stage_name(COMPRESS), t->key, t0, now_us() - t0);
        Sample(t, "COMPRESS");
        DoPush(t);
      });
    } else {
      DoPush(t);
    }
  }
}

void PSWorker::DoPush(const TaskPtr& t) {
  char* data = t->compressed ? (char*)t->compressed : (char*)t->host + t->offset;
  size_t len = t->compressed ? t->compressed_len : t->len;
  net::SArray<char> vals(data, len, false);
  int server = placer_->server_of(t->key, t->len);
  int cmd = command_encode(kDefaultPushPull, t->dtype);
  bytes_pushed_ += len;
  int64_t t0 = now_us();
  kv_->ZPush(server, t->key, vals, cmd, [this, t, t0] {
    if (timeline_ && timeline_->enabled()) timeline_->record(t->ctx->name, stage_name(PUSH), t->key, t0, now_us() - t0);
    push_q_->report_finish(t->len);   // the credit window covers data in flight to the server
    Sample(t, "PUSH");
    if (!eager_pull_) DoPull(t);
  });
  // The pull does not wait for the push acknowledgement (the reference's PUSH and PULL stages are two serial
  // round trips per partition): both requests travel on the same ordered connection, the server parks a pull
  // until the round's last push has been merged, and the payload has left the buffer when ZPush returns (or is
  // read by reference from a window the server only overwrites when it publishes the round).  One round trip
  // per partition instead of two.
  if (eager_pull_) DoPull(t);
}

void PSWorker::DoPull(const TaskPtr& t) {
  int server = placer_->server_of(t->key, t->len);
  int cmd = command_encode(kDefaultPushPull, t->dtype);
  char* dst = t->compressed ? (char*)t->compressed : (char*)t->host + t->offset;
  size_t cap = t->compressed ? CompressorOf(t->key)->max_compressed_bytes() : t->len;
  int64_t t0 = now_us();
  auto ts = std::make_shared<int>(-1);       // filled in by ZPull before the request is sent (see kv_app.h)
  // device-staged, uncompressed partitions can be DMA'd to the GPU straight out of a colocated server's store
  const bool want_ref = (t->dev_out || t->host_out) && !t->compressed && pull_by_ref_;
  kv_->ZPull(server, t->key, dst, cap, cmd, [this, t, t0, ts, want_ref] {
    if (timeline_ && timeline_->enabled()) timeline_->record(t->ctx->name, stage_name(PULL), t->key, t0, now_us() - t0);
    if (want_ref) {
      kv_->pulled_len(*ts);     // take-and-remove: asking for the timestamp also recorded the length
      net::KVWorker::PullRef ref = kv_->take_pull_ref(*ts);
      if (ref.ptr) {
        t->h2d_src = ref.ptr;
        t->h2d_region = ref.region;
        t->h2d_region_len = ref.region_len;
      }
    }
    if (t->host_out && !t->dev_out && !t->compressed) {
      // result wanted elsewhere: one copy, from the server's store if the pull was answered by reference, else from
      // the window the response landed in
      DeliverHost(t, t->h2d_src ? t->h2d_src : (char*)t->host + t->offset);
      return;
    }
    if (t->compressed) {
      size_t got = kv_->pulled_len(*ts);
      auto comp = CompressorOf(t->key);
      pool_->enqueue([this, t, comp, got] {
        int64_t t1 = now_us();
        comp->decompress(t->compressed, got, (char*)t->host + t->offset);
        if (timeline_ && timeline_->enabled())
          timeline_->record(t->ctx->name, stage_name(DECOMPRESS), t->key, t1, now_us() - t1);
        Sample(t, "DECOMPRESS");
        if (t->host_out && !t->dev_out) {
          DeliverHost(t, (char*)t->host + t->offset);
          return;
        }
        Finish(t);
      });
    } else {
      Sample(t, "PULL");
      Finish(t);
    }
  }, (t->compressed || want_ref) ? ts.get() : nullptr, want_ref);
}

void PSWorker::DeliverHost(const TaskPtr& t, const void* src) {
  const TaskPtr keep = t;
  pool_->enqueue([this, keep, src] {
    const TaskPtr& t = keep;
    char* dst = (char*)t->host_out + t->offset;
    const int es = dtype_size(t->dtype);
    memcpy(dst, src, t->len);         // partitions are independent: the pool runs several of these at once
    if (t->scale != 1.0) reducer_.scale(dst, (t->len / es) * es, t->dtype, t->scale);
    Sample(t, "PULL");
    Finish(t);
  });
}

void PSWorker::Finish(const TaskPtr& t) {
  if (t->dev_out) {
    // COPYH2D of this partition, issued from the pull completion (scaling on the host first): it overlaps
    // with the pulls / pushes / D2H copies of the partitions still in flight
    const TaskPtr keep = t;
    pool_->enqueue([this, keep] {
      const TaskPtr& t = keep;
      char* hp = (char*)t->host + t->offset;
      const int es = dtype_size(t->dtype);
      if (t->h2d_src) {
        // answered by reference: the source is the server's store; page-lock its mapping once, then DMA from it
        bool ok;
        {
          std::lock_guard<std::mutex> g(done_mu_);
          auto it = registered_.find(t->h2d_region);
          if (it == registered_.end()) {
            ok = gpu_->host_register(t->gpu_ctx, t->h2d_region, t->h2d_region_len) == 0;
            registered_[t->h2d_region] = ok;
          } else {
            ok = it->second;
          }
        }
        if (ok) {
          hp = (char*)t->h2d_src;
        } else {
          memcpy(hp, t->h2d_src, t->len);       // registration refused: stage through my own pinned window
        }
      }
      const bool dev_scale = t->scale != 1.0 && (t->dtype == F32 || t->dtype == F16 || t->dtype == BF16);
      if (t->scale != 1.0 && !dev_scale) {
        if (hp != (char*)t->host + t->offset) {   // never scale the server's store in place
          memcpy((char*)t->host + t->offset, hp, t->len);
          hp = (char*)t->host + t->offset;
        }
        reducer_.scale(hp, (t->len / es) * es, t->dtype, t->scale);
      }
      H2dTrace* tr = nullptr;
      if (timeline_ && timeline_->enabled()) tr = new H2dTrace{timeline_, t->ctx->name, t->key, now_us()};
      int rc = gpu_->h2d(t->gpu_ctx, (char*)t->dev_out + t->offset, hp, t->len, tr ? h2d_landed : nullptr, tr);
      // the 1/size of an average runs on the device behind the copy (no CPU pass over pinned memory)
      if (rc == 0 && dev_scale) rc = gpu_->scale(t->gpu_ctx, (char*)t->dev_out + t->offset, (t->len / es) * es, t->dtype, t->scale);
      Sample(t, "COPYH2D");
      uint32_t done = t->done_counter->fetch_add(1) + 1;
      if (done == t->total_parts && t->on_all_done)
        t->on_all_done(rc == 0 ? Status::OK() : Status::Error(ST_UNKNOWN, "cudaMemcpyAsync (COPYH2D) failed"));
    });
    return;
  }
  uint32_t done = t->done_counter->fetch_add(1) + 1;
  if (done == t->total_parts && t->on_all_done) t->on_all_done(Status::OK());
}

}  // namespace bps
