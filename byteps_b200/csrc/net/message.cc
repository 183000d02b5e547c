#include "net/message.h"

namespace bps {
namespace net {

namespace {

struct Writer {
  std::string s;
  void u8(uint8_t v) { s.push_back((char)v); }
  void i32(int32_t v) { s.append((const char*)&v, 4); }
  void u64(uint64_t v) { s.append((const char*)&v, 8); }
  void str(const std::string& v) {
    i32((int32_t)v.size());
    s.append(v);
  }
};

struct Reader {
  const char* p;
  const char* end;
  bool ok = true;
  bool need(size_t n) {
    if ((size_t)(end - p) < n) ok = false;
    return ok;
  }
  uint8_t u8() {
    if (!need(1)) return 0;
    return (uint8_t)*p++;
  }
  int32_t i32() {
    if (!need(4)) return 0;
    int32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint64_t u64() {
    if (!need(8)) return 0;
    uint64_t v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  std::string str() {
    int32_t n = i32();
    if (n < 0 || !need((size_t)n)) {
      ok = false;
      return "";
    }
    std::string v(p, p + n);
    p += n;
    return v;
  }
};

}  // namespace

std::string meta_pack(const Meta& m) {
  Writer w;
  w.i32(m.head);
  w.i32(m.app_id);
  w.i32(m.customer_id);
  w.i32(m.timestamp);
  w.i32(m.sender);
  w.i32(m.recver);
  uint8_t flags = (m.request ? 1 : 0) | (m.push ? 2 : 0) | (m.pull ? 4 : 0) | (m.simple_app ? 8 : 0);
  w.u8(flags);
  w.str(m.body);
  w.u64(m.key);
  w.i32(m.cmd);
  w.u64(m.val_len);
  w.u64(m.msg_sig);
  w.str(m.shm_name);
  w.u64(m.shm_offset);
  w.u64(m.shm_len);
  w.i32(m.src_dev);
  w.i32(m.src_id);
  w.i32(m.dst_dev);
  w.i32(m.dst_id);
  // control
  w.i32((int32_t)m.control.cmd);
  w.i32(m.control.barrier_group);
  w.u64(m.control.msg_sig);
  w.i32((int32_t)m.control.node.size());
  for (const auto& n : m.control.node) {
    w.i32((int32_t)n.role);
    w.i32(n.id);
    w.i32(n.customer_id);
    w.str(n.hostname);
    w.i32(n.port);
    w.u8(n.is_recovery ? 1 : 0);
    w.i32(n.aux_id);
  }
  return w.s;
}

bool meta_unpack(const char* buf, size_t len, Meta* m) {
  Reader r{buf, buf + len};
  m->head = r.i32();
  m->app_id = r.i32();
  m->customer_id = r.i32();
  m->timestamp = r.i32();
  m->sender = r.i32();
  m->recver = r.i32();
  uint8_t flags = r.u8();
  m->request = flags & 1;
  m->push = flags & 2;
  m->pull = flags & 4;
  m->simple_app = flags & 8;
  m->body = r.str();
  m->key = r.u64();
  m->cmd = r.i32();
  m->val_len = r.u64();
  m->msg_sig = r.u64();
  m->shm_name = r.str();
  m->shm_offset = r.u64();
  m->shm_len = r.u64();
  m->src_dev = r.i32();
  m->src_id = r.i32();
  m->dst_dev = r.i32();
  m->dst_id = r.i32();
  m->control.cmd = (Control::Command)r.i32();
  m->control.barrier_group = r.i32();
  m->control.msg_sig = r.u64();
  int32_t nn = r.i32();
  if (nn < 0 || nn > 65536) return false;
  m->control.node.clear();
  for (int i = 0; i < nn && r.ok; ++i) {
    Node n;
    n.role = (Role)r.i32();
    n.id = r.i32();
    n.customer_id = r.i32();
    n.hostname = r.str();
    n.port = r.i32();
    n.is_recovery = r.u8() != 0;
    n.aux_id = r.i32();
    m->control.node.push_back(n);
  }
  return r.ok;
}

}  // namespace net
}  // namespace bps
