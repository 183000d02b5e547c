#include "core/numa.h"

#include <dirent.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "core/env.h"

namespace bps {

namespace {
// <linux/mempolicy.h> values (the header is not guaranteed to be installed)
constexpr int kMpolPreferred = 1;
constexpr unsigned kMpolMfMove = 1u << 1;

std::string read_line(const std::string& path) {
  std::ifstream f(path);
  std::string s;
  if (f) std::getline(f, s);
  return s;
}
}  // namespace

std::vector<int> parse_cpu_list(const std::string& s) {
  std::vector<int> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && !isdigit((unsigned char)s[i])) ++i;
    if (i >= s.size()) break;
    long a = strtol(s.c_str() + i, nullptr, 10);
    while (i < s.size() && isdigit((unsigned char)s[i])) ++i;
    long b = a;
    if (i < s.size() && s[i] == '-') {
      ++i;
      b = strtol(s.c_str() + i, nullptr, 10);
      while (i < s.size() && isdigit((unsigned char)s[i])) ++i;
    }
    for (long c = a; c <= b && c < 65536; ++c) out.push_back((int)c);
  }
  return out;
}

int numa_num_nodes() {
  const long long fake = env_int("BYTEPS_NUMA_FAKE_NODES", 0);
  if (fake > 0) return (int)fake;
  static const int n = [] {
    int count = 0;
    if (DIR* d = opendir("/sys/devices/system/node")) {
      while (dirent* e = readdir(d))
        if (strncmp(e->d_name, "node", 4) == 0 && isdigit((unsigned char)e->d_name[4])) ++count;
      closedir(d);
    }
    return count > 0 ? count : 1;
  }();
  return n;
}

int numa_node_of_pci(const std::string& bus_id) {
  std::string id = bus_id;
  std::transform(id.begin(), id.end(), id.begin(), [](unsigned char c) { return (char)tolower(c); });
  std::string s = read_line("/sys/bus/pci/devices/" + id + "/numa_node");
  if (s.empty() && id.size() > 5 && id.compare(0, 4, "0000") != 0) s = read_line("/sys/bus/pci/devices/0000:" + id + "/numa_node");
  if (s.empty()) return -1;
  const long v = strtol(s.c_str(), nullptr, 10);
  return v >= 0 ? (int)v : -1;      // the kernel reports -1 on single-node hosts
}

std::vector<int> numa_cpus_of_node(int node) {
  if (node < 0) return {};
  return parse_cpu_list(read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"));
}

bool numa_bind_memory(void* p, size_t len, int node) {
#ifdef SYS_mbind
  if (!p || len == 0 || node < 0 || node >= 1024) return false;
  const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
  const uintptr_t lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + len + page - 1) & ~(page - 1);
  unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
  mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
  const long rc = syscall(SYS_mbind, (void*)lo, (unsigned long)(hi - lo), kMpolPreferred, mask,
                          (unsigned long)(node + 2), kMpolMfMove);
  return rc == 0;
#else
  (void)p; (void)len; (void)node;
  return false;
#endif
}

int numa_node_of_addr(const void* p) {
#ifdef SYS_move_pages
  const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
  void* pages[1] = {(void*)((uintptr_t)p & ~(page - 1))};
  int status[1] = {-1};
  if (syscall(SYS_move_pages, 0, 1ul, pages, nullptr, status, 0) != 0) return -1;
  return status[0] >= 0 ? status[0] : -1;
#else
  (void)p;
  return -1;
#endif
}

bool numa_pin_thread_to_node(int node) {
  const std::vector<int> cpus = numa_cpus_of_node(node);
  if (cpus.empty()) return false;
  cpu_set_t allowed, want;
  CPU_ZERO(&allowed);
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
  int n = 0;
  for (int c : cpus)
    if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) {
      CPU_SET(c, &want);
      ++n;
    }
  if (n == 0) return false;          // the process was confined elsewhere (taskset / cpuset): leave it alone
  return sched_setaffinity(0, sizeof want, &want) == 0;
}

int numa_prefer_node_for_process(int node) {
  int done = 0;
  if (numa_pin_thread_to_node(node)) done |= 1;
#ifdef SYS_set_mempolicy
  if (node >= 0 && node < 1024) {
    unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    if (syscall(SYS_set_mempolicy, kMpolPreferred, mask, (unsigned long)(node + 2)) == 0) done |= 2;
  }
#endif
  return done;
}

bool numa_aware() { return env_bool("BYTEPS_NUMA_AWARE", false); }   // read on set-up paths only: not cached

}  // namespace bps
