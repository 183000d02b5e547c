/* A complete BytePS job without Python: `capi_job server` runs the scheduler / server role named by DMLC_ROLE,
 * `capi_job worker` joins as a worker, sums float / int / compressed tensors through the servers and checks them.
 * Built and driven by tests/test_capi.py against byteps_b200/libbyteps_b200.so. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "capi/byteps_c_api.h"

#define CHECK(c)                                                                              \
  do {                                                                                        \
    if (!(c)) {                                                                               \
      fprintf(stderr, "%s:%d: CHECK(%s) failed: %s\n", __FILE__, __LINE__, #c, byteps_last_error()); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  if (strcmp(argv[1], "server") == 0) return byteps_server();
  CHECK(byteps_init() == 0);
  const int rank = byteps_rank(), size = byteps_size();
  CHECK(size == atoi(getenv("DMLC_NUM_WORKER")) && rank == atoi(getenv("DMLC_WORKER_ID")));
  CHECK(byteps_declare_tensor("grad") == byteps_declare_tensor("grad"));     /* idempotent */

  /* 10 MB of floats: three 4 MB partitions, averaged */
  const long n = 2500000;
  float* g = (float*)malloc(n * sizeof(float));
  for (int it = 0; it < 3; ++it) {
    for (long i = 0; i < n; ++i) g[i] = (float)((i % 97) * (rank + 1) + it);
    int h = byteps_push_pull("grad", g, n * (long)sizeof(float), BYTEPS_FLOAT32, 1, 0, it);
    CHECK(h >= 0);
    CHECK(byteps_wait(h) == 0);
    double tot = size * (size + 1) / 2.0;
    for (long i = 0; i < n; i += 9973) CHECK(fabs(g[i] - ((i % 97) * tot / size + it)) < 1e-3);
  }
  /* several tensors in flight, polled; integer average is a floor division */
  long long ints[4] = {7 * (rank + 1), -9 * (rank + 1), 100, 1};
  double d[1000];
  for (int i = 0; i < 1000; ++i) d[i] = 0.5 * i * (rank + 1);
  int h1 = byteps_push_pull("ints", ints, sizeof(ints), BYTEPS_INT64, 1, 0, 0);
  int h2 = byteps_push_pull("doubles", d, sizeof(d), BYTEPS_FLOAT64, 0, 5, 0);
  CHECK(h1 >= 0 && h2 >= 0);
  while (!byteps_poll(h2)) {
  }
  CHECK(byteps_wait(h2) == 0 && byteps_wait(h1) == 0);
  long long tot = (long long)size * (size + 1) / 2;
  CHECK(ints[0] == 7 * tot / size && ints[2] == 100 && ints[3] == 1);
  CHECK(ints[1] == (long long)floor(-9.0 * tot / size));
  for (int i = 0; i < 1000; i += 37) CHECK(fabs(d[i] - 0.5 * i * tot) < 1e-9);
  /* top-k compression declared with kwargs: each worker contributes one distinct spike */
  const char* keys[] = {"byteps_compressor_type", "byteps_compressor_k"};
  const char* vals[] = {"topk", "4"};
  CHECK(byteps_declare_tensor_kwargs("sparse", keys, vals, 2) >= 0);
  static float s[40000];
  memset(s, 0, sizeof(s));
  s[100 + rank] = 3.0f + rank;
  int h3 = byteps_push_pull("sparse", s, sizeof(s), BYTEPS_FLOAT32, 0, 0, 0);
  CHECK(h3 >= 0 && byteps_wait(h3) == 0);
  for (int r = 0; r < size; ++r) CHECK(s[100 + r] == 3.0f + r);
  CHECK(s[0] == 0.0f && s[39999] == 0.0f);
  /* a tensor in a registered window: pushed by reference when the transport allows it (shm van / IPC) */
  float* win = (float*)byteps_shm_alloc("win", 1 << 20);
  CHECK(win != NULL);
  for (int i = 0; i < (1 << 18); ++i) win[i] = (float)(i % 11) * (rank + 1);
  int h4 = byteps_push_pull("window", win, 1 << 20, BYTEPS_FLOAT32, 0, 0, 0);
  CHECK(h4 >= 0 && byteps_wait(h4) == 0);
  for (int i = 0; i < (1 << 18); i += 101) CHECK(win[i] == (float)(i % 11) * (size * (size + 1) / 2));
  /* elastic: leave and rejoin the same cluster; the names keep their keys */
  int key_before = byteps_declare_tensor("grad");
  CHECK(byteps_suspend() == 0);
  CHECK(byteps_declare_tensor("grad") == key_before);
  CHECK(byteps_shutdown() == 0);
  CHECK(byteps_shm_free("win") == 0);
  free(g);
  printf("capi worker %d/%d ok\n", rank, size);
  return 0;
}
