/* A BytePS job without Python whose tensors live in GPU memory: `capi_gpu_job server` runs the scheduler /
 * server role, `capi_gpu_job worker` joins as a worker, push_pulls device buffers through the servers
 * (byteps_push_pull_device: COPYD2H -> PUSH -> PULL -> COPYH2D per partition) and checks them.
 * Built and driven by tests/test_capi.py against libbyteps_b200.so + libbyteps_b200_cuda.so + libcudart. */
#include <cuda_runtime_api.h>
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
  int ndev = 0;
  CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0);
  CHECK(cudaSetDevice(rank % ndev) == cudaSuccess);
  cudaStream_t stream;
  CHECK(cudaStreamCreate(&stream) == cudaSuccess);
  /* 40 MB of floats: ten 4 MB partitions in flight at once */
  const long n = 10 * 1000 * 1000;
  float* host = (float*)malloc(n * sizeof(float));
  float* dev = NULL;
  CHECK(cudaMalloc((void**)&dev, n * sizeof(float)) == cudaSuccess);
  cudaEvent_t ready;
  CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) == cudaSuccess);
  for (int it = 0; it < 3; ++it) {
    for (long i = 0; i < n; ++i) host[i] = (float)((i % 89) * (rank + 1) + it);
    CHECK(cudaMemcpyAsync(dev, host, n * sizeof(float), cudaMemcpyHostToDevice, stream) == cudaSuccess);
    CHECK(cudaEventRecord(ready, stream) == cudaSuccess);          /* the data is ready when this event fires */
    int h = byteps_push_pull_device("dgrad", dev, n * (long)sizeof(float), BYTEPS_FLOAT32, 1, 0, it, ready);
    CHECK(h >= 0);
    CHECK(byteps_wait_device(h, stream) == 0);                     /* `stream` now waits for the H2D copies */
    memset(host, 0, n * sizeof(float));
    CHECK(cudaMemcpyAsync(host, dev, n * sizeof(float), cudaMemcpyDeviceToHost, stream) == cudaSuccess);
    CHECK(cudaStreamSynchronize(stream) == cudaSuccess);
    double tot = size * (size + 1) / 2.0;
    for (long i = 0; i < n; i += 4999) CHECK(fabs(host[i] - ((i % 89) * tot / size + it)) < 1e-3);
  }
  /* half precision sum, host-blocking wait */
  unsigned short* hh = (unsigned short*)malloc(4096 * 2);
  unsigned short* dh = NULL;
  CHECK(cudaMalloc((void**)&dh, 4096 * 2) == cudaSuccess);
  for (int i = 0; i < 4096; ++i) hh[i] = 0x3f80;                   /* bf16 1.0 */
  CHECK(cudaMemcpy(dh, hh, 4096 * 2, cudaMemcpyHostToDevice) == cudaSuccess);
  int h2 = byteps_push_pull_device("dbf16", dh, 4096 * 2, BYTEPS_BFLOAT16, 0, 0, 0, NULL);
  CHECK(h2 >= 0 && byteps_wait_device(h2, (void*)-1) == 0);
  CHECK(cudaMemcpy(hh, dh, 4096 * 2, cudaMemcpyDeviceToHost) == cudaSuccess);
  /* sum of `size` ones in bf16: 2.0 = 0x4000, 3.0 = 0x4040, 4.0 = 0x4080 */
  unsigned short want = size == 1 ? 0x3f80 : size == 2 ? 0x4000 : size == 3 ? 0x4040 : 0x4080;
  CHECK(hh[0] == want && hh[4095] == want);
  /* host pointers are rejected with a clear error */
  CHECK(byteps_push_pull_device("bad", host, 16, BYTEPS_FLOAT32, 0, 0, 0, NULL) < 0);
  CHECK(byteps_shutdown() == 0);
  cudaFree(dev);
  cudaFree(dh);
  free(host);
  free(hh);
  printf("capi gpu worker %d/%d ok\n", rank, size);
  return 0;
}
