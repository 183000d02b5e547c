/* C API of the native runtime (host and device tensors, CPU-server mode).
 *
 * Parity: the `extern "C"` surface of /root/reference/byteps/common/operations.h:28-84 that the reference's ctypes
 * loader (byteps/common/__init__.py:52-139) and its TensorFlow / MXNet C++ ops call.  Framework plugins written in
 * C/C++ link against (or dlopen) byteps_b200/_core*.so and drive the same registry, scheduler, KV transport and PS
 * worker pipeline the Python front ends use - no Python in the process.
 *
 * Scope: tensors in host memory, or in CUDA device memory (byteps_push_pull_device: the reference's EnqueueTensor
 * handles device tensors the same way in distributed mode - COPYD2H, PUSH, PULL, COPYH2D per partition,
 * operations.cc:182-281, core_loops.cc:378-443,650-753), summed by the servers (DMLC_NUM_SERVER >= 1); or a job of one
 * process (push_pull is then the identity).  The single-box NVLink engines are reached through byteps_b200.torch /
 * .dlpack.
 * All functions return 0 (or a non-negative id) on success and a negative value on failure; byteps_last_error() gives
 * the reason.  Every function is thread safe.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes (core/types.h): */
enum { BYTEPS_FLOAT32 = 0, BYTEPS_FLOAT64 = 1, BYTEPS_FLOAT16 = 2, BYTEPS_UINT8 = 3, BYTEPS_INT32 = 4, BYTEPS_INT8 = 5,
       BYTEPS_INT64 = 6, BYTEPS_BFLOAT16 = 7 };

int byteps_init(void);        /* reads DMLC_* / BYTEPS_* from the environment, joins the cluster (blocking) */
int byteps_lazy_init(void);   /* same; kept because the reference exposes both */
int byteps_shutdown(void);
int byteps_suspend(void);     /* leave the cluster, keep tensor names -> keys */
int byteps_resume(int num_workers, int num_servers, int global_rank);   /* rejoin with a new topology */

int byteps_rank(void);
int byteps_size(void);
int byteps_local_rank(void);
int byteps_local_size(void);

/* Name -> declared key (idempotent; declare in the same order on every worker). */
int byteps_declare_tensor(const char* name);
/* Same, with compressor configuration ("byteps_compressor_type" -> "topk", ...; see docs/gradient-compression.md). */
int byteps_declare_tensor_kwargs(const char* name, const char* const* keys, const char* const* values, int n);

/* In-place asynchronous push_pull of `nbytes` at `data`: sum over all workers, divided by size() if `average`.
 * Returns a handle.  The buffer must stay valid and untouched until the handle completes. */
int byteps_push_pull(const char* name, void* data, int64_t nbytes, int dtype, int average, int priority, int version);
int byteps_poll(int handle);   /* 1 = finished, 0 = in flight */
int byteps_wait(int handle);   /* blocks; releases the handle */

/* Same contract for `nbytes` of CUDA device memory at `dev` (in place).  `ready_event` is a cudaEvent_t that must have
 * completed before the data is read (NULL: the data is ready now).  The exchange is pipelined per partition: the D2H
 * copy of partition i, its push, the pulls and the H2D copy of partition j overlap on two side streams.  Needs
 * libbyteps_b200_cuda.so next to this library (or BYTEPS_CUDA_LIB=<path>).
 * byteps_wait_device() blocks until every partition's H2D copy has been ENQUEUED, then makes `stream` (a cudaStream_t;
 * NULL = the legacy default stream) wait for them: work enqueued on `stream` afterwards sees the result.  Passing
 * stream == (void*)-1 blocks the host until the copies have landed instead. */
int byteps_push_pull_device(const char* name, void* dev, int64_t nbytes, int dtype, int average, int priority,
                            int version, void* ready_event);
int byteps_wait_device(int handle, void* stream);

/* A buffer in a registered shared-memory window ("BytePS_ShM_<name>").  Tensors that live in such a buffer reach a
 * colocated server by reference (BYTEPS_ENABLE_IPC=1 or DMLC_PS_VAN_TYPE=shm): the server reads the window and writes
 * the result back into it, no payload crosses the transport.  Returns NULL on failure. */
void* byteps_shm_alloc(const char* name, int64_t nbytes);
int byteps_shm_free(const char* name);

/* Run the role named by DMLC_ROLE ("server" or "scheduler") to completion: join the cluster, serve until every node
 * has left (the reference's byteps_server(), server.cc:458-531).  Blocking. */
int byteps_server(void);

/* Oldest unread telemetry sample; (*ts_ms = 0, *mbps = -5.0) when there is none (the reference's sentinel). */
void byteps_get_pushpull_speed(int64_t* ts_ms, double* mbps);
const char* byteps_last_error(void);

#ifdef __cplusplus
}
#endif
