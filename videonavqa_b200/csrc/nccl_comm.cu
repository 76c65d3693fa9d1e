// K10: gradient all-reduce over NVLink 5 / NVSwitch through the C ABI; new work - the
// reference is single-GPU.  One process per GPU owns ONE ncclComm_t and ONE dedicated communication stream here:
//
//   vnqa_nccl_allreduce_sum_f32(buf, n, after)   records an event on `after` (the compute stream that produced buf),
//                                                makes the communication stream wait for it and enqueues an in-place
//                                                ncclAllReduce(SUM) there - the compute stream is never blocked, so
//                                                the collective overlaps the rest of the backward pass;
//   vnqa_nccl_join(stream)                       makes `stream` wait for everything queued on the communication stream.
//
// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on a particular libnccl:
// the caller passes the path of the NCCL build to use (videonavqa_b200.dp passes the one bundled with PyTorch).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.h"

namespace vnqa {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
std::mutex g_mu;
ncclComm_t g_comm = nullptr;
cudaStream_t g_stream = nullptr;      // the communication stream
cudaEvent_t g_ready = nullptr;        // "buffer is final" on the compute stream
cudaEvent_t g_done = nullptr;         // everything queued on the communication stream so far
int g_rank = -1, g_nranks = 0;

template <class F>
bool sym(F& fn, const char* name) {
  fn = reinterpret_cast<F>(dlsym(g_api.handle, name));
  return fn != nullptr;
}

#define VNQA_CHECK_NCCL(expr)                                                                          \
  do {                                                                                                 \
    ncclResult_t _r = (expr);                                                                          \
    if (_r != ncclSuccess) {                                                                           \
      ::vnqa::set_error("%s failed: %s", #expr, g_api.GetErrorString ? g_api.GetErrorString(_r) : "?"); \
      return -4;                                                                                       \
    }                                                                                                  \
  } while (0)

}  // namespace
}  // namespace vnqa

using namespace vnqa;

extern "C" int vnqa_nccl_load(const char* libpath) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_api.handle) return 0;
  const char* path = (libpath && *libpath) ? libpath : "libnccl.so.2";
  g_api.handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  VNQA_REQUIRE(g_api.handle != nullptr, "vnqa_nccl_load: dlopen(%s) failed: %s", path, dlerror());
  const bool ok = sym(g_api.GetVersion, "ncclGetVersion") && sym(g_api.GetUniqueId, "ncclGetUniqueId") &&
                  sym(g_api.CommInitRank, "ncclCommInitRank") && sym(g_api.CommDestroy, "ncclCommDestroy") &&
                  sym(g_api.CommAbort, "ncclCommAbort") && sym(g_api.AllReduce, "ncclAllReduce") &&
                  sym(g_api.Broadcast, "ncclBroadcast") && sym(g_api.AllGather, "ncclAllGather") &&
                  sym(g_api.GetErrorString, "ncclGetErrorString");
  if (!ok) {
    dlclose(g_api.handle);
    g_api = NcclApi{};
    set_error("vnqa_nccl_load: %s does not export the NCCL API", path);
    return -1;
  }
  return 0;
}

extern "C" int vnqa_nccl_version(void) {
  if (!g_api.handle) return 0;
  int v = 0;
  return g_api.GetVersion(&v) == ncclSuccess ? v : 0;
}

/* rank 0 creates the 128-byte id and hands it to the other ranks by any host channel (dp.py: torch.distributed store) */
extern "C" int vnqa_nccl_unique_id(void* out128) {
  VNQA_REQUIRE(g_api.handle, "vnqa_nccl_unique_id: call vnqa_nccl_load first");
  VNQA_REQUIRE(out128, "vnqa_nccl_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  VNQA_CHECK_NCCL(g_api.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

/* Collective over all ranks: communicator for the CURRENT CUDA device + the communication stream and its events. */
extern "C" int vnqa_nccl_init(const void* id128, int32_t rank, int32_t nranks) {
  VNQA_REQUIRE(g_api.handle, "vnqa_nccl_init: call vnqa_nccl_load first");
  VNQA_REQUIRE(id128 && nranks >= 1 && rank >= 0 && rank < nranks, "vnqa_nccl_init: bad arguments");
  VNQA_REQUIRE(g_comm == nullptr, "vnqa_nccl_init: communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  VNQA_CHECK_NCCL(g_api.CommInitRank(&g_comm, nranks, id, rank));
  int lo = 0, hi = 0;
  VNQA_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  // highest priority: a collective's CTAs are placed before the next persistent compute kernel's
  VNQA_CHECK_CUDA(cudaStreamCreateWithPriority(&g_stream, cudaStreamNonBlocking, hi));
  VNQA_CHECK_CUDA(cudaEventCreateWithFlags(&g_ready, cudaEventDisableTiming));
  VNQA_CHECK_CUDA(cudaEventCreateWithFlags(&g_done, cudaEventDisableTiming));
  g_rank = rank;
  g_nranks = nranks;
  return 0;
}

extern "C" int vnqa_nccl_rank(void) { return g_rank; }
extern "C" int vnqa_nccl_nranks(void) { return g_nranks; }
extern "C" void* vnqa_nccl_stream(void) { return g_stream; }

/* In-place SUM all-reduce of buf[0..count) on the communication stream, ordered after everything queued so far on
 * `after_stream`.  Returns immediately; vnqa_nccl_join orders a stream after the result. */
extern "C" int vnqa_nccl_allreduce_sum_f32(float* buf, int64_t count, void* after_stream) {
  VNQA_REQUIRE(g_comm, "vnqa_nccl_allreduce_sum_f32: communicator not initialised");
  VNQA_REQUIRE(buf && count > 0, "vnqa_nccl_allreduce_sum_f32: empty buffer");
  VNQA_CHECK_CUDA(cudaEventRecord(g_ready, as_stream(after_stream)));   // NULL is the legacy default stream
  VNQA_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_ready, 0));
  VNQA_CHECK_NCCL(g_api.AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, g_comm, g_stream));
  return 0;
}

/* fp64 variant for small checksums / metrics (same ordering rules) */
extern "C" int vnqa_nccl_allgather_f64(const double* send, double* recv, int64_t count, void* after_stream) {
  VNQA_REQUIRE(g_comm, "vnqa_nccl_allgather_f64: communicator not initialised");
  VNQA_CHECK_CUDA(cudaEventRecord(g_ready, as_stream(after_stream)));
  VNQA_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_ready, 0));
  VNQA_CHECK_NCCL(g_api.AllGather(send, recv, (size_t)count, ncclFloat64, g_comm, g_stream));
  return 0;
}

/* Broadcast from `root`, in place, on the communication stream after `after_stream` (start-up weight sync). */
extern "C" int vnqa_nccl_broadcast_f32(float* buf, int64_t count, int32_t root, void* after_stream) {
  VNQA_REQUIRE(g_comm, "vnqa_nccl_broadcast_f32: communicator not initialised");
  VNQA_REQUIRE(buf && count > 0 && root >= 0 && root < g_nranks, "vnqa_nccl_broadcast_f32: bad arguments");
  VNQA_CHECK_CUDA(cudaEventRecord(g_ready, as_stream(after_stream)));
  VNQA_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_ready, 0));
  VNQA_CHECK_NCCL(g_api.Broadcast(buf, buf, (size_t)count, ncclFloat32, root, g_comm, g_stream));
  return 0;
}

/* `stream` waits (on the device) for every collective queued so far. */
extern "C" int vnqa_nccl_join(void* stream) {
  VNQA_REQUIRE(g_comm, "vnqa_nccl_join: communicator not initialised");
  VNQA_CHECK_CUDA(cudaEventRecord(g_done, g_stream));
  VNQA_CHECK_CUDA(cudaStreamWaitEvent(as_stream(stream), g_done, 0));
  return 0;
}

extern "C" int vnqa_nccl_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(g_stream);
    g_api.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  if (g_ready) { cudaEventDestroy(g_ready); g_ready = nullptr; }
  if (g_done) { cudaEventDestroy(g_done); g_done = nullptr; }
  g_rank = -1;
  g_nranks = 0;
  return 0;
}

/* Tear the communicator down without waiting for in-flight collectives (a peer died). */
extern "C" int vnqa_nccl_abort(void) {
  if (g_comm) {
    g_api.CommAbort(g_comm);
    g_comm = nullptr;
  }
  return 0;
}
