// comm.cu — data-parallel gradient exchange over NCCL (NVLink 5 / NVSwitch).
//
// The reference GPU trainer is single-GPU (src/CuBaseLib/cudevice.cc:22-65); its CPU trainer sums per-thread
// gradients with a hand-rolled reduce (src/TNetLib/Platform.h:300-335, BiasedLinearity.cc:90-128).  Here every rank
// (one process per GPU) computes the gradient of its rows of the bunch and the per-layer [dW | db] buffers are
// summed with ncclAllReduce on a dedicated communication stream, ordered against the compute stream with events
// so that layer l's exchange overlaps layer l-1's backward GEMMs.
//
// libnccl is resolved with dlopen at first use: libtnetb200.so has no link-time dependency on it, so the library
// loads (and every symbol of include/tnet_b200.h is exported) on machines without NCCL.
#include <dlfcn.h>

#include "common.cuh"

namespace tnb {

typedef struct { char internal[TNB_NCCL_ID_BYTES]; } NcclUniqueId;  // ncclUniqueId: 128 opaque bytes (nccl.h)
typedef void *NcclComm;
enum { kNcclFloat = 7, kNcclSum = 0 };  // ncclFloat32, ncclSum (nccl.h enums, stable since NCCL 2.0)

static int (*p_GetUniqueId)(NcclUniqueId *) = nullptr;
static int (*p_CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
static int (*p_AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
static int (*p_CommDestroy)(NcclComm) = nullptr;
static const char *(*p_GetErrorString)(int) = nullptr;

static int load_nccl() {
  if (p_AllReduce) return TNB_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return TNB_ERR_NCCL; }
  p_GetUniqueId = (int (*)(NcclUniqueId *))dlsym(h, "ncclGetUniqueId");
  p_CommInitRank = (int (*)(NcclComm *, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
  p_AllReduce = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
  p_CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
  p_GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
  if (!p_GetUniqueId || !p_CommInitRank || !p_AllReduce || !p_CommDestroy) {
    p_AllReduce = nullptr;
    set_error("libnccl is missing required symbols");
    return TNB_ERR_NCCL;
  }
  return TNB_OK;
}

#define TNB_NCCL(call)                                                                            \
  do {                                                                                            \
    int r__ = (call);                                                                             \
    if (r__ != 0) {                                                                               \
      tnb::set_error("NCCL error %d (%s) at %s:%d", r__, p_GetErrorString ? p_GetErrorString(r__) : "?", __FILE__, __LINE__); \
      return TNB_ERR_NCCL;                                                                        \
    }                                                                                             \
  } while (0)

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_comm_unique_id(unsigned char id[TNB_NCCL_ID_BYTES]) {
  TNB_ARG(id, "null");
  int rc = load_nccl();
  if (rc != TNB_OK) return rc;
  NcclUniqueId u;
  TNB_NCCL(p_GetUniqueId(&u));
  memcpy(id, u.internal, TNB_NCCL_ID_BYTES);
  return TNB_OK;
}

int tnb_comm_init(TnbContext *ctx, const unsigned char id[TNB_NCCL_ID_BYTES], int rank, int world) {
  TNB_ARG(ctx && id, "null");
  TNB_ARG(world >= 1 && rank >= 0 && rank < world, "rank/world");
  int rc = load_nccl();
  if (rc != TNB_OK) return rc;
  TNB_CUDA(cudaSetDevice(ctx->device));
  NcclUniqueId u;
  memcpy(u.internal, id, TNB_NCCL_ID_BYTES);
  NcclComm comm = nullptr;
  TNB_NCCL(p_CommInitRank(&comm, world, u, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return TNB_OK;
}

int tnb_comm_destroy(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  if (ctx->nccl_comm && p_CommDestroy) {
    cudaStreamSynchronize(ctx->comm_stream);
    p_CommDestroy((NcclComm)ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
  return TNB_OK;
}

int tnb_comm_world(TnbContext *ctx, int *rank, int *world) {
  TNB_ARG(ctx && rank && world, "null");
  *rank = ctx->rank;
  *world = ctx->world;
  return TNB_OK;
}

int tnb_allreduce_sum(TnbContext *ctx, float *buf, size_t count) {
  TNB_ARG(ctx && buf, "null");
  if (ctx->world == 1 || count == 0) return TNB_OK;  // single rank: the sum is the buffer itself
  TNB_ARG(ctx->nccl_comm != nullptr, "communicator not initialised");
  // comm stream waits for everything enqueued so far on the compute stream (the producer of buf)
  TNB_CUDA(cudaEventRecord(ctx->ev_compute, ctx->stream));
  TNB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  TNB_NCCL(p_AllReduce(buf, buf, count, kNcclFloat, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->comm_stream));
  return TNB_OK;
}

int tnb_comm_wait(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  if (ctx->world == 1) return TNB_OK;
  TNB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  TNB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  return TNB_OK;
}

}  // extern "C"
