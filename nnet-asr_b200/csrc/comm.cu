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
#include <stdlib.h>

#include <condition_variable>
#include <mutex>

#include "common.cuh"

// ---- ranks as THREADS of one process (bin/TNetCu --GPUS=N): no NCCL, no CUDA IPC.  The ranks exchange bytes through a shared host
// buffer behind a generation barrier; with peer access enabled every rank's device pointers are valid on every other rank's GPU.
struct TnbLocalGroup_ {
  int world = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long generation = 0;
  std::vector<unsigned char> gather;  // [world][n] of the exchange in flight
  int devices[TNB_MAX_PEERS];
};

namespace tnb {

typedef struct { char internal[TNB_NCCL_ID_BYTES]; } NcclUniqueId;  // ncclUniqueId: 128 opaque bytes (nccl.h)
typedef void *NcclComm;
enum { kNcclFloat = 7, kNcclSum = 0 };  // ncclFloat32, ncclSum (nccl.h enums, stable since NCCL 2.0)

static int (*p_GetUniqueId)(NcclUniqueId *) = nullptr;
static int (*p_CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
static int (*p_AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
static int (*p_ReduceScatter)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
static int (*p_AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
static int (*p_GroupStart)() = nullptr;
static int (*p_GroupEnd)() = nullptr;
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
  p_ReduceScatter = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclReduceScatter");
  p_AllGather = (int (*)(const void *, void *, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
  p_GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  p_GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  p_CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
  p_GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
  if (!p_GetUniqueId || !p_CommInitRank || !p_AllReduce || !p_CommDestroy || !p_ReduceScatter || !p_AllGather) {
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

// all ranks of a local group meet here; the last one to arrive releases the others
static void local_barrier(TnbLocalGroup_ *g, std::unique_lock<std::mutex> &lk) {
  const unsigned long gen = g->generation;
  if (++g->arrived == g->world) {
    g->arrived = 0;
    g->generation++;
    g->cv.notify_all();
  } else {
    g->cv.wait(lk, [&] { return g->generation != gen; });
  }
}

int comm_allgather_bytes(TnbContext *ctx, unsigned char *all, size_t n) {
  TNB_ARG(ctx && all && n > 0, "null");
  if (ctx->world == 1) return TNB_OK;
  if (ctx->local_group) {
    TnbLocalGroup_ *g = (TnbLocalGroup_ *)ctx->local_group;
    std::unique_lock<std::mutex> lk(g->mu);
    if (g->gather.size() != n * (size_t)g->world) g->gather.assign(n * (size_t)g->world, 0);
    memcpy(&g->gather[n * (size_t)ctx->rank], all + n * (size_t)ctx->rank, n);
    local_barrier(g, lk);               // everybody has written its slice
    memcpy(all, g->gather.data(), n * (size_t)g->world);
    local_barrier(g, lk);               // everybody has read: the buffer may be reused
    return TNB_OK;
  }
  int rc = load_nccl();
  if (rc != TNB_OK) return rc;
  TNB_ARG(ctx->nccl_comm != nullptr, "communicator not initialised");
  unsigned char *d = nullptr;
  const size_t total = n * (size_t)ctx->world, mine = n * (size_t)ctx->rank;
  TNB_CUDA(cudaMalloc((void **)&d, total));
  TNB_CUDA(cudaMemcpyAsync(d + mine, all + mine, n, cudaMemcpyHostToDevice, ctx->comm_stream));
  int r = p_AllGather(d + mine, d, n, 0 /* ncclInt8 */, (NcclComm)ctx->nccl_comm, ctx->comm_stream);
  if (r != 0) { cudaFree(d); TNB_NCCL(r); }
  TNB_CUDA(cudaMemcpyAsync(all, d, total, cudaMemcpyDeviceToHost, ctx->comm_stream));
  TNB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  TNB_CUDA(cudaFree(d));
  return TNB_OK;
}

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
  // The backward GEMMs occupy 128 of the 148 SMs with CTAs that need a whole SM each; a collective launched next to them must
  // fit into the SMs that are left, or the next GEMM loses a wave waiting for the collective's CTAs to leave.  Measured on two
  // B200 (profiles/r01_dp_notes.md): NCCL moves ~21 GB/s of all-reduce algorithm bandwidth per CTA; 24 CTAs gave the best step
  // time (16: -1.5 %, 32: -4 %).  An explicit NCCL_MAX_CTAS wins.
  setenv("NCCL_MAX_CTAS", "24", 0);
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

int tnb_local_group_create(TnbLocalGroup **out, int world) {
  TNB_ARG(out && world >= 1 && world <= TNB_MAX_PEERS, "world");
  TnbLocalGroup_ *g = new TnbLocalGroup_();
  g->world = world;
  for (int i = 0; i < TNB_MAX_PEERS; i++) g->devices[i] = -1;
  *out = g;
  return TNB_OK;
}

int tnb_local_group_destroy(TnbLocalGroup *g) {
  delete g;
  return TNB_OK;
}

int tnb_comm_init_local(TnbContext *ctx, TnbLocalGroup *g, int rank) {
  TNB_ARG(ctx && g, "null");
  TNB_ARG(rank >= 0 && rank < g->world, "rank");
  TNB_ARG(ctx->nccl_comm == nullptr && ctx->local_group == nullptr, "the context already has a communicator");
  ctx->local_group = g;
  ctx->rank = rank;
  ctx->world = g->world;
  if (g->world == 1) return TNB_OK;
  // every rank learns every rank's device, then enables peer access to the others (NVLink / NVSwitch on an HGX box)
  std::vector<unsigned char> all(sizeof(int) * (size_t)g->world);
  memcpy(&all[sizeof(int) * (size_t)rank], &ctx->device, sizeof(int));
  int rc = comm_allgather_bytes(ctx, all.data(), sizeof(int));
  if (rc != TNB_OK) return rc;
  TNB_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < g->world; r++) {
    int dev;
    memcpy(&dev, &all[sizeof(int) * (size_t)r], sizeof(int));
    g->devices[r] = dev;
    if (r == rank) continue;
    if (dev == ctx->device) { set_error("local ranks %d and %d share GPU %d: kernels of different ranks wait for each other and need a GPU each", rank, r, dev); return TNB_ERR_ARG; }
    int can = 0;
    TNB_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, dev));
    if (!can) { set_error("GPU %d cannot access GPU %d directly (no peer access)", ctx->device, dev); return TNB_ERR_UNSUPPORTED; }
    cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TNB_CUDA(e);
    cudaGetLastError();
  }
  return TNB_OK;
}

int tnb_comm_destroy(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  ctx->local_group = nullptr;
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

int tnb_allreduce_sum(TnbContext *ctx, float *buf, size_t count) { return tnb_allreduce_sum_ev(ctx, buf, count, nullptr, nullptr); }

int tnb_allreduce_sum_ev(TnbContext *ctx, float *buf, size_t count, void *event, void *done) {
  TNB_ARG(ctx && buf, "null");
  if (count == 0) return TNB_OK;
  // comm stream waits for everything enqueued so far on the compute stream (the producer of buf) and for `event`
  TNB_CUDA(cudaEventRecord(ctx->ev_compute, ctx->main_stream));
  TNB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  if (event) TNB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, (cudaEvent_t)event, 0));
  static int skip = -1;  // TNB_DP_SKIP_COMM=1: timing experiment only (the data-parallel step without its collectives; results are wrong)
  if (skip < 0) { const char *e = getenv("TNB_DP_SKIP_COMM"); skip = (e && atoi(e) != 0) ? 1 : 0; }
  if (ctx->world > 1 && !skip) {  // single rank: the sum is the buffer itself
    TNB_ARG(ctx->nccl_comm != nullptr, "communicator not initialised");
    TNB_NCCL(p_AllReduce(buf, buf, count, kNcclFloat, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->comm_stream));
  }
  if (done) TNB_CUDA(cudaEventRecord((cudaEvent_t)done, ctx->comm_stream));
  return TNB_OK;
}

// Several buffers summed in ONE NCCL launch (ncclGroupStart/End).  A call costs a fixed latency on top of its bytes — on 8 B200
// an all-reduce of one 16.8 MB layer takes 101 us alone, the whole 112 MB model in one call 340 us — so the layers whose exchange
// is issued together anyway (CuNetwork's deferred middle layers) go out as one group.
int tnb_allreduce_sum_multi(TnbContext *ctx, float *const *bufs, const size_t *counts, int n, void *const *events, int n_events, void *done) {
  TNB_ARG(ctx && (n == 0 || (bufs && counts)) && n >= 0 && n_events >= 0 && (n_events == 0 || events), "null");
  TNB_CUDA(cudaEventRecord(ctx->ev_compute, ctx->main_stream));
  TNB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_compute, 0));
  for (int i = 0; i < n_events; i++)
    if (events[i]) TNB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, (cudaEvent_t)events[i], 0));
  static int skip = -1;
  if (skip < 0) { const char *e = getenv("TNB_DP_SKIP_COMM"); skip = (e && atoi(e) != 0) ? 1 : 0; }
  if (ctx->world > 1 && !skip && n > 0) {
    TNB_ARG(ctx->nccl_comm != nullptr, "communicator not initialised");
    const bool group = p_GroupStart && p_GroupEnd && n > 1;
    if (group) TNB_NCCL(p_GroupStart());
    for (int i = 0; i < n; i++) {
      if (counts[i] == 0) continue;
      int r = p_AllReduce(bufs[i], bufs[i], counts[i], kNcclFloat, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->comm_stream);
      if (r != 0) { if (group) p_GroupEnd(); TNB_NCCL(r); }
    }
    if (group) TNB_NCCL(p_GroupEnd());
  }
  if (done) TNB_CUDA(cudaEventRecord((cudaEvent_t)done, ctx->comm_stream));
  return TNB_OK;
}

// One layer's data-parallel update, entirely on the communication stream so that it overlaps the backward GEMMs of the layers
// below: reduce-scatter the local gradient (every rank receives the sum of ITS block of weight rows), apply
// CuBiasedLinearity::Update to that block only, all-gather the updated rows.  Compared with all-reduce + full update on every
// rank the bytes on the wire are the same, the update's HBM traffic (5 passes over the weights) is divided by the world size and
// nothing is left to do after the last layer but to wait.  This is the reference CPU trainer's scheme — every thread updates
// its slice of the rows from the summed gradient (TNetLib/BiasedLinearity.cc:133-178) — with NCCL collectives.
int tnb_dp_update(TnbContext *ctx, float *G, float *W, float *corrW, TnbMatrixDim dW, int rows_pad, float *gb, float *bias,
                  float *corrb, float lr, float mmt, float wc, int gdf, int n_frames_global) {
  TNB_ARG(ctx && G && W && corrW, "null");
  TNB_ARG((gb && bias && corrb) || (!gb && !bias && !corrb), "bias arguments go together");
  TNB_ARG(dW.rows > 0 && dW.cols > 0 && dW.stride >= dW.cols && n_frames_global > 0, "dims");
  const int world = ctx->world;
  TNB_ARG(rows_pad >= dW.rows && rows_pad % world == 0, "rows_pad must be a multiple of the world size, at least dW.rows");
  float scale, l2;
  update_scalars(lr, mmt, wc, gdf, n_frames_global, &scale, &l2);
  cudaStream_t cs = world > 1 ? ctx->comm_stream : ctx->main_stream;
  if (world > 1) {
    TNB_ARG(ctx->nccl_comm != nullptr, "communicator not initialised");
    TNB_CUDA(cudaEventRecord(ctx->ev_compute, ctx->main_stream));  // the gradient GEMM (and everything before it) is the producer
    TNB_CUDA(cudaStreamWaitEvent(cs, ctx->ev_compute, 0));
  }
  const int shard = rows_pad / world;
  const size_t shard_elems = (size_t)shard * (size_t)dW.stride;
  const size_t off = (size_t)ctx->rank * shard_elems;
  if (world > 1) {
    TNB_NCCL(p_ReduceScatter(G, G + off, shard_elems, kNcclFloat, kNcclSum, (NcclComm)ctx->nccl_comm, cs));
    if (gb) TNB_NCCL(p_AllReduce(gb, gb, (size_t)dW.cols, kNcclFloat, kNcclSum, (NcclComm)ctx->nccl_comm, cs));
  }
  int my_rows = dW.rows - ctx->rank * shard;   // rows of the logical matrix inside this rank's block
  if (my_rows > shard) my_rows = shard;
  int rc = launch_sgd_update(ctx, cs, G + off, W + off, corrW + off, my_rows, dW.cols, dW.stride, mmt, scale, l2);
  if (rc != TNB_OK) return rc;
  if (gb) {  // the bias is tiny: every rank applies the same update to its own copy
    rc = launch_sgd_update(ctx, cs, gb, bias, corrb, 1, dW.cols, dW.cols, mmt, scale, 0.0f);
    if (rc != TNB_OK) return rc;
  }
  if (world > 1) TNB_NCCL(p_AllGather(W + off, W, shard_elems, kNcclFloat, (NcclComm)ctx->nccl_comm, cs));
  return TNB_OK;
}

int tnb_comm_wait(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  if (ctx->world == 1) return TNB_OK;
  TNB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  TNB_CUDA(cudaStreamWaitEvent(ctx->main_stream, ctx->ev_comm, 0));
  return TNB_OK;
}

}  // extern "C"
