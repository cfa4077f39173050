// common.cuh — internal definitions shared by the .cu files of libtnetb200.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "tnet_b200.h"

namespace tnb {

void set_error(const char *fmt, ...);

#define TNB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      tnb::set_error("CUDA error %d (%s) at %s:%d: %s", (int)e__, cudaGetErrorString(e__),     \
                     __FILE__, __LINE__, #call);                                               \
      return TNB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define TNB_ARG(cond, msg)                                                                     \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      tnb::set_error("bad argument at %s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);        \
      return TNB_ERR_ARG;                                                                      \
    }                                                                                          \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define TNB_LAUNCHED(ctx)                                                                      \
  do {                                                                                         \
    (ctx)->launches++;                                                                         \
    cudaError_t e__ = cudaPeekAtLastError();                                                   \
    if (e__ != cudaSuccess) {                                                                  \
      cudaGetLastError();                                                                      \
      tnb::set_error("kernel launch failed %d (%s) at %s:%d", (int)e__,                        \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                             \
      return TNB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// tnb_gemm_batch (gemm_multi.cu): the cached plan of one launch
struct MgPlan {
  std::vector<int> src;          // GEMM slot -> index into must[] (i < n_must) or n_must + index into pool[]
  std::vector<int> pool_taken;   // tiles taken from each pool job
  void *dlist = nullptr;         // device copy of the per-pair tile lists [used_pairs][ipp] (int4)
  int used_pairs = 0, ipp = 0;
  long makespan = 0;
  double flops = 0.0;
};

struct TmapKey {
  const void *ptr;
  int rows, cols, stride, box_rows, box_cols, swizzle32;
  bool operator<(const TmapKey &o) const { return memcmp(this, &o, sizeof(TmapKey)) < 0; }
};

}  // namespace tnb

struct TnbContext_ {
  int device = 0;
  int sm_count = 148;
  int math_mode = TNB_MATH_3XTF32;
  cudaStream_t stream = nullptr;       // the stream entry points without a stream argument enqueue on (tnb_ctx_use_stream; default: main_stream)
  cudaStream_t main_stream = nullptr;  // TNB_STREAM_COMPUTE
  cudaStream_t comm_stream = nullptr;  // NCCL stream
  cudaStream_t copy_stream = nullptr;  // host<->device transfers that overlap compute (TNB_STREAM_COPY)
  cudaStream_t aux_stream = nullptr;   // small kernels next to the compute stream's GEMMs (TNB_STREAM_AUX)
  cudaStream_t aux2_stream = nullptr;  // (TNB_STREAM_AUX2)
  cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
  unsigned long long launches = 0;
  bool capturing = false;                   // between tnb_graph_begin and tnb_graph_end
  unsigned long long capture_base = 0;      // launch counter at tnb_graph_begin
  std::map<tnb::TmapKey, CUtensorMap> tmaps;  // TMA descriptors keyed by (ptr, dims, box)
  std::map<std::vector<int>, tnb::MgPlan> mg_plans;  // tnb_gemm_batch: launch plans keyed by the jobs' shapes and tile ranges
  int mg_pairs_limit = 0;                       // tnb_gemm_batch_set_pairs: CTA pairs a batch launch may use (0 = all that are co-resident)
  void *mg_trace = nullptr;                     // TNB_BATCH_TRACE=1: per-pair clock stamps of the last tnb_gemm_batch launch
  // scratch for deterministic per-row -> stats reductions
  float *row_scratch = nullptr;
  int *row_match = nullptr;
  int row_cap = 0;
  // column-sum scratch (bias gradient before the fused update)
  float *vec_scratch = nullptr;
  int vec_cap = 0;
  float *vec_scratch_side = nullptr;   // the same for column sums enqueued on another stream (they may run concurrently)
  int vec_cap_side = 0;
  // bf16 copies of fp32 GEMM operands for the generic entry points in TNB_MATH_BF16 (slot 0 = A, 1 = B)
  uint16_t *bf16_scratch[2] = {nullptr, nullptr};
  size_t bf16_cap[2] = {0, 0};
  // split-K GEMMs: accumulator exchange buffers, one per stream that has launched one (gemm_kernel.cuh)
  struct Xchg { cudaStream_t stream; float *ptr; size_t cap; };
  std::vector<Xchg> xchg;
  // GEMM profiling (tnb_ctx_profile_begin/end)
  bool profiling = false;
  bool pdl = true;  // programmatic dependent launch between consecutive GEMMs (TNB_PDL=0 disables)
  std::vector<cudaEvent_t> prof_events;  // start/stop pairs
  size_t prof_used = 0;
  double prof_flops = 0.0;
  // data-parallel
  void *nccl_comm = nullptr;
  void *local_group = nullptr;   // tnb_comm_init_local: the ranks are threads of this process (TnbLocalGroup)
  int rank = 0, world = 1;
  // peer-memory schedule (peer.cu): every rank's flag block as mapped into this process, and the launch counter
  unsigned *peer_flags[TNB_MAX_PEERS] = {};
  unsigned peer_seq = 0;
  unsigned push_seq = 0;   // tnb_peer_push_blocks calls (trace slots)
  cudaStream_t push_streams[TNB_MAX_PEERS] = {};  // tnb_peer_push_blocks: one stream (one copy engine at a time) per destination rank
  cudaEvent_t push_events[TNB_MAX_PEERS] = {};
  cudaEvent_t ev_push_fork = nullptr;
  cudaStream_t done_stream = nullptr;   // tnb_dp_peer_update_after: the kernels that wait for the other ranks' done flags
  cudaEvent_t ev_done_fork = nullptr;
  void *peer_trace = nullptr;   // TNB_DP_TRACE=1: %globaltimer stamps of the last 64 peer-memory kernels (tnb_peer_trace_read)
};

namespace tnb {
inline dim3 grid2d(int cols, int rows, int bx, int by) {
  return dim3((cols + bx - 1) / bx, (rows + by - 1) / by);
}
// the reference evaluates the update scalars in float (cuBiasedLinearity.cc:44-63): W += scale*corr ; W += l2*W
void update_scalars(float lr, float mmt, float wc, int gdf, int rows, float *scale, float *l2);
// corr = G + mmt*corr ; W += scale*corr ; W += l2*W over a [rows x cols] block, on the given stream
int launch_sgd_update(TnbContext *ctx, cudaStream_t stream, const float *G, float *W, float *corr, int rows, int cols, int stride,
                      float mmt, float scale, float l2);
int ensure_row_scratch(TnbContext *ctx, int rows);
int ensure_vec_scratch(TnbContext *ctx, int n);
int ensure_vec_scratch_side(TnbContext *ctx, int n);
// gathers n bytes per rank through the NCCL communicator: all[rank*n .. +n) in, every rank's bytes out (host memory; synchronous)
int comm_allgather_bytes(TnbContext *ctx, unsigned char *all, size_t n);
cudaStream_t stream_of(TnbContext *ctx, int stream_id);  // TNB_STREAM_* -> cudaStream_t (nullptr for an unknown id)
int get_tmap(TnbContext *ctx, const void *ptr, int rows, int cols, int stride, int box_rows,
             int box_cols, int swizzle32, CUtensorMap *out, int elem_bytes = 4);
}  // namespace tnb
