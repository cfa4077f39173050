// ctx.cu — context, device selection, memory plumbing and the TMA descriptor cache.
// Replaces CuDevice (reference: src/CuBaseLib/cudevice.cc:22-121) and the cudaMallocPitch /
// cudaMemcpy2D / cudaMemset calls inside CuMatrix/CuVector (src/CuBaseLib/cumatrix.tcc:16-190).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace tnb {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_row_scratch(TnbContext *ctx, int rows) {
  if (rows <= ctx->row_cap) return TNB_OK;
  int cap = rows < 1024 ? 1024 : rows;
  if (ctx->row_scratch) cudaFree(ctx->row_scratch);
  if (ctx->row_match) cudaFree(ctx->row_match);
  ctx->row_scratch = nullptr; ctx->row_match = nullptr; ctx->row_cap = 0;
  TNB_CUDA(cudaMalloc(&ctx->row_scratch, sizeof(float) * (size_t)cap));
  TNB_CUDA(cudaMalloc(&ctx->row_match, sizeof(int) * (size_t)cap));
  ctx->row_cap = cap;
  return TNB_OK;
}

int ensure_vec_scratch_side(TnbContext *ctx, int n) {
  if (n <= ctx->vec_cap_side) return TNB_OK;
  int cap = n < 4096 ? 4096 : n;
  if (ctx->vec_scratch_side) { TNB_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->vec_scratch_side); }
  ctx->vec_scratch_side = nullptr; ctx->vec_cap_side = 0;
  TNB_CUDA(cudaMalloc(&ctx->vec_scratch_side, sizeof(float) * (size_t)cap));
  ctx->vec_cap_side = cap;
  return TNB_OK;
}

cudaStream_t stream_of(TnbContext *ctx, int id) {
  switch (id) {
    case TNB_STREAM_COMPUTE: return ctx->main_stream;
    case TNB_STREAM_COPY: return ctx->copy_stream;
    case TNB_STREAM_COMM: return ctx->comm_stream;
    case TNB_STREAM_AUX: return ctx->aux_stream;
    case TNB_STREAM_AUX2: return ctx->aux2_stream;
    default: return nullptr;
  }
}

int ensure_vec_scratch(TnbContext *ctx, int n) {
  if (n <= ctx->vec_cap) return TNB_OK;
  int cap = n < 4096 ? 4096 : n;
  if (ctx->vec_scratch) cudaFree(ctx->vec_scratch);
  ctx->vec_scratch = nullptr; ctx->vec_cap = 0;
  TNB_CUDA(cudaMalloc(&ctx->vec_scratch, sizeof(float) * (size_t)cap));
  ctx->vec_cap = cap;
  return TNB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: libtnetb200.so does not link
// libcuda, so it loads (and exports its symbols) on a machine without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return TNB_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  TNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return TNB_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  return TNB_OK;
}

// 2-D fp32 tensor map over a row-major [rows x cols] matrix with pitch `stride` elements;
// box = box_rows x box_cols (box_cols*4 == 128 B), OOB elements read as zero.  swizzle32 = 0: 128-byte swizzle with
// 16-byte atoms (K-major UMMA operands); 1: 128-byte swizzle with 32-byte atoms, the only layout tcgen05 accepts for
// MN-major 32-bit (tf32) operands (UMMA LayoutType SWIZZLE_128B_BASE32B).
// elem_bytes = 2: the same over a bf16 matrix (box_cols = 64, plain 128-byte swizzle for both operand majors).
int get_tmap(TnbContext *ctx, const void *ptr, int rows, int cols, int stride, int box_rows, int box_cols,
             int swizzle32, CUtensorMap *out, int elem_bytes) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rows = rows; key.cols = cols; key.stride = stride;
  key.box_rows = box_rows; key.box_cols = box_cols; key.swizzle32 = swizzle32 | (elem_bytes << 8);
  auto it = ctx->tmaps.find(key);
  if (it != ctx->tmaps.end()) { *out = it->second; return TNB_OK; }
  int rc = load_encode();
  if (rc != TNB_OK) return rc;
  TNB_ARG(((uintptr_t)ptr & 15) == 0, "TMA needs a 16-byte aligned base");
  TNB_ARG(elem_bytes == 4 || elem_bytes == 2, "element size");
  TNB_ARG(((size_t)stride * elem_bytes) % 16 == 0, "TMA needs a row pitch that is a multiple of 16 bytes");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)stride * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d stride=%d box=%dx%d", (int)r, rows, cols,
              stride, box_rows, box_cols);
    return TNB_ERR_CUDA;
  }
  if (ctx->tmaps.size() > 8192) ctx->tmaps.clear();
  ctx->tmaps[key] = m;
  *out = m;
  return TNB_OK;
}

// bf16 copy of an fp32 GEMM operand in ctx-owned scratch (generic entry points in TNB_MATH_BF16).  Stream order makes the
// reuse safe: the conversion kernel is an ordinary launch, so it starts only after the GEMM that read the slot has completed.
int bf16_scratch(TnbContext *ctx, int slot, const float *src, int rows, int cols, int stride, uint16_t **out, int *out_stride) {
  const int st = ((cols + 63) / 64) * 64;
  const size_t need = (size_t)rows * (size_t)st;
  if (need > ctx->bf16_cap[slot]) {
    if (ctx->bf16_scratch[slot]) { TNB_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->bf16_scratch[slot]); }
    ctx->bf16_scratch[slot] = nullptr; ctx->bf16_cap[slot] = 0;
    TNB_CUDA(cudaMalloc(&ctx->bf16_scratch[slot], need * 2));
    ctx->bf16_cap[slot] = need;
  }
  *out = ctx->bf16_scratch[slot];
  *out_stride = st;
  return launch_to_bf16(ctx, *out, st, src, rows, cols, stride);
}

int xchg_buffer(TnbContext *ctx, cudaStream_t stream, size_t bytes, float **out) {
  *out = nullptr;
  static int use_dsmem = -1;
  if (use_dsmem < 0) { const char *e = getenv("TNB_GEMM_XCHG"); use_dsmem = (e && !strcmp(e, "dsmem")) ? 1 : 0; }
  if (use_dsmem) return TNB_OK;
  TnbContext::Xchg *x = nullptr;
  for (auto &e : ctx->xchg) if (e.stream == stream) x = &e;
  if (!x) { ctx->xchg.push_back(TnbContext::Xchg{stream, nullptr, 0}); x = &ctx->xchg.back(); }
  if (bytes > x->cap) {
    if (ctx->capturing) return TNB_OK;  // no allocation inside a stream capture: this launch exchanges through DSMEM
    if (x->ptr) { TNB_CUDA(cudaStreamSynchronize(stream)); cudaFree(x->ptr); x->ptr = nullptr; x->cap = 0; }
    const size_t want = bytes < ((size_t)10 << 20) ? ((size_t)10 << 20) : bytes;  // 148 CTAs x 64 KB fit the first allocation
    TNB_CUDA(cudaMalloc(&x->ptr, want));
    x->cap = want;
  }
  *out = x->ptr;
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

extern "C" {

const char *tnb_version(void) { return "tnet-b200 0.1 (sm_100a)"; }
const char *tnb_last_error(void) { return g_err; }

int tnb_device_count(int *count) {
  TNB_ARG(count != nullptr, "null");
  *count = 0;
  TNB_CUDA(cudaGetDeviceCount(count));
  return TNB_OK;
}

int tnb_ctx_create(TnbContext **out, int device) {
  TNB_ARG(out != nullptr, "null");
  *out = nullptr;
  int n = 0;
  TNB_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) { set_error("no CUDA device: libtnetb200 has no CPU fallback"); return TNB_ERR_CUDA; }
  if (device < 0) {
    // reference: cudevice.cc:27-56 — choose the GPU with the largest free/total memory ratio
    double best = -1.0;
    for (int d = 0; d < n; d++) {
      size_t fr = 0, tot = 0;
      if (cudaSetDevice(d) != cudaSuccess) continue;
      if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); continue; }
      double ratio = tot ? (double)fr / (double)tot : 0.0;
      if (ratio > best) { best = ratio; device = d; }
    }
    if (device < 0) device = 0;
  }
  TNB_ARG(device < n, "device index out of range");
  TNB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TNB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libtnetb200 is built for sm_100a only", device, prop.major, prop.minor);
    return TNB_ERR_UNSUPPORTED;
  }
  TnbContext *ctx = new TnbContext_();
  { const char *e = getenv("TNB_PDL"); if (e && atoi(e) == 0) ctx->pdl = false; }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  TNB_CUDA(cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->main_stream;
  {
    // TNB_COMM_PRIORITY=1: collectives on the highest-priority stream (their few CTAs are placed before a waiting GEMM wave)
    const char *e = getenv("TNB_COMM_PRIORITY");
    int lo = 0, hi = 0;
    TNB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TNB_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, (e && atoi(e) != 0) ? hi : 0));
  }
  TNB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  TNB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
  TNB_CUDA(cudaStreamCreateWithFlags(&ctx->aux2_stream, cudaStreamNonBlocking));
  TNB_CUDA(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
  TNB_CUDA(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
  *out = ctx;
  return TNB_OK;
}

int tnb_ctx_destroy(TnbContext *ctx) {
  if (!ctx) return TNB_OK;
  cudaSetDevice(ctx->device);
  tnb_comm_destroy(ctx);
  cudaStreamSynchronize(ctx->main_stream);
  if (ctx->row_scratch) cudaFree(ctx->row_scratch);
  if (ctx->row_match) cudaFree(ctx->row_match);
  if (ctx->vec_scratch) cudaFree(ctx->vec_scratch);
  for (int i = 0; i < 2; i++) if (ctx->bf16_scratch[i]) cudaFree(ctx->bf16_scratch[i]);
  for (auto &e : ctx->xchg) if (e.ptr) cudaFree(e.ptr);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev_compute);
  cudaEventDestroy(ctx->ev_comm);
  cudaStreamDestroy(ctx->main_stream);
  cudaStreamDestroy(ctx->comm_stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->aux_stream);
  cudaStreamDestroy(ctx->aux2_stream);
  if (ctx->vec_scratch_side) cudaFree(ctx->vec_scratch_side);
  for (auto &kv : ctx->mg_plans) if (kv.second.dlist) cudaFree(kv.second.dlist);
  if (ctx->mg_trace) cudaFree(ctx->mg_trace);
  if (ctx->peer_trace) cudaFree(ctx->peer_trace);
  if (ctx->done_stream) { cudaStreamDestroy(ctx->done_stream); cudaEventDestroy(ctx->ev_done_fork); }
  if (ctx->ev_push_fork) cudaEventDestroy(ctx->ev_push_fork);
  for (int i = 0; i < TNB_MAX_PEERS; i++) {
    if (ctx->push_streams[i]) cudaStreamDestroy(ctx->push_streams[i]);
    if (ctx->push_events[i]) cudaEventDestroy(ctx->push_events[i]);
  }
  delete ctx;
  return TNB_OK;
}

int tnb_ctx_device(TnbContext *ctx, int *device) { TNB_ARG(ctx && device, "null"); *device = ctx->device; return TNB_OK; }
int tnb_ctx_set_math(TnbContext *ctx, int m) {
  TNB_ARG(ctx, "null");
  TNB_ARG(m == TNB_MATH_3XTF32 || m == TNB_MATH_TF32 || m == TNB_MATH_FP32_SIMT || m == TNB_MATH_BF16, "unknown math mode");
  ctx->math_mode = m;
  return TNB_OK;
}
int tnb_ctx_get_math(TnbContext *ctx, int *m) { TNB_ARG(ctx && m, "null"); *m = ctx->math_mode; return TNB_OK; }
int tnb_ctx_stream(TnbContext *ctx, void **s) { TNB_ARG(ctx && s, "null"); *s = (void *)ctx->main_stream; return TNB_OK; }
int tnb_ctx_use_stream(TnbContext *ctx, int stream_id) {
  TNB_ARG(ctx && !ctx->capturing, "null / capturing");
  cudaStream_t st = stream_of(ctx, stream_id);
  TNB_ARG(st != nullptr && stream_id != TNB_STREAM_COMM, "stream");
  ctx->stream = st;
  return TNB_OK;
}
int tnb_ctx_sync(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  TNB_CUDA(cudaStreamSynchronize(ctx->main_stream));
  TNB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  TNB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  TNB_CUDA(cudaStreamSynchronize(ctx->aux_stream));
  TNB_CUDA(cudaStreamSynchronize(ctx->aux2_stream));
  if (ctx->done_stream) TNB_CUDA(cudaStreamSynchronize(ctx->done_stream));
  for (int i = 0; i < TNB_MAX_PEERS; i++)
    if (ctx->push_streams[i]) TNB_CUDA(cudaStreamSynchronize(ctx->push_streams[i]));
  return ctx->peer_flags[ctx->rank] ? tnb_peer_status(ctx) : TNB_OK;  // a peer-memory kernel that gave up waiting says so here
}
int tnb_ctx_free_memory(TnbContext *ctx, size_t *fr, size_t *tot) {
  TNB_ARG(ctx && fr && tot, "null");
  TNB_CUDA(cudaSetDevice(ctx->device));
  TNB_CUDA(cudaMemGetInfo(fr, tot));
  return TNB_OK;
}
int tnb_ctx_launch_count(TnbContext *ctx, unsigned long long *n) { TNB_ARG(ctx && n, "null"); *n = ctx->launches; return TNB_OK; }

int tnb_ctx_profile_begin(TnbContext *ctx) {
  TNB_ARG(ctx, "null");
  ctx->profiling = true;
  ctx->prof_used = 0;
  ctx->prof_flops = 0.0;
  return TNB_OK;
}
int tnb_ctx_profile_end(TnbContext *ctx, double *gemm_ms, unsigned long long *gemm_launches, double *gemm_flops) {
  TNB_ARG(ctx && gemm_ms && gemm_launches && gemm_flops, "null");
  ctx->profiling = false;
  TNB_CUDA(cudaStreamSynchronize(ctx->stream));
  double ms = 0.0;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    float t = 0.0f;
    TNB_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[i], ctx->prof_events[i + 1]));
    ms += t;
  }
  *gemm_ms = ms;
  *gemm_launches = ctx->prof_used / 2;
  *gemm_flops = ctx->prof_flops;
  ctx->prof_used = 0;
  return TNB_OK;
}

int tnb_malloc_pitch(TnbContext *ctx, void **ptr, int *stride_elems, int rows, int cols) {
  TNB_ARG(ctx && ptr && stride_elems, "null");
  TNB_ARG(rows >= 0 && cols >= 0, "negative dims");
  int stride = ((cols + 31) / 32) * 32;
  if (stride == 0) stride = 32;
  size_t bytes = (size_t)(rows > 0 ? rows : 1) * (size_t)stride * 4;
  TNB_CUDA(cudaSetDevice(ctx->device));
  TNB_CUDA(cudaMalloc(ptr, bytes));
  TNB_CUDA(cudaMemsetAsync(*ptr, 0, bytes, ctx->stream));
  *stride_elems = stride;
  return TNB_OK;
}
int tnb_malloc_pitch16(TnbContext *ctx, void **ptr, int *stride_elems, int rows, int cols) {
  TNB_ARG(ctx && ptr && stride_elems, "null");
  TNB_ARG(rows >= 0 && cols >= 0, "negative dims");
  int stride = ((cols + 63) / 64) * 64;  // 128-byte pitch
  if (stride == 0) stride = 64;
  size_t bytes = (size_t)(rows > 0 ? rows : 1) * (size_t)stride * 2;
  TNB_CUDA(cudaSetDevice(ctx->device));
  TNB_CUDA(cudaMalloc(ptr, bytes));
  TNB_CUDA(cudaMemsetAsync(*ptr, 0, bytes, ctx->stream));
  *stride_elems = stride;
  return TNB_OK;
}
int tnb_malloc(TnbContext *ctx, void **ptr, size_t bytes) {
  TNB_ARG(ctx && ptr, "null");
  TNB_CUDA(cudaSetDevice(ctx->device));
  TNB_CUDA(cudaMalloc(ptr, bytes ? bytes : 4));
  TNB_CUDA(cudaMemsetAsync(*ptr, 0, bytes ? bytes : 4, ctx->stream));
  return TNB_OK;
}
int tnb_free(TnbContext *ctx, void *ptr) {
  TNB_ARG(ctx, "null");
  if (!ptr) return TNB_OK;
  // drop cached TMA descriptors that point into this allocation
  for (auto it = ctx->tmaps.begin(); it != ctx->tmaps.end();) {
    if (it->first.ptr == ptr) it = ctx->tmaps.erase(it); else ++it;
  }
  TNB_CUDA(cudaStreamSynchronize(ctx->stream));
  TNB_CUDA(cudaFree(ptr));
  return TNB_OK;
}
int tnb_memset(TnbContext *ctx, void *ptr, int value, size_t bytes) {
  TNB_ARG(ctx && ptr, "null");
  TNB_CUDA(cudaMemsetAsync(ptr, value, bytes, ctx->stream));
  return TNB_OK;
}
static cudaMemcpyKind kind_of(int k) {
  return k == 0 ? cudaMemcpyHostToDevice : (k == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
}
int tnb_memcpy2d(TnbContext *ctx, void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t h, int kind) {
  TNB_ARG(ctx && dst && src, "null");
  TNB_ARG(kind >= 0 && kind <= 2, "kind");
  if (w == 0 || h == 0) return TNB_OK;
  TNB_CUDA(cudaMemcpy2DAsync(dst, dp, src, sp, w, h, kind_of(kind), ctx->stream));
  if (kind == 1) TNB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TNB_OK;
}
int tnb_memcpy(TnbContext *ctx, void *dst, const void *src, size_t bytes, int kind) {
  TNB_ARG(ctx && dst && src, "null");
  TNB_ARG(kind >= 0 && kind <= 2, "kind");
  if (bytes == 0) return TNB_OK;
  TNB_CUDA(cudaMemcpyAsync(dst, src, bytes, kind_of(kind), ctx->stream));
  if (kind == 1) TNB_CUDA(cudaStreamSynchronize(ctx->stream));
  return TNB_OK;
}
// ---- CUDA graphs: a launch-bound sequence of entry points (TRecurrentCu issues ~115 small kernels per frame) recorded once and
// replayed.  Between begin and end every tnb_* call on the compute stream is captured instead of executed; calls that allocate,
// synchronise or copy to/from pageable host memory must not be made there.  PDL attributes are left out of captured launches.
struct TnbGraph { cudaGraphExec_t exec; unsigned long long kernels; };
int tnb_graph_begin(TnbContext *ctx) {
  TNB_ARG(ctx && !ctx->capturing, "null / already capturing");
  TNB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  ctx->capture_base = ctx->launches;
  return TNB_OK;
}
int tnb_graph_end(TnbContext *ctx, void **graph) {
  TNB_ARG(ctx && graph && ctx->capturing, "null / not capturing");
  ctx->capturing = false;
  cudaGraph_t g = nullptr;
  TNB_CUDA(cudaStreamEndCapture(ctx->stream, &g));
  TnbGraph *t = new TnbGraph();
  t->kernels = ctx->launches - ctx->capture_base;
  ctx->launches = ctx->capture_base;  // nothing has run yet
  cudaError_t e = cudaGraphInstantiate(&t->exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { delete t; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return TNB_ERR_CUDA; }
  *graph = t;
  return TNB_OK;
}
int tnb_graph_launch(TnbContext *ctx, void *graph) {
  TNB_ARG(ctx && graph && !ctx->capturing, "null / capturing");
  TnbGraph *t = (TnbGraph *)graph;
  TNB_CUDA(cudaGraphLaunch(t->exec, ctx->stream));
  ctx->launches += t->kernels;
  return TNB_OK;
}
int tnb_graph_destroy(TnbContext *ctx, void *graph) {
  TNB_ARG(ctx, "null");
  if (graph) { TnbGraph *t = (TnbGraph *)graph; cudaGraphExecDestroy(t->exec); delete t; }
  return TNB_OK;
}

// ---- streams and events: what a host needs to overlap its transfers with the training step ----
int tnb_memcpy2d_on(TnbContext *ctx, int stream_id, void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t h, int kind) {
  TNB_ARG(ctx && dst && src, "null");
  TNB_ARG(kind >= 0 && kind <= 2 && stream_of(ctx, stream_id) != nullptr, "kind / stream");
  if (w == 0 || h == 0) return TNB_OK;
  TNB_CUDA(cudaMemcpy2DAsync(dst, dp, src, sp, w, h, kind_of(kind), stream_of(ctx, stream_id)));
  return TNB_OK;
}
int tnb_memcpy_on(TnbContext *ctx, int stream_id, void *dst, const void *src, size_t bytes, int kind) {
  TNB_ARG(ctx && dst && src, "null");
  TNB_ARG(kind >= 0 && kind <= 2 && stream_of(ctx, stream_id) != nullptr, "kind / stream");
  if (bytes == 0) return TNB_OK;
  TNB_CUDA(cudaMemcpyAsync(dst, src, bytes, kind_of(kind), stream_of(ctx, stream_id)));
  return TNB_OK;
}
int tnb_event_create(TnbContext *ctx, void **ev) {
  TNB_ARG(ctx && ev, "null");
  cudaEvent_t e;
  TNB_CUDA(cudaSetDevice(ctx->device));
  TNB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = (void *)e;
  return TNB_OK;
}
int tnb_event_destroy(TnbContext *ctx, void *ev) {
  TNB_ARG(ctx, "null");
  if (ev) TNB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return TNB_OK;
}
int tnb_event_record(TnbContext *ctx, void *ev, int stream_id) {
  TNB_ARG(ctx && ev && stream_of(ctx, stream_id) != nullptr, "null / stream");
  TNB_CUDA(cudaEventRecord((cudaEvent_t)ev, stream_of(ctx, stream_id)));
  return TNB_OK;
}
int tnb_stream_wait_event(TnbContext *ctx, int stream_id, void *ev) {
  TNB_ARG(ctx && ev && stream_of(ctx, stream_id) != nullptr, "null / stream");
  TNB_CUDA(cudaStreamWaitEvent(stream_of(ctx, stream_id), (cudaEvent_t)ev, 0));
  return TNB_OK;
}
int tnb_event_sync(TnbContext *ctx, void *ev) {
  TNB_ARG(ctx && ev, "null");
  TNB_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return TNB_OK;
}

int tnb_host_alloc(void **ptr, size_t bytes) {
  TNB_ARG(ptr, "null");
  TNB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 4));
  return TNB_OK;
}
int tnb_host_free(void *ptr) {
  if (ptr) TNB_CUDA(cudaFreeHost(ptr));
  return TNB_OK;
}

}  // extern "C"
