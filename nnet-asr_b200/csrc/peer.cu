// peer.cu — one layer's data-parallel update as ONE kernel over NVLink peer memory.
//
// The NCCL schedules of comm.cu spend a layer's exchange on three steps that each sweep the weight-sized arrays: the collective
// (all-reduce, or reduce-scatter + all-gather), then the update kernel (read G, corr, W; write corr, W).  Here every rank maps
// the other ranks' gradient and weight buffers into its address space (CUDA IPC, one process per GPU on one NVSwitch box) and a
// single kernel per layer does, for the block of weight rows this rank owns:
//     g = sum over ranks of G_r[block]            (peer loads over NVLink, fixed rank order -> deterministic)
//     corr = g + mmt*corr ; w = W + scale*corr ; w += l2*w        (CuBiasedLinearity::Update, cuBiasedLinearity.cc:55-63)
//     W_r[block] = w for every rank r                              (peer stores)
// i.e. reduce-scatter, the reference CPU trainer's "every worker updates its slice" (TNetLib/BiasedLinearity.cc:133-178) and
// all-gather fused; the summed gradient never exists in memory and the update's HBM traffic is divided by the world size.
// The bias (one row) is summed and updated redundantly by every rank.
//
// Ordering between the ranks uses two flag words per peer in a small flag block that every rank maps from every other rank:
//   ready[r] = s : rank r has entered its kernel number s — its gradient is complete (stream order) and it no longer reads the
//                  weights of this layer (its dX GEMM precedes its gradient GEMM);
//   done[r]  = s : rank r has finished its stores for kernel s (fence.sys before the flag).
// A kernel starts its loads when all ready flags have reached s and ends when all done flags have: after it, this rank's
// weights are complete and nobody reads its gradient buffer any more.  All ranks issue the same kernels in the same order on
// their communication stream (like a collective).  Every wait is bounded: a rank that waits longer than TNB_PEER_TIMEOUT_MS
// (default 10 s; 0 = no limit) writes {phase, peer it was waiting for, sequence number} into word 33 of its own flag block and
// leaves the kernel WITHOUT touching the weights; the CUDA context stays usable and tnb_peer_status / tnb_ctx_sync report which
// rank stalled (a __trap() here would poison every rank's context for a rank that was merely late, e.g. in file I/O).
//
// On ONE GPU several ranks must not be played as separate launches that wait for each other (nothing guarantees that they are
// co-resident): dp_peer_update_virtual_kernel runs all ranks' work as ONE cooperative grid (tnb_dp_peer_update_virtual), which is
// what the tests and the ncu capture of this kernel use.
#include <stdlib.h>

#include "common.cuh"

namespace tnb {

struct PeerArgs {
  int rank, world;
  const float *G[TNB_MAX_PEERS];
  float *W[TNB_MAX_PEERS];
  unsigned *flags[TNB_MAX_PEERS];  // per rank: [0,16) ready, [16,32) done, [32] arrival counter of the owner's own CTAs, [33] error word
  float *corr, *bias, *corrb;      // local
  int cols, stride, shard, rows_pad;
  float mmt, scale, l2;
  int pushed;        // the gradients were pushed into the owners' staging slices (tnb_affine_grad_scatter): G[rank] holds `world` local slices
  unsigned seq;
  long long timeout;  // cycles; 0 = wait for ever
  int defer_done;     // leave once this rank's done flag is published: a dp_peer_wait_done_kernel on another stream waits for the other ranks'
  long long *trace;   // TNB_DP_TRACE=1: [64 launches][4] %globaltimer stamps of CTA 0 (entry, all ranks ready, own rows done) and of the last CTA (all ranks done)
};

__device__ __forceinline__ void peer_stamp(const PeerArgs &a, int slot) {
  if (a.trace && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(a.seq & 63u) * 4 + slot] = t;
  }
}

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// threads [0, world) of the CTA each poll one flag until it has reached seq (wrap-safe compare).  Returns false for the whole CTA
// when a wait ran out of time; the error word of this rank's flag block then says what it was waiting for.
__device__ __forceinline__ bool wait_flags(unsigned *my, int off, int world, unsigned seq, long long timeout) {
  int bad = 0;
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(my + off + threadIdx.x) - seq) < 0) {
      if (timeout > 0 && clock64() - t0 > timeout) {
        // 0x80000000 | phase (0 = ready, 1 = done) << 24 | peer << 16 | low 16 bits of the sequence number
        atomicCAS(my + 33, 0u, 0x80000000u | ((unsigned)(off ? 1 : 0) << 24) | ((unsigned)threadIdx.x << 16) | (seq & 0xFFFFu));
        bad = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  return __syncthreads_or(bad) == 0;
}

__device__ __forceinline__ float4 ldg_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// WORLD > 0: compile-time rank count (all loads of U items issued before the first use); WORLD == 0: any rank count, one item at a time
// bid / nblk: this CTA's index among the CTAs working for rank a.rank and their number (the grid of dp_peer_update_kernel; one
// slice of the grid of dp_peer_update_virtual_kernel)
template <int WORLD, int U>
__device__ __forceinline__ void peer_update_body(const PeerArgs &a, const int bid, const int nblk) {
  const int world = WORLD > 0 ? WORLD : a.world;
  unsigned *my = a.flags[a.rank];
  if (bid == 0) peer_stamp(a, 0);
  if (bid == 0 && (int)threadIdx.x < world) st_release_sys(a.flags[threadIdx.x] + a.rank, a.seq);
  if (!wait_flags(my, 0, world, a.seq, a.timeout)) return;
  if (bid == 0) peer_stamp(a, 1);

  const int vcols = (a.cols + 3) >> 2;
  const long total = (long)a.shard * vcols;
  const size_t row0 = (size_t)a.rank * a.shard;
  const long step = (long)nblk * blockDim.x;
  if (WORLD > 0 && (a.cols & 3) == 0) {
    for (long i0 = (long)bid * blockDim.x + threadIdx.x; i0 < total; i0 += step * U) {
      float4 g[U][WORLD > 0 ? WORLD : 1], k[U], w[U];
      size_t base[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long i = i0 + u * step;
        const long ii = i < total ? i : i0;  // clamp: the duplicate's result is not stored
        base[u] = (row0 + (size_t)(ii / vcols)) * a.stride + ((size_t)(ii % vcols) << 2);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
#pragma unroll
        for (int r = 0; r < WORLD; r++)  // pushed: slice r of this rank's own staging buffer (local HBM); else rank r's buffer (NVLink)
          g[u][r] = a.pushed ? ldg_f4(a.G[a.rank] + (base[u] - row0 * a.stride) + (size_t)r * a.shard * a.stride) : ldg_f4(a.G[r] + base[u]);
        k[u] = ldg_f4(a.corr + base[u]);
        w[u] = ldg_f4(a.W[a.rank] + base[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (i0 + u * step >= total) continue;
        float4 s = g[u][0];
#pragma unroll
        for (int r = 1; r < WORLD; r++) { s.x += g[u][r].x; s.y += g[u][r].y; s.z += g[u][r].z; s.w += g[u][r].w; }
        float4 kk = k[u], ww = w[u];
        kk.x = s.x + a.mmt * kk.x; kk.y = s.y + a.mmt * kk.y; kk.z = s.z + a.mmt * kk.z; kk.w = s.w + a.mmt * kk.w;
        ww.x = a.scale * kk.x + ww.x; ww.y = a.scale * kk.y + ww.y; ww.z = a.scale * kk.z + ww.z; ww.w = a.scale * kk.w + ww.w;
        if (a.l2 != 0.0f) { ww.x = a.l2 * ww.x + ww.x; ww.y = a.l2 * ww.y + ww.y; ww.z = a.l2 * ww.z + ww.z; ww.w = a.l2 * ww.w + ww.w; }
        *reinterpret_cast<float4 *>(a.corr + base[u]) = kk;
#pragma unroll
        for (int r = 0; r < WORLD; r++) *reinterpret_cast<float4 *>(a.W[r] + base[u]) = ww;
      }
    }
  } else {  // any rank count / column count: scalar
    const long total_s = (long)a.shard * a.cols;
    for (long i = (long)bid * blockDim.x + threadIdx.x; i < total_s; i += step) {
      const size_t base = (row0 + (size_t)(i / a.cols)) * a.stride + (size_t)(i % a.cols);
      const size_t lbase = base - row0 * a.stride;  // position inside a staging slice
      float s = a.pushed ? a.G[a.rank][lbase] : a.G[0][base];
      for (int r = 1; r < world; r++) s += a.pushed ? a.G[a.rank][lbase + (size_t)r * a.shard * a.stride] : a.G[r][base];
      const float kk = s + a.mmt * a.corr[base];
      float ww = a.scale * kk + a.W[a.rank][base];
      if (a.l2 != 0.0f) ww = a.l2 * ww + ww;
      a.corr[base] = kk;
      for (int r = 0; r < world; r++) a.W[r][base] = ww;
    }
  }
  // bias: every rank sums all ranks' bias gradients (row rows_pad of the gradient buffers) in the same order and updates its own copy
  if (a.bias) {
    const size_t gb = (size_t)a.rows_pad * a.stride;
    for (int c = bid * blockDim.x + threadIdx.x; c < a.cols; c += (int)step) {
      float s = a.G[0][gb + c];
      for (int r = 1; r < world; r++) s += a.G[r][gb + c];
      const float kk = s + a.mmt * a.corrb[c];
      a.corrb[c] = kk;
      a.bias[c] = a.scale * kk + a.bias[c];
    }
  }

  if (bid == 0) peer_stamp(a, 2);
  // this rank is done when ALL its CTAs are: the last one to arrive publishes the flag and waits for the other ranks
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = (atomicAdd(my + 32, 1u) == (unsigned)nblk - 1) ? 1 : 0;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my[32] = 0;  // for the next launch (which starts after this kernel has ended)
  __threadfence_system();
  if ((int)threadIdx.x < world) st_release_sys(a.flags[threadIdx.x] + 16 + a.rank, a.seq);
  if (!a.defer_done) wait_flags(my, 16, world, a.seq, a.timeout);
  peer_stamp(a, 3);
}

// The second half of a kernel launched with defer_done: one warp that waits until every rank's done flag has reached `seq` (all
// blocks of this rank's weights have been written, nobody reads its gradient buffer any more).  It runs on its own stream, so the
// communication stream goes on with the next layer's kernel instead of idling through a cross-GPU round trip per layer.
__global__ void __launch_bounds__(32) dp_peer_wait_done_kernel(unsigned *my, int world, unsigned seq, long long timeout) {
  wait_flags(my, 16, world, seq, timeout);
}

// TNB_DP_TRACE=1: a one-thread kernel that leaves %globaltimer in *p (stream-ordered marker around the copy engines' pushes)
__global__ void peer_stamp_kernel(long long *p) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *p = t;
}

template <int WORLD, int U>
__global__ void __launch_bounds__(512) dp_peer_update_kernel(const __grid_constant__ PeerArgs a) {
  peer_update_body<WORLD, U>(a, (int)blockIdx.x, (int)gridDim.x);
}

// "Spread" shape (TNB_DP_PEER_SPREAD=1; measured equal or slightly slower than the shape above at 4 GPUs): one SMALL CTA per SM — 128 threads, at most 64 registers, no shared memory — so that a
// CTA of this kernel fits NEXT TO a GEMM CTA on the same SM (the GEMMs take 320 threads x <= 168 registers, ~200 KB of shared memory and
// the SM's TMEM; they leave ~10 K registers and all the issue slots their waiting warps do not use).  The 20 x 512-thread shape above
// needs SMs of its own: a GEMM grid that arrives while it runs finds 128 SMs instead of 148, which a 4-CTA-cluster launch does not
// always fit in one wave (measured: the forward GEMMs next to the deferred exchanges ran 20 % slower, profiles/r02_dp_timeline.md).
template <int WORLD>
__global__ void __launch_bounds__(128, 8) dp_peer_update_spread_kernel(const __grid_constant__ PeerArgs a) {
  peer_update_body<WORLD, 1>(a, (int)blockIdx.x, (int)gridDim.x);
}

// All ranks of a layer's exchange as ONE grid on one GPU (tests, ncu): CTAs [v*ctas, (v+1)*ctas) play rank v with args[v].  Launched
// cooperatively, so that every CTA is resident while the slices wait for each other's flags.
template <int WORLD, int U>
__global__ void __launch_bounds__(512) dp_peer_update_virtual_kernel(const PeerArgs *__restrict__ args, int ctas) {
  __shared__ PeerArgs a;
  const int v = (int)blockIdx.x / ctas;
  for (int i = threadIdx.x; i < (int)(sizeof(PeerArgs) / 4); i += blockDim.x) ((unsigned *)&a)[i] = ((const unsigned *)(args + v))[i];
  __syncthreads();
  peer_update_body<WORLD, U>(a, (int)blockIdx.x - v * ctas, ctas);
}

static int peer_ctas() {
  // next to the backward GEMMs (128 CTAs that need a whole SM each) only the 20 remaining SMs are free
  static int n = -1;
  if (n < 0) { const char *e = getenv("TNB_DP_PEER_CTAS"); n = e ? atoi(e) : 20; if (n < 1) n = 1; }
  return n;
}
static long long peer_timeout_cycles() {
  static long long t = -1;
  if (t < 0) {  // milliseconds -> cycles at ~2 GHz, evaluated in double (0.5 ms is 1e6 cycles, not 0); 0 or less: no limit
    const char *e = getenv("TNB_PEER_TIMEOUT_MS");
    const double ms = e ? atof(e) : 10000.0;
    t = ms > 0.0 ? (long long)(ms * 2.0e6) : 0;
    if (ms > 0.0 && t < 1) t = 1;
  }
  return t;
}

// argument block of one rank's kernel from its job description
static int fill_peer_args(const TnbPeerJob *job, int rank, int world, unsigned *const *flags, unsigned seq, PeerArgs *out, long *blocks) {
  TNB_ARG(job && flags, "null");
  TNB_ARG(world >= 1 && world <= TNB_MAX_PEERS && rank >= 0 && rank < world, "rank/world");
  const TnbMatrixDim d = job->dW;
  TNB_ARG(d.rows > 0 && d.cols > 0 && d.stride >= d.cols && (d.stride & 3) == 0 && job->n_frames > 0, "dims");
  TNB_ARG(job->rows_pad >= d.rows && job->rows_pad % world == 0, "rows_pad must be a multiple of the world size, at least dW.rows");
  TNB_ARG(job->corrW && ((job->bias && job->corrb) || (!job->bias && !job->corrb)), "null");
  PeerArgs &a = *out;
  memset(&a, 0, sizeof(a));
  a.rank = rank; a.world = world;
  for (int r = 0; r < world; r++) {
    TNB_ARG(job->G[r] && job->W[r] && flags[r], "null peer pointer");
    TNB_ARG(((uintptr_t)job->G[r] & 15) == 0 && ((uintptr_t)job->W[r] & 15) == 0, "peer buffers must be 16-byte aligned");
    a.G[r] = job->G[r]; a.W[r] = job->W[r]; a.flags[r] = flags[r];
  }
  TNB_ARG(((uintptr_t)job->corrW & 15) == 0, "corrW must be 16-byte aligned");
  a.corr = job->corrW; a.bias = job->bias; a.corrb = job->corrb;
  a.cols = d.cols; a.stride = d.stride; a.rows_pad = job->rows_pad; a.shard = job->rows_pad / world;
  a.mmt = job->mmt;
  a.pushed = job->pushed ? 1 : 0;
  update_scalars(job->lr, job->mmt, job->wc, job->grad_div_frm, job->n_frames, &a.scale, &a.l2);
  a.seq = seq;
  a.timeout = peer_timeout_cycles();
  a.trace = nullptr;
  const long items = (long)a.shard * ((d.cols + 3) / 4);
  long nb = (items + 511) / 512;
  if (nb > peer_ctas()) nb = peer_ctas();
  if (nb < 1) nb = 1;
  *blocks = nb;
  return TNB_OK;
}

// TNB_DP_TRACE=1: [64][4] stamps of the update kernels, then [64][2] stamps around the pushes (slot = push counter % 64)
static int peer_trace_buffer(TnbContext *ctx) {
  static int want_trace = -1;
  if (want_trace < 0) { const char *e = getenv("TNB_DP_TRACE"); want_trace = (e && atoi(e) != 0) ? 1 : 0; }
  if (want_trace && !ctx->capturing && !ctx->peer_trace) {
    TNB_CUDA(cudaMalloc(&ctx->peer_trace, sizeof(long long) * 384));
    TNB_CUDA(cudaMemset(ctx->peer_trace, 0, sizeof(long long) * 384));
  }
  return TNB_OK;
}

static int launch_peer_update(TnbContext *ctx, cudaStream_t stream, const TnbPeerJob *job, int rank, int world, unsigned *const *flags,
                              unsigned seq, int defer_done = 0) {
  TNB_ARG(ctx != nullptr, "null");
  PeerArgs a;
  long blocks = 1;
  int rc = fill_peer_args(job, rank, world, flags, seq, &a, &blocks);
  if (rc != TNB_OK) return rc;
  if (peer_trace_buffer(ctx) != TNB_OK) return TNB_ERR_CUDA;
  a.trace = (long long *)ctx->peer_trace;
  a.defer_done = defer_done;
  static int spread = -1;
  if (spread < 0) { const char *e = getenv("TNB_DP_PEER_SPREAD"); spread = e ? atoi(e) : 0; }
  if (spread && world > 1 && (world == 2 || world == 4 || world == 8)) {
    const long items = (long)a.shard * ((a.cols + 3) / 4);
    long nb = spread > 1 ? spread : ctx->sm_count;   // TNB_DP_PEER_SPREAD=n (> 1): n CTAs instead of one per SM
    if (nb > (items + 127) / 128) nb = (items + 127) / 128;
    if (nb < 1) nb = 1;
    const dim3 grid((unsigned)nb), block(128);
    if (world == 2) dp_peer_update_spread_kernel<2><<<grid, block, 0, stream>>>(a);
    else if (world == 4) dp_peer_update_spread_kernel<4><<<grid, block, 0, stream>>>(a);
    else dp_peer_update_spread_kernel<8><<<grid, block, 0, stream>>>(a);
    TNB_LAUNCHED(ctx);
    return TNB_OK;
  }
  const dim3 grid((unsigned)blocks), block(512);
  switch (world) {
    case 1: dp_peer_update_kernel<1, 4><<<grid, block, 0, stream>>>(a); break;
    case 2: dp_peer_update_kernel<2, 4><<<grid, block, 0, stream>>>(a); break;
    case 4: dp_peer_update_kernel<4, 2><<<grid, block, 0, stream>>>(a); break;
    case 8: dp_peer_update_kernel<8, 2><<<grid, block, 0, stream>>>(a); break;
    default: dp_peer_update_kernel<0, 1><<<grid, block, 0, stream>>>(a); break;
  }
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

// decode the error word a timed-out wait left in a rank's flag block (0 = none)
static void describe_peer_error(unsigned w, int rank, char *buf, size_t n) {
  snprintf(buf, n, "rank %d: peer-memory update kernel %u (low 16 bits of its sequence number) gave up waiting for rank %u's %s flag "
           "(TNB_PEER_TIMEOUT_MS): the ranks did not issue the same sequence of updates, or that rank is stalled",
           rank, w & 0xFFFFu, (w >> 16) & 0xFFu, ((w >> 24) & 1u) ? "done" : "ready");
}

// the per-context flag block, mapped from every rank (collective, first use)
static int ensure_peer_flags(TnbContext *ctx) {
  if (ctx->peer_flags[ctx->rank]) return TNB_OK;
  void *p = nullptr;
  TNB_CUDA(cudaMalloc(&p, 64 * sizeof(unsigned)));
  TNB_CUDA(cudaMemset(p, 0, 64 * sizeof(unsigned)));
  TNB_CUDA(cudaDeviceSynchronize());
  void *mapped[TNB_MAX_PEERS];
  int rc = tnb_peer_map(ctx, p, mapped);
  if (rc != TNB_OK) { cudaFree(p); return rc; }
  for (int r = 0; r < ctx->world; r++) ctx->peer_flags[r] = (unsigned *)mapped[r];
  ctx->peer_seq = 0;
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_peer_map(TnbContext *ctx, void *local, void **mapped) {
  TNB_ARG(ctx && local && mapped, "null");
  const int world = ctx->world, rank = ctx->rank;
  TNB_ARG(world <= TNB_MAX_PEERS, "too many ranks");
  for (int r = 0; r < world; r++) mapped[r] = nullptr;
  mapped[rank] = local;
  if (world == 1) return TNB_OK;
  TNB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->local_group) {
    // ranks are threads of this process with peer access enabled (tnb_comm_init_local): a rank's device pointer is valid as it is
    std::vector<unsigned char> ptrs((size_t)world * sizeof(void *));
    memcpy(&ptrs[(size_t)rank * sizeof(void *)], &local, sizeof(void *));
    int rcl = comm_allgather_bytes(ctx, ptrs.data(), sizeof(void *));
    if (rcl != TNB_OK) return rcl;
    for (int r = 0; r < world; r++) memcpy(&mapped[r], &ptrs[(size_t)r * sizeof(void *)], sizeof(void *));
    return TNB_OK;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == TNB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  std::vector<unsigned char> all((size_t)world * TNB_IPC_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  TNB_CUDA(cudaIpcGetMemHandle(&h, local));
  memcpy(&all[(size_t)rank * TNB_IPC_HANDLE_BYTES], &h, TNB_IPC_HANDLE_BYTES);
  int rc = comm_allgather_bytes(ctx, all.data(), TNB_IPC_HANDLE_BYTES);  // the communicator of tnb_comm_init carries the handles
  if (rc != TNB_OK) return rc;
  for (int r = 0; r < world; r++) {
    if (r == rank) continue;
    cudaIpcMemHandle_t hr;
    memcpy(&hr, &all[(size_t)r * TNB_IPC_HANDLE_BYTES], TNB_IPC_HANDLE_BYTES);
    TNB_CUDA(cudaIpcOpenMemHandle(&mapped[r], hr, cudaIpcMemLazyEnablePeerAccess));
  }
  return TNB_OK;
}

int tnb_peer_unmap(TnbContext *ctx, void *const *mapped) {
  TNB_ARG(ctx && mapped, "null");
  if (ctx->local_group) return TNB_OK;  // plain peer pointers: nothing was opened
  for (int r = 0; r < ctx->world; r++)
    if (r != ctx->rank && mapped[r]) cudaIpcCloseMemHandle(mapped[r]);
  return TNB_OK;
}

int tnb_dp_peer_update(TnbContext *ctx, const TnbPeerJob *job, void *wait_event, void *done_event) {
  TNB_ARG(ctx && job, "null");
  cudaStream_t cs = ctx->world > 1 ? ctx->comm_stream : ctx->main_stream;
  if (ctx->world > 1) {
    TNB_ARG(ctx->nccl_comm != nullptr || ctx->local_group != nullptr, "communicator not initialised");
    int rc = ensure_peer_flags(ctx);
    if (rc != TNB_OK) return rc;
    TNB_CUDA(cudaEventRecord(ctx->ev_compute, ctx->main_stream));  // the gradient GEMM (and this layer's dX before it) is the producer
    TNB_CUDA(cudaStreamWaitEvent(cs, ctx->ev_compute, 0));
    if (wait_event) TNB_CUDA(cudaStreamWaitEvent(cs, (cudaEvent_t)wait_event, 0));
  } else if (!ctx->peer_flags[0]) {
    void *p = nullptr;
    TNB_CUDA(cudaMalloc(&p, 64 * sizeof(unsigned)));
    TNB_CUDA(cudaMemset(p, 0, 64 * sizeof(unsigned)));
    ctx->peer_flags[0] = (unsigned *)p;
    ctx->peer_seq = 0;
  }
  int rc = launch_peer_update(ctx, cs, job, ctx->rank, ctx->world, ctx->peer_flags, ++ctx->peer_seq);
  if (rc != TNB_OK) return rc;
  if (done_event) TNB_CUDA(cudaEventRecord((cudaEvent_t)done_event, cs));
  return TNB_OK;
}

int tnb_dp_peer_update_after(TnbContext *ctx, const TnbPeerJob *job, void *const *wait_events, int n_wait, void *done_event) {
  TNB_ARG(ctx && job && (wait_events || n_wait == 0) && n_wait >= 0, "null");
  TNB_ARG(ctx->world > 1, "tnb_dp_peer_update_after is the multi-rank entry point (tnb_dp_peer_update covers one rank)");
  TNB_ARG(ctx->nccl_comm != nullptr || ctx->local_group != nullptr, "communicator not initialised");
  cudaStream_t cs = ctx->comm_stream;
  int rc = ensure_peer_flags(ctx);
  if (rc != TNB_OK) return rc;
  for (int i = 0; i < n_wait; i++)
    if (wait_events[i]) TNB_CUDA(cudaStreamWaitEvent(cs, (cudaEvent_t)wait_events[i], 0));
  static int split_done = -1;
  // TNB_DP_SPLIT_DONE=1: the wait for the other ranks' done flags as a separate one-warp kernel on its own stream.  Measured SLOWER at
  // 4 GPUs (1.063 against 1.040 ms per bunch): what a kernel spends after its own rows is mostly the tail of its own CTAs, and two
  // layers' kernels running at once slow each other down.  Off by default.
  if (split_done < 0) { const char *e = getenv("TNB_DP_SPLIT_DONE"); split_done = (e && atoi(e) != 0) ? 1 : 0; }
  const bool split = split_done && !ctx->capturing;
  rc = launch_peer_update(ctx, cs, job, ctx->rank, ctx->world, ctx->peer_flags, ++ctx->peer_seq, split ? 1 : 0);
  if (rc != TNB_OK) return rc;
  if (!split) {
    if (done_event) TNB_CUDA(cudaEventRecord((cudaEvent_t)done_event, cs));
    return TNB_OK;
  }
  if (!ctx->done_stream) {
    TNB_CUDA(cudaStreamCreateWithFlags(&ctx->done_stream, cudaStreamNonBlocking));
    TNB_CUDA(cudaEventCreateWithFlags(&ctx->ev_done_fork, cudaEventDisableTiming));
  }
  TNB_CUDA(cudaEventRecord(ctx->ev_done_fork, cs));
  TNB_CUDA(cudaStreamWaitEvent(ctx->done_stream, ctx->ev_done_fork, 0));
  dp_peer_wait_done_kernel<<<1, 32, 0, ctx->done_stream>>>(ctx->peer_flags[ctx->rank], ctx->world, ctx->peer_seq, peer_timeout_cycles());
  TNB_LAUNCHED(ctx);
  if (done_event) TNB_CUDA(cudaEventRecord((cudaEvent_t)done_event, ctx->done_stream));
  return TNB_OK;
}

int tnb_peer_push_blocks(TnbContext *ctx, int stream_id, const float *G, float *const *Gpeers, int world, int rank, TnbMatrixDim dG,
                         int rows_pad, void *wait_event, void *done_event) {
  TNB_ARG(ctx && G && Gpeers, "null");
  TNB_ARG(world >= 1 && world <= TNB_MAX_PEERS && rank >= 0 && rank < world, "rank/world");
  TNB_ARG(dG.rows > 0 && dG.cols > 0 && dG.stride >= dG.cols && rows_pad >= dG.rows && rows_pad % world == 0, "dims");
  cudaStream_t s = stream_of(ctx, stream_id);
  TNB_ARG(s != nullptr, "unknown stream id");
  if (wait_event) TNB_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)wait_event, 0));
  if (peer_trace_buffer(ctx) != TNB_OK) return TNB_ERR_CUDA;
  long long *tr = ctx->peer_trace ? (long long *)ctx->peer_trace + 256 + (ctx->push_seq++ & 63u) * 2 : nullptr;
  if (tr) peer_stamp_kernel<<<1, 1, 0, s>>>(tr);
  const size_t shard = (size_t)(rows_pad / world), block = shard * (size_t)dG.stride;  // floats: whole rows, hence contiguous
  // One stream per destination: copies to different ranks run on different copy engines at the same time (one stream alone moves a
  // 4 MB block at 150-250 GB/s, measured: profiles/r02_dp_timeline.md).  TNB_DP_PUSH_STREAMS=0: all copies on `s`, one after the other.
  static int fan = -1;
  if (fan < 0) { const char *e = getenv("TNB_DP_PUSH_STREAMS"); fan = (e && atoi(e) == 0) ? 0 : 1; }
  bool forked = false;
  for (int k = 1; k <= world; k++) {
    const int o = (rank + k) % world;  // rank r starts with r + 1: the ranks' copies go to different destinations at any one time
    TNB_ARG(Gpeers[o] != nullptr, "null peer pointer");
    cudaStream_t cs = s;
    if (fan && o != rank && !ctx->capturing) {
      if (!ctx->push_streams[o]) {
        TNB_CUDA(cudaStreamCreateWithFlags(&ctx->push_streams[o], cudaStreamNonBlocking));
        TNB_CUDA(cudaEventCreateWithFlags(&ctx->push_events[o], cudaEventDisableTiming));
      }
      cs = ctx->push_streams[o];
      if (!forked) {  // the destination streams start where `s` stands now (behind wait_event and everything enqueued on `s` before)
        if (!ctx->ev_push_fork) TNB_CUDA(cudaEventCreateWithFlags(&ctx->ev_push_fork, cudaEventDisableTiming));
        TNB_CUDA(cudaEventRecord(ctx->ev_push_fork, s));
        forked = true;
      }
      TNB_CUDA(cudaStreamWaitEvent(cs, ctx->ev_push_fork, 0));
    }
    TNB_CUDA(cudaMemcpyAsync(Gpeers[o] + (size_t)rank * block, G + (size_t)o * block, block * sizeof(float), cudaMemcpyDeviceToDevice, cs));
    if (cs != s) {
      TNB_CUDA(cudaEventRecord(ctx->push_events[o], cs));
      TNB_CUDA(cudaStreamWaitEvent(s, ctx->push_events[o], 0));
    }
  }
  if (tr) peer_stamp_kernel<<<1, 1, 0, s>>>(tr + 1);
  if (done_event) TNB_CUDA(cudaEventRecord((cudaEvent_t)done_event, s));
  return TNB_OK;
}

int tnb_dp_peer_update_on(TnbContext *ctx, int stream_id, const TnbPeerJob *job, int rank, int world, unsigned *const *flags, unsigned seq) {
  TNB_ARG(ctx != nullptr, "null");
  cudaStream_t s = stream_of(ctx, stream_id);
  TNB_ARG(s != nullptr, "unknown stream id");
  return launch_peer_update(ctx, s, job, rank, world, flags, seq);
}

int tnb_dp_peer_update_virtual(TnbContext *ctx, const TnbPeerJob *jobs, int world, unsigned *const *flags, unsigned seq, int ctas_per_rank) {
  TNB_ARG(ctx && jobs && flags, "null");
  TNB_ARG(world >= 1 && world <= TNB_MAX_PEERS, "world");
  std::vector<PeerArgs> args((size_t)world);
  long blocks = 1;
  for (int r = 0; r < world; r++) {
    long nb = 1;
    int rc = fill_peer_args(&jobs[r], r, world, flags, seq, &args[(size_t)r], &nb);
    if (rc != TNB_OK) return rc;
    if (r == 0) blocks = nb;
    TNB_ARG(nb == blocks, "the ranks' jobs must describe the same layer");
  }
  if (ctas_per_rank > 0 && ctas_per_rank < blocks) blocks = ctas_per_rank;
  while (blocks > 1 && blocks * world > ctx->sm_count) blocks--;  // one CTA per SM is always co-resident
  PeerArgs *dargs = nullptr;
  TNB_CUDA(cudaMalloc(&dargs, sizeof(PeerArgs) * (size_t)world));
  cudaError_t e = cudaMemcpyAsync(dargs, args.data(), sizeof(PeerArgs) * (size_t)world, cudaMemcpyHostToDevice, ctx->stream);
  int ctas = (int)blocks;
  void *kargs[2] = {(void *)&dargs, (void *)&ctas};
  const dim3 grid((unsigned)(blocks * world)), block(512);
  if (e == cudaSuccess) {
    switch (world) {
      case 1: e = cudaLaunchCooperativeKernel((const void *)dp_peer_update_virtual_kernel<1, 4>, grid, block, kargs, 0, ctx->stream); break;
      case 2: e = cudaLaunchCooperativeKernel((const void *)dp_peer_update_virtual_kernel<2, 4>, grid, block, kargs, 0, ctx->stream); break;
      case 4: e = cudaLaunchCooperativeKernel((const void *)dp_peer_update_virtual_kernel<4, 2>, grid, block, kargs, 0, ctx->stream); break;
      case 8: e = cudaLaunchCooperativeKernel((const void *)dp_peer_update_virtual_kernel<8, 2>, grid, block, kargs, 0, ctx->stream); break;
      default: e = cudaLaunchCooperativeKernel((const void *)dp_peer_update_virtual_kernel<0, 1>, grid, block, kargs, 0, ctx->stream); break;
    }
  }
  if (e == cudaSuccess) { ctx->launches++; e = cudaStreamSynchronize(ctx->stream); }
  cudaFree(dargs);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("virtual peer update: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
    return TNB_ERR_CUDA;
  }
  for (int r = 0; r < world; r++) {
    unsigned w = 0;
    TNB_CUDA(cudaMemcpy(&w, flags[r] + 33, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (w) {
      char buf[320];
      describe_peer_error(w, r, buf, sizeof(buf));
      set_error("%s", buf);
      return TNB_ERR_COMM;
    }
  }
  return TNB_OK;
}

int tnb_peer_trace_read(TnbContext *ctx, long long *out /* [64][4] + [64][2] */, unsigned *seq /* [2] */) {
  TNB_ARG(ctx && out && seq, "null");
  TNB_ARG(ctx->peer_trace != nullptr, "no trace: set TNB_DP_TRACE=1 before the first peer-memory update");
  TNB_CUDA(cudaDeviceSynchronize());
  TNB_CUDA(cudaMemcpy(out, ctx->peer_trace, sizeof(long long) * 384, cudaMemcpyDeviceToHost));
  seq[0] = ctx->peer_seq;
  seq[1] = ctx->push_seq;
  return TNB_OK;
}

int tnb_peer_status(TnbContext *ctx) {
  TNB_ARG(ctx != nullptr, "null");
  unsigned *my = ctx->peer_flags[ctx->rank];
  if (!my) return TNB_OK;
  unsigned w = 0;
  TNB_CUDA(cudaMemcpy(&w, my + 33, sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (!w) return TNB_OK;
  char buf[320];
  describe_peer_error(w, ctx->rank, buf, sizeof(buf));
  set_error("%s", buf);
  return TNB_ERR_COMM;
}

}  // extern "C"
