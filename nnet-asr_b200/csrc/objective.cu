// objective.cu — CuSoftmax + CuCrossEntropy / CuMeanSquareError + CheckClass for sm_100a.
//
// Reference (src/): CuTNetLib/cuActivation.cc:26-31 -> CuBaseLib/cumath.cc:42-74 -> cukernels.cu:222-242 (one THREAD per
// row when cols > 256) / :306-343 (smem trees, cols <= 256); cuObjectiveFunction.cc:48-84 (9 device ops + 2 blocking
// D2H copies per bunch); cukernels.cu:398-483 (_check_class / _check_class_reduce).
//
// Here one CTA owns one row: the row is staged once in shared memory, max / sum / arg-max are warp-shuffle +
// smem reductions, and softmax, err = y - t, the row's cross-entropy term and the frame-accuracy flag come out of
// a single pass over HBM (read a, t; write y, err).  Per-row results go to a scratch vector and are folded into the
// device-resident TnbObjStats by a fixed-order single-CTA kernel, so the epoch totals are deterministic and no
// host synchronisation happens per bunch.
//
// Frame accuracy is integer work and keeps the reference's tie rules exactly:
//   cols > 256 : sequential scan, strict '>' from -1e20  == lowest index among the maxima, NaN never wins;
//   cols <= 256: the pairwise index tree of _max_id_reduce (left slot kept on ties) — replicated step by step.
#include <float.h>

#include "common.cuh"

namespace tnb {

constexpr int ROW_THREADS = 256;
constexpr int ROW_SMEM_FLOATS = 8192;  // rows up to 8192 columns are staged in shared memory

struct ArgMax {
  float v;
  int i;
};
// sequential-scan semantics: a candidate is valid iff v > -1e20f; among valid ones larger v wins, then lower index
__device__ __forceinline__ ArgMax am_merge(ArgMax a, ArgMax b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax am_make(float v, int i) {
  ArgMax r;
  if (v > -1e20f) { r.v = v; r.i = i; } else { r.v = -1e20f; r.i = 0x7fffffff; }
  return r;
}

__device__ __forceinline__ float block_reduce_max(float v, float *red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < ROW_THREADS / 32; w++) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float *red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.0f;
  for (int w = 0; w < ROW_THREADS / 32; w++) r += red[w];  // fixed order
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_reduce_argmax_seq(ArgMax a, float *redv, int *redi, int dflt) {
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = am_merge(a, b);
  }
  if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = a.v; redi[threadIdx.x >> 5] = a.i; }
  __syncthreads();
  ArgMax r;
  r.v = redv[0]; r.i = redi[0];
  for (int w = 1; w < ROW_THREADS / 32; w++) { ArgMax b; b.v = redv[w]; b.i = redi[w]; r = am_merge(r, b); }
  __syncthreads();
  return (r.i == 0x7fffffff) ? dflt : r.i;
}
// _max_id_reduce (cukernels.cu:424-446) over val[0..n) with n <= 256 == blockDim: identical slot algebra
__device__ __forceinline__ int block_argmax_tree(const float *val, int *idx, int n) {
  if (threadIdx.x < n) idx[threadIdx.x] = threadIdx.x;
  __syncthreads();
  int nTotal = n;
  while (nTotal > 1) {
    int half = (1 + nTotal) >> 1;
    if (threadIdx.x < half && threadIdx.x + half < nTotal) {
      // the reference also evaluates the unpaired slot with temp = -1e20 and would then copy an UNINITIALISED index
      // when a value is below -1e20 (undefined behaviour upstream); an unpaired slot simply keeps its candidate here
      float temp = val[idx[threadIdx.x + half]];
      if (temp > val[idx[threadIdx.x]]) idx[threadIdx.x] = idx[threadIdx.x + half];
    }
    __syncthreads();
    nTotal = (1 + nTotal) >> 1;
  }
  int r = idx[0];
  __syncthreads();
  return r;
}

// MODE 0: softmax only (Y = softmax(A))
// MODE 1: fused softmax + xent (A = activations; Y optional)
// MODE 2: xent on given Y (A = Y input, no softmax)
// Targets are either a dense matrix T (the reference's interface) or, when T == NULL, one class id per row (labels[r*lstride]):
// the row of T is then the one-hot vector of that id, generated in registers — every value the kernel computes with is the one it
// would have loaded, so both forms give the same bits (an id outside [0, cols) stands for an all-zero row).
template <int MODE>
__global__ void __launch_bounds__(ROW_THREADS) row_objective_kernel(const float *__restrict__ A, const float *__restrict__ T,
                                                                    const int *__restrict__ labels, int lstride,
                                                                    float *__restrict__ Y, float *__restrict__ Err, int rows,
                                                                    int cols, int stride, float *__restrict__ row_xent,
                                                                    int *__restrict__ row_match) {
  __shared__ float srow[ROW_SMEM_FLOATS];
  __shared__ float red[ROW_THREADS / 32];
  __shared__ int redi[ROW_THREADS / 32];
  __shared__ int sidx[256];
  const bool staged = cols <= ROW_SMEM_FLOATS;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float *a = A + (size_t)r * stride;
    const int lab = (MODE != 0 && !T) ? labels[(size_t)r * lstride] : -1;
    // ---- load row, max ----
    float mx = -1e20f;
    for (int c = threadIdx.x; c < cols; c += ROW_THREADS) {
      float v = a[c];
      if (staged) srow[c] = v;
      if (MODE != 2) mx = (mx < v) ? v : mx;  // reference: if(max < x) max = x, from -1e20
    }
    float sum = 0.0f;
    if (MODE != 2) {
      mx = block_reduce_max(mx, red);
      // ---- exp, sum ----
      for (int c = threadIdx.x; c < cols; c += ROW_THREADS) {
        float v = staged ? srow[c] : a[c];
        float e = expf(v - mx);
        if (staged) srow[c] = e; else if (Y) Y[(size_t)r * stride + c] = e;  // unstaged rows park exp() in Y
        sum += e;
      }
      sum = block_reduce_sum(sum, red);
    } else {
      __syncthreads();
    }
    // ---- normalise, err, xent term, argmax candidates ----
    float xe = 0.0f;
    ArgMax ay = am_make(-1e30f, 0), at = am_make(-1e30f, 0);
    for (int c = threadIdx.x; c < cols; c += ROW_THREADS) {
      float y;
      if (MODE == 2) y = staged ? srow[c] : a[c];
      else {
        float e = staged ? srow[c] : (Y ? Y[(size_t)r * stride + c] : expf(a[c] - mx));
        y = e / sum;
      }
      if (MODE != 2 && Y) Y[(size_t)r * stride + c] = y;
      if (MODE != 0) {
        float t = T ? T[(size_t)r * stride + c] : (c == lab ? 1.0f : 0.0f);
        Err[(size_t)r * stride + c] = y - t;
        float ly = logf(y < FLT_MIN ? FLT_MIN : y);
        xe += ly * t;
        if (cols > 256) { ay = am_merge(ay, am_make(y, c)); at = am_merge(at, am_make(t, c)); }
        else if (staged) srow[c] = y;
      }
    }
    if (MODE != 0) {
      xe = block_reduce_sum(xe, red);
      int out_id, des_id;
      if (cols > 256) {
        out_id = block_reduce_argmax_seq(ay, red, redi, -1);
        des_id = block_reduce_argmax_seq(at, red, redi, -2);
      } else {
        __syncthreads();
        out_id = block_argmax_tree(srow, sidx, cols);  // srow holds y
        for (int c = threadIdx.x; c < cols; c += ROW_THREADS) srow[c] = T ? T[(size_t)r * stride + c] : (c == lab ? 1.0f : 0.0f);
        __syncthreads();
        des_id = block_argmax_tree(srow, sidx, cols);
      }
      if (threadIdx.x == 0) {
        row_xent[r] = -xe;
        row_match[r] = (out_id == des_id) ? 1 : 0;
      }
    }
    __syncthreads();
  }
}

// Wide rows (256 < cols <= 4096, cols and pitch multiples of 4) of the fused softmax + cross-entropy: the row never touches
// shared memory.  Every thread issues the 16-byte loads of its (up to 4) activation AND target quads before anything else —
// in the staged kernel above the target loads started after two block reductions and waited out a full DRAM latency
// (ncu: 37 % of the stall samples on `err = y - t`) — and keeps them in registers through max / exp / sum / normalise.
// Same arithmetic per element (expf(v - max), e / sum, logf(max(y, FLT_MIN)) * t) and the same first-maximum rule for the
// frame accuracy; only the grouping of the fixed-order partial sums differs from the staged kernel.
// MODE 0: softmax only (same arithmetic, so tnb_softmax and tnb_softmax_xent return identical y); MODE 1: fused.
constexpr int ROW_VEC_MAX = 4;
template <int MODE>
__global__ void __launch_bounds__(ROW_THREADS) row_softmax_xent_wide_kernel(const float *__restrict__ A, const float *__restrict__ T,
                                                                            const int *__restrict__ labels, int lstride,
                                                                            float *__restrict__ Y, float *__restrict__ Err, int rows,
                                                                            int cols, int stride, float *__restrict__ row_xent,
                                                                            int *__restrict__ row_match) {
  __shared__ float red[ROW_THREADS / 32];
  __shared__ int redi[ROW_THREADS / 32];
  const int nq = cols >> 2;  // quads per row
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4 *a4 = (const float4 *)(A + (size_t)r * stride);
    const float4 *t4 = (const float4 *)(T + (size_t)r * stride);
    const int lab = (MODE == 1 && !T) ? labels[(size_t)r * lstride] : -1;
    float4 av[ROW_VEC_MAX], tv[ROW_VEC_MAX];
#pragma unroll
    for (int i = 0; i < ROW_VEC_MAX; i++) {
      const int qd = threadIdx.x + i * ROW_THREADS;
      if (qd < nq) {
        av[i] = a4[qd];
        if (MODE == 1) {
          if (T) tv[i] = t4[qd];
          else { const int c = qd << 2; tv[i] = make_float4(c == lab ? 1.0f : 0.0f, c + 1 == lab ? 1.0f : 0.0f, c + 2 == lab ? 1.0f : 0.0f, c + 3 == lab ? 1.0f : 0.0f); }
        }
      }
    }
    float mx = -1e20f;
#pragma unroll
    for (int i = 0; i < ROW_VEC_MAX; i++) {
      if (threadIdx.x + i * ROW_THREADS < nq) {
        mx = (mx < av[i].x) ? av[i].x : mx; mx = (mx < av[i].y) ? av[i].y : mx;
        mx = (mx < av[i].z) ? av[i].z : mx; mx = (mx < av[i].w) ? av[i].w : mx;
      }
    }
    mx = block_reduce_max(mx, red);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < ROW_VEC_MAX; i++) {
      if (threadIdx.x + i * ROW_THREADS < nq) {
        av[i].x = expf(av[i].x - mx); av[i].y = expf(av[i].y - mx); av[i].z = expf(av[i].z - mx); av[i].w = expf(av[i].w - mx);
        sum += av[i].x; sum += av[i].y; sum += av[i].z; sum += av[i].w;
      }
    }
    sum = block_reduce_sum(sum, red);
    float xe = 0.0f;
    ArgMax ay = am_make(-1e30f, 0), at = am_make(-1e30f, 0);
#pragma unroll
    for (int i = 0; i < ROW_VEC_MAX; i++) {
      const int qd = threadIdx.x + i * ROW_THREADS;
      if (qd < nq) {
        float4 y, e;
        y.x = av[i].x / sum; y.y = av[i].y / sum; y.z = av[i].z / sum; y.w = av[i].w / sum;
        if (MODE == 0) { ((float4 *)(Y + (size_t)r * stride))[qd] = y; continue; }
        e.x = y.x - tv[i].x; e.y = y.y - tv[i].y; e.z = y.z - tv[i].z; e.w = y.w - tv[i].w;
        if (Y) ((float4 *)(Y + (size_t)r * stride))[qd] = y;
        ((float4 *)(Err + (size_t)r * stride))[qd] = e;
        if (T || (lab >> 2) == qd) {  // class ids: only the label's quad has a non-zero target (the others would add +-0)
          xe += logf(y.x < FLT_MIN ? FLT_MIN : y.x) * tv[i].x; xe += logf(y.y < FLT_MIN ? FLT_MIN : y.y) * tv[i].y;
          xe += logf(y.z < FLT_MIN ? FLT_MIN : y.z) * tv[i].z; xe += logf(y.w < FLT_MIN ? FLT_MIN : y.w) * tv[i].w;
        }
        const int c = qd << 2;
        ay = am_merge(ay, am_make(y.x, c)); ay = am_merge(ay, am_make(y.y, c + 1));
        ay = am_merge(ay, am_make(y.z, c + 2)); ay = am_merge(ay, am_make(y.w, c + 3));
        at = am_merge(at, am_make(tv[i].x, c)); at = am_merge(at, am_make(tv[i].y, c + 1));
        at = am_merge(at, am_make(tv[i].z, c + 2)); at = am_merge(at, am_make(tv[i].w, c + 3));
      }
    }
    if (MODE == 1) {
      xe = block_reduce_sum(xe, red);
      const int out_id = block_reduce_argmax_seq(ay, red, redi, -1);
      const int des_id = block_reduce_argmax_seq(at, red, redi, -2);
      if (threadIdx.x == 0) {
        row_xent[r] = -xe;
        row_match[r] = (out_id == des_id) ? 1 : 0;
      }
    }
    __syncthreads();
  }
}

// standalone CheckClass: match[r] written directly (reference API, cumath.cc:178-206)
__global__ void __launch_bounds__(ROW_THREADS) check_class_kernel(const float *__restrict__ out, const float *__restrict__ des,
                                                                  int *__restrict__ match, int rows, int cols, int stride) {
  __shared__ float sval[256];
  __shared__ float red[ROW_THREADS / 32];
  __shared__ int redi[ROW_THREADS / 32];
  __shared__ int sidx[256];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    int out_id, des_id;
    if (cols > 256) {
      ArgMax ay = am_make(-1e30f, 0), at = am_make(-1e30f, 0);
      for (int c = threadIdx.x; c < cols; c += ROW_THREADS) {
        ay = am_merge(ay, am_make(out[(size_t)r * stride + c], c));
        at = am_merge(at, am_make(des[(size_t)r * stride + c], c));
      }
      out_id = block_reduce_argmax_seq(ay, red, redi, -1);
      des_id = block_reduce_argmax_seq(at, red, redi, -2);
    } else {
      if (threadIdx.x < cols) sval[threadIdx.x] = out[(size_t)r * stride + threadIdx.x];
      __syncthreads();
      out_id = block_argmax_tree(sval, sidx, cols);
      if (threadIdx.x < cols) sval[threadIdx.x] = des[(size_t)r * stride + threadIdx.x];
      __syncthreads();
      des_id = block_argmax_tree(sval, sidx, cols);
    }
    if (threadIdx.x == 0) match[r] = (out_id == des_id) ? 1 : 0;
    __syncthreads();
  }
}

// MSE: err = y - t ; per-row sum of err^2
__global__ void __launch_bounds__(ROW_THREADS) row_mse_kernel(const float *__restrict__ Y, const float *__restrict__ T,
                                                              float *__restrict__ Err, int rows, int cols, int stride,
                                                              float *__restrict__ row_val) {
  __shared__ float red[ROW_THREADS / 32];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float s = 0.0f;
    for (int c = threadIdx.x; c < cols; c += ROW_THREADS) {
      size_t k = (size_t)r * stride + c;
      float e = Y[k] - T[k];
      Err[k] = e;
      s += e * e;
    }
    s = block_reduce_sum(s, red);
    if (threadIdx.x == 0) row_val[r] = s;
  }
}

// fixed-order fold of the per-row results into the device-resident accumulators (single CTA)
__global__ void __launch_bounds__(256) stats_fold_kernel(const float *__restrict__ row_val, const int *__restrict__ row_match,
                                                         int rows, TnbObjStats *stats) {
  __shared__ double sd[256];
  __shared__ int si[256];
  double s = 0.0;
  int m = 0;
  for (int r = threadIdx.x; r < rows; r += 256) {
    s += (double)row_val[r];
    if (row_match) m += row_match[r];
  }
  sd[threadIdx.x] = s; si[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sd[threadIdx.x] += sd[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats->error += sd[0];
    stats->frames += rows;
    stats->correct += si[0];
  }
}

template <int MODE>
static int launch_row_objective(TnbContext *ctx, const float *A, const float *T, float *Y, float *Err, TnbMatrixDim d,
                                TnbObjStats *stats, const int *labels = nullptr, int lstride = 1) {
  TNB_ARG(ctx && A, "null");
  TNB_ARG(d.rows >= 0 && d.cols > 0 && d.stride >= d.cols, "dims");
  if (MODE != 0) TNB_ARG((T || labels) && Err && stats, "null");
  if (labels) TNB_ARG(!T && lstride >= 1, "either a dense target matrix or class ids");
  if (MODE == 0) TNB_ARG(Y, "null");
  if (MODE == 1 && d.cols > ROW_SMEM_FLOATS) TNB_ARG(Y, "rows wider than 8192 need the Y buffer");
  if (d.rows == 0) return TNB_OK;
  int rc = ensure_row_scratch(ctx, d.rows);
  if (rc != TNB_OK) return rc;
  int blocks = d.rows < ctx->sm_count * 8 ? d.rows : ctx->sm_count * 8;
  const bool wide = MODE != 2 && d.cols > 256 && d.cols <= 4 * ROW_VEC_MAX * ROW_THREADS && (d.cols & 3) == 0 && (d.stride & 3) == 0 &&
                    ((uintptr_t)A & 15) == 0 && ((uintptr_t)T & 15) == 0 && ((uintptr_t)Err & 15) == 0 && (!Y || ((uintptr_t)Y & 15) == 0);
  if (wide)
    row_softmax_xent_wide_kernel<(MODE == 0 ? 0 : 1)><<<blocks, ROW_THREADS, 0, ctx->stream>>>(A, T, labels, lstride, Y, Err, d.rows, d.cols, d.stride,
                                                                                              ctx->row_scratch, ctx->row_match);
  else
    row_objective_kernel<MODE><<<blocks, ROW_THREADS, 0, ctx->stream>>>(A, T, labels, lstride, Y, Err, d.rows, d.cols, d.stride, ctx->row_scratch,
                                                                         ctx->row_match);
  TNB_LAUNCHED(ctx);
  if (MODE != 0) {
    stats_fold_kernel<<<1, 256, 0, ctx->stream>>>(ctx->row_scratch, ctx->row_match, d.rows, stats);
    TNB_LAUNCHED(ctx);
  }
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_softmax(TnbContext *ctx, float *y, const float *x, TnbMatrixDim d) {
  return launch_row_objective<0>(ctx, x, nullptr, y, nullptr, d, nullptr);
}
int tnb_softmax_xent(TnbContext *ctx, const float *A, const float *T, float *Y, float *Err, TnbMatrixDim d, TnbObjStats *stats) {
  return launch_row_objective<1>(ctx, A, T, Y, Err, d, stats);
}
int tnb_xent_eval(TnbContext *ctx, const float *Y, const float *T, float *Err, TnbMatrixDim d, TnbObjStats *stats) {
  return launch_row_objective<2>(ctx, Y, T, nullptr, Err, d, stats);
}
int tnb_softmax_xent_labels(TnbContext *ctx, const float *A, const int *labels, int label_stride, float *Y, float *Err, TnbMatrixDim d,
                            TnbObjStats *stats) {
  TNB_ARG(labels != nullptr, "null");
  return launch_row_objective<1>(ctx, A, nullptr, Y, Err, d, stats, labels, label_stride);
}
int tnb_xent_eval_labels(TnbContext *ctx, const float *Y, const int *labels, int label_stride, float *Err, TnbMatrixDim d, TnbObjStats *stats) {
  TNB_ARG(labels != nullptr, "null");
  return launch_row_objective<2>(ctx, Y, nullptr, nullptr, Err, d, stats, labels, label_stride);
}
int tnb_mse_eval(TnbContext *ctx, const float *Y, const float *T, float *Err, TnbMatrixDim d, TnbObjStats *stats) {
  TNB_ARG(ctx && Y && T && Err && stats, "null");
  TNB_ARG(d.rows >= 0 && d.cols > 0 && d.stride >= d.cols, "dims");
  if (d.rows == 0) return TNB_OK;
  int rc = ensure_row_scratch(ctx, d.rows);
  if (rc != TNB_OK) return rc;
  int blocks = d.rows < ctx->sm_count * 8 ? d.rows : ctx->sm_count * 8;
  row_mse_kernel<<<blocks, ROW_THREADS, 0, ctx->stream>>>(Y, T, Err, d.rows, d.cols, d.stride, ctx->row_scratch);
  TNB_LAUNCHED(ctx);
  stats_fold_kernel<<<1, 256, 0, ctx->stream>>>(ctx->row_scratch, nullptr, d.rows, stats);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int tnb_check_class(TnbContext *ctx, const float *out, const float *des, int *match, TnbMatrixDim d) {
  TNB_ARG(ctx && out && des && match, "null");
  TNB_ARG(d.rows >= 0 && d.cols > 0 && d.stride >= d.cols, "dims");
  if (d.rows == 0) return TNB_OK;
  int blocks = d.rows < ctx->sm_count * 8 ? d.rows : ctx->sm_count * 8;
  check_class_kernel<<<blocks, ROW_THREADS, 0, ctx->stream>>>(out, des, match, d.rows, d.cols, d.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

}  // extern "C"
