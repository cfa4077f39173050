// elementwise.cu — HBM-bound kernels of the hot path: the 1:1 replacements of the reference's
// cukernels.cu elementwise launchers, the column-sum reduction, the cache/splice gathers, the
// Hybrid-Taus RNG kernels and the fused SGD update.
//
// Design: one thread handles 4 consecutive columns of one row with 16-byte accesses whenever the
// matrix pitch allows it (tnb_malloc_pitch guarantees it), rows are walked by a grid-stride loop and the
// grid is sized in multiples of the SM count.  Arithmetic keeps the reference's float/double
// promotions where they are observable (diff-sigmoid's double product, the double column sums).
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace tnb {

// ---------------------------------------------------------------------------------------- generic map kernel
// F: __device__ float op(float dst_old, int row, int col, size_t idx)   — idx = col + row*stride
template <typename F>
__global__ void __launch_bounds__(256) map2d_kernel(float *__restrict__ dst, int rows, int cols, int stride, F f) {
  const int vcols = (cols + 3) >> 2;
  const long total = (long)rows * vcols;
  const bool vec = ((stride & 3) == 0) && (((uintptr_t)dst & 15) == 0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vcols);
    const int c = (int)(i % vcols) << 2;
    const size_t base = (size_t)r * stride + c;
    if (vec && c + 3 < cols) {
      float4 v = *(float4 *)(dst + base);
      v.x = f(v.x, r, c, base); v.y = f(v.y, r, c + 1, base + 1);
      v.z = f(v.z, r, c + 2, base + 2); v.w = f(v.w, r, c + 3, base + 3);
      *(float4 *)(dst + base) = v;
    } else {
      for (int t = 0; t < 4 && c + t < cols; t++) dst[base + t] = f(dst[base + t], r, c + t, base + t);
    }
  }
}

template <typename F>
static int launch_map(TnbContext *ctx, float *dst, TnbMatrixDim d, F f) {
  TNB_ARG(ctx && dst, "null");
  TNB_ARG(d.rows >= 0 && d.cols >= 0 && d.stride >= d.cols, "dims");
  if (d.rows == 0 || d.cols == 0) return TNB_OK;
  long total = (long)d.rows * ((d.cols + 3) / 4);
  long blocks = (total + 255) / 256;
  long cap = (long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  map2d_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(dst, d.rows, d.cols, d.stride, f);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

// ---------------------------------------------------------------------------------------- column sums
// vec[c] = alpha * sum_r mat[r,c] + beta*vec[c].  Reference: _add_col_sum (double, serial, cukernels.cu:149-164)
// and _add_col_sum_reduce (float tree, :169-187).  Here: 32 columns per CTA, 8 row-slices per column with double
// partials, fixed-order combine -> deterministic, and at least as accurate as either reference variant.
// phase 1: grid (ceil(cols/128), S): CTA (bx, by) sums rows [by*chunk, (by+1)*chunk) of 128 columns into part[by][col] (double).
// 32 column groups of 4 (one 16-byte load per thread and row) x 8 row lanes; two rows in flight per thread.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ mat, double *__restrict__ part, int rows, int cols,
                                                             int stride, int chunk) {
  __shared__ double sm[8][129];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cx * 4;
  const int r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  const bool vec = ((stride & 3) == 0) && (((uintptr_t)mat & 15) == 0) && (c + 3 < cols);
  if (vec) {
    int r = r0 + ry;
    for (; r + 8 < r1; r += 16) {
      const float4 a = *(const float4 *)(mat + (size_t)r * stride + c);
      const float4 b = *(const float4 *)(mat + (size_t)(r + 8) * stride + c);
      s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
      t0 += b.x; t1 += b.y; t2 += b.z; t3 += b.w;
    }
    for (; r < r1; r += 8) {
      const float4 a = *(const float4 *)(mat + (size_t)r * stride + c);
      s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
    }
  } else {
    for (int r = r0 + ry; r < r1; r += 8) {
      const float *p = mat + (size_t)r * stride + c;
      if (c < cols) s0 += p[0];
      if (c + 1 < cols) s1 += p[1];
      if (c + 2 < cols) s2 += p[2];
      if (c + 3 < cols) s3 += p[3];
    }
  }
  sm[ry][cx * 4 + 0] = s0 + t0; sm[ry][cx * 4 + 1] = s1 + t1; sm[ry][cx * 4 + 2] = s2 + t2; sm[ry][cx * 4 + 3] = s3 + t3;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < cols) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) t += sm[k][threadIdx.x];
      part[(size_t)blockIdx.y * cols + cc] = t;
    }
  }
}
// phase 2: fixed-order combine of the S partials.  With `upd` != NULL the vector is a momentum buffer and the bias update of
// CuBiasedLinearity::Update (cuBiasedLinearity.cc:56-59) is applied in the same pass: vec = sum + beta*vec ; upd += scale*vec.
__global__ void __launch_bounds__(256) colsum_final_kernel(float alpha, const double *__restrict__ part, float beta, float *vec, int cols, int S,
                                                           float *upd, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double t = 0.0;
  for (int k = 0; k < S; k++) t += part[(size_t)k * cols + c];
  float b = (beta == 0.0f) ? 0.0f : beta * vec[c];
  const float v = (float)((double)alpha * t + (double)b);
  vec[c] = v;
  if (upd) upd[c] = scale * v + upd[c];
}

int launch_colsum_update(TnbContext *ctx, float alpha, const float *mat, float beta, float *vec, int rows, int cols, int stride, float *upd,
                         float scale) {
  if (cols == 0) return TNB_OK;
  const int cb = (cols + 127) / 128;
  int S = (2 * ctx->sm_count + cb - 1) / cb;  // ~2 CTAs per SM in phase 1
  if (S > (rows + 31) / 32) S = (rows + 31) / 32;
  if (S < 1) S = 1;
  const int chunk = (rows + S - 1) / S;
  int rc = ensure_vec_scratch(ctx, 2 * S * cols);  // doubles
  if (rc != TNB_OK) return rc;
  double *part = (double *)ctx->vec_scratch;
  colsum_partial_kernel<<<dim3(cb, S), 256, 0, ctx->stream>>>(mat, part, rows, cols, stride, chunk);
  TNB_LAUNCHED(ctx);
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, ctx->stream>>>(alpha, part, beta, vec, cols, S, upd, scale);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int launch_colsum(TnbContext *ctx, float alpha, const float *mat, float beta, float *vec, int rows, int cols, int stride) {
  return launch_colsum_update(ctx, alpha, mat, beta, vec, rows, cols, stride, nullptr, 0.0f);
}

// The bias halves of several CuBiasedLinearity::Update calls in ONE pair of launches (blockIdx.z = layer): in a deep net the
// seven per-layer pairs cost more in launch gaps than in work.  Same arithmetic and summation order per layer as above.
struct BiasBatch {
  int n;
  struct { const float *E; int rows, cols, stride, S, chunk; long part_off; float *corrb, *bias; float mmt, scale; } j[TNB_MAX_BIAS_JOBS];
};
__global__ void __launch_bounds__(256) colsum_partial_batch_kernel(const __grid_constant__ BiasBatch b, double *part) {
  const auto &j = b.j[blockIdx.z];
  if ((int)blockIdx.x * 128 >= j.cols || (int)blockIdx.y >= j.S) return;
  __shared__ double sm[8][129];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cx * 4;
  const int r0 = blockIdx.y * j.chunk, r1 = min(j.rows, r0 + j.chunk);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  const float *mat = j.E;
  const int stride = j.stride, cols = j.cols;
  const bool vec = ((stride & 3) == 0) && (((uintptr_t)mat & 15) == 0) && (c + 3 < cols);
  if (vec) {
    int r = r0 + ry;
    for (; r + 24 < r1; r += 32) {  // four independent 16-byte loads in flight per thread
      const float4 a = *(const float4 *)(mat + (size_t)r * stride + c);
      const float4 bb = *(const float4 *)(mat + (size_t)(r + 8) * stride + c);
      const float4 cc4 = *(const float4 *)(mat + (size_t)(r + 16) * stride + c);
      const float4 dd = *(const float4 *)(mat + (size_t)(r + 24) * stride + c);
      s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
      t0 += bb.x; t1 += bb.y; t2 += bb.z; t3 += bb.w;
      s0 += cc4.x; s1 += cc4.y; s2 += cc4.z; s3 += cc4.w;
      t0 += dd.x; t1 += dd.y; t2 += dd.z; t3 += dd.w;
    }
    for (; r + 8 < r1; r += 16) {
      const float4 a = *(const float4 *)(mat + (size_t)r * stride + c);
      const float4 bb = *(const float4 *)(mat + (size_t)(r + 8) * stride + c);
      s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
      t0 += bb.x; t1 += bb.y; t2 += bb.z; t3 += bb.w;
    }
    for (; r < r1; r += 8) {
      const float4 a = *(const float4 *)(mat + (size_t)r * stride + c);
      s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
    }
  } else {
    for (int r = r0 + ry; r < r1; r += 8) {
      const float *p = mat + (size_t)r * stride + c;
      if (c < cols) s0 += p[0];
      if (c + 1 < cols) s1 += p[1];
      if (c + 2 < cols) s2 += p[2];
      if (c + 3 < cols) s3 += p[3];
    }
  }
  sm[ry][cx * 4 + 0] = s0 + t0; sm[ry][cx * 4 + 1] = s1 + t1; sm[ry][cx * 4 + 2] = s2 + t2; sm[ry][cx * 4 + 3] = s3 + t3;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < cols) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) t += sm[k][threadIdx.x];
      part[j.part_off + (size_t)blockIdx.y * cols + cc] = t;
    }
  }
}
__global__ void __launch_bounds__(256) colsum_final_batch_kernel(const __grid_constant__ BiasBatch b, const double *__restrict__ part) {
  const auto &j = b.j[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j.cols) return;
  double t = 0.0;
  for (int k = 0; k < j.S; k++) t += part[j.part_off + (size_t)k * j.cols + c];
  if (!j.bias) { j.corrb[c] = (float)t; return; }  // gradient only (data parallel: summed over the ranks before it is applied)
  const float bb = (j.mmt == 0.0f) ? 0.0f : j.mmt * j.corrb[c];
  const float v = (float)(t + (double)bb);
  j.corrb[c] = v;
  j.bias[c] = j.scale * v + j.bias[c];
}

// ---------------------------------------------------------------------------------------- first / second moments (TNormCu)
// sum[c] += sum_r x[r,c] ; sumsq[c] += sum_r (float)(x[r,c]*x[r,c])  in double (TNormCu.cc:268-272 accumulates float products into
// double vectors on the host after a D2H copy of every utterance).  One CTA per 128 columns x row slice, double atomics at the end:
// the accumulators are doubles, so the order of the additions moves the result by ~1e-16 relative.
__global__ void __launch_bounds__(256) moments_kernel(const float *__restrict__ x, int rows, int cols, int stride, int chunk,
                                                      double *__restrict__ sum, double *__restrict__ sumsq) {
  __shared__ double s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
  double a = 0.0, b = 0.0;
  if (c < cols)
    for (int r = r0 + ry; r < r1; r += 8) {
      const float v = x[(size_t)r * stride + c];
      a += (double)v;
      b += (double)(v * v);
    }
  s1[ry][cx] = a; s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && c < cols) {
    for (int k = 1; k < 8; k++) { a += s1[k][cx]; b += s2[k][cx]; }
    atomicAdd(&sum[c], a);
    atomicAdd(&sumsq[c], b);
  }
}

// ---------------------------------------------------------------------------------------- gathers
// _randomize (cukernels.cu:384-393): y[r,:] = x[perm[r],:]   — one warp per row, 16-byte copies
__global__ void __launch_bounds__(256) gather_rows_kernel(float *__restrict__ y, const float *__restrict__ x,
                                                          const int *__restrict__ perm, int nrows, int cols, int sy, int sx) {
  const int wpb = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool vec = ((sy & 3) == 0) && ((sx & 3) == 0) && (((uintptr_t)y & 15) == 0) && (((uintptr_t)x & 15) == 0);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nrows; r += gridDim.x * wpb) {
    const float *src = x + (size_t)perm[r] * sx;
    float *dst = y + (size_t)r * sy;
    if (vec) {
      const int nv = cols >> 2;
      for (int i = lane; i < nv; i += 32) ((float4 *)dst)[i] = ((const float4 *)src)[i];
      for (int i = (nv << 2) + lane; i < cols; i += 32) dst[i] = src[i];
    } else {
      for (int i = lane; i < cols; i += 32) dst[i] = src[i];
    }
  }
}

// _expand (cukernels.cu:349-361): y[r, k*D + c] = x[clamp(r + off[k]), c] — one CTA per output row block; the
// D-wide source rows a block touches stay in L1 (each is reused by up to K output rows)
__global__ void __launch_bounds__(256) expand_kernel(float *__restrict__ y, const float *__restrict__ x,
                                                     const int *__restrict__ off, int rows, int cols_out, int sy, int rows_in,
                                                     int cols_in, int sx) {
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    for (int i = threadIdx.x; i < cols_out; i += blockDim.x) {
      int k = i / cols_in, c = i - k * cols_in;
      int sr = r + off[k];
      sr = sr < 0 ? 0 : (sr >= rows_in ? rows_in - 1 : sr);
      y[(size_t)r * sy + i] = x[(size_t)sr * sx + c];
    }
  }
}

// The same through shared memory: a CTA owns a block of R consecutive output rows, stages the window of input rows they can
// touch (R + max(off) - min(off) rows, clamped at the ends exactly like the per-element rule) with coalesced loads, and writes
// the K-fold wider output rows with 16-byte stores from a per-CTA lookup table (output column -> window offset).  Each input
// row is read from HBM once per row block instead of K times, and the stores — (K-1)/K of the traffic — are full lines.
// R is chosen in the kernel from the fixed window budget because the offsets live in device memory; when the offsets span
// more rows than the budget holds, the CTA falls back to the direct rule above.
constexpr int EXPAND_SMEM_BYTES = 44 * 1024;
__global__ void __launch_bounds__(256) expand_smem_kernel(float *__restrict__ y, const float *__restrict__ x, const int *__restrict__ off,
                                                          int rows, int cols_out, int sy, int rows_in, int cols_in, int sx) {
  extern __shared__ int esm[];
  int *lut = esm;                              // [cols_out]: (off[k] - min_off) * cols_in + c
  float *win = (float *)(esm + cols_out);      // [R + span][cols_in]
  const int K = cols_out / cols_in;
  int mn = off[0], mxo = off[0];
  for (int k = 1; k < K; k++) { const int o = off[k]; mn = o < mn ? o : mn; mxo = o > mxo ? o : mxo; }
  const int span = mxo - mn;
  const int budget_rows = (EXPAND_SMEM_BYTES - 4 * cols_out) / (4 * cols_in);
  const int R = budget_rows - span;
  if (R < 8) {  // window does not fit: direct rule
    for (int r = blockIdx.x; r < rows; r += gridDim.x)
      for (int i = threadIdx.x; i < cols_out; i += blockDim.x) {
        const int k = i / cols_in, c = i - k * cols_in;
        int sr = r + off[k];
        sr = sr < 0 ? 0 : (sr >= rows_in ? rows_in - 1 : sr);
        y[(size_t)r * sy + i] = x[(size_t)sr * sx + c];
      }
    return;
  }
  for (int i = threadIdx.x; i < cols_out; i += blockDim.x) {
    const int k = i / cols_in, c = i - k * cols_in;
    lut[i] = (off[k] - mn) * cols_in + c;
  }
  const bool vec = ((sy & 3) == 0) && (((uintptr_t)y & 15) == 0);
  const int nq = cols_out >> 2;
  for (int r0 = blockIdx.x * R; r0 < rows; r0 += gridDim.x * R) {
    const int nr = (rows - r0 < R) ? rows - r0 : R;
    __syncthreads();  // lut ready / previous block's window consumed
    for (int idx = threadIdx.x; idx < (nr + span) * cols_in; idx += blockDim.x) {
      const int w = idx / cols_in, c = idx - w * cols_in;
      int sr = r0 + mn + w;
      sr = sr < 0 ? 0 : (sr >= rows_in ? rows_in - 1 : sr);
      win[idx] = x[(size_t)sr * sx + c];
    }
    __syncthreads();
    if (vec) {
      for (int idx = threadIdx.x; idx < nr * nq; idx += blockDim.x) {
        const int rr = idx / nq, q = idx - rr * nq;
        const float *wr = win + rr * cols_in;
        float4 v;
        v.x = wr[lut[4 * q]]; v.y = wr[lut[4 * q + 1]]; v.z = wr[lut[4 * q + 2]]; v.w = wr[lut[4 * q + 3]];
        *(float4 *)(y + (size_t)(r0 + rr) * sy + 4 * q) = v;
      }
      const int tail = cols_out - 4 * nq;
      for (int idx = threadIdx.x; idx < nr * tail; idx += blockDim.x) {
        const int rr = idx / tail, i = 4 * nq + (idx - rr * tail);
        y[(size_t)(r0 + rr) * sy + i] = win[rr * cols_in + lut[i]];
      }
    } else {
      for (int idx = threadIdx.x; idx < nr * cols_out; idx += blockDim.x) {
        const int rr = idx / cols_out, i = idx - rr * cols_out;
        y[(size_t)(r0 + rr) * sy + i] = win[rr * cols_in + lut[i]];
      }
    }
  }
}

// _rearrange (cukernels.cu:366-379): column gather, +inf for a bad index
__global__ void __launch_bounds__(256) rearrange_kernel(float *__restrict__ y, const float *__restrict__ x,
                                                        const int *__restrict__ copy_from, int rows, int cols_out, int sy,
                                                        int cols_in, int sx) {
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    for (int i = threadIdx.x; i < cols_out; i += blockDim.x) {
      int sc = copy_from[i];
      y[(size_t)r * sy + i] = (sc >= 0 && sc < cols_in) ? x[(size_t)r * sx + sc] : __int_as_float(0x7f800000);
    }
  }
}

// 16-byte stores where the row allows it (pitches are multiples of 32 floats; the pad columns of a row are never read as data)
__global__ void __launch_bounds__(256) onehot_kernel(float *__restrict__ T, const int *__restrict__ lab, int lstride, int rows, int cols,
                                                     int stride) {
  const bool vec = ((uintptr_t)T & 15) == 0 && (stride & 3) == 0;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int l = lab[(size_t)r * lstride];
    float *t = T + (size_t)r * stride;
    if (vec) {
      const int nq = cols >> 2;
      for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int c = q << 2;
        ((float4 *)t)[q] = make_float4(c == l ? 1.0f : 0.0f, c + 1 == l ? 1.0f : 0.0f, c + 2 == l ? 1.0f : 0.0f, c + 3 == l ? 1.0f : 0.0f);
      }
      for (int i = (nq << 2) + threadIdx.x; i < cols; i += blockDim.x) t[i] = (i == l) ? 1.0f : 0.0f;
    } else {
      for (int i = threadIdx.x; i < cols; i += blockDim.x) t[i] = (i == l) ? 1.0f : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------- Hybrid-Taus RNG
// curandkernels.cu:15-46.  Integer streams are bit-exact; the uniform is float(2.3283064365387e-10 * (double)u32).
__device__ __forceinline__ unsigned taus_step(unsigned &z, int S1, int S2, int S3, unsigned M) {
  unsigned b = (((z << S1) ^ z) >> S2);
  return z = (((z & M) << S3) ^ b);
}
__device__ __forceinline__ unsigned lcg_step(unsigned &z, unsigned A, unsigned C) { return z = (A * z + C); }
__device__ __forceinline__ float hybrid_taus(unsigned &z1, unsigned &z2, unsigned &z3, unsigned &z4) {
  float randval;
  do {
    randval = (float)(2.3283064365387e-10 * (double)(taus_step(z1, 13, 19, 12, 4294967294U) ^ taus_step(z2, 2, 25, 4, 4294967288U) ^
                                                     taus_step(z3, 3, 11, 17, 4294967280U) ^ lcg_step(z4, 1664525, 1013904223U)));
  } while (!(randval > 0.0f && randval < 1.0f));
  return randval;
}
__device__ __forceinline__ float box_muller(unsigned &z1, unsigned &z2, unsigned &z3, unsigned &z4) {
  // curandkernels.cu:72-82 (T = float): r = sqrt(-2.0*log(u0)) in double then float; sin in float
  const float M_2PI_F = 6.283185307179586476925286766558;
  float u0 = hybrid_taus(z1, z2, z3, z4), u1 = hybrid_taus(z1, z2, z3, z4);
  float r = (float)sqrt(-2.0 * (double)logf(u0));
  float theta = M_2PI_F * u1;
  return r * sinf(theta);
}

// MODE 0: mat = uniform ; 1: mat = gauss ; 2: states = probs > uniform ; 3: mat += gscale*gauss
template <int MODE>
__global__ void __launch_bounds__(256) rand_kernel(float *__restrict__ out, const float *__restrict__ probs, float gscale,
                                                   unsigned *__restrict__ z1, unsigned *__restrict__ z2,
                                                   unsigned *__restrict__ z3, unsigned *__restrict__ z4, int rows, int cols,
                                                   int stride) {
  const long total = (long)rows * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const size_t k = (size_t)r * stride + c;
    unsigned a = z1[k], b = z2[k], cc = z3[k], d = z4[k];
    float v;
    if (MODE == 0 || MODE == 2) v = hybrid_taus(a, b, cc, d); else v = box_muller(a, b, cc, d);
    z1[k] = a; z2[k] = b; z3[k] = cc; z4[k] = d;
    if (MODE == 0 || MODE == 1) out[k] = v;
    else if (MODE == 2) out[k] = (probs[k] > v) ? 1.0f : 0.0f;
    else out[k] = gscale * v + out[k];
  }
}

template <int MODE>
static int launch_rand(TnbContext *ctx, float *out, const float *probs, float gscale, unsigned *z1, unsigned *z2, unsigned *z3,
                       unsigned *z4, TnbMatrixDim d) {
  TNB_ARG(ctx && out && z1 && z2 && z3 && z4, "null");
  TNB_ARG(d.rows >= 0 && d.cols >= 0 && d.stride >= d.cols, "dims");
  if (d.rows == 0 || d.cols == 0) return TNB_OK;
  long total = (long)d.rows * d.cols;
  long blocks = (total + 255) / 256, cap = (long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  rand_kernel<MODE><<<(int)blocks, 256, 0, ctx->stream>>>(out, probs, gscale, z1, z2, z3, z4, d.rows, d.cols, d.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

// ---------------------------------------------------------------------------------------- fused SGD update
// cuBiasedLinearity.cc:55-63 given the summed gradient G:  corr = G + mmt*corr ; W += scale*corr ; W += l2*W
__global__ void __launch_bounds__(256) sgd_update_kernel(const float *__restrict__ G, float *__restrict__ W,
                                                         float *__restrict__ corr, int rows, int cols, int stride, float mmt,
                                                         float scale, float l2) {
  const int vcols = (cols + 3) >> 2;
  const long total = (long)rows * vcols;
  const bool vec = ((stride & 3) == 0) && (((uintptr_t)G & 15) == 0) && (((uintptr_t)W & 15) == 0) && (((uintptr_t)corr & 15) == 0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vcols);
    const int c = (int)(i % vcols) << 2;
    const size_t base = (size_t)r * stride + c;
    if (vec && c + 3 < cols) {
      float4 g = *(const float4 *)(G + base), k = *(float4 *)(corr + base), w = *(float4 *)(W + base);
      k.x = g.x + mmt * k.x; k.y = g.y + mmt * k.y; k.z = g.z + mmt * k.z; k.w = g.w + mmt * k.w;
      w.x = scale * k.x + w.x; w.y = scale * k.y + w.y; w.z = scale * k.z + w.z; w.w = scale * k.w + w.w;
      if (l2 != 0.0f) { w.x = l2 * w.x + w.x; w.y = l2 * w.y + w.y; w.z = l2 * w.z + w.z; w.w = l2 * w.w + w.w; }
      *(float4 *)(corr + base) = k;
      *(float4 *)(W + base) = w;
    } else {
      for (int t = 0; t < 4 && c + t < cols; t++) {
        float k = G[base + t] + mmt * corr[base + t];
        float w = scale * k + W[base + t];
        if (l2 != 0.0f) w = l2 * w + w;
        corr[base + t] = k; W[base + t] = w;
      }
    }
  }
}

// the reference evaluates the update scalars in float (cuBiasedLinearity.cc:44-63)
// the same for several parameter arrays in one launch (blockIdx.y = array): the data-parallel step applies all layers' updates
// after the last all-reduce, where seven pairs of launches cost more in gaps than in bandwidth.  Optionally refreshes the bf16
// twin of the updated weights (TNB_MATH_BF16).
struct SgdBatch {
  int n;
  struct { const float *G; float *W, *corr; uint16_t *W16; int rows, cols, stride, ldw16; float mmt, scale, l2; } j[2 * TNB_MAX_BIAS_JOBS];
};
__global__ void __launch_bounds__(256) sgd_update_batch_kernel(const __grid_constant__ SgdBatch b) {
  const auto &j = b.j[blockIdx.y];
  const int vcols = (j.cols + 3) >> 2;
  const long total = (long)j.rows * vcols;
  const bool vec = ((j.stride & 3) == 0) && (((uintptr_t)j.G & 15) == 0) && (((uintptr_t)j.W & 15) == 0) && (((uintptr_t)j.corr & 15) == 0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vcols);
    const int c = (int)(i % vcols) << 2;
    const size_t base = (size_t)r * j.stride + c;
    if (vec && c + 3 < j.cols) {
      float4 g = *(const float4 *)(j.G + base), k = *(float4 *)(j.corr + base), w = *(float4 *)(j.W + base);
      k.x = g.x + j.mmt * k.x; k.y = g.y + j.mmt * k.y; k.z = g.z + j.mmt * k.z; k.w = g.w + j.mmt * k.w;
      w.x = j.scale * k.x + w.x; w.y = j.scale * k.y + w.y; w.z = j.scale * k.z + w.z; w.w = j.scale * k.w + w.w;
      if (j.l2 != 0.0f) { w.x = j.l2 * w.x + w.x; w.y = j.l2 * w.y + w.y; w.z = j.l2 * w.z + w.z; w.w = j.l2 * w.w + w.w; }
      *(float4 *)(j.corr + base) = k;
      *(float4 *)(j.W + base) = w;
      if (j.W16) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
        uint2 u;
        u.x = *(const uint32_t *)&lo; u.y = *(const uint32_t *)&hi;
        *(uint2 *)(j.W16 + (size_t)r * j.ldw16 + c) = u;
      }
    } else {
      for (int t = 0; t < 4 && c + t < j.cols; t++) {
        float k = j.G[base + t] + j.mmt * j.corr[base + t];
        float w = j.scale * k + j.W[base + t];
        if (j.l2 != 0.0f) w = j.l2 * w + w;
        j.corr[base + t] = k; j.W[base + t] = w;
        if (j.W16) j.W16[(size_t)r * j.ldw16 + c + t] = __bfloat16_as_ushort(__float2bfloat16_rn(w));
      }
    }
  }
}

void update_scalars(float lr, float mmt, float wc, int gdf, int rows, float *scale, float *l2) {
  float N = 1;
  if (gdf) N = (float)rows;
  float mmt_gain = (float)(1.0 / (1.0 - mmt));
  N *= mmt_gain;
  *scale = -lr / N;
  *l2 = (float)(-lr * wc * (gdf ? 1.0 : rows));
}

int launch_sgd_update(TnbContext *ctx, cudaStream_t stream, const float *G, float *W, float *corr, int rows, int cols, int stride,
                      float mmt, float scale, float l2) {
  if (rows <= 0 || cols <= 0) return TNB_OK;
  long total = (long)rows * ((cols + 3) / 4);
  long blocks = (total + 255) / 256, cap = (long)ctx->sm_count * 8;
  // On the communication stream the update runs next to the backward GEMMs, whose CTAs need a whole SM each (196 KB of shared
  // memory, most of the register file): a few CTAs that fit into the SMs the GEMM grid leaves free, instead of a device-wide grid
  // that would stand in the way of the next GEMM's CTAs.
  if (stream != ctx->stream) {
    static int side_ctas = -1;
    if (side_ctas < 0) { const char *e = getenv("TNB_DP_UPDATE_CTAS"); side_ctas = e ? atoi(e) : 32; }
    cap = side_ctas;
  }
  if (blocks > cap) blocks = cap;
  sgd_update_kernel<<<(int)blocks, 256, 0, stream>>>(G, W, corr, rows, cols, stride, mmt, scale, l2);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

// ---------------------------------------------------------------------------------------- fp32 -> bf16 (TNB_MATH_BF16 operands)
// one thread converts 8 consecutive columns: two 16-byte loads, one 16-byte store; columns between `cols` and the pitch are
// written as zeros so that a later TMA box never reads uninitialised padding as NaN
__global__ void __launch_bounds__(256) to_bf16_kernel(uint16_t *__restrict__ dst, int dst_stride, const float *__restrict__ src,
                                                      int rows, int cols, int src_stride) {
  const int vcols = dst_stride >> 3;
  const long total = (long)rows * vcols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vcols), c = (int)(i % vcols) << 3;
    const float *sp = src + (size_t)r * src_stride + c;
    float v[8];
    if (c + 7 < cols && ((src_stride & 3) == 0) && (((uintptr_t)src & 15) == 0)) {
      const float4 a = *(const float4 *)sp, b = *(const float4 *)(sp + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int t = 0; t < 8; t++) v[t] = (c + t < cols) ? sp[t] : 0.0f;
    }
    uint4 o;
    __nv_bfloat162 p;
    p = __floats2bfloat162_rn(v[0], v[1]); o.x = *(const uint32_t *)&p;
    p = __floats2bfloat162_rn(v[2], v[3]); o.y = *(const uint32_t *)&p;
    p = __floats2bfloat162_rn(v[4], v[5]); o.z = *(const uint32_t *)&p;
    p = __floats2bfloat162_rn(v[6], v[7]); o.w = *(const uint32_t *)&p;
    *(uint4 *)(dst + (size_t)r * dst_stride + c) = o;
  }
}

int launch_to_bf16(TnbContext *ctx, uint16_t *dst, int dst_stride, const float *src, int rows, int cols, int src_stride) {
  TNB_ARG(ctx && dst && src, "null");
  TNB_ARG(rows >= 0 && cols >= 0 && dst_stride >= cols && src_stride >= cols, "dims");
  TNB_ARG((dst_stride % 8) == 0 && ((uintptr_t)dst % 16) == 0, "bf16 arrays need a 16-byte aligned base and a pitch multiple of 8");
  if (rows == 0 || cols == 0) return TNB_OK;
  const long total = (long)rows * (dst_stride >> 3);
  long blocks = (total + 255) / 256, cap = (long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  to_bf16_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(dst, dst_stride, src, rows, cols, src_stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

#define DIMCHK(d) TNB_ARG((d).rows >= 0 && (d).cols >= 0 && (d).stride >= (d).cols, "dims")

extern "C" {

int tnb_set_const(TnbContext *ctx, float *mat, float value, TnbMatrixDim d) {
  return launch_map(ctx, mat, d, [=] __device__(float, int, int, size_t) { return value; });
}
int tnb_apply_log(TnbContext *ctx, float *mat, TnbMatrixDim d) {
  return launch_map(ctx, mat, d, [=] __device__(float v, int, int, size_t) { return logf(v); });
}
int tnb_scale_cols(TnbContext *ctx, float *mat, const float *scale, TnbMatrixDim d) {
  TNB_ARG(scale, "null");
  return launch_map(ctx, mat, d, [=] __device__(float v, int, int c, size_t) { return v * scale[c]; });
}
int tnb_scale_rows(TnbContext *ctx, float *mat, const float *scale, TnbMatrixDim d) {
  TNB_ARG(scale, "null");
  return launch_map(ctx, mat, d, [=] __device__(float v, int r, int, size_t) { return v * scale[r]; });
}
int tnb_add_scaled(TnbContext *ctx, float alpha, const float *A, float beta, float *dst, TnbMatrixDim d) {
  TNB_ARG(A, "null");
  // reference reads dst even when beta == 0 (0*NaN = NaN); kept
  return launch_map(ctx, dst, d, [=] __device__(float v, int, int, size_t i) { return alpha * A[i] + beta * v; });
}
int tnb_add_scaled_row(TnbContext *ctx, float alpha, const float *row, float beta, float *dst, TnbMatrixDim d) {
  TNB_ARG(row, "null");
  return launch_map(ctx, dst, d, [=] __device__(float v, int, int c, size_t) { return alpha * row[c] + beta * v; });
}
int tnb_mul_elem(TnbContext *ctx, float *mat, const float *A, TnbMatrixDim d) {
  TNB_ARG(A, "null");
  return launch_map(ctx, mat, d, [=] __device__(float v, int, int, size_t i) { return v * A[i]; });
}
int tnb_log_elem(TnbContext *ctx, float *mat, TnbMatrixDim d) {
  return launch_map(ctx, mat, d, [=] __device__(float v, int, int, size_t) { return logf(v < FLT_MIN ? FLT_MIN : v); });
}
int tnb_sigmoid(TnbContext *ctx, float *y, const float *x, TnbMatrixDim d) {
  TNB_ARG(x, "null");
  // cukernels.cu:194-206: float exp, double add/divide, rounded to float
  return launch_map(ctx, y, d, [=] __device__(float, int, int, size_t i) { return (float)(1.0 / (1.0 + (double)expf(-x[i]))); });
}
int tnb_diff_sigmoid(TnbContext *ctx, float *eout, const float *e, const float *y, TnbMatrixDim d) {
  TNB_ARG(e && y, "null");
  // cukernels.cu:211-217: y*(1.0-y)*e evaluated in double
  return launch_map(ctx, eout, d, [=] __device__(float, int, int, size_t i) {
    double yy = (double)y[i];
    return (float)(yy * (1.0 - yy) * (double)e[i]);
  });
}

int tnb_add_col_sum(TnbContext *ctx, float alpha, const float *mat, float beta, float *vec, TnbMatrixDim d) {
  TNB_ARG(ctx && mat && vec, "null");
  DIMCHK(d);
  return launch_colsum(ctx, alpha, mat, beta, vec, d.rows, d.cols, d.stride);
}

int tnb_expand(TnbContext *ctx, float *y, const float *x, const int *off, TnbMatrixDim d_out, TnbMatrixDim d_in) {
  TNB_ARG(ctx && y && x && off, "null");
  DIMCHK(d_out); DIMCHK(d_in);
  TNB_ARG(d_in.cols > 0 && d_out.cols % d_in.cols == 0, "expand: output cols must be a multiple of input cols");
  TNB_ARG(d_out.rows == d_in.rows, "expand: rows");
  if (d_out.rows == 0) return TNB_OK;
  int blocks = d_out.rows < ctx->sm_count * 8 ? d_out.rows : ctx->sm_count * 8;
  if (d_in.cols > 0 && d_out.cols % d_in.cols == 0 && 4 * d_out.cols + 16 * 4 * d_in.cols <= EXPAND_SMEM_BYTES) {
    // shared-memory version: at most 5 CTAs per SM (44 KB each), row blocks walked by a grid-stride loop
    const int rb = (EXPAND_SMEM_BYTES - 4 * d_out.cols) / (4 * d_in.cols);  // upper bound of rows per block
    int nb = (d_out.rows + (rb > 16 ? rb / 2 : 8) - 1) / (rb > 16 ? rb / 2 : 8);
    if (nb > ctx->sm_count * 5) nb = ctx->sm_count * 5;
    if (nb < 1) nb = 1;
    expand_smem_kernel<<<nb, 256, EXPAND_SMEM_BYTES, ctx->stream>>>(y, x, off, d_out.rows, d_out.cols, d_out.stride, d_in.rows, d_in.cols,
                                                                    d_in.stride);
  } else {
    expand_kernel<<<blocks, 256, 0, ctx->stream>>>(y, x, off, d_out.rows, d_out.cols, d_out.stride, d_in.rows, d_in.cols, d_in.stride);
  }
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int tnb_rearrange(TnbContext *ctx, float *y, const float *x, const int *copy_from, TnbMatrixDim d_out, TnbMatrixDim d_in) {
  TNB_ARG(ctx && y && x && copy_from, "null");
  DIMCHK(d_out); DIMCHK(d_in);
  TNB_ARG(d_out.rows == d_in.rows, "rearrange: rows");
  if (d_out.rows == 0 || d_out.cols == 0) return TNB_OK;
  int blocks = d_out.rows < ctx->sm_count * 8 ? d_out.rows : ctx->sm_count * 8;
  rearrange_kernel<<<blocks, 256, 0, ctx->stream>>>(y, x, copy_from, d_out.rows, d_out.cols, d_out.stride, d_in.cols, d_in.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int tnb_randomize(TnbContext *ctx, float *y, const float *x, const int *copy_from, TnbMatrixDim d_out, TnbMatrixDim d_in) {
  TNB_ARG(ctx && y && x && copy_from, "null");
  DIMCHK(d_out); DIMCHK(d_in);
  TNB_ARG(d_out.cols == d_in.cols, "randomize: cols");
  if (d_out.rows == 0 || d_out.cols == 0) return TNB_OK;
  int blocks = (d_out.rows + 7) / 8;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  gather_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(y, x, copy_from, d_out.rows, d_out.cols, d_out.stride, d_in.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int tnb_onehot_strided(TnbContext *ctx, float *T, const int *labels, int label_stride, TnbMatrixDim d) {
  TNB_ARG(ctx && T && labels && label_stride >= 1, "null");
  DIMCHK(d);
  if (d.rows == 0 || d.cols == 0) return TNB_OK;
  int blocks = d.rows < ctx->sm_count * 8 ? d.rows : ctx->sm_count * 8;
  onehot_kernel<<<blocks, 256, 0, ctx->stream>>>(T, labels, label_stride, d.rows, d.cols, d.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}
int tnb_onehot(TnbContext *ctx, float *T, const int *labels, TnbMatrixDim d) { return tnb_onehot_strided(ctx, T, labels, 1, d); }

int tnb_rand(TnbContext *ctx, float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, TnbMatrixDim d) {
  return launch_rand<0>(ctx, mat, nullptr, 0.0f, z1, z2, z3, z4, d);
}
int tnb_gauss_rand(TnbContext *ctx, float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, TnbMatrixDim d) {
  return launch_rand<1>(ctx, mat, nullptr, 0.0f, z1, z2, z3, z4, d);
}
int tnb_binarize_probs(TnbContext *ctx, float *states, const float *probs, const float *rnd, TnbMatrixDim d) {
  TNB_ARG(probs && rnd, "null");
  return launch_map(ctx, states, d, [=] __device__(float, int, int, size_t i) { return (probs[i] > rnd[i]) ? 1.0f : 0.0f; });
}
int tnb_rand_binarize(TnbContext *ctx, float *states, const float *probs, unsigned *z1, unsigned *z2, unsigned *z3,
                      unsigned *z4, TnbMatrixDim d) {
  TNB_ARG(probs, "null");
  return launch_rand<2>(ctx, states, probs, 0.0f, z1, z2, z3, z4, d);
}
int tnb_add_gauss_noise(TnbContext *ctx, float *tgt, float gscale, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4,
                        TnbMatrixDim d) {
  return launch_rand<3>(ctx, tgt, nullptr, gscale, z1, z2, z3, z4, d);
}

int tnb_accum_moments(TnbContext *ctx, const float *X, TnbMatrixDim d, double *sum, double *sumsq) {
  TNB_ARG(ctx && X && sum && sumsq, "null");
  DIMCHK(d);
  if (d.rows == 0 || d.cols == 0) return TNB_OK;
  const int cb = (d.cols + 31) / 32;
  int S = (2 * ctx->sm_count + cb - 1) / cb;
  if (S > (d.rows + 63) / 64) S = (d.rows + 63) / 64;
  if (S < 1) S = 1;
  moments_kernel<<<dim3(cb, S), 256, 0, ctx->stream>>>(X, d.rows, d.cols, d.stride, (d.rows + S - 1) / S, sum, sumsq);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_to_bf16(TnbContext *ctx, uint16_t *dst, int dst_stride, const float *src, TnbMatrixDim d) {
  return launch_to_bf16(ctx, dst, dst_stride, src, d.rows, d.cols, d.stride);
}

int tnb_affine_grad(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *G, TnbMatrixDim dG,
                    float *gb) {
  TNB_ARG(ctx && X && E && G, "null");
  TNB_ARG(dX.rows == dE.rows && dG.rows == dX.cols && dG.cols == dE.cols, "dimension mismatch");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = G; ep.ldc = dG.stride; ep.alpha = 1.0f; ep.beta = 0.0f;
  int rc = launch_gemm(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X, dX.stride, E, dE.stride, ep);
  if (rc != TNB_OK) return rc;
  if (gb) return launch_colsum(ctx, 1.0f, E, 0.0f, gb, dE.rows, dE.cols, dE.stride);
  return TNB_OK;
}

// The data-parallel gradient GEMM with the reduce-scatter fused into its epilogue: G = X^T E is not written to this rank's memory
// but, row block by row block, straight into the staging slice the OWNING rank keeps for this rank (peer stores over NVLink, 128-byte
// segments from the transposed epilogue tile).  The owner's update kernel (tnb_dp_peer_update with job.pushed = 1) then sums its
// `world` slices from LOCAL memory.  Gpeers[o] = rank o's staging buffer [(world*shard + 1) x stride] as mapped into this process
// (tnb_peer_map); shard = rows_pad / world.  X16 / E16: the bf16 twins in TNB_MATH_BF16 (NULL otherwise).
int tnb_affine_grad_scatter(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, const uint16_t *X16, int ldx16,
                            const uint16_t *E16, int lde16, float *const *Gpeers, int world, int rank, TnbMatrixDim dG, int rows_pad) {
  TNB_ARG(ctx && X && E && Gpeers, "null");
  TNB_ARG(world >= 1 && world <= TNB_MAX_PEERS && rank >= 0 && rank < world, "rank/world");
  TNB_ARG(dX.rows == dE.rows && dG.rows == dX.cols && dG.cols == dE.cols, "dimension mismatch");
  TNB_ARG(rows_pad >= dG.rows && rows_pad % world == 0, "rows_pad must be a multiple of the world size, at least dG.rows");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  for (int r = 0; r < world; r++) { TNB_ARG(Gpeers[r] != nullptr && ((uintptr_t)Gpeers[r] & 15) == 0, "peer staging buffer"); ep.scat[r] = Gpeers[r]; }
  ep.scat_shard = rows_pad / world; ep.scat_rank = rank;
  ep.C = Gpeers[rank]; ep.ldc = dG.stride; ep.alpha = 1.0f; ep.beta = 0.0f;
  if (ctx->math_mode == TNB_MATH_BF16 && X16 && E16)
    return launch_gemm_bf16(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X16, ldx16, E16, lde16, ep);
  return launch_gemm(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X, dX.stride, E, dE.stride, ep);
}

int tnb_affine_update(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *W, TnbMatrixDim dW,
                      float *bias, float *corrW, float *corrb, float lr, float mmt, float wc, int gdf, int n_frames_global) {
  TNB_ARG(ctx && X && E && W && corrW && ((bias && corrb) || (!bias && !corrb)), "null");
  TNB_ARG(dX.rows == dE.rows && dW.rows == dX.cols && dW.cols == dE.cols, "dimension mismatch");
  const int rows = n_frames_global > 0 ? n_frames_global : dX.rows;
  float scale, l2;
  update_scalars(lr, mmt, wc, gdf, rows, &scale, &l2);
  // corrW = X^T E + mmt*corrW ; W += scale*corrW ; W += l2*W   — all in the dW GEMM epilogue
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = corrW; ep.ldc = dW.stride; ep.alpha = 1.0f; ep.beta = mmt;
  ep.W = W; ep.ldw = dW.stride; ep.w_scale = scale; ep.w_l2 = l2; ep.mode = EPI_UPD;
  int rc = launch_gemm(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X, dX.stride, E, dE.stride, ep);
  if (rc != TNB_OK || !bias) return rc;
  // corrb = colsum(E) + mmt*corrb ; b += scale*corrb   (one reduction + one combine/update kernel)
  return launch_colsum_update(ctx, 1.0f, E, mmt, corrb, dE.rows, dE.cols, dE.stride, bias, scale);
}

// ---- CuRbm::RbmUpdate (cuRbm.cc:131-174) ---------------------------------------------------------------------------------------
// One thread per column: the two serial double-precision column sums of the reference's _add_col_sum (cukernels.cu:155-167, the path
// AddColSum takes for more than 256 columns or more than 512 rows), then the three vector updates, in the reference's order:
//   corr = (-a)*colsum(neg) + mmt*corr ; corr = a*colsum(pos) + corr ; bias = corr + bias
__global__ void __launch_bounds__(128) rbm_bias_kernel(const float *__restrict__ pos, const float *__restrict__ neg, int rows, int cols, int stride,
                                                       float *__restrict__ bias, float *__restrict__ corr, float a, float mmt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double sp = 0.0, sn = 0.0;
  for (int r = 0; r < rows; r++) {
    sp += pos[(size_t)r * stride + c];
    sn += neg[(size_t)r * stride + c];
  }
  float k = corr[c];
  k = (float)((double)(-a) * sn + (double)(mmt * k));
  k = (float)((double)a * sp + (double)(1.0f * k));
  corr[c] = k;
  bias[c] = 1.0f * k + 1.0f * bias[c];
}

int tnb_rbm_cd1_update(TnbContext *ctx, const float *pos_vis, const float *neg_vis, TnbMatrixDim dV, const float *pos_hid, const float *neg_hid,
                       TnbMatrixDim dH, float *W, TnbMatrixDim dW, float *corrW, float *vis_bias, float *corr_vb, float *hid_bias, float *corr_hb,
                       float lr, float mmt, float wc) {
  TNB_ARG(ctx && pos_vis && neg_vis && pos_hid && neg_hid && W && corrW && vis_bias && corr_vb && hid_bias && corr_hb, "null");
  TNB_ARG(dV.rows == dH.rows && dV.rows > 0 && dW.rows == dV.cols && dW.cols == dH.cols, "dimension mismatch");
  const float a = lr / (float)dV.rows;  // the reference evaluates lr/N in float (cuRbm.cc:140)
  // corrW = -a * neg_vis^T neg_hid + mmt*corrW                       (cuRbm.cc:143)
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = corrW; ep.ldc = dW.stride; ep.alpha = -a; ep.beta = mmt;
  int rc = launch_gemm(ctx, 'T', 'N', dV.cols, dH.cols, dV.rows, neg_vis, dV.stride, neg_hid, dH.stride, ep);
  if (rc != TNB_OK) return rc;
  // corrW = +a * pos_vis^T pos_hid + corrW ; corrW += -lr*wc*W ; W += corrW      (cuRbm.cc:144-146: one GEMM with the two sweeps in its epilogue)
  ep.alpha = a; ep.beta = 1.0f;
  ep.W = W; ep.ldw = dW.stride; ep.w_scale = 1.0f; ep.w_l2 = 0.0f; ep.c_wdecay = -lr * wc;
  rc = launch_gemm(ctx, 'T', 'N', dV.cols, dH.cols, dV.rows, pos_vis, dV.stride, pos_hid, dH.stride, ep);
  if (rc != TNB_OK) return rc;
  // the two biases (cuRbm.cc:148-154)
  const float *pp[2] = {pos_vis, pos_hid}, *nn[2] = {neg_vis, neg_hid};
  const TnbMatrixDim dd[2] = {dV, dH};
  float *bb[2] = {vis_bias, hid_bias}, *kk[2] = {corr_vb, corr_hb};
  for (int i = 0; i < 2; i++) {
    if (dd[i].rows > 512 || dd[i].cols > 256) {
      rbm_bias_kernel<<<(dd[i].cols + 127) / 128, 128, 0, ctx->stream>>>(pp[i], nn[i], dd[i].rows, dd[i].cols, dd[i].stride, bb[i], kk[i], a, mmt);
      TNB_LAUNCHED(ctx);
    } else {  // small shapes take the reference's float tree reduction: the 1:1 entry points reproduce it
      TnbMatrixDim dv = {1, dd[i].cols, dd[i].cols};
      rc = tnb_add_col_sum(ctx, -a, nn[i], mmt, kk[i], dd[i]);
      if (rc == TNB_OK) rc = tnb_add_col_sum(ctx, a, pp[i], 1.0f, kk[i], dd[i]);
      if (rc == TNB_OK) rc = tnb_add_scaled(ctx, 1.0f, kk[i], 1.0f, bb[i], dv);
      if (rc != TNB_OK) return rc;
    }
  }
  return TNB_OK;
}

int tnb_bias_update_batch(TnbContext *ctx, const TnbBiasJob *jobs, int n) { return tnb_bias_update_batch_on(ctx, TNB_STREAM_COMPUTE, jobs, n); }

int tnb_bias_update_batch_on(TnbContext *ctx, int stream_id, const TnbBiasJob *jobs, int n) {
  TNB_ARG(ctx && (jobs || n == 0), "null");
  cudaStream_t st = stream_of(ctx, stream_id);
  TNB_ARG(st != nullptr, "stream");
  TNB_ARG(n >= 0 && n <= TNB_MAX_BIAS_JOBS, "between 0 and TNB_MAX_BIAS_JOBS jobs per call");
  if (n == 0) return TNB_OK;
  BiasBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  long off = 0;
  int max_cb = 1, max_S = 1, max_cols = 1;
  for (int i = 0; i < n; i++) {
    const TnbBiasJob &q = jobs[i];
    TNB_ARG(q.E && q.corrb, "null");
    TNB_ARG(q.dE.rows >= 0 && q.dE.cols >= 0 && q.dE.stride >= q.dE.cols, "dims");
    const int cb = (q.dE.cols + 127) / 128;
    int S = (6 * ctx->sm_count + cb * n - 1) / (cb * n);  // ~6 CTAs of 256 threads per SM over the whole batch (the kernel is latency-bound)
    if (S > (q.dE.rows + 31) / 32) S = (q.dE.rows + 31) / 32;
    if (S < 1) S = 1;
    float scale, l2;
    update_scalars(q.lr, q.mmt, 0.0f, q.grad_div_frm, q.n_frames_global > 0 ? q.n_frames_global : q.dE.rows, &scale, &l2);
    auto &j = b.j[i];
    j.E = q.E; j.rows = q.dE.rows; j.cols = q.dE.cols; j.stride = q.dE.stride; j.S = S; j.chunk = (q.dE.rows + S - 1) / S;
    j.part_off = off; j.corrb = q.corrb; j.bias = q.bias; j.mmt = q.mmt; j.scale = scale;
    off += (long)S * q.dE.cols;
    if (cb > max_cb) max_cb = cb;
    if (S > max_S) max_S = S;
    if (q.dE.cols > max_cols) max_cols = q.dE.cols;
  }
  const bool side = st != ctx->stream;
  int rc = side ? ensure_vec_scratch_side(ctx, (int)(2 * off)) : ensure_vec_scratch(ctx, (int)(2 * off));  // doubles
  if (rc != TNB_OK) return rc;
  double *part = (double *)(side ? ctx->vec_scratch_side : ctx->vec_scratch);
  colsum_partial_batch_kernel<<<dim3(max_cb, max_S, n), 256, 0, st>>>(b, part);
  TNB_LAUNCHED(ctx);
  colsum_final_batch_kernel<<<dim3((max_cols + 255) / 256, n), 256, 0, st>>>(b, part);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_affine_grad_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *E16, int lde16,
                         const float *E, TnbMatrixDim dE, float *G, TnbMatrixDim dG, float *gb) {
  TNB_ARG(ctx && X16 && E16 && G, "null");
  TNB_ARG(dX.rows == dE.rows && dG.rows == dX.cols && dG.cols == dE.cols, "dimension mismatch");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = G; ep.ldc = dG.stride; ep.alpha = 1.0f; ep.beta = 0.0f;
  int rc = launch_gemm_bf16(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X16, ldx16, E16, lde16, ep);
  if (rc != TNB_OK) return rc;
  if (gb) { TNB_ARG(E, "the bias gradient is summed from the fp32 error"); return launch_colsum(ctx, 1.0f, E, 0.0f, gb, dE.rows, dE.cols, dE.stride); }
  return TNB_OK;
}

int tnb_affine_update_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *E16, int lde16,
                           const float *E, TnbMatrixDim dE, float *W, TnbMatrixDim dW, uint16_t *W16, int ldw16, float *bias,
                           float *corrW, float *corrb, float lr, float mmt, float wc, int gdf, int n_frames_global) {
  TNB_ARG(ctx && X16 && E16 && W && corrW && ((bias && corrb && E) || (!bias && !corrb)), "null");
  TNB_ARG(dX.rows == dE.rows && dW.rows == dX.cols && dW.cols == dE.cols, "dimension mismatch");
  const int rows = n_frames_global > 0 ? n_frames_global : dX.rows;
  float scale, l2;
  update_scalars(lr, mmt, wc, gdf, rows, &scale, &l2);
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = corrW; ep.ldc = dW.stride; ep.alpha = 1.0f; ep.beta = mmt;
  ep.W = W; ep.ldw = dW.stride; ep.w_scale = scale; ep.w_l2 = l2; ep.mode = EPI_UPD;
  ep.W16 = W16; ep.ldw16 = ldw16;
  int rc = launch_gemm_bf16(ctx, 'T', 'N', dX.cols, dE.cols, dX.rows, X16, ldx16, E16, lde16, ep);
  if (rc != TNB_OK || !bias) return rc;
  return launch_colsum_update(ctx, 1.0f, E, mmt, corrb, dE.rows, dE.cols, dE.stride, bias, scale);
}

int tnb_sgd_update_batch(TnbContext *ctx, const TnbSgdJob *jobs, int n) { return tnb_sgd_update_batch_on(ctx, TNB_STREAM_COMPUTE, jobs, n); }

int tnb_sgd_update_batch_on(TnbContext *ctx, int stream_id, const TnbSgdJob *jobs, int n) {
  TNB_ARG(ctx && (jobs || n == 0), "null");
  TNB_ARG(stream_of(ctx, stream_id) != nullptr, "stream");
  TNB_ARG(n >= 0 && n <= TNB_MAX_BIAS_JOBS, "between 0 and TNB_MAX_BIAS_JOBS jobs per call");
  if (n == 0) return TNB_OK;
  SgdBatch b;
  memset(&b, 0, sizeof(b));
  long max_total = 1;
  for (int i = 0; i < n; i++) {
    const TnbSgdJob &q = jobs[i];
    TNB_ARG(q.G && q.W && q.corrW, "null");
    TNB_ARG((q.gb && q.bias && q.corrb) || (!q.gb && !q.bias && !q.corrb), "bias arguments go together");
    TNB_ARG(q.dW.rows >= 0 && q.dW.cols >= 0 && q.dW.stride >= q.dW.cols, "dims");
    TNB_ARG(!q.W16 || ((uintptr_t)q.W16 % 8 == 0 && q.ldw16 % 4 == 0 && q.ldw16 >= q.dW.cols), "bf16 twin alignment");
    float scale, l2;
    update_scalars(q.lr, q.mmt, q.wc, q.grad_div_frm, q.n_frames, &scale, &l2);
    auto &w = b.j[b.n++];
    w.G = q.G; w.W = q.W; w.corr = q.corrW; w.W16 = q.W16; w.ldw16 = q.ldw16;
    w.rows = q.dW.rows; w.cols = q.dW.cols; w.stride = q.dW.stride; w.mmt = q.mmt; w.scale = scale; w.l2 = l2;
    const long total = (long)q.dW.rows * ((q.dW.cols + 3) / 4);
    if (total > max_total) max_total = total;
    if (q.gb) {
      auto &v = b.j[b.n++];
      v.G = q.gb; v.W = q.bias; v.corr = q.corrb; v.W16 = nullptr; v.ldw16 = 0;
      v.rows = 1; v.cols = q.dW.cols; v.stride = q.dW.cols; v.mmt = q.mmt; v.scale = scale; v.l2 = 0.0f;
    }
  }
  long blocks = (max_total + 255) / 256, cap = ((long)ctx->sm_count * 8 + b.n - 1) / b.n;
  // next to the compute stream's GEMMs (whose CTAs take most of an SM's registers): one small CTA per SM fits beside a GEMM CTA,
  // a device-filling grid would keep the next GEMM's CTAs waiting for a free SM
  if (stream_of(ctx, stream_id) != ctx->stream) cap = (ctx->sm_count + b.n - 1) / b.n;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sgd_update_batch_kernel<<<dim3((unsigned)blocks, (unsigned)b.n), 256, 0, stream_of(ctx, stream_id)>>>(b);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_sgd_update(TnbContext *ctx, const float *G, float *W, float *corrW, TnbMatrixDim dW, const float *gb, float *bias,
                   float *corrb, float lr, float mmt, float wc, int gdf, int n_frames) {
  TNB_ARG(ctx && G && W && corrW, "null");
  DIMCHK(dW);
  float scale, l2;
  update_scalars(lr, mmt, wc, gdf, n_frames, &scale, &l2);
  int rc = launch_sgd_update(ctx, ctx->stream, G, W, corrW, dW.rows, dW.cols, dW.stride, mmt, scale, l2);
  if (rc != TNB_OK) return rc;
  if (gb) {
    TNB_ARG(bias && corrb, "null bias");
    return launch_sgd_update(ctx, ctx->stream, gb, bias, corrb, 1, dW.cols, dW.cols, mmt, scale, 0.0f);
  }
  return TNB_OK;
}

}  // extern "C"
