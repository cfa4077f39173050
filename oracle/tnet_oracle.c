/*
 * tnet_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the TNetCu hot path of troylee/nnet-asr (TNet v1.8):
 * the CuBaseLib kernels, CuBiasedLinearity/CuSigmoid/CuSoftmax, CuCrossEntropy /
 * CuMeanSquareError, CuCache (shuffle + bunch slicing), CuRand (Hybrid-Taus),
 * CuRbm CD-1 and CuRecurrent BPTT.  Every function cites the reference
 * file:line it restates (paths relative to /root/reference/src).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library, and only as the checker.  The product (libtnetb200.so) never
 * links, loads or calls it.
 *
 * Parity pin: see oracle/README.md — pinned against the reference's own CPU
 * trainer (oracle/_ref/TNet, run in the build container) and against the
 * reference's GPU trainer (oracle/_ref/TNetCu, run on a B200 through gpurun);
 * the committed fixtures live in tests/golden/.
 *
 * Arithmetic follows the reference kernel by kernel: float storage, float or
 * double intermediates exactly where cukernels.cu has them (C's usual
 * arithmetic conversions reproduce CUDA's for the `T = float` instantiation).
 * Transcendentals use libm (expf/logf/exp), which differ from CUDA's device
 * versions by <= 2 ulp: that is part of the stated tolerance, not of the oracle.
 */
#define _XOPEN_SOURCE 700
#define _DEFAULT_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define IDX(r, c, s) ((size_t)(c) + (size_t)(r) * (size_t)(s))

/* ------------------------------------------------------------------------- */
/* CuMatrix elementwise kernels                                              */
/* ------------------------------------------------------------------------- */

/* cukernels.cu:13-19 _set_const */
void orc_set_const(float *mat, float value, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) mat[IDX(j, i, stride)] = value;
}

/* cukernels.cu:25-31 _apply_log : mat = log(mat)  (float overload) */
void orc_apply_log(float *mat, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) mat[IDX(j, i, stride)] = logf(mat[IDX(j, i, stride)]);
}

/* cukernels.cu:67-73 _scale_cols : mat[j,i] *= scale[i] */
void orc_scale_cols(float *mat, const float *scale, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) mat[IDX(j, i, stride)] *= scale[i];
}

/* cukernels.cu:78-84 _scale_rows : mat[j,i] *= scale[j] */
void orc_scale_rows(float *mat, const float *scale, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) mat[IDX(j, i, stride)] *= scale[j];
}

/* cukernels.cu:89-95 _add_scaled : dst = alpha*A + beta*dst  (all float) */
void orc_add_scaled(float alpha, const float *A, float beta, float *dst, int rows, int cols,
                    int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t x = IDX(j, i, stride);
      float a = alpha * A[x];
      float b = beta * dst[x];
      dst[x] = a + b;
    }
}

/* cukernels.cu:100-119 _add_scaled_row : dst = alpha*row[i] + beta*dst */
void orc_add_scaled_row(float alpha, const float *row, float beta, float *dst, int rows, int cols,
                        int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t x = IDX(j, i, stride);
      float a = alpha * row[i];
      float b = beta * dst[x];
      dst[x] = a + b;
    }
}

/* cukernels.cu:122-128 _mul_elem : mat = mat * A */
void orc_mul_elem(float *mat, const float *A, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) mat[IDX(j, i, stride)] = mat[IDX(j, i, stride)] * A[IDX(j, i, stride)];
}

/* cukernels.cu:133-142 _log_elem : floor at FLT_MIN then log */
void orc_log_elem(float *mat, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t x = IDX(j, i, stride);
      if (mat[x] < FLT_MIN) mat[x] = FLT_MIN;
      mat[x] = logf(mat[x]);
    }
}

/* ------------------------------------------------------------------------- */
/* tree reductions shared by the *_reduce kernels                            */
/* ------------------------------------------------------------------------- */

/* cukernels.cu:278-301 _sum_reduce : pairwise tree in shared memory (float) */
static float sum_reduce_tree(float *buffer, int n) {
  int nTotalThreads = n;
  while (nTotalThreads > 1) {
    int halfPoint = ((1 + nTotalThreads) >> 1);
    for (int t = 0; t < halfPoint; t++) {
      float temp = 0.0f;
      if (t + halfPoint < nTotalThreads) temp = buffer[t + halfPoint];
      buffer[t] += temp;
    }
    nTotalThreads = ((1 + nTotalThreads) >> 1);
  }
  return buffer[0];
}

/* cukernels.cu:249-272 _max_reduce (float, init -1e20) */
static float max_reduce_tree(float *buffer, int n) {
  int nTotalThreads = n;
  while (nTotalThreads > 1) {
    int halfPoint = ((1 + nTotalThreads) >> 1);
    for (int t = 0; t < halfPoint; t++) {
      float temp = -1e20;
      if (t + halfPoint < nTotalThreads) temp = buffer[t + halfPoint];
      if (temp > buffer[t]) buffer[t] = temp;
    }
    nTotalThreads = ((1 + nTotalThreads) >> 1);
  }
  return buffer[0];
}

/* cukernels.cu:424-446 _max_id_reduce : index tree, left slot kept on ties.
 * Upstream, an unpaired slot compares temp = -1e20 and, if its value is below -1e20, copies
 * idx[t+halfPoint] which was never written (undefined behaviour).  Defined inputs (softmax
 * outputs, one-hot targets) never go there; the restatement leaves an unpaired slot unchanged. */
static int max_id_reduce_tree(const float *val, int *idx, int n) {
  int nTotalThreads = n;
  while (nTotalThreads > 1) {
    int halfPoint = ((1 + nTotalThreads) >> 1);
    for (int t = 0; t < halfPoint; t++) {
      if (t + halfPoint >= nTotalThreads) continue;
      float temp = val[idx[t + halfPoint]];
      if (temp > val[idx[t]]) idx[t] = idx[t + halfPoint];
    }
    nTotalThreads = ((1 + nTotalThreads) >> 1);
  }
  return idx[0];
}

/* ------------------------------------------------------------------------- */
/* CuVector::AddColSum                                                        */
/* ------------------------------------------------------------------------- */

/* cukernels.cu:149-164 _add_col_sum : double accumulator, serial over rows */
void orc_add_col_sum_serial(float alpha, const float *mat, float beta, float *vec, int rows,
                            int cols, int stride) {
  for (int i = 0; i < cols; i++) {
    double sum = 0.0;
    for (int k = 0; k < rows; k++) sum += mat[IDX(k, i, stride)];
    vec[i] = alpha * sum + beta * vec[i];
  }
}

/* cukernels.cu:169-187 _add_col_sum_reduce : float tree over rows (rows<=512) */
void orc_add_col_sum_reduce(float alpha, const float *mat, float beta, float *vec, int rows,
                            int cols, int stride) {
  float aux[512];
  for (int i = 0; i < cols; i++) {
    for (int t = 0; t < rows; t++) aux[t] = mat[IDX(t, i, stride)];
    float sum = sum_reduce_tree(aux, rows);
    float a = alpha * sum;
    float b = beta * vec[i];
    vec[i] = a + b;
  }
}

/* cuvector.tcc:164-191 CuVector<float>::AddColSum — shape-dependent dispatch */
void orc_add_col_sum(float alpha, const float *mat, float beta, float *vec, int rows, int cols,
                     int stride) {
  if (rows > 512 || cols > 256)
    orc_add_col_sum_serial(alpha, mat, beta, vec, rows, cols, stride);
  else
    orc_add_col_sum_reduce(alpha, mat, beta, vec, rows, cols, stride);
}

/* ------------------------------------------------------------------------- */
/* CuMath                                                                     */
/* ------------------------------------------------------------------------- */

/* cukernels.cu:194-206 _sigmoid : T res = 1.0 / (1.0 + exp(-x)) ; exp is the
 * float overload, the add/divide are double, the result is rounded to float */
void orc_sigmoid(float *y, const float *x, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t k = IDX(j, i, stride);
      float res = 1.0 / (1.0 + expf(-x[k]));
      y[k] = res;
    }
}

/* cukernels.cu:211-217 _diff_sigmoid : eout = y*(1.0-y)*e  (double product) */
void orc_diff_sigmoid(float *eout, const float *e, const float *y, int rows, int cols,
                      int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t k = IDX(j, i, stride);
      eout[k] = y[k] * (1.0 - y[k]) * e[k];
    }
}

/* cukernels.cu:222-242 _softmax : one thread per row, double max/sum, exp in
 * double because (float - double) promotes */
void orc_softmax_serial(float *y, const float *x, int rows, int cols, int stride) {
  for (int j = 0; j < rows; j++) {
    double max = -1e20;
    double sum = 0.0;
    for (int i = 0; i < cols; i++) {
      if (max < x[IDX(j, i, stride)]) max = x[IDX(j, i, stride)];
      y[IDX(j, i, stride)] = x[IDX(j, i, stride)];
    }
    for (int i = 0; i < cols; i++) {
      y[IDX(j, i, stride)] = exp(y[IDX(j, i, stride)] - max);
      sum += y[IDX(j, i, stride)];
    }
    for (int i = 0; i < cols; i++) y[IDX(j, i, stride)] /= sum;
  }
}

/* cukernels.cu:306-343 _softmax_reduce : cols<=256, float trees */
void orc_softmax_reduce(float *y, const float *x, int rows, int cols, int stride) {
  float row_data[256], aux[256];
  for (int j = 0; j < rows; j++) {
    for (int i = 0; i < cols; i++) row_data[i] = x[IDX(j, i, stride)];
    for (int i = 0; i < cols; i++) aux[i] = row_data[i];
    float max = max_reduce_tree(aux, cols);
    for (int i = 0; i < cols; i++) row_data[i] = expf(row_data[i] - max);
    for (int i = 0; i < cols; i++) aux[i] = row_data[i];
    float sum = sum_reduce_tree(aux, cols);
    for (int i = 0; i < cols; i++) {
      row_data[i] /= sum;
      y[IDX(j, i, stride)] = row_data[i];
    }
  }
}

/* cumath.cc:42-74 CuMath<float>::Softmax dispatch */
void orc_softmax(float *y, const float *x, int rows, int cols, int stride) {
  if (cols > 256)
    orc_softmax_serial(y, x, rows, cols, stride);
  else
    orc_softmax_reduce(y, x, rows, cols, stride);
}

/* cukernels.cu:349-361 _expand ; cumath.cc:118-133 */
void orc_expand(float *y, const float *x, const int *off, int rows, int cols_out, int stride_out,
                int rows_in, int cols_in, int stride_in) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols_out; i++) {
      int src_col = i % cols_in;
      int src_row = j + off[i / cols_in];
      if (src_row < 0) src_row = 0;
      if (src_row >= rows_in) src_row = rows_in - 1;
      y[IDX(j, i, stride_out)] = x[IDX(src_row, src_col, stride_in)];
    }
}

/* cukernels.cu:366-379 _rearrange ; bad index -> +inf */
void orc_rearrange(float *y, const float *x, const int *copy_from, int rows, int cols_out,
                   int stride_out, int cols_in, int stride_in) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols_out; i++) {
      int src_col = copy_from[i];
      if (src_col >= 0 && src_col < cols_in)
        y[IDX(j, i, stride_out)] = x[IDX(j, src_col, stride_in)];
      else
        y[IDX(j, i, stride_out)] = INFINITY;
    }
}

/* cukernels.cu:384-393 _randomize ; cumath.cc:155-174 (rows = perm length) */
void orc_randomize(float *y, const float *x, const int *copy_from, int n_perm, int cols,
                   int stride_out, int stride_in) {
  for (int j = 0; j < n_perm; j++) {
    int src_row = copy_from[j];
    for (int i = 0; i < cols; i++) y[IDX(j, i, stride_out)] = x[IDX(src_row, i, stride_in)];
  }
}

/* cukernels.cu:398-419 _check_class : sequential scan, strict >, init -1e20 */
void orc_check_class_serial(const float *out, const float *des, int *match, int rows, int cols,
                            int stride) {
  for (int i = 0; i < rows; i++) {
    int out_id = -1, des_id = -2;
    float out_max = -1e20, des_max = -1e20;
    for (int k = 0; k < cols; k++) {
      float val = out[IDX(i, k, stride)];
      if (val > out_max) { out_max = val; out_id = k; }
    }
    for (int k = 0; k < cols; k++) {
      float val = des[IDX(i, k, stride)];
      if (val > des_max) { des_max = val; des_id = k; }
    }
    match[i] = ((out_id == des_id) ? 1 : 0);
  }
}

/* cukernels.cu:455-483 _check_class_reduce : index tree (cols<=256) */
void orc_check_class_reduce(const float *out, const float *des, int *match, int rows, int cols,
                            int stride) {
  float value[256];
  int index[256];
  for (int j = 0; j < rows; j++) {
    for (int t = 0; t < cols; t++) { value[t] = out[IDX(j, t, stride)]; index[t] = t; }
    int out_max = max_id_reduce_tree(value, index, cols);
    for (int t = 0; t < cols; t++) { value[t] = des[IDX(j, t, stride)]; index[t] = t; }
    int des_max = max_id_reduce_tree(value, index, cols);
    match[j] = ((out_max == des_max) ? 1 : 0);
  }
}

/* cumath.cc:178-206 CuMath<float>::CheckClass dispatch */
void orc_check_class(const float *out, const float *des, int *match, int rows, int cols,
                     int stride) {
  if (cols > 256)
    orc_check_class_serial(out, des, match, rows, cols, stride);
  else
    orc_check_class_reduce(out, des, match, rows, cols, stride);
}

/* ------------------------------------------------------------------------- */
/* GEMM  (cumatrix.tcc:335-370 CuMatrix<float>::Gemm -> legacy cublasSgemm)   */
/* Row-major C[m x n] = alpha*op(A)*op(B) + beta*C.  cuBLAS leaves the        */
/* summation order unspecified; the oracle accumulates k ascending, in float  */
/* (acc_double=0, what an fp32 SGEMM does) or in double (acc_double=1, the    */
/* exact product rounded once — the tighter anchor for tolerance tests).      */
/* ------------------------------------------------------------------------- */
void orc_gemm(char transa, char transb, int m, int n, int k, float alpha, const float *A, int lda,
              const float *B, int ldb, float beta, float *C, int ldc, int acc_double) {
  int ta = (transa == 'T' || transa == 't');
  int tb = (transb == 'T' || transb == 't');
  if (!tb) {
    /* op(B) = B is [k x n] row-major: walk it row by row with one accumulator per output column.  Every C[i][j] still sums its
     * products in the order p = 0..k-1 (one sequential chain per element), so the result is the plain triple loop's bit for bit;
     * only the memory access pattern differs (the full-size parity cases of tests/ finish in seconds instead of minutes). */
#pragma omp parallel
    {
      double *accd = acc_double ? (double *)malloc(sizeof(double) * (size_t)n) : NULL;
      float *accf = acc_double ? NULL : (float *)malloc(sizeof(float) * (size_t)n);
#pragma omp for schedule(static)
      for (int i = 0; i < m; i++) {
        if (acc_double) {
          for (int j = 0; j < n; j++) accd[j] = 0.0;
          for (int p = 0; p < k; p++) {
            const double a = (double)(ta ? A[IDX(p, i, lda)] : A[IDX(i, p, lda)]);
            const float *brow = B + (size_t)p * (size_t)ldb;
            for (int j = 0; j < n; j++) accd[j] += a * (double)brow[j];
          }
        } else {
          for (int j = 0; j < n; j++) accf[j] = 0.0f;
          for (int p = 0; p < k; p++) {
            const float a = ta ? A[IDX(p, i, lda)] : A[IDX(i, p, lda)];
            const float *brow = B + (size_t)p * (size_t)ldb;
            for (int j = 0; j < n; j++) accf[j] = fmaf(a, brow[j], accf[j]);
          }
        }
        for (int j = 0; j < n; j++) {
          float r = acc_double ? (float)accd[j] : accf[j];
          float old = (beta == 0.0f) ? 0.0f : beta * C[IDX(i, j, ldc)];
          C[IDX(i, j, ldc)] = alpha * r + old;
        }
      }
      free(accd);
      free(accf);
    }
    return;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < n; j++) {
      float r;
      if (acc_double) {
        double acc = 0.0;
        for (int p = 0; p < k; p++) {
          float a = ta ? A[IDX(p, i, lda)] : A[IDX(i, p, lda)];
          float b = B[IDX(j, p, ldb)];
          acc += (double)a * (double)b;
        }
        r = (float)acc;
      } else {
        float acc = 0.0f;
        for (int p = 0; p < k; p++) {
          float a = ta ? A[IDX(p, i, lda)] : A[IDX(i, p, lda)];
          float b = B[IDX(j, p, ldb)];
          acc = fmaf(a, b, acc);
        }
        r = acc;
      }
      float old = (beta == 0.0f) ? 0.0f : beta * C[IDX(i, j, ldc)];
      C[IDX(i, j, ldc)] = alpha * r + old;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* CuBiasedLinearity  (cuBiasedLinearity.cc)                                  */
/* W is [nin x nout] row-major in memory (stored transposed on disk :70-78)   */
/* ------------------------------------------------------------------------- */

/* cuBiasedLinearity.cc:11-16 PropagateFnc */
void orc_affine_fwd(const float *X, int ldx, const float *W, int ldw, const float *bias, float *Y,
                    int ldy, int rows, int nin, int nout, int acc_double) {
  orc_add_scaled_row(1.0f, bias, 0.0f, Y, rows, nout, ldy);
  orc_gemm('N', 'N', rows, nout, nin, 1.0f, X, ldx, W, ldw, 1.0f, Y, ldy, acc_double);
}

/* cuBiasedLinearity.cc:20-25 BackpropagateFnc : Eprev = E * W^T */
void orc_affine_bwd(const float *E, int lde, const float *W, int ldw, float *Eprev, int ldp,
                    int rows, int nin, int nout, int acc_double) {
  orc_gemm('N', 'T', rows, nin, nout, 1.0f, E, lde, W, ldw, 0.0f, Eprev, ldp, acc_double);
}

/* cuBiasedLinearity.cc:44-64 Update (the "#if 1 new implementation") */
void orc_affine_update(const float *X, int ldx, const float *E, int lde, float *W, int ldw,
                       float *bias, float *corrW, int ldc, float *corrb, int rows, int nin,
                       int nout, float lr, float mmt, float wc, int grad_div_frm, int acc_double) {
  float N = 1;
  if (grad_div_frm) N = (float)rows;
  float mmt_gain = (float)(1.0 / (1.0 - mmt));
  N *= mmt_gain;
  orc_gemm('T', 'N', nin, nout, rows, 1.0f, X, ldx, E, lde, mmt, corrW, ldc, acc_double);
  orc_add_col_sum(1.0f, E, mmt, corrb, rows, nout, lde);
  orc_add_scaled(-lr / N, corrW, 1.0f, W, nin, nout, ldw); /* needs ldw==ldc */
  orc_add_scaled(-lr / N, corrb, 1.0f, bias, 1, nout, nout);
  float L2_decay = -lr * wc * (grad_div_frm ? 1.0 : rows);
  orc_add_scaled(L2_decay, W, 1.0f, W, nin, nout, ldw);
}

/* ------------------------------------------------------------------------- */
/* CuMath::OffsetGemm (cumath.cc:210-244): GEMM on sub-blocks addressed by a    */
/* column offset into each operand; extents clipped to op(B) and C.  With the  */
/* row-major layout an offset is a pointer bump, the leading dimensions stay.  */
/* ------------------------------------------------------------------------- */
static void offset_gemm(char ta, char tb, int rows, int cols, int k, float alpha, const float *A,
                        int lda, int offA, const float *B, int ldb, int offB, float beta, float *C,
                        int ldc, int offC, int acc_double) {
  orc_gemm(ta, tb, rows, cols, k, alpha, A + offA, lda, B + offB, ldb, beta, C + offC, ldc, acc_double);
}

/* ------------------------------------------------------------------------- */
/* CuSharedLinearity (cuSharedLinearity.cc): W [bi x bo] (bi = nin/K,          */
/* bo = nout/K) shared by the K column blocks of the input; bias [bo]          */
/* ------------------------------------------------------------------------- */

/* cuSharedLinearity.cc:9-25 PropagateFnc; VecExpand = cumath.cc:366-384 (bias tiled K times) */
void orc_shared_fwd(const float *X, int ldx, const float *W, const float *bias, float *Y, int ldy,
                    int rows, int bi, int bo, int K, int acc_double) {
  float *bexp = (float *)malloc(sizeof(float) * (size_t)bo * K);
  for (int i = 0; i < bo * K; i++) bexp[i] = bias[i % bo];
  orc_add_scaled_row(1.0f, bexp, 0.0f, Y, rows, bo * K, ldy);
  for (int i = 0; i < K; i++)
    offset_gemm('N', 'N', rows, bo, bi, 1.0f, X, ldx, i * bi, W, bo, 0, 1.0f, Y, ldy, i * bo, acc_double);
  free(bexp);
}

/* cuSharedLinearity.cc:28-36 BackpropagateFnc */
void orc_shared_bwd(const float *E, int lde, const float *W, float *Eprev, int ldp, int rows, int bi,
                    int bo, int K, int acc_double) {
  for (int i = 0; i < K; i++)
    offset_gemm('N', 'T', rows, bi, bo, 1.0f, E, lde, i * bo, W, bo, 0, 0.0f, Eprev, ldp, i * bi, acc_double);
}

/* cuSharedLinearity.cc:62-93 Update ("#if 1 new implementation"); VecAddColSum = cumath.cc:388-404,
 * which always launches the serial double-accumulator kernel on the [K x bo] view */
void orc_shared_update(const float *X, int ldx, const float *E, int lde, float *W, float *bias,
                       float *corrW, float *corrb, int rows, int bi, int bo, int K, float lr,
                       float mmt, float wc, int grad_div_frm, int acc_double) {
  float N = 1;
  if (grad_div_frm) N = (float)rows;
  float mmt_gain = (float)(1.0 / (1.0 - mmt));
  N *= mmt_gain;
  N *= (float)K;
  for (int i = 0; i < K; i++)
    offset_gemm('T', 'N', bi, bo, rows, 1.0f, X, ldx, i * bi, E, lde, i * bo, (i == 0) ? mmt : 1.0f,
                corrW, bo, 0, acc_double);
  float *cexp = (float *)calloc((size_t)bo * K, sizeof(float));
  orc_add_col_sum(1.0f, E, 0.0f, cexp, rows, bo * K, lde);
  orc_add_col_sum_serial(1.0f, cexp, mmt, corrb, K, bo, bo);
  free(cexp);
  orc_add_scaled(-lr / N, corrW, 1.0f, W, bi, bo, bo);
  orc_add_scaled(-lr / N, corrb, 1.0f, bias, 1, bo, bo);
  orc_add_scaled(-lr * wc, W, 1.0f, W, bi, bo, bo);
}

/* ------------------------------------------------------------------------- */
/* CuDiscreteLinearity (cuDiscreteLinearity.cc): block-diagonal affine layer;  */
/* block i is Wb[i] [bin[i] x bout[i]], one bias over all outputs             */
/* ------------------------------------------------------------------------- */

/* cuDiscreteLinearity.cc:7-25 PropagateFnc */
void orc_discrete_fwd(const float *X, int ldx, float *const *Wb, const int *bin, const int *bout,
                      int nblocks, const float *bias, float *Y, int ldy, int rows, int nout,
                      int acc_double) {
  orc_add_scaled_row(1.0f, bias, 0.0f, Y, rows, nout, ldy);
  int oi = 0, oo = 0;
  for (int i = 0; i < nblocks; i++) {
    offset_gemm('N', 'N', rows, bout[i], bin[i], 1.0f, X, ldx, oi, Wb[i], bout[i], 0, 1.0f, Y, ldy, oo, acc_double);
    oi += bin[i]; oo += bout[i];
  }
}

/* cuDiscreteLinearity.cc:28-41 BackpropagateFnc */
void orc_discrete_bwd(const float *E, int lde, float *const *Wb, const int *bin, const int *bout,
                      int nblocks, float *Eprev, int ldp, int rows, int acc_double) {
  int oi = 0, oo = 0;
  for (int i = 0; i < nblocks; i++) {
    offset_gemm('N', 'T', rows, bin[i], bout[i], 1.0f, E, lde, oi, Wb[i], bout[i], 0, 0.0f, Eprev, ldp, oo, acc_double);
    oi += bout[i]; oo += bin[i];
  }
}

/* cuDiscreteLinearity.cc:44-76 Update */
void orc_discrete_update(const float *X, int ldx, const float *E, int lde, float **Wb, float **corrWb,
                         const int *bin, const int *bout, int nblocks, float *bias, float *corrb,
                         int rows, int nout, float lr, float mmt, float wc, int grad_div_frm,
                         int acc_double) {
  float N = 1;
  if (grad_div_frm) N = (float)rows;
  float mmt_gain = (float)(1.0 / (1.0 - mmt));
  N *= mmt_gain;
  int oi = 0, oo = 0;
  for (int i = 0; i < nblocks; i++) {
    offset_gemm('T', 'N', bin[i], bout[i], rows, 1.0f, X, ldx, oi, E, lde, oo, mmt, corrWb[i], bout[i], 0, acc_double);
    oi += bin[i]; oo += bout[i];
  }
  for (int i = 0; i < nblocks; i++) {
    orc_add_scaled(-lr / N, corrWb[i], 1.0f, Wb[i], bin[i], bout[i], bout[i]);
    orc_add_scaled(-lr * wc, Wb[i], 1.0f, Wb[i], bin[i], bout[i], bout[i]);
  }
  orc_add_col_sum(1.0f, E, mmt, corrb, rows, nout, lde);
  orc_add_scaled(-lr / N, corrb, 1.0f, bias, 1, nout, nout);
}

/* ------------------------------------------------------------------------- */
/* Objective functions (cuObjectiveFunction.cc)                               */
/* ------------------------------------------------------------------------- */
typedef struct {
  double error;        /* mError   */
  long long frames;    /* mFrames  */
  long long correct;   /* mCorrect */
} OrcObjStats;

/* cuObjectiveFunction.cc:48-84 CuCrossEntropy::Evaluate */
void orc_xent_evaluate(const float *Y, const float *T, float *Err, int rows, int cols, int stride,
                       OrcObjStats *st) {
  /* err = y ; err = -1*t + 1*err */
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) Err[IDX(j, i, stride)] = Y[IDX(j, i, stride)];
  orc_add_scaled(-1.0f, T, 1.0f, Err, rows, cols, stride);
  /* classification */
  int *match = (int *)malloc(sizeof(int) * (size_t)rows);
  orc_check_class(Y, T, match, rows, cols, stride);
  int msum = 0;
  for (int j = 0; j < rows; j++) msum += match[j];
  st->correct += msum;
  free(match);
  /* xent */
  float *aux = (float *)malloc(sizeof(float) * (size_t)rows * (size_t)stride);
  memcpy(aux, Y, sizeof(float) * (size_t)rows * (size_t)stride);
  orc_log_elem(aux, rows, cols, stride);
  orc_mul_elem(aux, T, rows, cols, stride);
  float *vec = (float *)calloc((size_t)cols, sizeof(float));
  orc_add_col_sum(-1.0f, aux, 0.0f, vec, rows, cols, stride);
  double s = 0.0; /* Vector<float>::Sum, KaldiLib/Vector.tcc:266-278 (double acc, float result) */
  for (int i = 0; i < cols; i++) s += vec[i];
  st->error += (float)s;
  st->frames += rows;
  free(aux);
  free(vec);
}

/* cuObjectiveFunction.cc:26-46 CuMeanSquareError::Evaluate (no 1/2 factor) */
void orc_mse_evaluate(const float *Y, const float *T, float *Err, int rows, int cols, int stride,
                      OrcObjStats *st) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) Err[IDX(j, i, stride)] = Y[IDX(j, i, stride)];
  orc_add_scaled(-1.0f, T, 1.0f, Err, rows, cols, stride);
  float *aux = (float *)malloc(sizeof(float) * (size_t)rows * (size_t)stride);
  memcpy(aux, Err, sizeof(float) * (size_t)rows * (size_t)stride);
  orc_mul_elem(aux, aux, rows, cols, stride);
  float *vec = (float *)calloc((size_t)cols, sizeof(float));
  orc_add_col_sum(1.0f, aux, 0.0f, vec, rows, cols, stride);
  double s = 0.0;
  for (int i = 0; i < cols; i++) s += vec[i];
  st->error += (float)s;
  st->frames += rows;
  free(aux);
  free(vec);
}

/* ------------------------------------------------------------------------- */
/* CuCache permutation  (cuCache.cc:124-153, cuCache.h:47-48)                 */
/* std::random_shuffle(p, p+n, rng) with rng(k) = lrand48() % k; libstdc++    */
/* (bits/stl_algo.h, __random_shuffle with generator):                        */
/*   for (i = first+1; i != last; ++i) { j = first + rng((i-first)+1);        */
/*                                       if (i != j) iter_swap(i, j); }       */
/* ------------------------------------------------------------------------- */
void orc_srand48(long seed) { srand48(seed); }
long orc_lrand48(void) { return lrand48(); }

void orc_shuffle_perm(int *perm, int n) {
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int i = 1; i < n; i++) {
    int j = (int)(lrand48() % (long)(i + 1));
    if (i != j) { int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
  }
}

/* ---- CuCache state machine (cuCache.cc:21-200) on host buffers ---------- */
typedef struct {
  int state; /* 0 EMPTY 1 INTAKE 2 FULL 3 EXHAUST */
  size_t intake, exhaust, cachesize, bunchsize;
  int discarded, randomized;
  int fdim, ddim;
  float *feat, *des, *feat_r, *des_r;
  float *feat_left, *des_left;
  int left_rows;
  int *last_perm; int last_perm_n;
} OrcCache;

OrcCache *orc_cache_new(size_t cachesize, size_t bunchsize) {
  if (cachesize % bunchsize != 0) return NULL; /* cuCache.cc:25-27 Error */
  OrcCache *c = (OrcCache *)calloc(1, sizeof(OrcCache));
  c->cachesize = cachesize; c->bunchsize = bunchsize;
  return c;
}
void orc_cache_free(OrcCache *c) {
  if (!c) return;
  free(c->feat); free(c->des); free(c->feat_r); free(c->des_r);
  free(c->feat_left); free(c->des_left); free(c->last_perm); free(c);
}
int orc_cache_full(OrcCache *c) { return c->state == 2; }
int orc_cache_empty(OrcCache *c) { return c->state == 0 || c->intake < c->bunchsize; }
int orc_cache_discarded(OrcCache *c) { return c->discarded; }
int orc_cache_intake(OrcCache *c) { return (int)c->intake; }
const int *orc_cache_last_perm(OrcCache *c, int *n) { *n = c->last_perm_n; return c->last_perm; }

/* cuCache.cc:41-120 AddData (dense row-major inputs, ld = dim) */
void orc_cache_add(OrcCache *c, const float *F, const float *D, int rows, int fdim, int ddim) {
  if (!c->feat) {
    c->fdim = fdim; c->ddim = ddim;
    c->feat = (float *)calloc(c->cachesize * fdim, sizeof(float));
    c->des = (float *)calloc(c->cachesize * ddim, sizeof(float));
    c->feat_r = (float *)calloc(c->cachesize * fdim, sizeof(float));
    c->des_r = (float *)calloc(c->cachesize * ddim, sizeof(float));
  }
  if (c->state == 0) {
    c->state = 1; c->intake = 0;
    int leftover = c->left_rows;
    if ((size_t)leftover > c->cachesize) leftover = (int)c->cachesize;
    if (leftover > 0) {
      memcpy(c->feat, c->feat_left, sizeof(float) * (size_t)leftover * fdim);
      memcpy(c->des, c->des_left, sizeof(float) * (size_t)leftover * ddim);
      free(c->feat_left); free(c->des_left);
      c->feat_left = c->des_left = NULL; c->left_rows = 0;
      c->intake += leftover;
    }
  }
  int cache_space = (int)(c->cachesize - c->intake);
  int fill_rows = cache_space < rows ? cache_space : rows;
  int leftover = rows - fill_rows;
  memcpy(c->feat + c->intake * fdim, F, sizeof(float) * (size_t)fill_rows * fdim);
  memcpy(c->des + c->intake * ddim, D, sizeof(float) * (size_t)fill_rows * ddim);
  if (leftover > 0) {
    free(c->feat_left); free(c->des_left);
    c->feat_left = (float *)malloc(sizeof(float) * (size_t)leftover * fdim);
    c->des_left = (float *)malloc(sizeof(float) * (size_t)leftover * ddim);
    memcpy(c->feat_left, F + (size_t)fill_rows * fdim, sizeof(float) * (size_t)leftover * fdim);
    memcpy(c->des_left, D + (size_t)fill_rows * ddim, sizeof(float) * (size_t)leftover * ddim);
    c->left_rows = leftover;
  }
  c->intake += fill_rows;
  if (c->intake == c->cachesize) c->state = 2;
}

/* cuCache.cc:124-153 Randomize */
void orc_cache_randomize(OrcCache *c) {
  int n = (int)c->intake;
  free(c->last_perm);
  c->last_perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  c->last_perm_n = n;
  orc_shuffle_perm(c->last_perm, n);
  orc_randomize(c->feat_r, c->feat, c->last_perm, n, c->fdim, c->fdim, c->fdim);
  orc_randomize(c->des_r, c->des, c->last_perm, n, c->ddim, c->ddim, c->ddim);
  c->randomized = 1;
}

/* cuCache.cc:155-200 GetBunch ; returns 0 ok, -1 if cache EMPTY (reference: Error) */
int orc_cache_get_bunch(OrcCache *c, float *F, float *D) {
  if (c->state == 0) return -1;
  if (c->state == 2) { c->state = 3; c->exhaust = 0; }
  if (c->state == 1) { c->state = 3; c->exhaust = 0; }
  const float *fs = c->randomized ? c->feat_r : c->feat;
  const float *ds = c->randomized ? c->des_r : c->des;
  memcpy(F, fs + c->exhaust * c->fdim, sizeof(float) * c->bunchsize * c->fdim);
  memcpy(D, ds + c->exhaust * c->ddim, sizeof(float) * c->bunchsize * c->ddim);
  c->exhaust += c->bunchsize;
  /* NB: size_t arithmetic as in the reference (mIntakePos-mBunchsize unsigned) */
  if (c->exhaust > c->intake - c->bunchsize) {
    c->discarded += (int)(c->intake - c->exhaust);
    c->state = 0;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* CuRand  (curandkernels.cu:15-103, curand.tcc:13-60)                        */
/* ------------------------------------------------------------------------- */
static unsigned taus_step(unsigned *z, int S1, int S2, int S3, unsigned M) {
  unsigned b = (((*z << S1) ^ *z) >> S2);
  return *z = (((*z & M) << S3) ^ b);
}
static unsigned lcg_step(unsigned *z, unsigned A, unsigned C) { return *z = (A * *z + C); }

/* curandkernels.cu:30-46 HybridTaus<float> : the product is double*unsigned,
 * rounded to float on assignment; rejection loop on (0,1) */
static float hybrid_taus(unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4) {
  float randval;
  do {
    randval = 2.3283064365387e-10 *
              (taus_step(z1, 13, 19, 12, 4294967294U) ^ taus_step(z2, 2, 25, 4, 4294967288U) ^
               taus_step(z3, 3, 11, 17, 4294967280U) ^ lcg_step(z4, 1664525, 1013904223U));
  } while (!(randval > 0.0 && randval < 1.0));
  return randval;
}

/* curand.tcc:42-49 SeedRandom : value = lrand48() until > 128, row-major order;
 * curand.tcc:13-24 SeedGpu fills z1,z2,z3,z4 one whole matrix after another */
void orc_rand_seed(unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, int rows, int cols,
                   int stride) {
  unsigned *zs[4] = {z1, z2, z3, z4};
  for (int q = 0; q < 4; q++)
    for (int j = 0; j < rows; j++)
      for (int i = 0; i < cols; i++) {
        unsigned value = 0;
        while (value <= 128) value = (unsigned)lrand48();
        zs[q][IDX(j, i, stride)] = value;
      }
}

/* curandkernels.cu:50-59 _rand */
void orc_rand(float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, int rows,
              int cols, int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t k = IDX(j, i, stride);
      mat[k] = hybrid_taus(&z1[k], &z2[k], &z3[k], &z4[k]);
    }
}

/* curandkernels.cu:72-95 BoxMuller/_gauss_rand : r*sin(theta), float math
 * (sqrt/log/sin resolve to the float overloads for T=float; -2.0*log(u0) is
 * double*float -> double, sqrt(double) -> double, assigned to T r = float) */
void orc_gauss_rand(float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, int rows,
                    int cols, int stride) {
  const float M_2PI = 6.283185307179586476925286766558;
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t k = IDX(j, i, stride);
      float u0 = hybrid_taus(&z1[k], &z2[k], &z3[k], &z4[k]);
      float u1 = hybrid_taus(&z1[k], &z2[k], &z3[k], &z4[k]);
      float r = sqrt(-2.0 * logf(u0));
      float theta = M_2PI * u1;
      mat[k] = r * sinf(theta);
    }
}

/* curandkernels.cu:99-108 _binarize_probs */
void orc_binarize_probs(float *states, const float *probs, const float *rnd, int rows, int cols,
                        int stride) {
  for (int j = 0; j < rows; j++)
    for (int i = 0; i < cols; i++) {
      size_t k = IDX(j, i, stride);
      states[k] = ((probs[k] > rnd[k]) ? 1.0 : 0.0);
    }
}

/* ------------------------------------------------------------------------- */
/* MLP network trainer: CuNetwork::Propagate / Backpropagate                  */
/* (cuNetwork.h:137-194) + TNetCu.cc:427-441 bunch loop                       */
/* ------------------------------------------------------------------------- */
enum { ORC_AFFINE = 0, ORC_SIGMOID = 1, ORC_SOFTMAX = 2, ORC_SHARED = 3, ORC_DISCRETE = 4 };

typedef struct {
  int type, nin, nout;
  float *W, *b, *corrW, *corrb; /* affine: W [nin x nout]; shared: W [nin/K x nout/K], b [nout/K]; discrete: b only */
  int K;                        /* shared: instances; discrete: blocks */
  float **Wb, **corrWb;         /* discrete: block i [bin[i] x bout[i]] */
  int *bin, *bout;
  float lr, mmt, wc;
  int gdf;
  float *out, *eout; /* [rows x nout], [rows x nin] */
  int rows_alloc;
} OrcLayer;

typedef struct {
  OrcLayer *L;
  int n;
  int stopper; /* index of mpPropagErrorStopper or -1 */
  int acc_double;
  OrcObjStats st;
  float *err;
  int err_rows;
} OrcNet;

OrcNet *orc_net_new(int acc_double) {
  OrcNet *h = (OrcNet *)calloc(1, sizeof(OrcNet));
  h->stopper = -1; h->acc_double = acc_double;
  return h;
}
static OrcLayer *net_push(OrcNet *h, int type, int nin, int nout) {
  h->L = (OrcLayer *)realloc(h->L, sizeof(OrcLayer) * (size_t)(h->n + 1));
  OrcLayer *l = &h->L[h->n++];
  memset(l, 0, sizeof(*l));
  l->type = type; l->nin = nin; l->nout = nout; l->gdf = 1;
  return l;
}
/* Wt is the ON-DISK layout [nout x nin] (cuBiasedLinearity.cc:70-78) */
void orc_net_add_affine(OrcNet *h, int nin, int nout, const float *Wt, const float *b) {
  OrcLayer *l = net_push(h, ORC_AFFINE, nin, nout);
  l->W = (float *)malloc(sizeof(float) * (size_t)nin * nout);
  l->b = (float *)malloc(sizeof(float) * (size_t)nout);
  l->corrW = (float *)calloc((size_t)nin * nout, sizeof(float));
  l->corrb = (float *)calloc((size_t)nout, sizeof(float));
  for (int o = 0; o < nout; o++)
    for (int i = 0; i < nin; i++) l->W[IDX(i, o, nout)] = Wt[IDX(o, i, nin)];
  memcpy(l->b, b, sizeof(float) * (size_t)nout);
}
/* Wt [nout/K x nin/K] on-disk layout (cuSharedLinearity.cc:112-121) */
void orc_net_add_shared(OrcNet *h, int nin, int nout, int K, const float *Wt, const float *b) {
  OrcLayer *l = net_push(h, ORC_SHARED, nin, nout);
  int bi = nin / K, bo = nout / K;
  l->K = K;
  l->W = (float *)malloc(sizeof(float) * (size_t)bi * bo);
  l->b = (float *)malloc(sizeof(float) * (size_t)bo);
  l->corrW = (float *)calloc((size_t)bi * bo, sizeof(float));
  l->corrb = (float *)calloc((size_t)bo, sizeof(float));
  for (int o = 0; o < bo; o++)
    for (int i = 0; i < bi; i++) l->W[IDX(i, o, bo)] = Wt[IDX(o, i, bi)];
  memcpy(l->b, b, sizeof(float) * (size_t)bo);
}
/* blocks concatenated in Wt_all, each in on-disk layout [bout[i] x bin[i]] (cuDiscreteLinearity.cc:91-105) */
void orc_net_add_discrete(OrcNet *h, int nblocks, const int *bin, const int *bout, const float *Wt_all,
                          const float *b) {
  int nin = 0, nout = 0;
  for (int i = 0; i < nblocks; i++) { nin += bin[i]; nout += bout[i]; }
  OrcLayer *l = net_push(h, ORC_DISCRETE, nin, nout);
  l->K = nblocks;
  l->bin = (int *)malloc(sizeof(int) * (size_t)nblocks);
  l->bout = (int *)malloc(sizeof(int) * (size_t)nblocks);
  l->Wb = (float **)calloc((size_t)nblocks, sizeof(float *));
  l->corrWb = (float **)calloc((size_t)nblocks, sizeof(float *));
  const float *src = Wt_all;
  for (int k = 0; k < nblocks; k++) {
    l->bin[k] = bin[k]; l->bout[k] = bout[k];
    l->Wb[k] = (float *)malloc(sizeof(float) * (size_t)bin[k] * bout[k]);
    l->corrWb[k] = (float *)calloc((size_t)bin[k] * bout[k], sizeof(float));
    for (int o = 0; o < bout[k]; o++)
      for (int i = 0; i < bin[k]; i++) l->Wb[k][IDX(i, o, bout[k])] = src[IDX(o, i, bin[k])];
    src += (size_t)bin[k] * bout[k];
  }
  l->b = (float *)malloc(sizeof(float) * (size_t)nout);
  l->corrb = (float *)calloc((size_t)nout, sizeof(float));
  memcpy(l->b, b, sizeof(float) * (size_t)nout);
}
void orc_net_add_sigmoid(OrcNet *h, int n) { net_push(h, ORC_SIGMOID, n, n); }
void orc_net_add_softmax(OrcNet *h, int n) { net_push(h, ORC_SOFTMAX, n, n); }

static int layer_updatable(const OrcLayer *l) {
  return l->type == ORC_AFFINE || l->type == ORC_SHARED || l->type == ORC_DISCRETE;
}

/* cuNetwork.cc:80-135 SetLearnRate (factors==NULL => scale 1) + SetMomentum etc. */
void orc_net_set_hyper(OrcNet *h, float lr, const float *factors, int nfactors, float mmt, float wc,
                       int gdf) {
  int k = 0, given = 0;
  h->stopper = -1;
  for (int i = 0; i < h->n; i++) {
    OrcLayer *l = &h->L[i];
    if (!layer_updatable(l)) continue;
    float scale = 1.0f;
    if (factors && k < nfactors) scale = factors[k];
    k++;
    l->lr = lr * scale; l->mmt = mmt; l->wc = wc; l->gdf = gdf;
    if (!given && (lr * scale > 0.0)) { h->stopper = i; given = 1; }
  }
}

static void layer_bufs(OrcLayer *l, int rows) {
  if (l->rows_alloc == rows) return;
  free(l->out); free(l->eout);
  l->out = (float *)calloc((size_t)rows * l->nout, sizeof(float));
  l->eout = (float *)calloc((size_t)rows * l->nin, sizeof(float));
  l->rows_alloc = rows;
}

/* cuNetwork.h:137-165 ; out may be NULL */
void orc_net_propagate(OrcNet *h, const float *X, int rows, float *out) {
  const float *in = X;
  for (int i = 0; i < h->n; i++) {
    OrcLayer *l = &h->L[i];
    layer_bufs(l, rows);
    switch (l->type) {
      case ORC_AFFINE:
        orc_affine_fwd(in, l->nin, l->W, l->nout, l->b, l->out, l->nout, rows, l->nin, l->nout,
                       h->acc_double);
        break;
      case ORC_SHARED:
        orc_shared_fwd(in, l->nin, l->W, l->b, l->out, l->nout, rows, l->nin / l->K, l->nout / l->K, l->K,
                       h->acc_double);
        break;
      case ORC_DISCRETE:
        orc_discrete_fwd(in, l->nin, l->Wb, l->bin, l->bout, l->K, l->b, l->out, l->nout, rows, l->nout,
                         h->acc_double);
        break;
      case ORC_SIGMOID: orc_sigmoid(l->out, in, rows, l->nout, l->nout); break;
      case ORC_SOFTMAX: orc_softmax(l->out, in, rows, l->nout, l->nout); break;
    }
    in = l->out;
  }
  if (out) memcpy(out, in, sizeof(float) * (size_t)rows * h->L[h->n - 1].nout);
}

/* cuNetwork.h:170-194 */
void orc_net_backpropagate(OrcNet *h, const float *X, const float *globerr, int rows) {
  const float *ein = globerr;
  for (int i = h->n - 1; i >= 0; i--) {
    OrcLayer *l = &h->L[i];
    const float *in = (i == 0) ? X : h->L[i - 1].out;
    if (i != h->stopper) {
      switch (l->type) {
        case ORC_AFFINE:
          orc_affine_bwd(ein, l->nout, l->W, l->nout, l->eout, l->nin, rows, l->nin, l->nout,
                         h->acc_double);
          break;
        case ORC_SHARED:
          orc_shared_bwd(ein, l->nout, l->W, l->eout, l->nin, rows, l->nin / l->K, l->nout / l->K, l->K,
                         h->acc_double);
          break;
        case ORC_DISCRETE:
          orc_discrete_bwd(ein, l->nout, l->Wb, l->bin, l->bout, l->K, l->eout, l->nin, rows, h->acc_double);
          break;
        case ORC_SIGMOID: /* cuActivation.cc:17-22 */
          orc_diff_sigmoid(l->eout, ein, l->out, rows, l->nout, l->nout);
          break;
        case ORC_SOFTMAX: /* cuActivation.cc:35-41 identity */
          memcpy(l->eout, ein, sizeof(float) * (size_t)rows * l->nout);
          break;
      }
    }
    if (l->type == ORC_AFFINE && l->lr > 0.0f) {
      orc_affine_update(in, l->nin, ein, l->nout, l->W, l->nout, l->b, l->corrW, l->nout, l->corrb,
                        rows, l->nin, l->nout, l->lr, l->mmt, l->wc, l->gdf, h->acc_double);
    }
    if (l->type == ORC_SHARED && l->lr > 0.0f)
      orc_shared_update(in, l->nin, ein, l->nout, l->W, l->b, l->corrW, l->corrb, rows, l->nin / l->K,
                        l->nout / l->K, l->K, l->lr, l->mmt, l->wc, l->gdf, h->acc_double);
    if (l->type == ORC_DISCRETE && l->lr > 0.0f)
      orc_discrete_update(in, l->nin, ein, l->nout, l->Wb, l->corrWb, l->bin, l->bout, l->K, l->b, l->corrb,
                          rows, l->nout, l->lr, l->mmt, l->wc, l->gdf, h->acc_double);
    if (i == h->stopper) break;
    ein = l->eout;
  }
}

/* TNetCu.cc:427-441 one bunch: propagate, xent evaluate, backpropagate */
void orc_net_train_bunch(OrcNet *h, const float *X, const float *T, int rows, int cross_validate) {
  int nout = h->L[h->n - 1].nout;
  if (h->err_rows != rows) {
    free(h->err);
    h->err = (float *)calloc((size_t)rows * nout, sizeof(float));
    h->err_rows = rows;
  }
  orc_net_propagate(h, X, rows, NULL);
  orc_xent_evaluate(h->L[h->n - 1].out, T, h->err, rows, nout, nout, &h->st);
  if (!cross_validate) orc_net_backpropagate(h, X, h->err, rows);
}

void orc_net_stats(OrcNet *h, double *error, long long *frames, long long *correct) {
  *error = h->st.error; *frames = h->st.frames; *correct = h->st.correct;
}
/* returns W in on-disk layout [nout x nin] */
void orc_net_get_affine(OrcNet *h, int layer, float *Wt, float *b) {
  OrcLayer *l = &h->L[layer];
  for (int o = 0; o < l->nout; o++)
    for (int i = 0; i < l->nin; i++) Wt[IDX(o, i, l->nin)] = l->W[IDX(i, o, l->nout)];
  memcpy(b, l->b, sizeof(float) * (size_t)l->nout);
}
/* shared: Wt [nout/K x nin/K], b [nout/K] */
void orc_net_get_shared(OrcNet *h, int layer, float *Wt, float *b) {
  OrcLayer *l = &h->L[layer];
  int bi = l->nin / l->K, bo = l->nout / l->K;
  for (int o = 0; o < bo; o++)
    for (int i = 0; i < bi; i++) Wt[IDX(o, i, bi)] = l->W[IDX(i, o, bo)];
  memcpy(b, l->b, sizeof(float) * (size_t)bo);
}
/* discrete: the blocks concatenated, each [bout[i] x bin[i]]; b [nout] */
void orc_net_get_discrete(OrcNet *h, int layer, float *Wt_all, float *b) {
  OrcLayer *l = &h->L[layer];
  float *dst = Wt_all;
  for (int k = 0; k < l->K; k++) {
    for (int o = 0; o < l->bout[k]; o++)
      for (int i = 0; i < l->bin[k]; i++) dst[IDX(o, i, l->bin[k])] = l->Wb[k][IDX(i, o, l->bout[k])];
    dst += (size_t)l->bin[k] * l->bout[k];
  }
  memcpy(b, l->b, sizeof(float) * (size_t)l->nout);
}
const float *orc_net_layer_out(OrcNet *h, int layer) { return h->L[layer].out; }
const float *orc_net_layer_eout(OrcNet *h, int layer) { return h->L[layer].eout; }
const float *orc_net_err(OrcNet *h) { return h->err; }
void orc_net_free(OrcNet *h) {
  if (!h) return;
  for (int i = 0; i < h->n; i++) {
    OrcLayer *l = &h->L[i];
    free(l->W); free(l->b); free(l->corrW); free(l->corrb); free(l->out); free(l->eout);
    if (l->type == ORC_DISCRETE)
      for (int k = 0; k < l->K; k++) { free(l->Wb[k]); free(l->corrWb[k]); }
    free(l->Wb); free(l->corrWb); free(l->bin); free(l->bout);
  }
  free(h->L); free(h->err); free(h);
}

/* ------------------------------------------------------------------------- */
/* CuRbm  (cuRbm.cc) + TRbmCu.cc:326-354 CD-1 step                            */
/* ------------------------------------------------------------------------- */
typedef struct {
  int nvis, nhid, vis_gauss, hid_gauss;
  float *W;       /* [nvis x nhid] */
  float *vb, *hb, *cW, *cvb, *chb;
  float lr, mmt, wc;
  int acc_double;
  OrcObjStats st;
  /* CuRbmSparse (cuRbmSparse.h:64-70,92-105): off unless orc_rbm_set_sparse() was called */
  int sparse;
  float sp_prior, sp_lambda, sp_cost;
  float *sp_q, *sp_qcur, *vis_mean;
} OrcRbm;

/* Wt on-disk [nhid x nvis] (cuRbm.cc:198-207) */
OrcRbm *orc_rbm_new(int nvis, int nhid, int vis_gauss, int hid_gauss, const float *Wt,
                    const float *vb, const float *hb, float lr, float mmt, float wc,
                    int acc_double) {
  OrcRbm *r = (OrcRbm *)calloc(1, sizeof(OrcRbm));
  r->nvis = nvis; r->nhid = nhid; r->vis_gauss = vis_gauss; r->hid_gauss = hid_gauss;
  r->lr = lr; r->mmt = mmt; r->wc = wc; r->acc_double = acc_double;
  r->W = (float *)malloc(sizeof(float) * (size_t)nvis * nhid);
  for (int h = 0; h < nhid; h++)
    for (int v = 0; v < nvis; v++) r->W[IDX(v, h, nhid)] = Wt[IDX(h, v, nvis)];
  r->vb = (float *)malloc(sizeof(float) * nvis); memcpy(r->vb, vb, sizeof(float) * nvis);
  r->hb = (float *)malloc(sizeof(float) * nhid); memcpy(r->hb, hb, sizeof(float) * nhid);
  r->cW = (float *)calloc((size_t)nvis * nhid, sizeof(float));
  r->cvb = (float *)calloc(nvis, sizeof(float));
  r->chb = (float *)calloc(nhid, sizeof(float));
  return r;
}
/* cuRbmSparse.h:92-105 constructor defaults (prior 1e-4, lambda 0.95); the cost comes from the file (cuRbmSparse.cc:198-199) */
void orc_rbm_set_sparse(OrcRbm *r, float cost) {
  r->sparse = 1; r->sp_prior = 0.0001f; r->sp_lambda = 0.95f; r->sp_cost = cost;
  r->sp_q = (float *)malloc(sizeof(float) * r->nhid);
  r->sp_qcur = (float *)calloc(r->nhid, sizeof(float));
  r->vis_mean = (float *)calloc(r->nvis, sizeof(float));
  for (int h = 0; h < r->nhid; h++) r->sp_q[h] = r->sp_prior;
}
void orc_rbm_free(OrcRbm *r) {
  if (!r) return;
  free(r->W); free(r->vb); free(r->hb); free(r->cW); free(r->cvb); free(r->chb);
  free(r->sp_q); free(r->sp_qcur); free(r->vis_mean); free(r);
}
/* cuRbm.cc:15-23 PropagateFnc / :104-115 Propagate */
void orc_rbm_propagate(OrcRbm *r, const float *vis, float *hid, int rows) {
  orc_set_const(hid, 0.0f, rows, r->nhid, r->nhid);
  orc_add_scaled_row(1.0f, r->hb, 0.0f, hid, rows, r->nhid, r->nhid);
  orc_gemm('N', 'N', rows, r->nhid, r->nvis, 1.0f, vis, r->nvis, r->W, r->nhid, 1.0f, hid, r->nhid,
           r->acc_double);
  if (!r->hid_gauss) orc_sigmoid(hid, hid, rows, r->nhid, r->nhid);
}
/* cuRbm.cc:118-128 Reconstruct */
void orc_rbm_reconstruct(OrcRbm *r, const float *hid, float *vis, int rows) {
  orc_set_const(vis, 0.0f, rows, r->nvis, r->nvis);
  orc_add_scaled_row(1.0f, r->vb, 0.0f, vis, rows, r->nvis, r->nvis);
  orc_gemm('N', 'T', rows, r->nvis, r->nhid, 1.0f, hid, r->nhid, r->W, r->nhid, 1.0f, vis, r->nvis,
           r->acc_double);
  if (!r->vis_gauss) orc_sigmoid(vis, vis, rows, r->nvis, r->nvis);
}
/* cuRbm.cc:131-174 RbmUpdate */
void orc_rbm_update(OrcRbm *r, const float *pos_vis, const float *pos_hid, const float *neg_vis,
                    const float *neg_hid, int rows) {
  float N = (float)rows;
  int V = r->nvis, H = r->nhid;
  const int sp = r->sparse && !r->hid_gauss;
  if (sp) { /* cuRbmSparse.cc:139-145 : q = lambda*q + (1-lambda)*mean(pos_hid); qcur = q - prior; mean visible */
    orc_add_col_sum((float)(1.0 / rows), pos_hid, 0.0f, r->sp_qcur, rows, H, H);
    orc_add_scaled((float)(1.0 - r->sp_lambda), r->sp_qcur, r->sp_lambda, r->sp_q, 1, H, H);
    orc_set_const(r->sp_qcur, -r->sp_prior, 1, H, H);
    orc_add_scaled(1.0f, r->sp_q, 1.0f, r->sp_qcur, 1, H, H);
    orc_add_col_sum((float)(1.0 / rows), pos_vis, 0.0f, r->vis_mean, rows, V, V);
  }
  orc_gemm('T', 'N', V, H, rows, -r->lr / N, neg_vis, V, neg_hid, H, r->mmt, r->cW, H, r->acc_double);
  orc_gemm('T', 'N', V, H, rows, +r->lr / N, pos_vis, V, pos_hid, H, 1.0f, r->cW, H, r->acc_double);
  orc_add_scaled(-r->lr * r->wc, r->W, 1.0f, r->cW, V, H, H);
  if (sp) /* cuRbmSparse.cc:151-153 : cW += -cost * vis_mean * qcur^T (cublasSger) */
    for (int v = 0; v < V; v++)
      for (int h = 0; h < H; h++) r->cW[IDX(v, h, H)] += -r->sp_cost * r->vis_mean[v] * r->sp_qcur[h];
  orc_add_scaled(1.0f, r->cW, 1.0f, r->W, V, H, H);
  orc_add_col_sum(-r->lr / N, neg_vis, r->mmt, r->cvb, rows, V, V);
  orc_add_col_sum(+r->lr / N, pos_vis, 1.0f, r->cvb, rows, V, V);
  orc_add_scaled(1.0f, r->cvb, 1.0f, r->vb, 1, V, V);
  orc_add_col_sum(-r->lr / N, neg_hid, r->mmt, r->chb, rows, H, H);
  orc_add_col_sum(+r->lr / N, pos_hid, 1.0f, r->chb, rows, H, H);
  if (sp) orc_add_scaled(-r->sp_cost, r->sp_qcur, 1.0f, r->chb, 1, H, H); /* cuRbmSparse.cc:162-164 */
  orc_add_scaled(1.0f, r->chb, 1.0f, r->hb, 1, H, H);
}
/* TRbmCu.cc:326-354 : one CD-1 bunch.  z1..z4 are the [rows x nhid] RNG state
 * (stride nhid).  Scratch buffers supplied by the caller ([rows x nhid] x3,
 * [rows x nvis] x2). */
void orc_rbm_cd1_bunch(OrcRbm *r, const float *pos_vis, int rows, unsigned *z1, unsigned *z2,
                       unsigned *z3, unsigned *z4, float *pos_hid, float *neg_hid, float *rnd,
                       float *neg_vis, float *err) {
  orc_rbm_propagate(r, pos_vis, pos_hid, rows);
  if (!r->hid_gauss) {
    orc_rand(rnd, z1, z2, z3, z4, rows, r->nhid, r->nhid);
    orc_binarize_probs(neg_hid, pos_hid, rnd, rows, r->nhid, r->nhid);
  } else {
    /* TRbmCu.cc:337-339 : neg_hid = pos_hid ; AddGaussNoise(neg_hid) */
    memcpy(neg_hid, pos_hid, sizeof(float) * (size_t)rows * r->nhid);
    orc_gauss_rand(rnd, z1, z2, z3, z4, rows, r->nhid, r->nhid);
    orc_add_scaled(1.0f, rnd, 1.0f, neg_hid, rows, r->nhid, r->nhid);
  }
  orc_rbm_reconstruct(r, neg_hid, neg_vis, rows);
  orc_rbm_propagate(r, neg_vis, neg_hid, rows);
  orc_rbm_update(r, pos_vis, pos_hid, neg_vis, neg_hid, rows);
  orc_mse_evaluate(neg_vis, pos_vis, err, rows, r->nvis, r->nvis, &r->st);
}
void orc_rbm_get(OrcRbm *r, float *Wt, float *vb, float *hb) {
  for (int h = 0; h < r->nhid; h++)
    for (int v = 0; v < r->nvis; v++) Wt[IDX(h, v, r->nvis)] = r->W[IDX(v, h, r->nhid)];
  memcpy(vb, r->vb, sizeof(float) * r->nvis);
  memcpy(hb, r->hb, sizeof(float) * r->nhid);
}
void orc_rbm_stats(OrcRbm *r, double *error, long long *frames) {
  *error = r->st.error; *frames = r->st.frames;
}

/* ------------------------------------------------------------------------- */
/* CuRecurrent (cuRecurrent.cc:16-153)  — one frame at a time                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int nin, nout, bptt;
  float *W;  /* [(nin+nout) x nout] */
  float *b, *cW, *cb;
  float *hist; /* [(bptt+1) x (nin+nout)] */
  float *out;  /* [nout] persistent output (y_{t-1} feeds y_t) */
  float lr, mmt, wc;
} OrcRnn;

OrcRnn *orc_rnn_new(int nin, int nout, int bptt, const float *Wt, const float *b, float lr,
                    float mmt, float wc) {
  OrcRnn *r = (OrcRnn *)calloc(1, sizeof(OrcRnn));
  int K = nin + nout;
  r->nin = nin; r->nout = nout; r->bptt = bptt; r->lr = lr; r->mmt = mmt; r->wc = wc;
  r->W = (float *)malloc(sizeof(float) * (size_t)K * nout);
  for (int o = 0; o < nout; o++)
    for (int i = 0; i < K; i++) r->W[IDX(i, o, nout)] = Wt[IDX(o, i, K)];
  r->b = (float *)malloc(sizeof(float) * nout); memcpy(r->b, b, sizeof(float) * nout);
  r->cW = (float *)calloc((size_t)K * nout, sizeof(float));
  r->cb = (float *)calloc(nout, sizeof(float));
  r->hist = (float *)calloc((size_t)(bptt + 1) * K, sizeof(float));
  r->out = (float *)calloc(nout, sizeof(float));
  return r;
}
void orc_rnn_free(OrcRnn *r) {
  if (!r) return;
  free(r->W); free(r->b); free(r->cW); free(r->cb); free(r->hist); free(r->out); free(r);
}
/* cuRecurrent.h:36-41 ClearHistory */
void orc_rnn_clear(OrcRnn *r) {
  memset(r->hist, 0, sizeof(float) * (size_t)(r->bptt + 1) * (r->nin + r->nout));
  memset(r->out, 0, sizeof(float) * r->nout);
}
/* cuRecurrent.cc:16-54 PropagateFnc : y = sigmoid(b + [x ; y_prev] * W) */
void orc_rnn_propagate(OrcRnn *r, const float *x, float *y) {
  int K = r->nin + r->nout;
  memmove(r->hist + K, r->hist, sizeof(float) * (size_t)r->bptt * K);
  memcpy(r->hist, x, sizeof(float) * r->nin);
  memcpy(r->hist + r->nin, r->out, sizeof(float) * r->nout);
  for (int o = 0; o < r->nout; o++) {
    float acc = 0.0f;
    for (int i = 0; i < K; i++) acc = fmaf(r->hist[i], r->W[IDX(i, o, r->nout)], acc);
    r->out[o] = 1.0f * acc + 1.0f * r->b[o];
  }
  orc_sigmoid(r->out, r->out, 1, r->nout, r->nout);
  memcpy(y, r->out, sizeof(float) * r->nout);
}
/* cuRecurrent.cc:57-83 BackpropagateFnc : eout += (e .* dsig) * W[0:nin,:]^T
 * (beta=1 into the persistent buffer — quirk SURVEY A.5) */
void orc_rnn_backpropagate(OrcRnn *r, const float *e, float *eout_accum) {
  float *ds = (float *)malloc(sizeof(float) * r->nout);
  orc_diff_sigmoid(ds, e, r->out, 1, r->nout, r->nout);
  for (int i = 0; i < r->nin; i++) {
    float acc = 0.0f;
    for (int o = 0; o < r->nout; o++) acc = fmaf(r->W[IDX(i, o, r->nout)], ds[o], acc);
    eout_accum[i] = acc + eout_accum[i];
  }
  free(ds);
}
/* cuRecurrent.cc:86-153 Update */
void orc_rnn_update(OrcRnn *r, const float *e) {
  int K = r->nin + r->nout, H = r->nout;
  float *ds = (float *)malloc(sizeof(float) * H);
  float *ep = (float *)malloc(sizeof(float) * H);
  orc_diff_sigmoid(ds, e, r->out, 1, H, H);
  orc_set_const(r->cW, 0.0f, K, H, H);
  for (int i = 0; i < K; i++)
    for (int o = 0; o < H; o++) r->cW[IDX(i, o, H)] += (-r->lr * r->hist[i]) * ds[o];
  orc_add_col_sum(-r->lr, ds, r->mmt, r->cb, 1, H, H);
  for (int t = 1; t <= r->bptt; t++) {
    for (int i = 0; i < H; i++) {
      float acc = 0.0f;
      for (int o = 0; o < H; o++) acc = fmaf(r->W[IDX(r->nin + i, o, H)], ds[o], acc);
      ep[i] = acc;
    }
    const float *hout = r->hist + (size_t)(t - 1) * K + r->nin;
    orc_diff_sigmoid(ds, ep, hout, 1, H, H);
    const float *hrow = r->hist + (size_t)t * K;
    for (int i = 0; i < K; i++)
      for (int o = 0; o < H; o++) r->cW[IDX(i, o, H)] += (-r->lr * hrow[i]) * ds[o];
    orc_add_col_sum(-r->lr, ds, 1.0f, r->cb, 1, H, H);
  }
  orc_add_scaled(-r->lr * r->wc, r->W, 1.0f, r->cW, K, H, H);
  orc_add_scaled(1.0f, r->cW, 1.0f, r->W, K, H, H);
  orc_add_scaled(1.0f, r->cb, 1.0f, r->b, 1, H, H);
  free(ds); free(ep);
}
void orc_rnn_get(OrcRnn *r, float *Wt, float *b) {
  int K = r->nin + r->nout;
  for (int o = 0; o < r->nout; o++)
    for (int i = 0; i < K; i++) Wt[IDX(o, i, K)] = r->W[IDX(i, o, r->nout)];
  memcpy(b, r->b, sizeof(float) * r->nout);
}
