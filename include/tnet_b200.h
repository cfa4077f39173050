/*
 * tnet_b200.h — C ABI of libtnetb200.so: the Blackwell (sm_100a) replacement for the bottom seam of
 * TNet's GPU training path (troylee/nnet-asr).
 *
 * What it replaces in the reference (paths relative to /root/reference/src):
 *   - CuBaseLib/cukernels.h:5-79      the `cudaF_*` extern "C" kernel launchers
 *   - CuBaseLib/curandkernels.h:8-31  `cudaF_rand / cudaF_gauss_rand / cudaF_binarize_probs`
 *   - legacy cuBLAS v1 calls: cublasSgemm (CuBaseLib/cumatrix.tcc:363, cumath.cc:105,237),
 *     cublasSgemv (cumath.cc:334), cublasSger (cumath.cc:358, cumatrix.tcc:384), cublasInit/Shutdown
 *     (cudevice.cc:60,66)
 *   - the cudaMallocPitch/cudaMemcpy2D/cudaMemset plumbing of CuMatrix/CuVector
 *     (cumatrix.tcc:16-190, cuvector.tcc:14-120)
 *
 * Differences from the reference seam: operands keep the `(pointer, MatrixDim)` form but the
 * caller no longer chooses grid/block; every call takes an explicit context (device, stream, math mode,
 * workspace) and returns an int status instead of relying on cudaGetLastError(); ops that always follow
 * one another in the reference (bias + GEMM + sigmoid; GEMM + diff-sigmoid; GEMM + momentum/L2 update;
 * softmax + cross-entropy + frame accuracy) have fused entry points next to the 1:1 ones.
 *
 * All matrix pointers are DEVICE pointers to row-major fp32 with leading dimension `stride` (in elements).
 * The tensor-core GEMMs need `stride` to be a multiple of 4 and every base 16-byte aligned (tnb_malloc_pitch
 * guarantees a 128-byte pitch); GEMMs on views that are not (column blocks at odd offsets, CuMath::OffsetGemm) run on a
 * plain fp32 FMA kernel on the GPU.  Nothing here falls back to the CPU: without a CUDA device every compute
 * entry point returns TNB_ERR_CUDA.
 */
#ifndef TNET_B200_H_
#define TNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as the reference's MatrixDim (CuBaseLib/cukernels.h:12-16). */
typedef struct TnbMatrixDim_ {
  int rows;
  int cols;
  int stride;
} TnbMatrixDim;

typedef struct TnbContext_ TnbContext; /* opaque: device, stream(s), tensor-map cache, counters */

enum {
  TNB_OK = 0,
  TNB_ERR_CUDA = 1,     /* CUDA runtime/driver error; text via tnb_last_error() */
  TNB_ERR_ARG = 2,      /* bad dimension / alignment / null pointer */
  TNB_ERR_UNSUPPORTED = 3,
  TNB_ERR_NCCL = 4,
  TNB_ERR_COMM = 5      /* a data-parallel rank gave up waiting for a peer (peer-memory schedule); text says which */
};

/* GEMM arithmetic (north_star: "fp32-equivalent via 3xTF32 as the default, bf16 reported separately") */
enum {
  TNB_MATH_3XTF32 = 0, /* split fp32 = hi + lo (tf32 each); acc += lo*hi + hi*lo + hi*hi  (default) */
  TNB_MATH_TF32 = 1,   /* single tf32 product (operands rounded to 10-bit mantissa) */
  TNB_MATH_FP32_SIMT = 2, /* plain fp32 FMA on CUDA cores: debug / cross-check path, no tensor cores */
  TNB_MATH_BF16 = 3    /* operands rounded to bf16 (RN), products accumulated in fp32 (TMEM); weights, activations, errors and
                          momentum buffers stay fp32.  Reported separately from the fp32-equivalent default. */
};

/* ---- context / device (replaces CuDevice, CuBaseLib/cudevice.cc:22-121) -------------------------- */
const char *tnb_version(void);
const char *tnb_last_error(void);                       /* thread-local text of the last failure */
int tnb_device_count(int *count);                       /* TNB_ERR_CUDA when no driver/GPU */
/* device < 0 : pick the GPU with the largest free-memory ratio (cudevice.cc:27-56). */
int tnb_ctx_create(TnbContext **ctx, int device);
int tnb_ctx_destroy(TnbContext *ctx);
int tnb_ctx_device(TnbContext *ctx, int *device);
int tnb_ctx_set_math(TnbContext *ctx, int math_mode);
int tnb_ctx_get_math(TnbContext *ctx, int *math_mode);
int tnb_ctx_stream(TnbContext *ctx, void **cuda_stream); /* the cudaStream_t every op is enqueued on */
/* make the entry points that take no stream argument enqueue on another stream of the context (TNB_STREAM_COMPUTE restores the
 * default).  The host mirror uses it to run a layer's weight-gradient GEMM next to the dX chain of the layers below. */
int tnb_ctx_use_stream(TnbContext *ctx, int stream_id);
int tnb_ctx_sync(TnbContext *ctx);                      /* reference semantics: cudaThreadSynchronize() */
int tnb_ctx_free_memory(TnbContext *ctx, size_t *free_bytes, size_t *total_bytes); /* cudevice.cc:100-118 */
/* number of kernels of THIS library launched through ctx since creation (graph replays included) */
int tnb_ctx_launch_count(TnbContext *ctx, unsigned long long *launches);

/* Per-kernel device timing of the GEMM launches (CUDA events on the ctx stream around every tcgen05 GEMM launched
 * between begin and end): total milliseconds, launch count and algorithmic flops (2*M*N*K per launch).  This is what
 * bench.py's roofline object is computed from. */
int tnb_ctx_profile_begin(TnbContext *ctx);
int tnb_ctx_profile_end(TnbContext *ctx, double *gemm_ms, unsigned long long *gemm_launches, double *gemm_flops);

/* ---- memory (replaces cudaMallocPitch/cudaFree/cudaMemcpy2D/cudaMemset in cumatrix.tcc) ----------- */
/* rows x cols fp32 (or any 4-byte type), zero-filled like CuMatrix::Init (cumatrix.tcc:16-34);
 * *stride_elems is a multiple of 32 elements (128 B). */
int tnb_malloc_pitch(TnbContext *ctx, void **ptr, int *stride_elems, int rows, int cols);
/* rows x cols 2-byte elements (bf16 twins of fp32 matrices, TNB_MATH_BF16), zero-filled; *stride_elems is a multiple of 64. */
int tnb_malloc_pitch16(TnbContext *ctx, void **ptr, int *stride_elems, int rows, int cols);
int tnb_malloc(TnbContext *ctx, void **ptr, size_t bytes); /* zero-filled */
int tnb_free(TnbContext *ctx, void *ptr);
int tnb_memset(TnbContext *ctx, void *ptr, int value, size_t bytes);
/* kind: 0 H2D, 1 D2H, 2 D2D.  Asynchronous on the ctx stream for pinned host memory. */
int tnb_memcpy2d(TnbContext *ctx, void *dst, size_t dpitch_bytes, const void *src, size_t spitch_bytes,
                 size_t width_bytes, size_t height, int kind);
int tnb_memcpy(TnbContext *ctx, void *dst, const void *src, size_t bytes, int kind);
int tnb_host_alloc(void **ptr, size_t bytes); /* pinned host memory */
/* ---- CUDA graphs: record a launch-bound sequence of calls once, replay it (TRecurrentCu: ~115 small kernels per frame).
 * No counterpart in the reference, which launches every kernel synchronously (cuSafeCall + cudaThreadSynchronize,
 * src/CuBaseLib/cucommon.h:13-22; per-frame sequence: cuRecurrent.cc:16-153).
 * Between begin and end every call on the compute stream is captured instead of executed; do not allocate, synchronise or copy
 * to/from pageable host memory there (run the sequence once eagerly first so that every buffer and scratch exists). */
int tnb_graph_begin(TnbContext *ctx);
int tnb_graph_end(TnbContext *ctx, void **graph);
int tnb_graph_launch(TnbContext *ctx, void *graph);   /* counts the graph's kernels in tnb_ctx_launch_count */
int tnb_graph_destroy(TnbContext *ctx, void *graph);

/* ---- streams and events (the reference copies synchronously on the default stream: cumatrix.tcc:68-118).  A host that wants
 * its transfers to overlap the training step enqueues them on the context's copy stream and orders the two streams with
 * events; nothing below blocks the host except tnb_event_sync. */
enum {
  TNB_STREAM_COMPUTE = 0, /* every entry point without a stream argument */
  TNB_STREAM_COPY = 1,    /* host<->device transfers */
  TNB_STREAM_COMM = 2,    /* the collectives (tnb_allreduce_sum, tnb_dp_update) */
  TNB_STREAM_AUX = 3,     /* side streams for small kernels that run next to the compute stream's GEMMs: the data-parallel step */
  TNB_STREAM_AUX2 = 4     /*   puts the bias-gradient column sums on one and the per-layer updates on the other */
};
int tnb_memcpy2d_on(TnbContext *ctx, int stream_id, void *dst, size_t dpitch_bytes, const void *src, size_t spitch_bytes,
                    size_t width_bytes, size_t height, int kind); /* host side must be pinned; never synchronises */
int tnb_memcpy_on(TnbContext *ctx, int stream_id, void *dst, const void *src, size_t bytes, int kind);
int tnb_event_create(TnbContext *ctx, void **event);
int tnb_event_destroy(TnbContext *ctx, void *event);
int tnb_event_record(TnbContext *ctx, void *event, int stream_id);      /* after everything enqueued so far on that stream */
int tnb_stream_wait_event(TnbContext *ctx, int stream_id, void *event); /* device-side wait */
int tnb_event_sync(TnbContext *ctx, void *event);                       /* host-side wait */
int tnb_host_free(void *ptr);

/* ---- 1:1 replacements of cukernels.h (float instances) ------------------------------------------- */
int tnb_set_const(TnbContext *ctx, float *mat, float value, TnbMatrixDim d);                 /* cukernels.h:22 */
int tnb_apply_log(TnbContext *ctx, float *mat, TnbMatrixDim d);                              /* :23 */
int tnb_scale_cols(TnbContext *ctx, float *mat, const float *scale, TnbMatrixDim d);         /* :26 */
int tnb_scale_rows(TnbContext *ctx, float *mat, const float *scale, TnbMatrixDim d);         /* :27 */
int tnb_add_scaled(TnbContext *ctx, float alpha, const float *A, float beta, float *dst, TnbMatrixDim d); /* :28 */
int tnb_add_scaled_row(TnbContext *ctx, float alpha, const float *row, float beta, float *dst, TnbMatrixDim d); /* :29 */
int tnb_mul_elem(TnbContext *ctx, float *mat, const float *A, TnbMatrixDim d);               /* :30 */
int tnb_log_elem(TnbContext *ctx, float *mat, TnbMatrixDim d);                               /* :31 */
/* vec[c] = alpha * sum_r mat[r,c] + beta * vec[c]; both reference variants (:34-35) collapse into one
 * deterministic kernel with double partial sums. */
int tnb_add_col_sum(TnbContext *ctx, float alpha, const float *mat, float beta, float *vec, TnbMatrixDim d);
int tnb_sigmoid(TnbContext *ctx, float *y, const float *x, TnbMatrixDim d);                  /* :40 */
int tnb_diff_sigmoid(TnbContext *ctx, float *eout, const float *e, const float *y, TnbMatrixDim d); /* :41 */
int tnb_softmax(TnbContext *ctx, float *y, const float *x, TnbMatrixDim d);                  /* :38-39 */
int tnb_expand(TnbContext *ctx, float *y, const float *x, const int *off, TnbMatrixDim d_out, TnbMatrixDim d_in); /* :43 */
int tnb_rearrange(TnbContext *ctx, float *y, const float *x, const int *copy_from, TnbMatrixDim d_out, TnbMatrixDim d_in); /* :44 */
/* y[r,:] = x[copy_from[r],:] for r < d_out.rows (= permutation length), cumath.cc:155-174 */
int tnb_randomize(TnbContext *ctx, float *y, const float *x, const int *copy_from, TnbMatrixDim d_out, TnbMatrixDim d_in); /* :45 */
/* sum[c] += sum_r X[r,c] ; sumsq[c] += sum_r (float)(X[r,c]*X[r,c]); sum/sumsq: DEVICE doubles [cols], accumulated across calls
 * (the global mean/variance statistics TNormCu.cc:268-272 gathers on the host after copying every utterance back) */
int tnb_accum_moments(TnbContext *ctx, const float *X, TnbMatrixDim d, double *sum, double *sumsq);
/* match[r] = argmax(out[r,:]) == argmax(des[r,:]) with the reference's tie rules (bit-exact):
 * cols > 256 sequential first-max; cols <= 256 the index tree of _max_id_reduce (cukernels.cu:424-446). */
int tnb_check_class(TnbContext *ctx, const float *out, const float *des, int *match, TnbMatrixDim d); /* :47-48 */

/* ---- replacements of curandkernels.h (per-element Hybrid-Taus state z1..z4, bit-exact streams) ----- */
int tnb_rand(TnbContext *ctx, float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, TnbMatrixDim d);
int tnb_gauss_rand(TnbContext *ctx, float *mat, unsigned *z1, unsigned *z2, unsigned *z3, unsigned *z4, TnbMatrixDim d);
int tnb_binarize_probs(TnbContext *ctx, float *states, const float *probs, const float *rnd, TnbMatrixDim d);
/* fused CuRand::BinarizeProbs (curand.tcc:136-155): states = probs > HybridTaus(z) ? 1 : 0, no tmp matrix */
int tnb_rand_binarize(TnbContext *ctx, float *states, const float *probs, unsigned *z1, unsigned *z2, unsigned *z3,
                      unsigned *z4, TnbMatrixDim d);
/* fused CuRand::AddGaussNoise (curand.tcc:56-60): tgt += gscale * BoxMuller(z) */
int tnb_add_gauss_noise(TnbContext *ctx, float *tgt, float gscale, unsigned *z1, unsigned *z2, unsigned *z3,
                        unsigned *z4, TnbMatrixDim d);

/* ---- GEMM: replaces cublasSgemm as called by CuMatrix<float>::Gemm (cumatrix.tcc:335-370) --------- */
/* Row-major  C[m x n] = alpha * op(A) * op(B) + beta * C ; transa/transb in {'N','T'} exactly as
 * CuMatrix::Gemm(transa, transb, alpha, A, B, beta) sees them.  tcgen05/TMEM kernel fed by TMA;
 * arithmetic per tnb_ctx_set_math. */
int tnb_gemm(TnbContext *ctx, char transa, char transb, int m, int n, int k, float alpha, const float *A, int lda,
             const float *B, int ldb, float beta, float *C, int ldc);
/* y = alpha*op(A)[offset rows/cols]*x + beta*y  — CuMath::OffsetGemv (cumath.cc:283-340) */
int tnb_offset_gemv(TnbContext *ctx, char trans, float alpha, const float *A, TnbMatrixDim dA, const float *x, int dimX,
                    float beta, float *y, int dimY, int offsetY);
/* A += alpha * x * y^T  — CuMath::BlasGer / CuMatrix::BlasGer (cumath.cc:344-362) */
int tnb_ger(TnbContext *ctx, float alpha, const float *x, int dimX, const float *y, int dimY, float *A, TnbMatrixDim dA);

/* CuRecurrent::Update fused (cuRecurrent.cc:92-153; W is [(nin + H) x H], the last H rows are the recurrent weights):
 * tnb_rnn_bptt_step: one step of the BPTT chain — d_out = diffsigmoid(W[nin.., :] * d_prev, y_hist) and bcorr += -lr * d_out — in one
 *   launch (the reference: OffsetGemv + DiffSigmoid + BlasGer + AddColSum per step);
 * tnb_rnn_apply: W += sum_i (-lr * hist[i]) (x) d[i] + (-lr*wc) * W over the stored steps i = 0..nsteps-1, summed in the reference's
 *   order, in one pass over W (the reference: SetConst + nsteps BlasGer + 2 AddScaled). */
int tnb_rnn_bptt_step(TnbContext *ctx, const float *W, TnbMatrixDim dW, int nin, const float *d_prev, const float *y_hist, float *d_out,
                      float *bcorr, float lr);
int tnb_rnn_apply(TnbContext *ctx, float *W, TnbMatrixDim dW, const float *hist, int ld_hist, const float *d, int ld_d, int nsteps, float lr,
                  float wc);

/* ---- fused hot-path ops ----------------------------------------------------------------------------- */
enum { TNB_ACT_NONE = 0, TNB_ACT_SIGMOID = 1 };
/* CuBiasedLinearity::PropagateFnc (+ CuSigmoid::PropagateFnc when act = SIGMOID):
 *   Y[rows x nout] = act( X[rows x nin] * W[nin x nout] + bias )      cuBiasedLinearity.cc:11-16, cuActivation.cc:9-14 */
int tnb_affine_fwd(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *W, TnbMatrixDim dW, const float *bias,
                   float *Y, TnbMatrixDim dY, int act);
/* CuBiasedLinearity::BackpropagateFnc (+ CuSigmoid::BackpropagateFnc of the layer below when Yprev != NULL):
 *   Eprev[rows x nin] = (E[rows x nout] * W^T) (.* Yprev .* (1 - Yprev))   cuBiasedLinearity.cc:20-25, cuActivation.cc:17-22 */
int tnb_affine_bwd_dx(TnbContext *ctx, const float *E, TnbMatrixDim dE, const float *W, TnbMatrixDim dW, const float *Yprev,
                      TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev);
/* weight gradient only:  G[nin x nout] = X^T * E ;  gb[nout] = colsum(E)   (for data-parallel: allreduce G,gb, then
 * tnb_sgd_update).  gb may be NULL. */
int tnb_affine_grad(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *G,
                    TnbMatrixDim dG, float *gb);
/* CuBiasedLinearity::Update (cuBiasedLinearity.cc:44-64) in one pass over the weights:
 *   corrW = X^T*E + mmt*corrW ; corrb = colsum(E) + mmt*corrb ; W += (-lr/N)*corrW ; b += (-lr/N)*corrb ;
 *   W += (-lr*wc*(gdf?1:rows))*W        with N = (gdf?rows:1)/(1-mmt), scalars evaluated in float as the reference.
 * n_frames_global > 0 overrides `rows` in N and in the L2 factor (data-parallel shards of one bunch). */
int tnb_affine_update(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *W,
                      TnbMatrixDim dW, float *bias, float *corrW, float *corrb, float lr, float mmt, float wc,
                      int grad_div_frm, int n_frames_global);
/* bias == NULL && corrb == NULL in tnb_affine_update (and its _bf16 twin): weight half only.  The bias halves of several layers
 * are then applied together by tnb_bias_update_batch — same arithmetic per layer, two launches for the whole stack instead of two
 * per layer (CuNetwork::Backpropagate defers them to the end of the pass; nothing reads a bias during backpropagation):
 *   corrb = colsum(E) + mmt*corrb ; b += (-lr/N)*corrb        with N as in tnb_affine_update */
#define TNB_MAX_BIAS_JOBS 16
typedef struct TnbBiasJob_ {
  const float *E;      /* error at the layer's output [rows x nout] */
  TnbMatrixDim dE;
  float *bias, *corrb; /* [nout]; bias == NULL: gradient only, corrb[c] = colsum(E)[c] (summed over ranks before it is applied) */
  float lr, mmt;
  int grad_div_frm, n_frames_global;
} TnbBiasJob;
int tnb_bias_update_batch(TnbContext *ctx, const TnbBiasJob *jobs, int n); /* cuBiasedLinearity.cc:56-59 for n layers */
/* the same on another stream of the context, with its own reduction scratch (may run next to compute-stream column sums) */
int tnb_bias_update_batch_on(TnbContext *ctx, int stream_id, const TnbBiasJob *jobs, int n);
/* the same update (cuBiasedLinearity.cc:55-63) given an already summed gradient (after the NCCL allreduce):
 *   corrW = G + mmt*corrW ; ... as above.  gb/bias/corrb may be NULL together. */
int tnb_sgd_update(TnbContext *ctx, const float *G, float *W, float *corrW, TnbMatrixDim dW, const float *gb, float *bias,
                   float *corrb, float lr, float mmt, float wc, int grad_div_frm, int n_frames);

/* ---- TNB_MATH_BF16 with resident bf16 twins -----------------------------------------------------------
 * In bf16 mode every entry point above still takes fp32 arrays (operands are rounded into context scratch per call).  The hot
 * path avoids that conversion by keeping a bf16 twin next to each fp32 matrix a GEMM reads: the three fused layer ops below
 * READ twins (X16, W16, E16: row-major bf16 bit patterns, pitch in elements, 16-byte aligned, pitch a multiple of 8) and WRITE
 * the twin of what they produce (Y16, Eprev16, W16; may be NULL) from the same fp32 value they store.  Arithmetic and fp32
 * results are identical to the fp32-array entry points in bf16 mode; the reference code they stand for is the same
 * (cuBiasedLinearity.cc:11-16 forward, :20-25 dX, :44-64 update; cuActivation.cc:9-22 sigmoid / diff-sigmoid). */
int tnb_to_bf16(TnbContext *ctx, uint16_t *dst, int dst_stride, const float *src, TnbMatrixDim d); /* dst = bf16_rn(src); pad columns zeroed */
int tnb_affine_fwd_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *W16, int ldw16,
                        TnbMatrixDim dW, const float *bias, float *Y, TnbMatrixDim dY, uint16_t *Y16, int ldy16, int act);
int tnb_affine_bwd_dx_bf16(TnbContext *ctx, const uint16_t *E16, int lde16, TnbMatrixDim dE, const uint16_t *W16, int ldw16,
                           TnbMatrixDim dW, const float *Yprev, TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev,
                           uint16_t *Eprev16, int ldep16);
/* E (fp32) is read for the bias gradient (column sums stay fp32/double as in the default mode) */
int tnb_affine_grad_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *E16, int lde16,
                         const float *E, TnbMatrixDim dE, float *G, TnbMatrixDim dG, float *gb);
int tnb_affine_update_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *E16, int lde16,
                           const float *E, TnbMatrixDim dE, float *W, TnbMatrixDim dW, uint16_t *W16, int ldw16, float *bias,
                           float *corrW, float *corrb, float lr, float mmt, float wc, int grad_div_frm, int n_frames_global);

/* ---- several independent GEMMs (each with its fused epilogue) as ONE persistent launch (csrc/gemm_multi.cu) -----------------------
 * In CuNetwork::Backpropagate (cuNetwork.h:170-194) the input-gradient GEMM of layer l (cuBiasedLinearity.cc:24) and the
 * weight-gradient + update GEMMs of the layers above it (cuBiasedLinearity.cc:44-64) do not depend on each other: the reference
 * issues them one cublasSgemm after the other, each leaving most of the chip idle at a 1024-frame bunch.  A job is one such GEMM
 * C = epilogue(op(A) * op(B)) cut into 256 x 256 output tiles; tnb_gemm_batch runs all tiles of the `must` jobs plus as many tiles
 * of the `pool` jobs (in order; tile_first / tile_count of a pool job record its progress and are updated) as fit without
 * lengthening the launch — everything left in the pool when there is no `must` job.  Results per element are those of the single
 * GEMM entry points (same mainloop, same epilogue arithmetic).  The caller orders conflicting jobs: a pool job that updates W in
 * place must not share a launch with a job that reads W.  3xTF32 and bf16 modes (bf16: the twins A16 / B16 are read). */
enum { TNB_EPI_STORE = 0, TNB_EPI_FWD = 1, TNB_EPI_DX = 2, TNB_EPI_UPDATE = 3 };
typedef struct TnbGemmJob_ {
  int transa, transb;                 /* 0: op(X) = X, 1: op(X) = X^T — operands row-major as CuMatrix::Gemm receives them (cumatrix.tcc:335-370) */
  int m, n, k;
  const float *A; int lda;
  const float *B; int ldb;
  const uint16_t *A16; int lda16;     /* bf16 twins of A and B (TNB_MATH_BF16) */
  const uint16_t *B16; int ldb16;
  int epilogue;                       /* TNB_EPI_*: which fused layer op this is (selects the specialised epilogue; STORE = any combination) */
  float alpha, beta;                  /* out = alpha*acc + beta*C_old (+ bias) (sigmoid) (.* y(1-y)) */
  float *C; int ldc;
  const float *bias; int act;
  const float *mulY; int ldy;
  float *W; int ldw; float w_scale, w_l2;   /* fused SGD: W += w_scale*out ; W += w_l2*W */
  uint16_t *C16; int ldc16;           /* optional bf16 twins of the stored output / the updated weights */
  uint16_t *W16; int ldw16;
  int tile_first, tile_count;         /* tiles [tile_first, tile_first + tile_count) of the row-major 256 x 256 tile grid still to run */
} TnbGemmJob;
/* the three layer ops as jobs (same arguments as tnb_affine_bwd_dx / tnb_affine_update / tnb_affine_grad; the bias halves of the
 * update stay with tnb_bias_update_batch).  tile_count is set to the job's number of tiles. */
int tnb_job_affine_bwd_dx(TnbGemmJob *job, const float *E, TnbMatrixDim dE, const float *W, TnbMatrixDim dW, const float *Yprev,
                          TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev);
int tnb_job_affine_update(TnbGemmJob *job, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *W, TnbMatrixDim dW,
                          float *corrW, float lr, float mmt, float wc, int grad_div_frm, int n_frames_global);
int tnb_job_affine_grad(TnbGemmJob *job, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *G, TnbMatrixDim dG);
/* bf16 mode: the twins the job reads (A16, B16) and keeps current (C16, W16; may be NULL) */
int tnb_job_set_twins(TnbGemmJob *job, const uint16_t *A16, int lda16, const uint16_t *B16, int ldb16, uint16_t *C16, int ldc16,
                      uint16_t *W16, int ldw16);
int tnb_gemm_job_tiles(const TnbGemmJob *job);
/* 1 if the job suits the batch kernel in the context's math mode (operand alignment; output and contraction at least a tile wide:
 * smaller GEMMs are better served by the tile shapes of the single-GEMM path), else 0 */
int tnb_gemm_batch_ok(TnbContext *ctx, const TnbGemmJob *job);
int tnb_gemm_batch(TnbContext *ctx, const TnbGemmJob *must, int n_must, TnbGemmJob *pool, int n_pool);
/* CTA pairs (2 SMs each) a batch launch may occupy; 0 = all that can be co-resident (74 on a B200).  Fewer leaves SMs to kernels
 * that run beside the batch, e.g. the data-parallel gradient exchange. */
int tnb_gemm_batch_set_pairs(TnbContext *ctx, int pairs);
/* developer aid (TNB_BATCH_TRACE=1): clock64() stamps [pair][64] of the last tnb_gemm_batch launch (slots: csrc/gemm_multi.cu) */
int tnb_gemm_batch_trace_read(TnbContext *ctx, long long *out);

/* CuRbm::RbmUpdate, the CD-1 update of TRbmCu (cuRbm.cc:131-174), for the weights and both biases: the reference issues two GEMMs,
 * two weight-sized AddScaled sweeps, four column sums and two vector adds; here the second GEMM's epilogue does the sweeps
 * (corrW += -lr*wc*W ; W += corrW on the tile it has just produced) and one kernel per bias does its two column sums and updates:
 *   corrW = mmt*corrW + (lr/N) * (pos_vis^T pos_hid - neg_vis^T neg_hid) - lr*wc*W ;  W += corrW
 *   corr_b = mmt*corr_b + (lr/N) * (colsum(pos) - colsum(neg)) ;  b += corr_b          (visible and hidden)
 * in the reference's operation order.  N = rows of the bunch. */
int tnb_rbm_cd1_update(TnbContext *ctx, const float *pos_vis, const float *neg_vis, TnbMatrixDim dV, const float *pos_hid, const float *neg_hid,
                       TnbMatrixDim dH, float *W, TnbMatrixDim dW, float *corrW, float *vis_bias, float *corr_vb, float *hid_bias, float *corr_hb,
                       float lr, float mmt, float wc);

/* tnb_sgd_update (cuBiasedLinearity.cc:55-63) for several layers in one launch (the data-parallel step applies them after the last all-reduce); W16/ldw16:
 * optional bf16 twin of W to refresh (NULL otherwise). */
typedef struct TnbSgdJob_ {
  const float *G;
  float *W, *corrW;
  TnbMatrixDim dW;
  const float *gb;       /* may be NULL together with bias, corrb */
  float *bias, *corrb;
  float lr, mmt, wc;
  int grad_div_frm, n_frames;
  uint16_t *W16;
  int ldw16;
} TnbSgdJob;
int tnb_sgd_update_batch(TnbContext *ctx, const TnbSgdJob *jobs, int n); /* n <= TNB_MAX_BIAS_JOBS */
/* the same enqueued on another stream of the context (next to the backward GEMMs of the layers below); the caller orders it
 * behind the all-reduce that produced G and the compute stream behind it with events */
int tnb_sgd_update_batch_on(TnbContext *ctx, int stream_id, const TnbSgdJob *jobs, int n);

/* Objective accumulators kept on the device (read once per epoch instead of 2 blocking D2H per bunch,
 * cuObjectiveFunction.cc:61-80).  Layout is ABI. */
typedef struct TnbObjStats_ {
  double error;      /* sum of -t*log(max(y,FLT_MIN))  (xent)  or  sum err^2 (mse)  */
  long long frames;  /* rows evaluated */
  long long correct; /* rows with argmax(y)==argmax(t) (xent only) */
} TnbObjStats;
/* CuSoftmax::PropagateFnc + CuCrossEntropy::Evaluate fused (cuActivation.cc:26-31, cuObjectiveFunction.cc:48-84):
 *   Y = softmax(A) (Y may be NULL), Err = Y - T, stats += {xent, rows, correct}.  `stats` is a DEVICE pointer.
 *   ONE TnbMatrixDim describes A, T, Y and Err: all four have d.rows rows and d.stride floats of pitch (the callers check, as the
 *   host mirror's CheckTargetLayout does; the same holds for tnb_xent_eval and tnb_mse_eval). */
int tnb_softmax_xent(TnbContext *ctx, const float *A, const float *T, float *Y, float *Err, TnbMatrixDim d, TnbObjStats *stats);
/* CuCrossEntropy::Evaluate on an existing softmax output */
int tnb_xent_eval(TnbContext *ctx, const float *Y, const float *T, float *Err, TnbMatrixDim d, TnbObjStats *stats);
/* The same two with the targets given as ONE CLASS ID PER ROW (labels[r * label_stride], int32) instead of a dense matrix — what
 * the reference builds its one-hot rows from (KaldiLib/Labels.cc:66,156).  The one-hot row is generated in registers, so the 4*cols
 * bytes per frame of the target read (12 kB at 3000 classes) and the dense matrix itself disappear; results are bit-identical to
 * the dense form on the one-hot matrix of the same ids.  An id outside [0, cols) stands for an all-zero target row. */
int tnb_softmax_xent_labels(TnbContext *ctx, const float *A, const int *labels, int label_stride, float *Y, float *Err, TnbMatrixDim d,
                            TnbObjStats *stats);
int tnb_xent_eval_labels(TnbContext *ctx, const float *Y, const int *labels, int label_stride, float *Err, TnbMatrixDim d, TnbObjStats *stats);
/* CuMeanSquareError::Evaluate (cuObjectiveFunction.cc:26-46; no 1/2 factor) */
int tnb_mse_eval(TnbContext *ctx, const float *Y, const float *T, float *Err, TnbMatrixDim d, TnbObjStats *stats);
/* dense one-hot targets from class ids (Labels.cc:66,156 builds them on the host): T[r, lab[r]] = 1 else 0 */
int tnb_onehot(TnbContext *ctx, float *T, const int *labels, TnbMatrixDim d);
/* the same with the ids label_stride ints apart (the [frames x 1] label column CuCache keeps: stride = its pitch) */
int tnb_onehot_strided(TnbContext *ctx, float *T, const int *labels, int label_stride, TnbMatrixDim d);

/* ---- data-parallel exchange (no counterpart in the reference GPU path; mirrors the CPU trainer's
 *      gradient reduce, TNetLib/Platform.h:300-335, BiasedLinearity.cc:90-178) ------------------------ */
#define TNB_NCCL_ID_BYTES 128
int tnb_comm_unique_id(unsigned char id[TNB_NCCL_ID_BYTES]);             /* rank 0 creates, caller broadcasts */
int tnb_comm_init(TnbContext *ctx, const unsigned char id[TNB_NCCL_ID_BYTES], int rank, int world);
int tnb_comm_destroy(TnbContext *ctx);
int tnb_comm_world(TnbContext *ctx, int *rank, int *world);
/* in-place sum over ranks of `count` floats, enqueued on the ctx's communication stream after everything
 * already enqueued on the compute stream; tnb_comm_wait() makes the compute stream wait for it. */
int tnb_allreduce_sum(TnbContext *ctx, float *buf, size_t count);
/* the same, additionally ordered behind `event` (e.g. a bias gradient computed on a side stream into the same buffer); `done`, if
 * not NULL, is recorded on the communication stream behind the all-reduce (a side stream can then apply the update) */
int tnb_allreduce_sum_ev(TnbContext *ctx, float *buf, size_t count, void *event, void *done);
/* n buffers in one NCCL launch (ncclGroupStart/End): ordered behind the compute stream and the given events, `done` recorded behind it */
int tnb_allreduce_sum_multi(TnbContext *ctx, float *const *bufs, const size_t *counts, int n, void *const *events, int n_events, void *done);
/* One layer's data-parallel update, enqueued on the communication stream behind everything already on the compute stream:
 * reduce-scatter G over the ranks (rank r receives the sum of rows [r*rows_pad/world, (r+1)*rows_pad/world)), apply
 * CuBiasedLinearity::Update to those rows of W/corrW, all-gather the updated rows of W; the bias gradient gb[ncols] is all-reduced
 * and applied by every rank.  G, W and corrW must have capacity for rows_pad rows (a multiple of the world size, >= dW.rows; the
 * rows beyond dW.rows of G must be zero).  tnb_comm_wait() afterwards makes the compute stream wait for it.  With a single rank
 * it is tnb_sgd_update. */
int tnb_dp_update(TnbContext *ctx, float *G, float *W, float *corrW, TnbMatrixDim dW, int rows_pad, float *gb, float *bias,
                  float *corrb, float lr, float mmt, float wc, int grad_div_frm, int n_frames_global);
int tnb_comm_wait(TnbContext *ctx);
/* Ranks as THREADS of one process, one GPU each (bin/TNetCu --GPUS=N): the group object is shared by the threads; each calls
 * tnb_comm_init_local with its context and rank (collective: returns when all have).  Peer access is enabled between the group's
 * GPUs, so the peer-memory schedule below works on plain device pointers (no NCCL, no CUDA IPC).  The NCCL schedules
 * (tnb_allreduce_sum*, tnb_dp_update) need tnb_comm_init instead. */
typedef struct TnbLocalGroup_ TnbLocalGroup;
int tnb_local_group_create(TnbLocalGroup **out, int world);
int tnb_local_group_destroy(TnbLocalGroup *group);
int tnb_comm_init_local(TnbContext *ctx, TnbLocalGroup *group, int rank);

/* ---- the same exchange without NCCL: one kernel per layer over NVLink peer memory (csrc/peer.cu) ------------------------------
 * One process per GPU on one NVSwitch box.  tnb_peer_map is collective over the ranks of tnb_comm_init: it exports `local` (the base
 * pointer of a device allocation of this library: tnb_malloc / tnb_malloc_pitch) with CUDA IPC, gathers the ranks' handles through
 * the communicator and maps the other ranks' allocations into this process; mapped[r] is rank r's buffer as addressable from this
 * GPU's kernels (mapped[rank] == local).  All ranks call it for the same buffers in the same order. */
#define TNB_MAX_PEERS 16
#define TNB_IPC_HANDLE_BYTES 64
int tnb_peer_map(TnbContext *ctx, void *local, void **mapped /* [world] */);
int tnb_peer_unmap(TnbContext *ctx, void *const *mapped);
typedef struct TnbPeerJob_ {
  float *G[TNB_MAX_PEERS]; /* every rank's [dW ; db] gradient buffer, (rows_pad + 1) rows of pitch dW.stride, as mapped into this process */
  float *W[TNB_MAX_PEERS]; /* every rank's weights [rows_pad x stride] */
  float *corrW;            /* this rank's momentum buffer [rows_pad x stride] */
  float *bias, *corrb;     /* this rank's bias and its momentum buffer [cols] (both NULL: no bias) */
  TnbMatrixDim dW;         /* logical weight matrix */
  int rows_pad;            /* multiple of the world size, >= dW.rows; rows beyond dW.rows are zero in G and W */
  float lr, mmt, wc;
  int grad_div_frm, n_frames; /* n_frames = frames of the GLOBAL bunch */
  int pushed;              /* 1: the ranks' gradients were PUSHED into the owners' staging slices by tnb_affine_grad_scatter: G[rank] is
                            * [world slices of rows_pad/world rows ; this rank's bias gradient row] and is summed from local memory */
} TnbPeerJob;
/* GEMM -> reduce-scatter in one kernel: the gradient GEMM of a data-parallel rank (cuBiasedLinearity.cc:55 without the update) whose
 * epilogue stores row block o of dW = X^T E into slice `rank` of rank o's staging buffer Gpeers[o] (peer memory, as returned by
 * tnb_peer_map for the ranks' gradient buffers: [(rows_pad + 1) x dG.stride] each), tile by tile while the MMAs of the next tile run
 * on the other SMs.  X16 / E16: bf16 twins (TNB_MATH_BF16) or NULL.  Pair with TnbPeerJob.pushed = 1. */
int tnb_affine_grad_scatter(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, const uint16_t *X16, int ldx16,
                            const uint16_t *E16, int lde16, float *const *Gpeers, int world, int rank, TnbMatrixDim dG, int rows_pad);
/* One layer's data-parallel update as ONE kernel on the communication stream, behind everything already on the compute stream and
 * behind `wait_event` (may be NULL): this rank sums its block of rows [rank*rows_pad/world, +rows_pad/world) of all ranks' gradients
 * through peer loads (rank order 0..world-1), applies CuBiasedLinearity::Update to that block of its corrW, and stores the updated
 * weights into EVERY rank's W; the bias row is summed and updated by every rank for itself.  Ranks synchronise through flag words in
 * peer memory at the start (all gradients complete, nobody reads W any more) and at the end (all blocks of this rank's W written)
 * of the kernel; `done_event` (may be NULL) is recorded behind it.  With one rank it is tnb_sgd_update on the compute stream. */
int tnb_dp_peer_update(TnbContext *ctx, const TnbPeerJob *job, void *wait_event, void *done_event);
/* The same kernel on the communication stream behind the given events only (nothing is recorded on the compute stream, so that
 * consecutive GEMMs there keep their programmatic-dependent-launch overlap); several ranks only. */
int tnb_dp_peer_update_after(TnbContext *ctx, const TnbPeerJob *job, void *const *wait_events, int n_wait, void *done_event);
/* Gradient push by the copy engines instead of by the GEMM epilogue: block o (rows [o*rows_pad/world, (o+1)*rows_pad/world)) of
 * this rank's full gradient G [rows_pad x dG.stride] goes to slice `rank` of rank o's staging buffer Gpeers[o] (tnb_peer_map of the
 * ranks' [(rows_pad + 1) x dG.stride] buffers), as `world` device-to-device copies on stream `stream_id` behind `wait_event`;
 * `done_event` is recorded behind them.  The copies are ordered behind everything enqueued on `stream_id` before the call (they may
 * run on per-destination streams of the context, which fork from and join `stream_id`).  No SM is involved and the gradient GEMM's
 * epilogue stores stay local. */
int tnb_peer_push_blocks(TnbContext *ctx, int stream_id, const float *G, float *const *Gpeers, int world, int rank, TnbMatrixDim dG,
                         int rows_pad, void *wait_event, void *done_event);
/* the same kernel for an explicit rank / world / flag blocks (64 zero-initialised words per rank) / sequence number (1, 2, ... per
 * launch) on a given stream of the context: lets a test play several ranks on one GPU */
int tnb_dp_peer_update_on(TnbContext *ctx, int stream_id, const TnbPeerJob *job, int rank, int world, unsigned *const *flags, unsigned seq);
/* ALL `world` ranks of one layer's exchange as ONE cooperative grid on this context's GPU (jobs[r] = rank r's job; CTAs
 * [r*c, (r+1)*c) play rank r, c = min(ctas_per_rank or the default, what is co-resident)): the way to run the multi-rank kernel on
 * a single GPU — separate launches that wait for each other's flags are not guaranteed to be co-resident.  Synchronous; returns
 * TNB_ERR_COMM if a rank's wait ran out of time. */
int tnb_dp_peer_update_virtual(TnbContext *ctx, const TnbPeerJob *jobs, int world, unsigned *const *flags, unsigned seq, int ctas_per_rank);
/* TNB_OK, or TNB_ERR_COMM with a message naming the rank and flag a peer-memory kernel of this context gave up waiting for (a wait
 * that exceeds TNB_PEER_TIMEOUT_MS ends the kernel without touching the weights and without poisoning the CUDA context; 0 = no
 * limit).  Synchronises nothing itself: call it after tnb_ctx_sync / an event wait.  tnb_ctx_sync calls it. */
int tnb_peer_status(TnbContext *ctx);
/* Debugging (TNB_DP_TRACE=1): %globaltimer stamps (ns).  out[0..255]: this context's last 64 peer-memory kernels, slot = sequence
 * number % 64: {entry, all ranks ready, CTA 0's rows done, all ranks done}; out[256..383]: the last 64 tnb_peer_push_blocks calls,
 * slot = call counter % 64: {copies start, copies done}; seq[0] = sequence number of the latest kernel, seq[1] = number of push calls. */
int tnb_peer_trace_read(TnbContext *ctx, long long *out /* [384] */, unsigned *seq /* [2] */);

#ifdef __cplusplus
}
#endif
#endif /* TNET_B200_H_ */
