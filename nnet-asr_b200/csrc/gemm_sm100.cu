// gemm_sm100.cu — the three CuBiasedLinearity contractions as one hand-written sm_100a kernel family.
//
// Replaces cublasSgemm as called from CuMatrix<float>::Gemm (reference: src/CuBaseLib/cumatrix.tcc:335-370):
//   forward  Y  = X * W            ('N','N')   cuBiasedLinearity.cc:15
//   dX       Ep = E * W^T          ('N','T')   cuBiasedLinearity.cc:24
//   dW       cW = X^T * E + m*cW   ('T','N')   cuBiasedLinearity.cc:55
// and fuses what the reference runs as separate kernels into the epilogue: bias row + sigmoid
// (cukernels.cu:100-119,194-206), diff-sigmoid (:211-217), momentum / learning-rate / L2 update (:89-95).
//
// Design (B200):
//   * operands stay fp32 in HBM, row-major with a 128-byte pitch; TMA (cp.async.bulk.tensor.2d, 128B swizzle)
//     stages 128 x 32 (A) and BN x 32 (B) fp32 tiles into shared memory.  Row-major operands that are
//     contracted over their ROW index (W in forward, X and E in dW) are loaded as MN-major tiles
//     (32-float column chunks) so no transposed copy ever exists in HBM.
//   * tcgen05.mma.kind::tf32, M=128, N=BN, K=8 per instruction, accumulator in TMEM (BN columns).
//   * 3xTF32 (default): the tensor core reads fp32 words as tf32 by TRUNCATION (measured on B200: tools/probe_tf32.py,
//     profiles/r01_tf32_probe.txt), so the staged fp32 tile itself is the `hi` operand; converter warps only write
//     lo = rna_tf32(x - trunc_tf32(x)) into a second buffer (x - trunc(x) is exact in fp32).  Per K-step the MMA warp
//     issues lo*hi, hi*lo, hi*hi.
//   * warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issue, warps2-9 = converters, then epilogue
//     (tcgen05.ld 32x32b -> registers -> fused epilogue -> global).
//   * every mbarrier spin is bounded: a protocol bug traps instead of hanging the GPU.
#include "common.cuh"
#include "gemm.cuh"

namespace tnb {

// ----------------------------------------------------------------------------------------------- SIMT cross-check
// plain fp32 FMA GEMM with the same epilogue (TNB_MATH_FP32_SIMT, and shapes the TMA path cannot take)
template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float *__restrict__ A, int lda, const float *__restrict__ B,
                                                        int ldb, int M, int N, int K, EpiParams ep) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      int kk, mm;
      if (TA) { kk = i / 64; mm = i % 64; } else { mm = i / 16; kk = i % 16; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.0f;
      if (gm < M && gk < K) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[kk][mm] = v;
      int nn;
      if (TB) { nn = i / 16; kk = i % 16; } else { kk = i / 64; nn = i % 64; }
      int gn = n0 + nn; gk = k0 + kk;
      v = 0.0f;
      if (gn < N && gk < K) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; i++) {
    int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      size_t ci = (size_t)row * ep.ldc + n;
      float *cb = ep.C;
      if (ep.scat_shard > 0) {
        const int own = row / ep.scat_shard;
        cb = ep.scat[own];
        ci = (size_t)(ep.scat_rank * ep.scat_shard + (row - own * ep.scat_shard)) * ep.ldc + n;
      }
      float cold = (ep.beta != 0.0f) ? cb[ci] : 0.0f;
      float bv = ep.bias ? ep.bias[n] : 0.0f;
      float yv = ep.mulY ? ep.mulY[(size_t)row * ep.ldy + n] : 0.0f;
      float o = epi_one(ep, acc[i][j], cold, bv, yv);
      if (ep.W && ep.c_wdecay != 0.0f) o = ep.c_wdecay * ep.W[(size_t)row * ep.ldw + n] + o;
      cb[ci] = o;
      if (ep.W) {
        float *wp = ep.W + (size_t)row * ep.ldw + n;
        float w = ep.w_scale * o + *wp;
        if (ep.w_l2 != 0.0f) w = ep.w_l2 * w + w;
        *wp = w;
      }
    }
  }
}

// How many CTAs of a cluster launch can be resident at once (clusters must sit inside one GPC, so this can be below the SM count).
static int cluster_capacity(TnbContext *ctx, int cluster) {
  static int cap[64][5] = {};
  int &c = cap[ctx->device & 63][cluster];
  if (c == 0) {
    c = ctx->sm_count - ctx->sm_count % cluster;
    int n = 0;
    const int rc = (cluster == 4) ? tc_max_active_clusters<256, 3, 2, 2>(&n) : tc_max_active_clusters<256, 3, 2, 1>(&n);
    if (rc == TNB_OK && n > 0) c = n * cluster;
    if (getenv("TNB_GEMM_DEBUG")) fprintf(stderr, "[tnb] cluster size %d: %d co-resident CTAs\n", cluster, c);
  }
  return c;
}

// C[M x N] (+epilogue) = op(A) * op(B); A, B row-major as CuMatrix::Gemm receives them.
int launch_gemm(TnbContext *ctx, char transa, char transb, int M, int N, int K, const float *A, int lda, const float *B,
                int ldb, const EpiParams &ep) {
  TNB_ARG(ctx != nullptr, "null ctx");
  TNB_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  TNB_ARG(A && B && ep.C, "null operand");
  const int ta = (transa == 'T' || transa == 't'), tb = (transb == 'T' || transb == 't');
  TNB_ARG(ta || transa == 'N' || transa == 'n', "transa");
  TNB_ARG(tb || transb == 'N' || transb == 'n', "transb");
  const bool vec_ok = ((uintptr_t)ep.C % 16 == 0) && (ep.ldc % 4 == 0) && (!ep.bias || (uintptr_t)ep.bias % 16 == 0) &&
                      (!ep.mulY || ((uintptr_t)ep.mulY % 16 == 0 && ep.ldy % 4 == 0)) &&
                      (!ep.W || ((uintptr_t)ep.W % 16 == 0 && ep.ldw % 4 == 0));
  const bool tma_ok = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && vec_ok;
  if (ctx->math_mode == TNB_MATH_FP32_SIMT || !tma_ok) {
    // Sub-matrix views at column offsets that are not multiples of 4 floats cannot be described to TMA (16-byte alignment):
    // CuMath::OffsetGemm / <blocklinearity> slice a [T x 51*23] matrix into 51-column blocks (cumath.cc:88-113, example 01's
    // Hamm_dct_norm).  Those GEMMs are tiny feature-transform work; they run on the plain fp32 FMA kernel in every math mode
    // (exact fp32 products, i.e. at least as accurate as the tensor-core modes).
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    if (!ta && !tb) gemm_simt_kernel<0, 0><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else if (!ta && tb) gemm_simt_kernel<0, 1><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else if (ta && !tb) gemm_simt_kernel<1, 0><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else gemm_simt_kernel<1, 1><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    TNB_LAUNCHED(ctx);
    return TNB_OK;
  }
  if (ctx->math_mode == TNB_MATH_BF16) {
    // fp32 operands in bf16 mode: round both to bf16 in ctx scratch, then the bf16 tensor-core path (the fused layer ops have
    // *_bf16 entry points that take resident bf16 twins instead)
    uint16_t *a16 = nullptr, *b16 = nullptr;
    int lda16 = 0, ldb16 = 0;
    int rc = bf16_scratch(ctx, 0, A, ta ? K : M, ta ? M : K, lda, &a16, &lda16);
    if (rc != TNB_OK) return rc;
    rc = bf16_scratch(ctx, 1, B, tb ? N : K, tb ? K : N, ldb, &b16, &ldb16);
    if (rc != TNB_OK) return rc;
    return launch_gemm_bf16(ctx, transa, transb, M, N, K, a16, lda16, b16, ldb16, ep);
  }
  // operand majors: op(A)=A  -> A is [M x K], contraction contiguous -> K-major ; op(A)=A^T -> A is [K x M] -> MN-major
  //                 op(B)=B  -> B is [K x N], N contiguous -> MN-major         ; op(B)=B^T -> B is [N x K] -> K-major
  const int a_mn = ta ? 1 : 0;
  const int b_mn = tb ? 0 : 1;
  // Tile choice over {1 CTA, CTA pair} x BN: minimise waves * tile_time.  tile_time is the shared-memory-bandwidth model of
  // DESIGN.md 3.1 — bytes through one SM's smem per K block (TMA fill + converter read/write + operand reads of the MMAs, at
  // 128 B/clk) — plus a fixed prologue and an epilogue proportional to BN.  A pair stages only half of B per SM.
  const bool three = ctx->math_mode == TNB_MATH_3XTF32;
  const int num_kb = (K + BK - 1) / BK;
  int bn = 128, cg = 1, split = 1;
  double best = 1e300;
  static int force_split = -1;
  if (force_split < 0) { const char *e = getenv("TNB_GEMM_SPLIT"); force_split = e ? atoi(e) : 0; }  // 1 / 2 force (debugging)
  static int force_cg = -1;
  if (force_cg < 0) { const char *e = getenv("TNB_GEMM_CG"); force_cg = e ? atoi(e) : 0; }  // 1 / 2 force a mode (debugging)
  const int cands[4] = {64, 128, 192, 256};
  static int force_bn = -1;
  if (force_bn < 0) { const char *e = getenv("TNB_GEMM_BN"); force_bn = e ? atoi(e) : 0; }
  for (int g = 1; g <= 2; g++) {
    if (force_cg && g != force_cg) continue;
    if (g == 2 && M <= BM && !force_cg) continue;  // a pair needs two 128-row blocks
    if (g == 2 && !three) continue;   // single-pass tf32 is L2/latency-bound: pairs measured 4 % slower there (not instantiated)
    for (int ci = 0; ci < 4; ci++) {
      const int c = cands[ci];
      if (force_bn && c != force_bn) continue;
      if (g == 2 && c == 64) continue;
      if (c > 64 && N <= c / 2) continue;  // do not pad N by more than 2x
      int mt = (M + BM - 1) / BM;
      if (g == 2) mt = (mt + 1) & ~1;
      const double a_b = 16.0 * 1024, b_b = 128.0 * c / g;        // staged bytes per K block
      const double mma_reads = (three ? 12.0 : 4.0) * (4096.0 + 32.0 * c / g);
      const double smem_bytes = (three ? 3.0 : 1.0) * (a_b + b_b) + mma_reads;
      const int stage_bytes = (int)((a_b + b_b) * (three ? 2 : 1));
      for (int sp = 1; sp <= 2; sp++) {
        // split-K: two pairs share one output tile, each over half of K, and swap half an accumulator through DSMEM at the end
        if (force_split == 1 && sp == 2) continue;
        if (force_split == 2 && sp == 1 && g == 2 && three && c != 64 && num_kb >= 8) continue;
        if (sp == 2 && (g != 2 || !three || c == 64 || num_kb < 8)) continue;
        const long ctas = (long)mt * ((N + c - 1) / c) * sp;
        const long cap = (g * sp == 1) ? ctx->sm_count : cluster_capacity(ctx, g * sp);
        const long waves = (ctas + cap - 1) / cap;
        double t = ((num_kb + sp - 1) / sp) * smem_bytes / 128.0 + 4000.0 + 30.0 * c / sp + (sp == 2 ? 2500.0 + 8.0 * c : 0.0);
        if ((196 * 1024) / stage_bytes < 3) t *= 1.05;  // only 2 pipeline stages fit
        const double cost = waves * t;
        if (cost < best * 0.999) { best = cost; bn = c; cg = g; split = sp; }
      }
    }
  }
  const int bh = bn / cg;  // B rows staged per CTA = TMA box height of a K-major B
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = get_tmap(ctx, A, M, K, lda, BM, BK, 0, &tmA);   // rows = m, cols = k, box 128 x 32
  else rc = get_tmap(ctx, A, K, M, lda, BK, 32, 1, &tmA);         // rows = k, cols = m, box 32 x 32
  if (rc != TNB_OK) return rc;
  if (!b_mn) rc = get_tmap(ctx, B, N, K, ldb, bh, BK, 0, &tmB);   // rows = n, cols = k, box BH x 32
  else rc = get_tmap(ctx, B, K, N, ldb, BK, 32, 1, &tmB);         // rows = k, cols = n, box 32 x 32
  if (rc != TNB_OK) return rc;
  if (getenv("TNB_GEMM_DEBUG"))
    fprintf(stderr, "[tnb] gemm %c%c M=%d N=%d K=%d -> BN=%d CG=%d SPLIT=%d\n", transa, transb, M, N, K, bn, cg, split);
  const int nt = three ? 3 : 1;
#define TNB_TRY(BN_, NT_, CG_, SP_) \
  if (bn == BN_ && nt == NT_ && cg == CG_ && split == SP_) return launch_tc_major<BN_, NT_, CG_, SP_>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  TNB_TRY(256, 3, 2, 2) TNB_TRY(192, 3, 2, 2) TNB_TRY(128, 3, 2, 2)
  TNB_TRY(256, 3, 2, 1) TNB_TRY(192, 3, 2, 1) TNB_TRY(128, 3, 2, 1)
  TNB_TRY(256, 3, 1, 1) TNB_TRY(192, 3, 1, 1) TNB_TRY(128, 3, 1, 1) TNB_TRY(64, 3, 1, 1)
  TNB_TRY(256, 1, 1, 1) TNB_TRY(192, 1, 1, 1) TNB_TRY(128, 1, 1, 1) TNB_TRY(64, 1, 1, 1)
#undef TNB_TRY
  set_error("no GEMM instance for BN=%d terms=%d CG=%d SPLIT=%d", bn, nt, cg, split);
  return TNB_ERR_UNSUPPORTED;
}

// ----------------------------------------------------------------------------------------------- bf16 operands
// Same kernel family with NTERMS = 16: 64-deep K blocks of bf16, tcgen05.mma kind::f16, no converter warps.  Tile choice: the
// MMA needs 2*BN cycles per K block (4 instructions of 128 x BN x 16 per SM at 8192 flop/clk), the operand fetch
// (16 KB + 128 B * BH per CTA and K block) arrives at about 42 B/clk/SM when every SM pulls from L2 at once (LTS cap / 148).
int launch_gemm_bf16(TnbContext *ctx, char transa, char transb, int M, int N, int K, const uint16_t *A, int lda,
                     const uint16_t *B, int ldb, const EpiParams &ep) {
  TNB_ARG(ctx != nullptr, "null ctx");
  TNB_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  TNB_ARG(A && B && ep.C, "null operand");
  const int ta = (transa == 'T' || transa == 't'), tb = (transb == 'T' || transb == 't');
  TNB_ARG(ta || transa == 'N' || transa == 'n', "transa");
  TNB_ARG(tb || transb == 'N' || transb == 'n', "transb");
  TNB_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && (lda % 8 == 0) && (ldb % 8 == 0),
          "bf16 operands must be 16-byte aligned with a pitch multiple of 8 elements");
  TNB_ARG(((uintptr_t)ep.C % 16 == 0) && (ep.ldc % 4 == 0) && (!ep.bias || (uintptr_t)ep.bias % 16 == 0) &&
              (!ep.mulY || ((uintptr_t)ep.mulY % 16 == 0 && ep.ldy % 4 == 0)) && (!ep.W || ((uintptr_t)ep.W % 16 == 0 && ep.ldw % 4 == 0)) &&
              (!ep.C16 || ((uintptr_t)ep.C16 % 8 == 0 && ep.ldc16 % 4 == 0)) && (!ep.W16 || ((uintptr_t)ep.W16 % 8 == 0 && ep.ldw16 % 4 == 0)),
          "epilogue arrays must be 16-byte aligned with a pitch multiple of 4 elements");
  const int a_mn = ta ? 1 : 0, b_mn = tb ? 0 : 1;
  const int num_kb = (K + 63) / 64;
  static int force_bn = -1, force_cg = -1, force_split = -1;
  if (force_bn < 0) { const char *e = getenv("TNB_GEMM_BN"); force_bn = e ? atoi(e) : 0; }
  if (force_cg < 0) { const char *e = getenv("TNB_GEMM_CG"); force_cg = e ? atoi(e) : 0; }
  if (force_split < 0) { const char *e = getenv("TNB_GEMM_SPLIT"); force_split = e ? atoi(e) : 0; }
  // experiment knob: extra cycles charged to a split-K choice (the model's per-K-block fetch term is pessimistic for bf16 tiles, which
  // may hide that 256x128 pairs without the accumulator exchange are as fast as split 256x256 ones)
  static double split_penalty = -1.0;
  if (split_penalty < 0.0) { const char *e = getenv("TNB_GEMM_BF16_SPLIT_PENALTY"); split_penalty = e ? atof(e) : 0.0; }
  int bn = 128, cg = 1, split = 1;
  double best = 1e300;
  const int cands[3] = {64, 128, 256};
  for (int g = 1; g <= 2; g++) {
    if (force_cg && g != force_cg) continue;
    if (g == 2 && M <= BM && !force_cg) continue;
    for (int ci = 0; ci < 3; ci++) {
      const int c = cands[ci];
      if (force_bn && c != force_bn) continue;
      if (g == 2 && c == 64) continue;
      if (c > 64 && N <= c / 2) continue;
      int mt = (M + BM - 1) / BM;
      if (g == 2) mt = (mt + 1) & ~1;
      const double fetch = (16384.0 + 128.0 * c / g) / 42.0, mma = 2.0 * c;
      for (int sp = 1; sp <= 2; sp++) {
        if (force_split == 1 && sp == 2) continue;
        if (sp == 2 && (g != 2 || num_kb < 8)) continue;
        if (force_split == 2 && sp == 1 && g == 2 && num_kb >= 8) continue;
        const long ctas = (long)mt * ((N + c - 1) / c) * sp;
        const long cap = (g * sp == 1) ? ctx->sm_count : cluster_capacity(ctx, g * sp);
        const long waves = (ctas + cap - 1) / cap;
        const double t = ((num_kb + sp - 1) / sp) * (fetch > mma ? fetch : mma) + 4000.0 + 30.0 * c / sp + (sp == 2 ? 2500.0 + 8.0 * c + split_penalty : 0.0);
        const double cost = waves * t;
        if (cost < best * 0.999) { best = cost; bn = c; cg = g; split = sp; }
      }
    }
  }
  const int bh = bn / cg;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = get_tmap(ctx, A, M, K, lda, BM, 64, 0, &tmA, 2);   // rows = m, cols = k, box 128 x 64
  else rc = get_tmap(ctx, A, K, M, lda, 64, 64, 0, &tmA, 2);         // rows = k, cols = m, box 64 x 64
  if (rc != TNB_OK) return rc;
  if (!b_mn) rc = get_tmap(ctx, B, N, K, ldb, bh, 64, 0, &tmB, 2);   // rows = n, cols = k, box BH x 64
  else rc = get_tmap(ctx, B, K, N, ldb, 64, 64, 0, &tmB, 2);         // rows = k, cols = n, box 64 x 64
  if (rc != TNB_OK) return rc;
  if (getenv("TNB_GEMM_DEBUG"))
    fprintf(stderr, "[tnb] gemm bf16 %c%c M=%d N=%d K=%d -> BN=%d CG=%d SPLIT=%d\n", transa, transb, M, N, K, bn, cg, split);
#define TNB_TRY(BN_, CG_, SP_) \
  if (bn == BN_ && cg == CG_ && split == SP_) return launch_tc_major<BN_, 16, CG_, SP_>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  TNB_TRY(256, 2, 2) TNB_TRY(128, 2, 2) TNB_TRY(256, 2, 1) TNB_TRY(128, 2, 1)
  TNB_TRY(256, 1, 1) TNB_TRY(128, 1, 1) TNB_TRY(64, 1, 1)
#undef TNB_TRY
  set_error("no bf16 GEMM instance for BN=%d CG=%d SPLIT=%d", bn, cg, split);
  return TNB_ERR_UNSUPPORTED;
}

// ----------------------------------------------------------------------------------------------- gemv / ger
// reference: CuMath<float>::OffsetGemv (cumath.cc:283-340), used only by CuRecurrent (batch-1, frame-serial)
__global__ void gemv_n_kernel(float alpha, const float *__restrict__ A, int lda, int row_off, int nrows, int ncols,
                              const float *__restrict__ x, float beta, float *y) {
  // y[r] = alpha * sum_c A[row_off + r, c] * x[c] + beta*y[r]   (one warp per row)
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const float *a = A + (size_t)(row_off + r) * lda;
  float s = 0.0f;
  for (int c = lane; c < ncols; c += 32) s = fmaf(a[c], x[c], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[r] = alpha * s + (beta == 0.0f ? 0.0f : beta * y[r]);
}
__global__ void gemv_t_kernel(float alpha, const float *__restrict__ A, int lda, int col_off, int nrows, int ncols,
                              const float *__restrict__ x, float beta, float *y) {
  // y[c] = alpha * sum_r A[r, col_off + c] * x[r] + beta*y[c]   (one thread per column, coalesced over c)
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  float s = 0.0f;
  for (int r = 0; r < nrows; r++) s = fmaf(A[(size_t)r * lda + col_off + c], x[r], s);
  y[c] = alpha * s + (beta == 0.0f ? 0.0f : beta * y[c]);
}
__global__ void ger_kernel(float alpha, const float *__restrict__ x, int dimX, const float *__restrict__ y, int dimY, float *A,
                           int lda) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c < dimY && r < dimX) A[(size_t)r * lda + c] += (alpha * x[r]) * y[c];
}

// ---- CuRecurrent::Update, fused (cuRecurrent.cc:92-153) --------------------------------------------------------------------
// One step of the BPTT chain in ONE launch instead of five (gemv, diff-sigmoid, rank-1 update, two column-sum kernels):
//   e[r]    = sum_c W[nin + r, c] * d_prev[c]                    (OffsetGemv 'N' on the recurrent rows, warp per row as gemv_n_kernel)
//   d[r]    = (float)(y[r] * (1 - y[r]) * e[r])  in double        (DiffSigmoid with the history frame's activations)
//   bcorr[r] = (float)(-lr * d[r] + bcorr[r])    in double        (AddColSum(-lr, d, 1.0) of a one-row matrix)
// The rank-1 updates of all steps are applied together by rnn_apply_kernel from the stored d vectors.
__global__ void __launch_bounds__(256) rnn_bptt_step_kernel(const float *__restrict__ W, int ldw, int nin, int H, const float *__restrict__ d_prev,
                                                            const float *__restrict__ y, float *__restrict__ d_out, float *__restrict__ bcorr,
                                                            float neg_lr) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= H) return;
  const float *a = W + (size_t)(nin + r) * ldw;
  float s = 0.0f;
  for (int c = lane; c < H; c += 32) s = fmaf(a[c], d_prev[c], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double yy = (double)y[r];
    const float d = (float)(yy * (1.0 - yy) * (double)s);
    d_out[r] = d;
    bcorr[r] = (float)((double)neg_lr * (double)d + (double)bcorr[r]);
  }
}
// W[k,h] += sum_i (-lr * hist[i][k]) * d[i][h] + (-lr*wc) * W[k,h], the sum in the order of the reference's sequence of rank-1
// updates into a zeroed correction matrix (i = 0 .. nsteps-1), then corr += l2*W, then W += corr: one pass over W instead of
// nsteps + 3 (BlasGer x nsteps, SetConst, AddScaled x 2).
__global__ void __launch_bounds__(256) rnn_apply_kernel(float *__restrict__ W, int ldw, int K, int H, const float *__restrict__ hist, int ldh,
                                                        const float *__restrict__ d, int ldd, int nsteps, float neg_lr, float l2) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int k = blockIdx.y;
  if (c >= H || k >= K) return;
  float corr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int i = 0; i < nsteps; i++) {
    const float t = neg_lr * hist[(size_t)i * ldh + k];
    const float *di = d + (size_t)i * ldd + c;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (c + j < H) corr[j] = fmaf(t, di[j], corr[j]);
  }
  float *w = W + (size_t)k * ldw + c;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (c + j < H) {
      const float cj = fmaf(l2, w[j], corr[j]);  // corr = l2*W + 1*corr
      w[j] = cj + w[j];                          // W = 1*corr + 1*W
    }
}

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_rnn_bptt_step(TnbContext *ctx, const float *W, TnbMatrixDim dW, int nin, const float *d_prev, const float *y_hist, float *d_out,
                      float *bcorr, float lr) {
  TNB_ARG(ctx && W && d_prev && y_hist && d_out && bcorr, "null");
  const int H = dW.cols;
  TNB_ARG(nin >= 0 && dW.rows == nin + H && dW.stride >= H, "W must be [(nin + H) x H]");
  rnn_bptt_step_kernel<<<(H + 7) / 8, 256, 0, ctx->stream>>>(W, dW.stride, nin, H, d_prev, y_hist, d_out, bcorr, -lr);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_rnn_apply(TnbContext *ctx, float *W, TnbMatrixDim dW, const float *hist, int ld_hist, const float *d, int ld_d, int nsteps, float lr,
                  float wc) {
  TNB_ARG(ctx && W && hist && d, "null");
  TNB_ARG(nsteps >= 1 && dW.rows > 0 && dW.cols > 0 && ld_hist >= dW.rows && ld_d >= dW.cols, "dims");
  dim3 grid(((dW.cols + 3) / 4 + 255) / 256, dW.rows);
  rnn_apply_kernel<<<grid, 256, 0, ctx->stream>>>(W, dW.stride, dW.rows, dW.cols, hist, ld_hist, d, ld_d, nsteps, -lr, -lr * wc);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_gemm(TnbContext *ctx, char transa, char transb, int m, int n, int k, float alpha, const float *A, int lda,
             const float *B, int ldb, float beta, float *C, int ldc) {
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = C; ep.ldc = ldc; ep.alpha = alpha; ep.beta = beta;
  return launch_gemm(ctx, transa, transb, m, n, k, A, lda, B, ldb, ep);
}

int tnb_affine_fwd(TnbContext *ctx, const float *X, TnbMatrixDim dX, const float *W, TnbMatrixDim dW, const float *bias,
                   float *Y, TnbMatrixDim dY, int act) {
  TNB_ARG(ctx && X && W && bias && Y, "null");
  TNB_ARG(dX.cols == dW.rows && dY.cols == dW.cols && dY.rows == dX.rows, "dimension mismatch");
  TNB_ARG(act == TNB_ACT_NONE || act == TNB_ACT_SIGMOID, "act");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = Y; ep.ldc = dY.stride; ep.alpha = 1.0f; ep.beta = 0.0f; ep.bias = bias; ep.act = act; ep.mode = EPI_FWD;
  return launch_gemm(ctx, 'N', 'N', dX.rows, dW.cols, dX.cols, X, dX.stride, W, dW.stride, ep);
}

int tnb_affine_bwd_dx(TnbContext *ctx, const float *E, TnbMatrixDim dE, const float *W, TnbMatrixDim dW, const float *Yprev,
                      TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev) {
  TNB_ARG(ctx && E && W && Eprev, "null");
  TNB_ARG(dE.cols == dW.cols && dEprev.cols == dW.rows && dEprev.rows == dE.rows, "dimension mismatch");
  if (Yprev) TNB_ARG(dYprev.rows == dEprev.rows && dYprev.cols == dEprev.cols, "Yprev dims");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = Eprev; ep.ldc = dEprev.stride; ep.alpha = 1.0f; ep.beta = 0.0f;
  ep.mulY = Yprev; ep.ldy = dYprev.stride;
  if (Yprev) ep.mode = EPI_DX;
  return launch_gemm(ctx, 'N', 'T', dE.rows, dW.rows, dE.cols, E, dE.stride, W, dW.stride, ep);
}

int tnb_affine_fwd_bf16(TnbContext *ctx, const uint16_t *X16, int ldx16, TnbMatrixDim dX, const uint16_t *W16, int ldw16,
                        TnbMatrixDim dW, const float *bias, float *Y, TnbMatrixDim dY, uint16_t *Y16, int ldy16, int act) {
  TNB_ARG(ctx && X16 && W16 && bias && Y, "null");
  TNB_ARG(dX.cols == dW.rows && dY.cols == dW.cols && dY.rows == dX.rows, "dimension mismatch");
  TNB_ARG(act == TNB_ACT_NONE || act == TNB_ACT_SIGMOID, "act");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = Y; ep.ldc = dY.stride; ep.alpha = 1.0f; ep.beta = 0.0f; ep.bias = bias; ep.act = act; ep.mode = EPI_FWD;
  ep.C16 = Y16; ep.ldc16 = ldy16;
  return launch_gemm_bf16(ctx, 'N', 'N', dX.rows, dW.cols, dX.cols, X16, ldx16, W16, ldw16, ep);
}

int tnb_affine_bwd_dx_bf16(TnbContext *ctx, const uint16_t *E16, int lde16, TnbMatrixDim dE, const uint16_t *W16, int ldw16,
                           TnbMatrixDim dW, const float *Yprev, TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev,
                           uint16_t *Eprev16, int ldep16) {
  TNB_ARG(ctx && E16 && W16 && Eprev, "null");
  TNB_ARG(dE.cols == dW.cols && dEprev.cols == dW.rows && dEprev.rows == dE.rows, "dimension mismatch");
  if (Yprev) TNB_ARG(dYprev.rows == dEprev.rows && dYprev.cols == dEprev.cols, "Yprev dims");
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.C = Eprev; ep.ldc = dEprev.stride; ep.alpha = 1.0f; ep.beta = 0.0f;
  ep.mulY = Yprev; ep.ldy = dYprev.stride;
  if (Yprev) ep.mode = EPI_DX;
  ep.C16 = Eprev16; ep.ldc16 = ldep16;
  return launch_gemm_bf16(ctx, 'N', 'T', dE.rows, dW.rows, dE.cols, E16, lde16, W16, ldw16, ep);
}

int tnb_offset_gemv(TnbContext *ctx, char trans, float alpha, const float *A, TnbMatrixDim dA, const float *x, int dimX,
                    float beta, float *y, int dimY, int offsetY) {
  TNB_ARG(ctx && A && x && y, "null");
  if (trans == 'N' || trans == 'n') {
    // y[dimY] = A[offsetY : offsetY+dimY, :] * x[dA.cols]
    TNB_ARG(dimX == dA.cols && dA.rows >= dimY + offsetY, "gemv N dims");
    int wpb = 8;
    gemv_n_kernel<<<(dimY + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(alpha, A, dA.stride, offsetY, dimY, dA.cols, x, beta, y);
  } else if (trans == 'T' || trans == 't') {
    // y[dimY] = A[:, offsetY : offsetY+dimY]^T * x[dA.rows]
    TNB_ARG(dimX == dA.rows && dA.cols >= dimY + offsetY, "gemv T dims");
    gemv_t_kernel<<<(dimY + 127) / 128, 128, 0, ctx->stream>>>(alpha, A, dA.stride, offsetY, dA.rows, dimY, x, beta, y);
  } else {
    TNB_ARG(false, "trans");
  }
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

int tnb_ger(TnbContext *ctx, float alpha, const float *x, int dimX, const float *y, int dimY, float *A, TnbMatrixDim dA) {
  TNB_ARG(ctx && x && y && A, "null");
  TNB_ARG(dimX == dA.rows && dimY == dA.cols, "ger dims");
  dim3 grid((dimY + 255) / 256, dimX);
  ger_kernel<<<grid, 256, 0, ctx->stream>>>(alpha, x, dimX, y, dimY, A, dA.stride);
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

}  // extern "C"
