// gemm.cuh — epilogue description shared by the GEMM launcher and the fused layer ops
#pragma once
#include "common.cuh"

namespace tnb {

struct EpiParams {
  float *C; int ldc;
  float alpha, beta;            // out = alpha*acc + beta*C_old
  const float *bias;            // out += bias[n]
  int act;                      // TNB_ACT_*
  const float *mulY; int ldy;   // out *= y*(1-y)              (CuSigmoid::BackpropagateFnc)
  float *W; int ldw;            // fused SGD: W += w_scale*out ; W += w_l2*W
  float w_scale, w_l2;
  float c_wdecay;               // generic epilogue only: out += c_wdecay * W_old before the store and the W update (CD-1: corr += -lr*wc*W)
  uint16_t *C16; int ldc16;     // optional bf16 copy of the stored output (TNB_MATH_BF16 shadows)
  uint16_t *W16; int ldw16;     // optional bf16 copy of the updated weights
  // generic epilogue only (data-parallel gradient GEMM): row i of the output goes to scat[i / scat_shard], at row
  // scat_rank * scat_shard + i % scat_shard — the block owner's staging slice for this rank (peer memory); scat_shard == 0: plain C
  float *scat[TNB_MAX_PEERS];
  int scat_shard, scat_rank;
  float *xchg;                  // split-K kernels: L2 exchange buffer of the launch (set by the launcher; NULL = exchange through DSMEM)
  int mode;                     // EPI_*: which specialised epilogue the fused entry point asks for (EPI_GENERIC = any combination)
};

// Specialised epilogues (template parameter of the kernel).  The generic one evaluates every field of EpiParams at run time; the
// three fused layer ops get compact code paths whose global reads are issued in one batch before the accumulator is touched.
enum { EPI_GENERIC = 0, EPI_FWD = 1 /* C = act(acc + bias) */, EPI_DX = 2 /* C = acc .* y(1-y) */,
       EPI_UPD = 3 /* C = acc + beta*C ; W += s*C ; W += l2*W */ };

__device__ __forceinline__ float sigmoidf_ref(float x) {
  // reference: 1.0/(1.0+exp(-x)) with a float exp and a double divide rounded to float (cukernels.cu:194-206).
  // Branch-free on purpose: an IEEE division carries a slow-path branch per element, which splits the unrolled epilogue into
  // basic blocks the scheduler cannot interleave (the sigmoid was 4000 of the 5400 cycles per 32-column chunk).  Full-precision
  // expf, then the reciprocal as MUFU.RCP + one Newton step: within 1 ulp of the correctly rounded quotient.  The clamp keeps
  // 1+e finite (e <= 6.1e37) so the Newton step is well defined; sigmoid(-87) = 1.6e-38 is already below any tolerance.
  const float d = 1.0f + expf(-fmaxf(x, -87.0f));
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}

__device__ __forceinline__ float epi_one(const EpiParams &ep, float acc, float cold, float bias, float y) {
  float o = ep.alpha * acc;
  if (ep.beta != 0.0f) o += ep.beta * cold;
  o += bias;
  if (ep.act == TNB_ACT_SIGMOID) o = sigmoidf_ref(o);
  if (ep.mulY) o = (y * (1.0f - y)) * o;
  return o;
}

constexpr int BM = 128;  // rows of the output tile one CTA owns
constexpr int BK = 32;   // fp32 elements per K block = 128 bytes = one swizzle span
constexpr int CONV_WARPS = 8;                       // converter / epilogue warps
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int GEMM_THREADS = 64 + CONV_THREADS;     // + TMA warp + MMA warp

// Defined in gemm_kernel.cuh and explicitly instantiated, tile shape by tile shape, in the gemm_inst_*.cu translation units.
// BN: tile width; NTERMS: 3 = 3xTF32, 1 = single tf32 pass, 16 = bf16 operands (16-bit arrays in HBM); CG: CTAs per tile (2 = tcgen05 cta_group::2 pair); SPLIT: split-K factor.
template <int BN, int NTERMS, int CG, int SPLIT>
int launch_tc_major(TnbContext *ctx, int a_mn, int b_mn, const CUtensorMap &tmA, const CUtensorMap &tmB, int M, int N, int K,
                    const EpiParams &ep);
template <int BN, int NTERMS, int CG, int SPLIT>
int tc_max_active_clusters(int *clusters);

// C[M x N] (+epilogue) = op(A) * op(B); A, B row-major exactly as CuMatrix::Gemm receives them
// (reference: src/CuBaseLib/cumatrix.tcc:335-370).
int launch_gemm(TnbContext *ctx, char transa, char transb, int M, int N, int K, const float *A, int lda, const float *B,
                int ldb, const EpiParams &ep);
// the same contraction with bf16 operands (row-major bf16 arrays, pitch in elements a multiple of 8; TMA + kind::f16 MMA)
int launch_gemm_bf16(TnbContext *ctx, char transa, char transb, int M, int N, int K, const uint16_t *A, int lda,
                     const uint16_t *B, int ldb, const EpiParams &ep);
// A16 = bf16_rn(A) for a [rows x cols] fp32 matrix (elementwise.cu)
int launch_to_bf16(TnbContext *ctx, uint16_t *dst, int dst_stride, const float *src, int rows, int cols, int src_stride);
// split-K accumulator exchange buffer of the stream the launch goes to (ctx-owned, grown on demand; *out = NULL when it cannot be
// provided — inside a graph capture before the first allocation — and the kernel then exchanges through DSMEM)
int xchg_buffer(TnbContext *ctx, cudaStream_t stream, size_t bytes, float **out);
// ctx-owned bf16 scratch copies of fp32 operands (generic entry points in TNB_MATH_BF16): slot 0 / 1
int bf16_scratch(TnbContext *ctx, int slot, const float *src, int rows, int cols, int stride, uint16_t **out, int *out_stride);

}  // namespace tnb
