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
};

// C[M x N] (+epilogue) = op(A) * op(B); A, B row-major exactly as CuMatrix::Gemm receives them
// (reference: src/CuBaseLib/cumatrix.tcc:335-370).
int launch_gemm(TnbContext *ctx, char transa, char transb, int M, int N, int K, const float *A, int lda, const float *B,
                int ldb, const EpiParams &ep);

}  // namespace tnb
