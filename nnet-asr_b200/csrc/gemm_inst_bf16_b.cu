// gemm_inst_bf16_b.cu — explicit instantiations of the bf16-operand GEMM tile shapes, CTA pairs and split-K (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace tnb {
TNB_GEMM_INSTANTIATE(128, 16, 2, 1)
TNB_GEMM_INSTANTIATE(256, 16, 2, 1)
TNB_GEMM_INSTANTIATE(128, 16, 2, 2)
TNB_GEMM_INSTANTIATE(256, 16, 2, 2)
}  // namespace tnb

TNB_GEMM_TRACE_READERS(bf16_b)
