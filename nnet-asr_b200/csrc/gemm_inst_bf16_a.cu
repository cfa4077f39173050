// gemm_inst_bf16_a.cu — explicit instantiations of the bf16-operand GEMM tile shapes, single CTA (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace tnb {
TNB_GEMM_INSTANTIATE(64, 16, 1, 1)
TNB_GEMM_INSTANTIATE(128, 16, 1, 1)
TNB_GEMM_INSTANTIATE(256, 16, 1, 1)
}  // namespace tnb

TNB_GEMM_TRACE_READERS(bf16_a)
