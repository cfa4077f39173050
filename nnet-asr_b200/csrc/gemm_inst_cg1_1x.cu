// gemm_inst_cg1_1x.cu — explicit instantiations of one group of GEMM tile shapes (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace tnb {
TNB_GEMM_INSTANTIATE(64, 1, 1, 1)
TNB_GEMM_INSTANTIATE(128, 1, 1, 1)
TNB_GEMM_INSTANTIATE(192, 1, 1, 1)
TNB_GEMM_INSTANTIATE(256, 1, 1, 1)
}  // namespace tnb

TNB_GEMM_TRACE_READERS(cg1_1x)
