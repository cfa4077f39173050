// gemm_inst_split.cu — explicit instantiations of one group of GEMM tile shapes (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace tnb {
TNB_GEMM_INSTANTIATE(128, 3, 2, 2)
TNB_GEMM_INSTANTIATE(192, 3, 2, 2)
TNB_GEMM_INSTANTIATE(256, 3, 2, 2)
}  // namespace tnb

TNB_GEMM_TRACE_READERS(split)
