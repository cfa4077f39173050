// store_bench.cu — how fast can ONE CTA per SM drain a 128 KB fp32 tile to global memory?  (epilogue design probe)
// variants: 0 = 8 warps, STG.128, a warp instruction covers 4 rows x 128 B (the GEMM epilogue's pattern)
//           1 = 8 warps, STG.128, a warp instruction covers 1 row x 512 B
//           2 = as 0 with st.global.cs      3 = 1-D bulk copies smem->global (cp.async.bulk), 512 B per row chunk
//           4 = 2-D TMA tensor store of 32-column x 128-row boxes (needs a tensor map)   5 = as 0 with 16 warps
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o store_bench store_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ROWS = 128, COLS = 256;  // tile per CTA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VAR>
__global__ void __launch_bounds__(512, 1) store_kernel(float *C, int ldc, int tiles_n, const __grid_constant__ CUtensorMap tm, long long *cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float *tile = (float *)smem;  // [ROWS][COLS] (variants 3,4 read it)
  const int m0 = (blockIdx.x / tiles_n) * ROWS, n0 = (blockIdx.x % tiles_n) * COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < ROWS * COLS; i += blockDim.x) tile[i] = (float)i;
  __syncthreads();
  long long t0 = clock64();
  if (VAR == 0 || VAR == 2 || VAR == 5) {
    const int cg4 = (lane & 7) * 4, r8 = lane >> 3;
    for (int chunk = warp; chunk < (ROWS / 32) * (COLS / 32); chunk += nw) {
      const int q = chunk % (ROWS / 32), c = chunk / (ROWS / 32);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int row = m0 + q * 32 + r8 + 4 * k, n = n0 + c * 32 + cg4;
        const float4 v = *(const float4 *)(tile + (q * 32 + r8 + 4 * k) * COLS + c * 32 + cg4);
        if (VAR == 2) __stcs((float4 *)(C + (size_t)row * ldc + n), v);
        else *(float4 *)(C + (size_t)row * ldc + n) = v;
      }
    }
  } else if (VAR == 1) {
    for (int idx = warp; idx < ROWS * (COLS / 128); idx += nw) {
      const int r = idx / (COLS / 128), c = (idx % (COLS / 128)) * 128 + lane * 4;
      const float4 v = *(const float4 *)(tile + r * COLS + c);
      *(float4 *)(C + (size_t)(m0 + r) * ldc + n0 + c) = v;
    }
  } else if (VAR == 3) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(C + (size_t)(m0 + r) * ldc + n0),
                   "r"(smem_u32(tile + r * COLS)), "r"(COLS * 4) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (VAR == 4) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < COLS / 32) {  // one 32-col x 128-row box (16 KB) per thread; smem box layout = [128 rows][32 floats]
      const int c = threadIdx.x;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm),
                   "r"(smem_u32(smem + c * (ROWS * 128))), "r"(n0 + c * 32), "r"(m0) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int VAR>
void run(const char *name, float *C, int M, int N, const CUtensorMap &tm, long long *cyc, int threads) {
  const int tiles_n = N / COLS, ctas = (M / ROWS) * tiles_n;
  auto k = store_kernel<VAR>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * COLS * 4 + 64 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) k<<<ctas, threads, ROWS * COLS * 4 + 64 * 1024>>>(C, N, tiles_n, tm, cyc);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 10; i++) k<<<ctas, threads, ROWS * COLS * 4 + 64 * 1024>>>(C, N, tiles_n, tm, cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[1024];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  long long mx = 0, sum = 0;
  for (int i = 0; i < ctas; i++) { if (h[i] > mx) mx = h[i]; sum += h[i]; }
  printf("%-46s %4d CTAs x %3d thr: store phase avg %6lld max %6lld cycles (%.1f B/clk/SM), kernel %.1f us\n", name, ctas, threads,
         sum / ctas, mx, (double)ROWS * COLS * 4 / (double)(sum / ctas), ms * 100.0);
}

int main(int argc, char **argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 2048, N = argc > 2 ? atoi(argv[2]) : 2048;
  float *C;
  long long *cyc;
  CK(cudaMalloc(&C, (size_t)M * N * 4));
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M}, gstr[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {32, ROWS}, estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  printf("C = %d x %d fp32 (%.1f MB)\n", M, N, M * (double)N * 4 / 1e6);
  run<0>("0: 8 warps STG.128, 4 rows x 128 B per instr", C, M, N, tm, cyc, 256);
  run<1>("1: 8 warps STG.128, 1 row x 512 B per instr", C, M, N, tm, cyc, 256);
  run<2>("2: as 0 with st.global.cs", C, M, N, tm, cyc, 256);
  run<5>("5: as 0 with 16 warps", C, M, N, tm, cyc, 512);
  run<3>("3: cp.async.bulk 1-D, 1 KB rows", C, M, N, tm, cyc, 256);
  run<4>("4: TMA 2-D tensor store, 16 KB boxes", C, M, N, tm, cyc, 256);
  return 0;
}
