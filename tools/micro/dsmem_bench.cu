// dsmem_bench.cu — how fast can a CTA hand 64 KB to its cluster partner?  (split-K accumulator exchange probe)
//   0 = st.shared::cluster.v4 from 8 warps (what the GEMM does)   1 = cp.async.bulk shared::cta -> shared::cluster, 16 KB pieces
//   2 = through global memory: STG.128, cluster barrier, LDG.128
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_bench dsmem_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int BYTES = 64 * 1024;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int VAR>
__global__ void __launch_bounds__(256, 1) xchg(float *scratch, long long *cyc, float *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float *src = (float *)smem, *dst = (float *)(smem + BYTES);
  uint64_t *bar = (uint64_t *)(smem + 2 * BYTES);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  for (int i = threadIdx.x; i < BYTES / 4; i += blockDim.x) src[i] = (float)(i + rank);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  csync();
  long long t0 = clock64();
  if (VAR == 0) {
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) {
      const float4 v = ((const float4 *)src)[i];
      asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mapa(smem_u32(dst + 4 * i), peer)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    csync();
  } else if (VAR == 1) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BYTES) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    csync();  // the peer's barrier expects the bytes before any of them can land
    if (threadIdx.x < 4) {
      const uint32_t off = threadIdx.x * (BYTES / 4);
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(mapa(smem_u32((uint8_t *)dst + off), peer)), "r"(smem_u32((uint8_t *)src + off)), "r"(BYTES / 4), "r"(mapa(smem_u32(bar), peer)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
  } else {
    float *mine = scratch + (size_t)blockIdx.x * (BYTES / 4), *theirs = scratch + (size_t)(blockIdx.x ^ 1) * (BYTES / 4);
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) ((float4 *)mine)[i] = ((const float4 *)src)[i];
    csync();
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) {
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"((const float4 *)theirs + i));
      ((float4 *)dst)[i] = v;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  csync();
  if (dst[threadIdx.x] != (float)(threadIdx.x + peer)) sink[0] = -1.0f;  // data check
}

template <int VAR>
void run(const char *name, float *scratch, long long *cyc, float *sink, int ctas) {
  auto k = xchg<VAR>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BYTES + 64));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 2 * BYTES + 64;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaMemset(sink, 0, 4));
  for (int i = 0; i < 5; i++) CK(cudaLaunchKernelEx(&cfg, k, scratch, cyc, sink));
  CK(cudaDeviceSynchronize());
  long long h[1024]; float s;
  CK(cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&s, sink, 4, cudaMemcpyDeviceToHost));
  long long mx = 0, sum = 0;
  for (int i = 0; i < ctas; i++) { if (h[i] > mx) mx = h[i]; sum += h[i]; }
  printf("%-52s %3d CTAs: avg %6lld max %6lld cycles for 64 KB each way (%.1f B/clk/SM) %s\n", name, ctas, sum / ctas, mx, (double)BYTES / (double)(sum / ctas), s < 0 ? "DATA MISMATCH" : "ok");
}
int main() {
  float *scratch, *sink; long long *cyc;
  CK(cudaMalloc(&scratch, (size_t)256 * BYTES)); CK(cudaMalloc(&cyc, 8192)); CK(cudaMalloc(&sink, 4));
  for (int ctas : {2, 128}) {
    run<0>("0: st.shared::cluster.v4, 8 warps", scratch, cyc, sink, ctas);
    run<1>("1: cp.async.bulk smem->peer smem, 4 x 16 KB", scratch, cyc, sink, ctas);
    run<2>("2: STG -> cluster barrier -> LDG.cg (L2)", scratch, cyc, sink, ctas);
  }
  return 0;
}
