// gemm_kernel.cuh — the tcgen05/TMEM/TMA GEMM kernel template and its launcher (see gemm_sm100.cu for the design notes).
// Included by gemm_sm100.cu (dispatcher) and by the gemm_inst_*.cu translation units that instantiate groups of tile shapes
// (split so that `make -j` compiles them in parallel).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm.cuh"

namespace tnb {




// ----------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: ~seconds of spinning means the pipeline protocol is broken -> trap (error), never hang
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { printf("tnb gemm: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
  }
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  // default .release.cta semantics (what CUTLASS' ClusterBarrier::arrive(cta) issues): the staged data never crosses the CTA
  // boundary through the generic proxy — each SM's tensor core reads its own smem after the local fence.proxy.async — so only
  // the ordering travels.  A .release.cluster here costs MEMBAR.ALL.GPU per arrive (measured: pair mode slower than 1 CTA).
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// 16-byte store into the shared memory of CTA `cta` of the cluster (same offset as the local address `p`)
__device__ __forceinline__ void st_remote_f4(float *p, uint32_t cta, float4 v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(p)), "r"(cta));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// kind::f16 (bf16 operands, fp32 accumulate): M x N x 16 per instruction
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// completion of all MMAs issued so far by this thread -> arrive on `bar` (in both CTAs of the pair when CG == 2)
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t *bar, uint16_t mask = 3) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout type [61,64): SWIZZLE_128B = 2 (K-major tiles), SWIZZLE_128B_BASE32B = 1 (the only
// layout tcgen05 takes for MN-major 32-bit operands: 32-byte swizzle atoms, 4-row groups)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// low part of the 3xTF32 split: hi is what the tensor core sees when it reads x (the top 19 bits), lo = rna_tf32(x - hi).
// Inf/NaN stay in hi only.
__device__ __forceinline__ float lo_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  const float r = x - __uint_as_float(u & 0xFFFFE000u);
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  return ((u & 0x7F800000u) == 0x7F800000u) ? 0.0f : __uint_as_float(l);
}

// Pipeline tracing (compile with -DTNB_GEMM_TRACE): CTA (0,0) records clock64() at every hand-off of the mainloop
// ([event][k block]) and at entry / setup / epilogue start / epilogue end / exit ([1][0..4]); tools/dbg_timeline.py prints it.
// This is how the per-K-block costs quoted in DESIGN.md 3.1 were measured.
#ifdef TNB_GEMM_TRACE
__device__ long long g_dbg_ts[8 * 256];
#define DBG_TS(ev, kb) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (kb) < 256) g_dbg_ts[(ev) * 256 + (kb)] = clock64(); } while (0)
// every CTA also records %globaltimer at entry / epilogue start / exit and its SM id ([cta][4]): launch skew and tail of the grid
__device__ long long g_dbg_cta[4 * 1024];
__device__ __forceinline__ long long dbg_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int dbg_smid() { int t; asm volatile("mov.u32 %0, %%smid;" : "=r"(t)); return t; }
#define DBG_CTA(slot) do { const int cta_ = blockIdx.y * gridDim.x + blockIdx.x; if (cta_ < 1024) { \
  g_dbg_cta[cta_ * 4 + (slot)] = dbg_gtime(); if ((slot) == 0) g_dbg_cta[cta_ * 4 + 3] = dbg_smid(); } } while (0)
#else
#define DBG_TS(ev, kb) do { } while (0)
#define DBG_CTA(slot) do { } while (0)
#endif

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile: each CTA
// stages its own 128 rows of A and HALF of the B tile (BN/2), the pair's tensor cores read both halves.
template <int BN_, int NTERMS_, int CG_>
struct GemmCfg {
  static constexpr int BN = BN_;
  static constexpr int BH = BN_ / CG_;  // B rows (n) staged by one CTA
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BH * BK * 4;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (NTERMS_ == 3 ? 2 : 1);
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};


// ----------------------------------------------------------------------------------------------- kernel
// A_MN / B_MN: 0 = K-major tile (operand rows are the M/N index, contraction index contiguous),
//              1 = MN-major tile (operand rows are the contraction index, M/N index contiguous).
// SPLIT = 2 (pair mode only): a cluster of 4 = two pairs working on the SAME 256 x BN tile, each over half of the K blocks; after
// the mainloop the pairs swap half of their accumulator columns through distributed shared memory and each finishes (adds,
// fused epilogue, store) the half it keeps.  Lets a 1024-row bunch use the MMA-bound 256 x 256 tile on 128 SMs.
template <int BN, int A_MN, int B_MN, int NTERMS, int CG, int SPLIT, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, EpiParams ep) {
  using Cfg = GemmCfg<BN, NTERMS, CG>;
  constexpr int BH = Cfg::BH;
  constexpr int STAGES = Cfg::STAGES;
  // NTERMS == 16: bf16 operands in HBM (TNB_MATH_BF16).  A 128-byte smem row then holds 64 elements instead of 32, so a K block
  // is 64 deep and an MN-major chunk 64 wide; the stage geometry IN BYTES is the same as for fp32/tf32.
  constexpr bool BF = (NTERMS == 16);
  constexpr int BKE = BF ? 64 : 32;   // elements per K block (one 128-byte swizzle span)
  constexpr int CW = BF ? 64 : 32;    // M/N elements per 128-byte row of an MN-major chunk
  constexpr int CHUNK_BYTES = BKE * 128;
  static_assert(!B_MN || BH % CW == 0, "an MN-major B tile is staged in whole 128-byte chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t *full_bar = bars;
  uint64_t *conv_bar = bars + STAGES;
  uint64_t *empty_bar = bars + 2 * STAGES;
  uint64_t *tmem_full_bar = bars + 3 * STAGES;
  uint32_t *tmem_ptr_smem = (uint32_t *)(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { DBG_TS(1, 0); DBG_CTA(0); }
  // Programmatic dependent launch: let the next GEMM of the stream be scheduled right away (its CTAs run their setup, then
  // block in griddepcontrol.wait until this grid has completed and flushed); no-op when launched without the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  static_assert(SPLIT == 1 || CG == 2, "split-K is built on CTA pairs");
  const uint32_t crank = (CG * SPLIT > 1) ? cluster_ctarank() : 0u;  // rank in the cluster of CG*SPLIT CTAs
  const uint32_t rank = crank & (CG - 1);    // rank in the pair: 0 = leader (issues the MMAs of the pair)
  const uint32_t split = crank / CG;         // which half of the K blocks this pair accumulates
  // consecutive CTAs (a pair when CG == 2) take consecutive 128-row blocks of the same N tile; with SPLIT the next pair repeats them
  const int m0 = (int)(blockIdx.x / (CG * SPLIT)) * (BM * CG) + (int)rank * BM;
  const int n0 = blockIdx.y * BN;
  const int total_kb = (K + BKE - 1) / BKE;
  const int kb_begin = (SPLIT == 1) ? 0 : (int)split * ((total_kb + 1) / 2);
  const int num_kb = (SPLIT == 1) ? total_kb : (split == 0 ? (total_kb + 1) / 2 : total_kb / 2);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      // arrivals per phase: every converter warp (3xTF32) or one forwarding warp (single pass) of each CTA of the pair
      mbar_init(&conv_bar[s], (NTERMS == 3 ? CONV_WARPS : 1) * CG);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 256 ? 256 : 512)));
  if (warp == 1) tmem_alloc<CG>(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  if (CG * SPLIT > 1) cluster_sync_all(); else __syncthreads();  // the peer's barriers must exist before any remote arrive / multicast commit
  tc_fence_after();
  if (threadIdx.x == 0) DBG_TS(1, 1);
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above touched no global memory: only now wait for the grid this one depends on (PDL)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  auto stage_a = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        DBG_TS(0, kb);
        mbar_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
        const int k0 = (kb_begin + kb) * BKE;
        if (A_MN == 0) {
          tma_load_2d(stage_a(s), &tmA, &full_bar[s], k0, m0);  // box 32(k) x 128(m)
        } else {
#pragma unroll
          for (int j = 0; j < BM / CW; j++)  // box CW(m) x BKE(k): one chunk of 128-byte rows per CW m
            tma_load_2d(stage_a(s) + j * CHUNK_BYTES, &tmA, &full_bar[s], m0 + CW * j, k0);
        }
        const int nb = n0 + (int)rank * BH;  // this CTA's part of the B tile
        if (B_MN == 0) {
          tma_load_2d(stage_b(s), &tmB, &full_bar[s], k0, nb);  // box 32(k) x BH(n)
        } else {
#pragma unroll
          for (int j = 0; j < BH / CW; j++)
            tma_load_2d(stage_b(s) + j * CHUNK_BYTES, &tmB, &full_bar[s], nb + CW * j, k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, a_major [15], b_major [16],
      // N>>3 [17,23), M>>4 [24,29)
      // (bf16: A = B = 1, kind::f16)
      constexpr uint32_t FMT = BF ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); a K step (8 floats / 16 bf16) = +32 B
      // MN-major fp32: 32-float chunks CHUNK_BYTES apart (LBO), 4-k-row swizzle groups 512 B apart (SBO), layout SWIZZLE_128B_BASE32B;
      //                a K step of 8 rows = +1024 B
      // MN-major bf16: 64-element chunks CHUNK_BYTES apart (LBO), 8-k-row swizzle groups 1024 B apart (SBO), layout SWIZZLE_128B;
      //                a K step of 16 rows = +2048 B
      constexpr uint32_t MN_SBO = BF ? 1024 : 512, MN_LT = BF ? 2 : 1, MN_KSTEP = BF ? 2048 : 1024;
      const uint32_t a_lbo = A_MN ? CHUNK_BYTES : 16, b_lbo = B_MN ? CHUNK_BYTES : 16;
      const uint32_t a_sbo = A_MN ? MN_SBO : 1024, b_sbo = B_MN ? MN_SBO : 1024;
      const uint32_t a_lt = A_MN ? MN_LT : 2, b_lt = B_MN ? MN_LT : 2;
      const uint32_t a_kstep = A_MN ? MN_KSTEP : 32, b_kstep = B_MN ? MN_KSTEP : 32;
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait((NTERMS == 3 || CG == 2) ? &conv_bar[s] : &full_bar[s], ph);
        DBG_TS(4, kb);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage_a(s)), b_hi = smem_u32(stage_b(s));
        const uint32_t a_lo = smem_u32(stage_alo(s)), b_lo = smem_u32(stage_blo(s));
#pragma unroll
        for (int ks = 0; ks < BK / 8; ks++) {
          const uint64_t dah = make_desc(a_hi + ks * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t dbh = make_desc(b_hi + ks * b_kstep, b_lbo, b_sbo, b_lt);
          const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
          if (NTERMS == 3) {
            const uint64_t dal = make_desc(a_lo + ks * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t dbl = make_desc(b_lo + ks * b_kstep, b_lbo, b_sbo, b_lt);
            umma_tf32<CG>(tmem_base, dal, dbh, idesc, first);
            umma_tf32<CG>(tmem_base, dah, dbl, idesc, 1u);
            umma_tf32<CG>(tmem_base, dah, dbh, idesc, 1u);
          } else if (BF) {
            umma_bf16<CG>(tmem_base, dah, dbh, idesc, first);
          } else {
            umma_tf32<CG>(tmem_base, dah, dbh, idesc, first);
          }
        }
        umma_commit<CG>(&empty_bar[s], (uint16_t)(3u << crank));  // smem slot (of both CTAs of the pair) reusable once these MMAs have read it
        DBG_TS(5, kb);
      }
      umma_commit<CG>(tmem_full_bar, (uint16_t)(3u << crank));  // accumulator complete (in both CTAs' TMEM)
    }
    __syncwarp();
  } else {
    // ===================== converters (3xTF32) then epilogue =====================
    const int ct = threadIdx.x - 64;  // 0..CONV_THREADS-1
    // The fused epilogue re-reads C (momentum buffer, beta != 0) and W: pull this CTA's tiles of both into L2 now, so that the
    // epilogue, which all CTAs reach at the same time, is served from L2 instead of queueing on HBM.
    if (ep.beta != 0.0f || ep.W || ep.mulY) {
      for (int i = ct; i < BM * (BN / 32); i += CONV_THREADS) {
        const int r = m0 + i / (BN / 32), cc = n0 + (i % (BN / 32)) * 32;
        if (r < M && cc < N) {
          if (ep.beta != 0.0f) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.C + (size_t)r * ep.ldc + cc));
          if (ep.W) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.W + (size_t)r * ep.ldw + cc));
          if (ep.mulY) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.mulY + (size_t)r * ep.ldy + cc));
        }
      }
    }
    // pair mode: these warps also forward "my stage has landed" to the leader's barrier (one warp is enough without conversion)
    if (NTERMS == 3 || (CG == 2 && warp == 2)) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);  // all lanes poll (a single polling lane + __syncwarp measured 1.5x slower)
        if (threadIdx.x == 64) DBG_TS(2, kb);
        if (NTERMS == 3) {
          // A and B tiles are contiguous ([A_hi][B_hi] -> [A_lo][B_lo]): one linear pass, 16 B per thread per step
          const float4 *src = (const float4 *)stage_a(s);
          float4 *dst = (float4 *)stage_alo(s);
          constexpr int NV = (Cfg::A_BYTES + Cfg::B_BYTES) / 16;
          static_assert(NV % CONV_THREADS == 0, "tile bytes must split evenly over the converter threads");
#pragma unroll
          for (int i = 0; i < NV / CONV_THREADS; i++) {
            const float4 x = src[ct + CONV_THREADS * i];
            float4 l;
            l.x = lo_tf32(x.x); l.y = lo_tf32(x.y); l.z = lo_tf32(x.z); l.w = lo_tf32(x.w);
            dst[ct + CONV_THREADS * i] = l;
          }
          fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        }
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_remote(&conv_bar[s], crank & ~1u); else mbar_arrive(&conv_bar[s]); }
        if (threadIdx.x == 64) DBG_TS(3, kb);
      }
    }
  }

  // ---- epilogue: TMEM -> registers -> smem transpose -> fused ops with COALESCED global accesses ----
  // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns).  Storing from that layout makes each warp
  // instruction touch 32 different 128-byte lines; going through a padded 32x36 smem tile per warp re-maps lanes so that
  // 8 consecutive lanes cover one 128-byte row segment (4 lines per instruction instead of 32) for every array the fused
  // epilogue reads or writes (C, C_old, bias, Yprev, W).  The stage buffers are free once tmem_full has fired.
  // The specialised epilogues (EPI_DX, EPI_UPD) issue ALL global reads of a 32-column chunk (8 float4 per array and thread)
  // before the accumulator is read, the first chunk's even before the last MMAs have completed: the reads of one chunk
  // used to be 8 dependent round trips to L2/HBM per array (ncu: 30 % of the update kernel's stall samples).
  const int q = warp & 3;              // TMEM lane quarter this warp may read
  const int chalf = (warp - 2) >> 2;   // two warps share a quarter: even / odd 32-column chunks
  constexpr int HALFC = BN / 64;       // 32-column chunks per half tile (split-K)
  constexpr int RS = BN / 2 + 4;       // row pitch (floats) of the split-K receive buffer: 16-byte aligned, conflict-free
  float *scratch = (float *)smem + (warp >= 2 ? warp - 2 : 0) * (32 * 36);
  float *recv = (float *)smem + CONV_WARPS * (32 * 36);  // [128][RS] floats, behind the per-warp transpose tiles
  const int cg4 = (lane & 7) * 4;      // column offset of this lane inside the 32-column chunk
  const int r8 = lane >> 3;            // row offset (0..3) inside a group of 4 rows
  const int nchunks = (SPLIT == 2) ? HALFC : BN / 32;
  const int row0 = m0 + q * 32 + r8;   // this lane's rows: row0 + 4k, k = 0..7
  // optional bf16 twins of the outputs (TNB_MATH_BF16: the next GEMM reads these instead of converting the fp32 arrays)
  auto st16 = [&](uint16_t *base, int ld, int row, int n, const float4 &o) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    uint2 u;
    u.x = *(const uint32_t *)&lo; u.y = *(const uint32_t *)&hi;
    *(uint2 *)(base + (size_t)row * ld + n) = u;
  };
  float4 pa[8], pb[8];                 // prefetched epilogue operands of the current chunk: EPI_DX: pa = y ; EPI_UPD: pa = C_old, pb = W
  auto chunk_col = [&](int ci) { return n0 + ((SPLIT == 2) ? (int)split * HALFC + ci : ci) * 32 + cg4; };
  auto prefetch_chunk = [&](int ci) {
    if (EPI != EPI_DX && EPI != EPI_UPD) return;
    const int n = chunk_col(ci);
    if (n + 3 >= N) return;  // the ragged last columns take the scalar path
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int row = row0 + 4 * k;
      if (row < M) {
        if (EPI == EPI_DX) pa[k] = *(const float4 *)(ep.mulY + (size_t)row * ep.ldy + n);
        if (EPI == EPI_UPD) {
          pa[k] = *(const float4 *)(ep.C + (size_t)row * ep.ldc + n);
          pb[k] = *(const float4 *)(ep.W + (size_t)row * ep.ldw + n);
        }
      }
    }
  };
  if (warp >= 2) {
    if (chalf < nchunks) prefetch_chunk(chalf);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) { DBG_TS(1, 2); DBG_CTA(1); }
  }
  // split-K: which of this warp's chunks of the exchanged half (chunk ci <-> slot (ci - chalf) / XSTEP of the warp's exchange area)
  constexpr int XSTEP = CONV_WARPS / 4;
  constexpr int XITERS = (HALFC + XSTEP - 1) / XSTEP;
  const float4 *xmine = nullptr;       // this CTA's slot of the L2 exchange buffer: what the partner pair has left for it
  if (SPLIT == 2 && ep.xchg) {
    // Accumulator exchange THROUGH L2 (default): every epilogue thread stores the 32-column TMEM rows of the half the OTHER pair
    // finishes into the partner CTA's slot, in the register layout tcgen05.ld produced them in ([warp][chunk][j][lane] float4: a warp
    // instruction covers 512 contiguous bytes); the partner thread at the same (warp, lane) — same TMEM lanes, same rows — reads them
    // back the same way and adds them to its own accumulator before the transpose.  One cluster barrier (release / acquire at cluster
    // scope) orders the stores before the loads.  Distributed shared memory moves 17-21 B/clk per SM; L2 takes the 64 KB of a CTA at
    // the rate of its ordinary stores, and the read-back overlaps the epilogue (measured: profiles/r02_split_exchange.md).
    const size_t slot_f4 = (size_t)CONV_WARPS * XITERS * 8 * 32;
    const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    xmine = (const float4 *)ep.xchg + cta * slot_f4;
    if (threadIdx.x == 64) DBG_TS(7, 0);
    if (warp >= 2) {
      float4 *xdst = (float4 *)ep.xchg + (cta ^ 2u) * slot_f4 + (size_t)(warp - 2) * XITERS * 256 + lane;
#pragma unroll 1
      for (int ci = chalf, it = 0; ci < HALFC; ci += XSTEP, it++) {
        const int c = (1 - (int)split) * HALFC + ci;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 8; j++)
          __stcg(xdst + (it * 8 + j) * 32, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                                      __uint_as_float(v[4 * j + 3])));
      }
    }
    if (threadIdx.x == 64) DBG_TS(7, 2);
    cluster_sync_all();  // release (my stores) / acquire (the partner's) at cluster scope
    if (threadIdx.x == 64) DBG_TS(7, 3);
  } else if (SPLIT == 2) {
    // exchange through distributed shared memory (TNB_GEMM_XCHG=dsmem; the first version, kept for A/B measurements)
    if (threadIdx.x == 64) DBG_TS(7, 0);
    cluster_sync_all();  // every pair of the cluster has finished its MMAs: all four CTAs' stage buffers are free
    if (threadIdx.x == 64) DBG_TS(7, 1);
    if (warp >= 2) {
      // send the half of the accumulator columns the OTHER pair finishes to the CTA holding the same rows there
      const uint32_t partner = crank ^ 2u;
      float *dst_row = recv + (q * 32 + lane) * RS;
#pragma unroll 1
      for (int ci = chalf; ci < HALFC; ci += CONV_WARPS / 4) {
        const int c = (1 - (int)split) * HALFC + ci;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 8; j++)
          st_remote_f4(dst_row + ci * 32 + 4 * j, partner,
                       make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3])));
      }
    }
    if (threadIdx.x == 64) DBG_TS(7, 2);
    cluster_sync_all();  // the partner's half has landed in my receive buffer
    if (threadIdx.x == 64) DBG_TS(7, 3);
  }
  if (warp >= 2) {
#pragma unroll 1
    for (int ci = chalf; ci < nchunks; ci += CONV_WARPS / 4) {
      const int c = (SPLIT == 2) ? (int)split * HALFC + ci : ci;
      if (n0 + c * 32 >= N) break;
      float4 px[8];  // split-K through L2: the partner pair's sums for this chunk
      if (SPLIT == 2 && xmine) {
        const float4 *xsrc = xmine + ((size_t)(warp - 2) * XITERS + (ci - chalf) / XSTEP) * 256 + lane;
#pragma unroll
        for (int j = 0; j < 8; j++) px[j] = __ldcg(xsrc + j * 32);
      }
      if (ci != chalf) prefetch_chunk(ci);
      if (threadIdx.x == 64) DBG_TS(6, 4 * (ci / 2) + 0);
      const int n = chunk_col(ci);
      const bool vec = n + 3 < N;
      float4 bv = make_float4(0, 0, 0, 0);
      if ((EPI == EPI_FWD || EPI == EPI_GENERIC) && ep.bias && vec) bv = *(const float4 *)(ep.bias + n);
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (threadIdx.x == 64) DBG_TS(6, 4 * (ci / 2) + 1);
      float4 *srow = (float4 *)(scratch + lane * 36);
      if (SPLIT == 2 && xmine) {
        // own K half + the partner's (a + b is commutative: both halves of the tile agree bit for bit)
#pragma unroll
        for (int j = 0; j < 8; j++)
          srow[j] = make_float4(__uint_as_float(v[4 * j]) + px[j].x, __uint_as_float(v[4 * j + 1]) + px[j].y, __uint_as_float(v[4 * j + 2]) + px[j].z,
                                __uint_as_float(v[4 * j + 3]) + px[j].w);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++)
          srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                __uint_as_float(v[4 * j + 3]));
      }
      __syncwarp();
      if (threadIdx.x == 64) DBG_TS(6, 4 * (ci / 2) + 2);
      // accumulator value (both K halves when split) of this lane's 4 columns in row r8 + 4k
      auto acc_at = [&](int k) {
        const int r = r8 + 4 * k;
        float4 a4 = *(const float4 *)(scratch + r * 36 + cg4);
        if (SPLIT == 2 && !xmine) {  // other K half, received through distributed shared memory
          const float4 p4 = *(const float4 *)(recv + (q * 32 + r) * RS + ci * 32 + cg4);
          a4.x += p4.x; a4.y += p4.y; a4.z += p4.z; a4.w += p4.w;
        }
        return a4;
      };
      if (vec) {
        if (EPI == EPI_FWD) {
          const bool sig = ep.act == TNB_ACT_SIGMOID;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int row = row0 + 4 * k;
            const float4 a4 = acc_at(k);
            float4 o = make_float4(a4.x + bv.x, a4.y + bv.y, a4.z + bv.z, a4.w + bv.w);
            if (sig) { o.x = sigmoidf_ref(o.x); o.y = sigmoidf_ref(o.y); o.z = sigmoidf_ref(o.z); o.w = sigmoidf_ref(o.w); }
            if (row < M) {
              *(float4 *)(ep.C + (size_t)row * ep.ldc + n) = o;
              if (ep.C16) st16(ep.C16, ep.ldc16, row, n, o);
            }
          }
        } else if (EPI == EPI_DX) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int row = row0 + 4 * k;
            const float4 a4 = acc_at(k);
            if (row < M) {
              const float4 y = pa[k];
              float4 o;
              o.x = (y.x * (1.0f - y.x)) * a4.x; o.y = (y.y * (1.0f - y.y)) * a4.y;
              o.z = (y.z * (1.0f - y.z)) * a4.z; o.w = (y.w * (1.0f - y.w)) * a4.w;
              *(float4 *)(ep.C + (size_t)row * ep.ldc + n) = o;
              if (ep.C16) st16(ep.C16, ep.ldc16, row, n, o);
            }
          }
        } else if (EPI == EPI_UPD) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int row = row0 + 4 * k;
            const float4 a4 = acc_at(k);
            if (row < M) {
              const float4 cold = pa[k];
              float4 w = pb[k], o;
              // same operation order as epi_one + the generic update below (alpha == 1 on this path)
              o = a4;
              if (ep.beta != 0.0f) {
                o.x = a4.x + ep.beta * cold.x; o.y = a4.y + ep.beta * cold.y;
                o.z = a4.z + ep.beta * cold.z; o.w = a4.w + ep.beta * cold.w;
              }
              *(float4 *)(ep.C + (size_t)row * ep.ldc + n) = o;
              w.x = ep.w_scale * o.x + w.x; w.y = ep.w_scale * o.y + w.y;
              w.z = ep.w_scale * o.z + w.z; w.w = ep.w_scale * o.w + w.w;
              if (ep.w_l2 != 0.0f) {
                w.x = ep.w_l2 * w.x + w.x; w.y = ep.w_l2 * w.y + w.y;
                w.z = ep.w_l2 * w.z + w.z; w.w = ep.w_l2 * w.w + w.w;
              }
              *(float4 *)(ep.W + (size_t)row * ep.ldw + n) = w;
              if (ep.W16) st16(ep.W16, ep.ldw16, row, n, w);
            }
          }
        } else {
#pragma unroll 4
          for (int k = 0; k < 8; k++) {
            const int row = row0 + 4 * k;
            const float4 a4 = acc_at(k);
            if (row < M) {
              size_t crow = (size_t)row * (size_t)ep.ldc;
              float *cbase = ep.C;
              if (ep.scat_shard > 0) {  // data-parallel gradient: straight into the owning rank's staging slice (peer stores over NVLink)
                const int own = row / ep.scat_shard;
                cbase = ep.scat[own];
                crow = (size_t)(ep.scat_rank * ep.scat_shard + (row - own * ep.scat_shard)) * (size_t)ep.ldc;
              }
              float4 cold = make_float4(0, 0, 0, 0), yv = make_float4(0, 0, 0, 0);
              if (ep.beta != 0.0f) cold = *(const float4 *)(cbase + crow + n);
              if (ep.mulY) yv = *(const float4 *)(ep.mulY + (size_t)row * ep.ldy + n);
              float4 o;
              o.x = epi_one(ep, a4.x, cold.x, bv.x, yv.x);
              o.y = epi_one(ep, a4.y, cold.y, bv.y, yv.y);
              o.z = epi_one(ep, a4.z, cold.z, bv.z, yv.z);
              o.w = epi_one(ep, a4.w, cold.w, bv.w, yv.w);
#ifdef TNB_GEMM_TRACE
              if (ep.alpha == -77.0f) { asm volatile("" ::"f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)); continue; }  // probe: epilogue without its stores
#endif
              float4 *wp = (float4 *)(ep.W + (size_t)row * ep.ldw + n);
              float4 w = make_float4(0, 0, 0, 0);
              if (ep.W) {
                w = *wp;
                if (ep.c_wdecay != 0.0f) { o.x = ep.c_wdecay * w.x + o.x; o.y = ep.c_wdecay * w.y + o.y; o.z = ep.c_wdecay * w.z + o.z; o.w = ep.c_wdecay * w.w + o.w; }
              }
              *(float4 *)(cbase + crow + n) = o;
              if (ep.C16) st16(ep.C16, ep.ldc16, row, n, o);
              if (ep.W) {
                w.x = ep.w_scale * o.x + w.x; w.y = ep.w_scale * o.y + w.y;
                w.z = ep.w_scale * o.z + w.z; w.w = ep.w_scale * o.w + w.w;
                if (ep.w_l2 != 0.0f) {
                  w.x = ep.w_l2 * w.x + w.x; w.y = ep.w_l2 * w.y + w.y;
                  w.z = ep.w_l2 * w.z + w.z; w.w = ep.w_l2 * w.w + w.w;
                }
                *wp = w;
                if (ep.W16) st16(ep.W16, ep.ldw16, row, n, w);
              }
            }
          }
        }
      } else if (n < N) {
        // ragged last columns of the matrix (N not a multiple of 4): element by element, any epilogue
#pragma unroll 1
        for (int k = 0; k < 8; k++) {
          const int row = row0 + 4 * k;
          const float4 a4 = acc_at(k);
          if (row >= M) continue;
          const float acc[4] = {a4.x, a4.y, a4.z, a4.w};
          size_t crow = (size_t)row * (size_t)ep.ldc;
          float *cbase = ep.C;
          if (ep.scat_shard > 0) {
            const int own = row / ep.scat_shard;
            cbase = ep.scat[own];
            crow = (size_t)(ep.scat_rank * ep.scat_shard + (row - own * ep.scat_shard)) * (size_t)ep.ldc;
          }
          for (int t = 0; t < 4 && n + t < N; t++) {
            float cold = (ep.beta != 0.0f) ? cbase[crow + n + t] : 0.0f;
            float bs = ep.bias ? ep.bias[n + t] : 0.0f;
            float yv = ep.mulY ? ep.mulY[(size_t)row * ep.ldy + n + t] : 0.0f;
            float o = epi_one(ep, acc[t], cold, bs, yv);
            if (ep.W && ep.c_wdecay != 0.0f) o = ep.c_wdecay * ep.W[(size_t)row * ep.ldw + n + t] + o;
            cbase[crow + n + t] = o;
            if (ep.C16) ep.C16[(size_t)row * ep.ldc16 + n + t] = __bfloat16_as_ushort(__float2bfloat16_rn(o));
            if (ep.W) {
              float *wp = ep.W + (size_t)row * ep.ldw + n + t;
              float w = ep.w_scale * o + *wp;
              if (ep.w_l2 != 0.0f) w = ep.w_l2 * w + w;
              *wp = w;
              if (ep.W16) ep.W16[(size_t)row * ep.ldw16 + n + t] = __bfloat16_as_ushort(__float2bfloat16_rn(w));
            }
          }
        }
      }
      if (threadIdx.x == 64) DBG_TS(6, 4 * (ci / 2) + 3);
      __syncwarp();  // the next chunk overwrites the scratch tile
    }
    if (threadIdx.x == 64) DBG_TS(1, 3);
  }
  tc_fence_before();
  if (CG * SPLIT > 1) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) { DBG_TS(1, 4); DBG_CTA(2); }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------- host launch
template <int BN, int A_MN, int B_MN, int NTERMS, int CG, int SPLIT, int EPI>
static int launch_tc(TnbContext *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, int M, int N, int K, const EpiParams &ep) {
  using Cfg = GemmCfg<BN, NTERMS, CG>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, NTERMS, CG, SPLIT, EPI>;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    TNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[ctx->device & 63] = true;
  }
  int mtiles = (M + BM - 1) / BM;
  if (CG == 2) mtiles = (mtiles + 1) & ~1;  // whole pairs; a pair's second CTA may be entirely out of range (zero-filled by TMA)
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->profiling) {
    while (ctx->prof_events.size() < ctx->prof_used + 2) {
      cudaEvent_t e;
      TNB_CUDA(cudaEventCreate(&e));
      ctx->prof_events.push_back(e);
    }
    e0 = ctx->prof_events[ctx->prof_used];
    e1 = ctx->prof_events[ctx->prof_used + 1];
    ctx->prof_used += 2;
    ctx->prof_flops += 2.0 * (double)M * (double)N * (double)K;
    TNB_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(mtiles * SPLIT, (N + BN - 1) / BN);  // split-K: each pair of row blocks appears once per K half
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * SPLIT;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (ctx->pdl && !ctx->profiling && !ctx->capturing) {
    // programmatic dependent launch: this grid may be scheduled while the previous kernel of the stream is still running; the
    // kernel orders itself behind it with griddepcontrol.wait before its first global access
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  EpiParams epl = ep;
  epl.xchg = nullptr;
  if (SPLIT == 2) {
    // one slot per CTA: [8 warps][chunks per warp][8][32] float4 (see the kernel's exchange step)
    constexpr size_t SLOT = (size_t)CONV_WARPS * (((BN / 64) + CONV_WARPS / 4 - 1) / (CONV_WARPS / 4)) * 8 * 32 * 16;
    const int rc = xchg_buffer(ctx, ctx->stream, (size_t)cfg.gridDim.x * cfg.gridDim.y * SLOT, &epl.xchg);
    if (rc != TNB_OK) return rc;
  }
  TNB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, epl));
  if (e1) TNB_CUDA(cudaEventRecord(e1, ctx->stream));
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

template <int BN, int NTERMS, int CG, int SPLIT>
int launch_tc_major(TnbContext *ctx, int a_mn, int b_mn, const CUtensorMap &tmA, const CUtensorMap &tmB, int M, int N, int K,
                    const EpiParams &ep) {
  // the specialised epilogues exist for the operand majors their fused layer op uses (forward NN, dX NT, update TN)
  if (!a_mn && b_mn && ep.mode == EPI_FWD) return launch_tc<BN, 0, 1, NTERMS, CG, SPLIT, EPI_FWD>(ctx, tmA, tmB, M, N, K, ep);
  if (!a_mn && !b_mn && ep.mode == EPI_DX) return launch_tc<BN, 0, 0, NTERMS, CG, SPLIT, EPI_DX>(ctx, tmA, tmB, M, N, K, ep);
  if (a_mn && b_mn && ep.mode == EPI_UPD) return launch_tc<BN, 1, 1, NTERMS, CG, SPLIT, EPI_UPD>(ctx, tmA, tmB, M, N, K, ep);
  if (!a_mn && !b_mn) return launch_tc<BN, 0, 0, NTERMS, CG, SPLIT, EPI_GENERIC>(ctx, tmA, tmB, M, N, K, ep);
  if (!a_mn && b_mn) return launch_tc<BN, 0, 1, NTERMS, CG, SPLIT, EPI_GENERIC>(ctx, tmA, tmB, M, N, K, ep);
  if (a_mn && !b_mn) return launch_tc<BN, 1, 0, NTERMS, CG, SPLIT, EPI_GENERIC>(ctx, tmA, tmB, M, N, K, ep);
  return launch_tc<BN, 1, 1, NTERMS, CG, SPLIT, EPI_GENERIC>(ctx, tmA, tmB, M, N, K, ep);
}

// co-resident clusters of this tile shape (clusters must sit inside one GPC, so CTAs * clusters can be below the SM count)
template <int BN, int NTERMS, int CG, int SPLIT>
int tc_max_active_clusters(int *clusters) {
  using Cfg = GemmCfg<BN, NTERMS, CG>;
  auto kern = gemm_tcgen05_kernel<BN, 0, 0, NTERMS, CG, SPLIT, EPI_GENERIC>;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CG * SPLIT * 64, 1);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * SPLIT;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  cudaError_t e = cudaOccupancyMaxActiveClusters(clusters, kern, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); *clusters = 0; return TNB_ERR_CUDA; }
  return TNB_OK;
}

// Tracing builds: device globals are private to a translation unit, so every instantiation unit exports its own readers
// (tnb_dbg_read_ts_<tag>, tnb_dbg_read_cta_<tag>); tools/dbg_*.py pick the unit whose buffers hold the latest timestamps.
#ifdef TNB_GEMM_TRACE
#define TNB_GEMM_TRACE_READERS(tag)                                                                                              \
  extern "C" int tnb_dbg_read_ts_##tag(long long *out) {                                                                         \
    return cudaMemcpyFromSymbol(out, tnb::g_dbg_ts, sizeof(long long) * 8 * 256) == cudaSuccess ? 0 : 1;                         \
  }                                                                                                                              \
  extern "C" int tnb_dbg_read_cta_##tag(long long *out) {                                                                        \
    return cudaMemcpyFromSymbol(out, tnb::g_dbg_cta, sizeof(long long) * 4 * 1024) == cudaSuccess ? 0 : 1;                       \
  }
#else
#define TNB_GEMM_TRACE_READERS(tag)
#endif

#define TNB_GEMM_INSTANTIATE(BN, NTERMS, CG, SPLIT)                                                                              \
  template int launch_tc_major<BN, NTERMS, CG, SPLIT>(TnbContext *, int, int, const CUtensorMap &, const CUtensorMap &, int, int, \
                                                      int, const EpiParams &);                                                    \
  template int tc_max_active_clusters<BN, NTERMS, CG, SPLIT>(int *);

}  // namespace tnb
