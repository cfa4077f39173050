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

constexpr int BM = 128;
constexpr int BK = 32;  // fp32 elements per K block = 128 bytes = one swizzle span
constexpr int CONV_WARPS = 8;                       // converter / epilogue warps
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int GEMM_THREADS = 64 + CONV_THREADS;     // + TMA warp + MMA warp


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
__device__ __forceinline__ float sigmoidf_ref(float x) {
  // reference: 1.0/(1.0+exp(-x)) with a float exp and a double divide (cukernels.cu:194-206); the float
  // evaluation below differs by <= 1 ulp
  return 1.0f / (1.0f + expf(-x));
}

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

__device__ __forceinline__ float epi_one(const EpiParams &ep, float acc, float cold, float bias, float y) {
  float o = ep.alpha * acc;
  if (ep.beta != 0.0f) o += ep.beta * cold;
  o += bias;
  if (ep.act == TNB_ACT_SIGMOID) o = sigmoidf_ref(o);
  if (ep.mulY) o = (y * (1.0f - y)) * o;
  return o;
}

// ----------------------------------------------------------------------------------------------- kernel
// A_MN / B_MN: 0 = K-major tile (operand rows are the M/N index, contraction index contiguous),
//              1 = MN-major tile (operand rows are the contraction index, M/N index contiguous).
// SPLIT = 2 (pair mode only): a cluster of 4 = two pairs working on the SAME 256 x BN tile, each over half of the K blocks; after
// the mainloop the pairs swap half of their accumulator columns through distributed shared memory and each finishes (adds,
// fused epilogue, store) the half it keeps.  Lets a 1024-row bunch use the MMA-bound 256 x 256 tile on 128 SMs.
template <int BN, int A_MN, int B_MN, int NTERMS, int CG, int SPLIT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, EpiParams ep) {
  using Cfg = GemmCfg<BN, NTERMS, CG>;
  constexpr int BH = Cfg::BH;
  constexpr int STAGES = Cfg::STAGES;
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
  static_assert(SPLIT == 1 || CG == 2, "split-K is built on CTA pairs");
  const uint32_t crank = (CG * SPLIT > 1) ? cluster_ctarank() : 0u;  // rank in the cluster of CG*SPLIT CTAs
  const uint32_t rank = crank & (CG - 1);    // rank in the pair: 0 = leader (issues the MMAs of the pair)
  const uint32_t split = crank / CG;         // which half of the K blocks this pair accumulates
  // consecutive CTAs (a pair when CG == 2) take consecutive 128-row blocks of the same N tile; with SPLIT the next pair repeats them
  const int m0 = (int)(blockIdx.x / (CG * SPLIT)) * (BM * CG) + (int)rank * BM;
  const int n0 = blockIdx.y * BN;
  const int total_kb = (K + BK - 1) / BK;
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
        const int k0 = (kb_begin + kb) * BK;
        if (A_MN == 0) {
          tma_load_2d(stage_a(s), &tmA, &full_bar[s], k0, m0);  // box 32(k) x 128(m)
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; j++)  // box 32(m) x 32(k), one 4 KB chunk per 32 m
            tma_load_2d(stage_a(s) + j * (BK * 128), &tmA, &full_bar[s], m0 + 32 * j, k0);
        }
        const int nb = n0 + (int)rank * BH;  // this CTA's part of the B tile
        if (B_MN == 0) {
          tma_load_2d(stage_b(s), &tmB, &full_bar[s], k0, nb);  // box 32(k) x BH(n)
        } else {
#pragma unroll
          for (int j = 0; j < BH / 32; j++)
            tma_load_2d(stage_b(s) + j * (BK * 128), &tmB, &full_bar[s], nb + 32 * j, k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, a_major [15], b_major [16],
      // N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); a K step of 8 floats = +32 B
      // MN-major: 32-float chunks BK*128 B apart (LBO), 4-k-row swizzle groups 512 B apart (SBO); a K step of 8 rows = +1024 B
      const uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
      const uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
      const uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
      const uint32_t a_kstep = A_MN ? 1024 : 32, b_kstep = B_MN ? 1024 : 32;
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
    if (ep.beta != 0.0f || ep.W) {
      for (int i = ct; i < BM * (BN / 32); i += CONV_THREADS) {
        const int r = m0 + i / (BN / 32), cc = n0 + (i % (BN / 32)) * 32;
        if (r < M && cc < N) {
          if (ep.beta != 0.0f) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.C + (size_t)r * ep.ldc + cc));
          if (ep.W) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.W + (size_t)r * ep.ldw + cc));
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
  const int q = warp & 3;              // TMEM lane quarter this warp may read
  const int chalf = (warp - 2) >> 2;   // two warps share a quarter: even / odd 32-column chunks
  constexpr int HALFC = BN / 64;       // 32-column chunks per half tile (split-K)
  constexpr int RS = BN / 2 + 4;       // row pitch (floats) of the split-K receive buffer: 16-byte aligned, conflict-free
  float *scratch = (float *)smem + (warp >= 2 ? warp - 2 : 0) * (32 * 36);
  float *recv = (float *)smem + CONV_WARPS * (32 * 36);  // [128][RS] floats, behind the per-warp transpose tiles
  if (warp >= 2) {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) { DBG_TS(1, 2); DBG_CTA(1); }
  }
  if (SPLIT == 2) {
    cluster_sync_all();  // every pair of the cluster has finished its MMAs: all four CTAs' stage buffers are free
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
    cluster_sync_all();  // the partner's half has landed in my receive buffer
  }
  if (warp >= 2) {
    const int cg4 = (lane & 7) * 4;      // column offset of this lane inside the 32-column chunk
    const int r8 = lane >> 3;            // row offset (0..3) inside a group of 4 rows
    const int nchunks = (SPLIT == 2) ? HALFC : BN / 32;
#pragma unroll 1
    for (int ci = chalf; ci < nchunks; ci += CONV_WARPS / 4) {
      const int c = (SPLIT == 2) ? (int)split * HALFC + ci : ci;
      const int nc0 = n0 + c * 32;
      if (nc0 >= N) break;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      float4 *srow = (float4 *)(scratch + lane * 36);
#pragma unroll
      for (int j = 0; j < 8; j++)
        srow[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      const int n = nc0 + cg4;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int r = r8 + 4 * k;
        const int row = m0 + q * 32 + r;
        float4 a4 = *(const float4 *)(scratch + r * 36 + cg4);
        if (SPLIT == 2) {  // other K half, computed by the partner pair (a + b is commutative: both halves of the tile agree)
          const float4 p4 = *(const float4 *)(recv + (q * 32 + r) * RS + ci * 32 + cg4);
          a4.x += p4.x; a4.y += p4.y; a4.z += p4.z; a4.w += p4.w;
        }
        if (row < M && n < N) {
          const float acc[4] = {a4.x, a4.y, a4.z, a4.w};
          const size_t crow = (size_t)row * (size_t)ep.ldc;
          if (n + 3 < N) {
            float4 cold = make_float4(0, 0, 0, 0), yv = make_float4(0, 0, 0, 0), bv = make_float4(0, 0, 0, 0);
            if (ep.beta != 0.0f) cold = *(const float4 *)(ep.C + crow + n);
            if (ep.bias) bv = *(const float4 *)(ep.bias + n);
            if (ep.mulY) yv = *(const float4 *)(ep.mulY + (size_t)row * ep.ldy + n);
            float4 o;
            o.x = epi_one(ep, acc[0], cold.x, bv.x, yv.x);
            o.y = epi_one(ep, acc[1], cold.y, bv.y, yv.y);
            o.z = epi_one(ep, acc[2], cold.z, bv.z, yv.z);
            o.w = epi_one(ep, acc[3], cold.w, bv.w, yv.w);
            *(float4 *)(ep.C + crow + n) = o;
            if (ep.W) {
              float4 *wp = (float4 *)(ep.W + (size_t)row * ep.ldw + n);
              float4 w = *wp;
              w.x = ep.w_scale * o.x + w.x; w.y = ep.w_scale * o.y + w.y;
              w.z = ep.w_scale * o.z + w.z; w.w = ep.w_scale * o.w + w.w;
              if (ep.w_l2 != 0.0f) {
                w.x = ep.w_l2 * w.x + w.x; w.y = ep.w_l2 * w.y + w.y;
                w.z = ep.w_l2 * w.z + w.z; w.w = ep.w_l2 * w.w + w.w;
              }
              *wp = w;
            }
          } else {
            for (int t = 0; t < 4 && n + t < N; t++) {
              float cold = (ep.beta != 0.0f) ? ep.C[crow + n + t] : 0.0f;
              float bv = ep.bias ? ep.bias[n + t] : 0.0f;
              float yv = ep.mulY ? ep.mulY[(size_t)row * ep.ldy + n + t] : 0.0f;
              float o = epi_one(ep, acc[t], cold, bv, yv);
              ep.C[crow + n + t] = o;
              if (ep.W) {
                float *wp = ep.W + (size_t)row * ep.ldw + n + t;
                float w = ep.w_scale * o + *wp;
                if (ep.w_l2 != 0.0f) w = ep.w_l2 * w + w;
                *wp = w;
              }
            }
          }
        }
      }
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
      float cold = (ep.beta != 0.0f) ? ep.C[ci] : 0.0f;
      float bv = ep.bias ? ep.bias[n] : 0.0f;
      float yv = ep.mulY ? ep.mulY[(size_t)row * ep.ldy + n] : 0.0f;
      float o = epi_one(ep, acc[i][j], cold, bv, yv);
      ep.C[ci] = o;
      if (ep.W) {
        float *wp = ep.W + (size_t)row * ep.ldw + n;
        float w = ep.w_scale * o + *wp;
        if (ep.w_l2 != 0.0f) w = ep.w_l2 * w + w;
        *wp = w;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------- host launch
template <int BN, int A_MN, int B_MN, int NTERMS, int CG, int SPLIT>
static int launch_tc(TnbContext *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, int M, int N, int K, const EpiParams &ep) {
  using Cfg = GemmCfg<BN, NTERMS, CG>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, NTERMS, CG, SPLIT>;
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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * SPLIT;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TNB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, ep));
  if (e1) TNB_CUDA(cudaEventRecord(e1, ctx->stream));
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

template <int BN, int NTERMS, int CG, int SPLIT = 1>
static int launch_tc_major(TnbContext *ctx, int a_mn, int b_mn, const CUtensorMap &tmA, const CUtensorMap &tmB, int M, int N,
                           int K, const EpiParams &ep) {
  if (!a_mn && !b_mn) return launch_tc<BN, 0, 0, NTERMS, CG, SPLIT>(ctx, tmA, tmB, M, N, K, ep);
  if (!a_mn && b_mn) return launch_tc<BN, 0, 1, NTERMS, CG, SPLIT>(ctx, tmA, tmB, M, N, K, ep);
  if (a_mn && !b_mn) return launch_tc<BN, 1, 0, NTERMS, CG, SPLIT>(ctx, tmA, tmB, M, N, K, ep);
  return launch_tc<BN, 1, 1, NTERMS, CG, SPLIT>(ctx, tmA, tmB, M, N, K, ep);
}

template <int BN, int CG>
static int launch_tc_terms(TnbContext *ctx, bool three, int a_mn, int b_mn, const CUtensorMap &tmA, const CUtensorMap &tmB, int M,
                           int N, int K, const EpiParams &ep) {
  return three ? launch_tc_major<BN, 3, CG>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep)
               : launch_tc_major<BN, 1, CG>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
}

// How many CTAs of a cluster launch can be resident at once (clusters must sit inside one GPC, so this can be below the SM count).
static int cluster_capacity(TnbContext *ctx, int cluster) {
  static int cap[64][5] = {};
  int &c = cap[ctx->device & 63][cluster];
  if (c == 0) {
    c = ctx->sm_count - ctx->sm_count % cluster;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cluster * 64, 1);
    cfg.blockDim = dim3(GEMM_THREADS);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e;
    if (cluster == 4) {
      using Cfg = GemmCfg<256, 3, 2>;
      auto kern = gemm_tcgen05_kernel<256, 0, 0, 3, 2, 2>;
      cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    } else {
      using Cfg = GemmCfg<256, 3, 2>;
      auto kern = gemm_tcgen05_kernel<256, 0, 0, 3, 2, 1>;
      cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    }
    if (e == cudaSuccess && n > 0) c = n * cluster; else cudaGetLastError();
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
    if (ctx->math_mode != TNB_MATH_FP32_SIMT && !tma_ok) {
      set_error("GEMM operands must be 16-byte aligned with a pitch multiple of 4 floats for the tensor-core path");
      return TNB_ERR_ARG;
    }
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    if (!ta && !tb) gemm_simt_kernel<0, 0><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else if (!ta && tb) gemm_simt_kernel<0, 1><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else if (ta && !tb) gemm_simt_kernel<1, 0><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    else gemm_simt_kernel<1, 1><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, M, N, K, ep);
    TNB_LAUNCHED(ctx);
    return TNB_OK;
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
    if (g == 2 && !three && !force_cg) continue;   // single-pass tf32 is L2/latency-bound: pairs measured 4 % slower there
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
  if (split == 2) {  // 3xTF32 CTA pairs only
    if (bn == 256) return launch_tc_major<256, 3, 2, 2>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
    if (bn == 192) return launch_tc_major<192, 3, 2, 2>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
    return launch_tc_major<128, 3, 2, 2>(ctx, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  }
  if (cg == 2) {
    if (bn == 256) return launch_tc_terms<256, 2>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
    if (bn == 192) return launch_tc_terms<192, 2>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
    return launch_tc_terms<128, 2>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  }
  if (bn == 256) return launch_tc_terms<256, 1>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  if (bn == 192) return launch_tc_terms<192, 1>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  if (bn == 128) return launch_tc_terms<128, 1>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
  return launch_tc_terms<64, 1>(ctx, three, a_mn, b_mn, tmA, tmB, M, N, K, ep);
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

}  // namespace tnb

using namespace tnb;

extern "C" {

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
  ep.C = Y; ep.ldc = dY.stride; ep.alpha = 1.0f; ep.beta = 0.0f; ep.bias = bias; ep.act = act;
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
  return launch_gemm(ctx, 'N', 'T', dE.rows, dW.rows, dE.cols, E, dE.stride, W, dW.stride, ep);
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

#ifdef TNB_GEMM_TRACE
// read back the pipeline timestamps (tracing builds only; not part of the ABI)
extern "C" int tnb_dbg_read_ts(long long *out) { return cudaMemcpyFromSymbol(out, tnb::g_dbg_ts, sizeof(long long) * 8 * 256) == cudaSuccess ? 0 : 1; }
extern "C" int tnb_dbg_read_cta(long long *out) { return cudaMemcpyFromSymbol(out, tnb::g_dbg_cta, sizeof(long long) * 4 * 1024) == cudaSuccess ? 0 : 1; }
#endif
