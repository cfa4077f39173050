// gemm_multi.cu — several INDEPENDENT GEMMs (each with its fused epilogue) as ONE persistent launch: tnb_gemm_batch.
//
// Why: a 1024-frame bunch gives a 2048-wide layer only 32 MMA-bound 256 x 256 tiles per GEMM — a quarter of the
// chip's CTA pairs — and every launch pays setup, pipeline fill, epilogue and teardown un-overlapped (about 45 % of a launch in
// 3xTF32, 75 % in bf16 mode).  In the backward pass the input-gradient GEMM of layer l (cuBiasedLinearity.cc:24) and the
// weight-gradient GEMMs of the layers above it (cuBiasedLinearity.cc:55, whose only consumer is the next forward pass) are
// independent of each other.  This kernel runs such a set as one grid of CTA pairs, each pair walking its own list of 256 x 256
// output tiles (host-side longest-first packing), with
//   * ONE shared-memory operand ring that keeps running across tiles (the TMA producer is already loading the next tile's
//     operands while the last MMAs of the current one execute: no pipeline fill between tiles),
//   * TWO 256-column TMEM accumulators (all 512 columns): dedicated epilogue warps drain tile i (tcgen05.ld -> fused epilogue ->
//     global) while the MMA lane accumulates tile i+1 into the other buffer,
//   * operand majors and epilogue chosen per tile at run time (forward NN, dX NT, dW TN in one grid),
//   * no split-K, hence no accumulator exchange: the independent GEMMs fill the pairs instead.
// Mainloop, descriptors, 3xTF32 split and epilogue arithmetic are those of gemm_kernel.cuh (same results bit for bit per tile).
#include <algorithm>

#include "gemm_kernel.cuh"

namespace tnb {

constexpr int MG_MAX_GEMMS = 6;
constexpr int MG_MAX_MAPS = 12;
constexpr int MG_BN = 256;

struct MgGemm {
  int M, N, K;
  int a_mn, b_mn;   // operand majors (0 = K-major, 1 = MN-major)
  int tm_a, tm_b;   // indices into MgParams::maps
  int pad_;
  EpiParams ep;
};

struct alignas(64) MgParams {
  CUtensorMap maps[MG_MAX_MAPS];
  MgGemm g[MG_MAX_GEMMS];
  const int4 *items;  // [pairs][ipp]: x = GEMM index (-1 ends the list), y = first row of the pair's 256-row tile, z = first column
  int ipp;
  long long *trace;   // TNB_BATCH_TRACE=1: [pairs][64] clock64() stamps of the leader CTA (tools/dbg/batch_timeline.py); NULL otherwise
};
// trace slots: 0 entry, 1 set-up done, 2 exit; per tile i < 8: 8+6i first operands converted, 9+6i last MMA issued, 10+6i accumulator
// complete (seen by the epilogue), 11+6i epilogue done, 12+6i TMA issued the tile's first load, 13+6i TMA issued its last load
#define MG_TRACE(slot) do { if (p.trace && rank == 0) p.trace[(size_t)(blockIdx.x >> 1) * 64 + (slot)] = clock64(); } while (0)

// Warp roles are laid out by WARPGROUPS (4 warps) because setmaxnreg moves registers between whole warpgroups:
//   3xTF32 (640 threads, launched at 96 registers): warpgroup 0 = TMA warp, MMA warp, 2 idle (-> 56 registers); warpgroups 1-2 = the 8
//   converter warps (-> 72); warpgroups 3-4 = 8 epilogue warps (-> 152: two epilogue operand tiles of a chunk in flight per thread).
//   bf16 (384 threads, 168 registers, no setmaxnreg): warpgroup 0 = TMA, MMA, stage forwarder, L2 prefetcher; warpgroups 1-2 = epilogue.
template <int NTERMS>
struct MgCfg {
  static constexpr bool BF = (NTERMS == 16);
  static constexpr int CONV_WARP0 = BF ? 2 : 4;               // bf16: warp 2 forwards "stage landed", warp 3 prefetches epilogue operands
  static constexpr int CONVW = BF ? 2 : CONV_WARPS;
  static constexpr int EPIW = 8;                              // epilogue warps (any warp w may read TMEM lanes 32*(w%4)..+31)
  static constexpr int EPI_WARP0 = BF ? 4 : 12;
  static constexpr int THREADS = 32 * (EPI_WARP0 + EPIW);     // 640 (3xTF32) / 384 (bf16)
  static constexpr int REGS_CTRL = 48, REGS_CONV = 64, REGS_EPI = 152;   // 3xTF32: setmaxnreg moves registers WITHIN the CTA's launch allocation: 128 * (48 + 2*64 + 2*152) = 61440 = 640 * 96
  static constexpr int A_BYTES = BM * BK * 4;                 // 16 KB in either element type (128 rows x 128 bytes)
  static constexpr int BH = MG_BN / 2;
  static constexpr int B_BYTES = BH * BK * 4;                 // 16 KB: this CTA's half of the B tile
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (NTERMS == 3 ? 2 : 1);
  static constexpr int STAGES = BF ? 6 : 3;
  static constexpr int SCRATCH_BYTES = EPIW * 32 * 32 * 4;    // one XOR-swizzled 32 x 32 transpose tile per epilogue warp
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + SCRATCH_BYTES;  // the dynamic array is declared 1024-byte aligned
};
static_assert(MgCfg<3>::SMEM_BYTES <= 232448 && MgCfg<16>::SMEM_BYTES <= 232448, "shared memory budget");

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------- epilogue of one tile
// The code of gemm_kernel.cuh's epilogue without split-K: q = TMEM lane quarter of this warp, chunks chalf, chalf+STEP, ... of the
// tile's eight 32-column chunks.  m0 = first row of THIS CTA's 128 rows.
template <int EPI, int STEP>
__device__ __forceinline__ void mg_epilogue(const EpiParams &ep, const int M, const int N, const int m0, const int n0, const uint32_t tmem_acc,
                                            float *scratch, const int q, const int chalf, const int lane) {
  const int cg4 = (lane & 7) * 4;
  const int r8 = lane >> 3;
  const int row0 = m0 + q * 32 + r8;
  auto st16 = [&](uint16_t *base, int ld, int row, int n, const float4 &o) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    uint2 u;
    u.x = *(const uint32_t *)&lo; u.y = *(const uint32_t *)&hi;
    *(uint2 *)(base + (size_t)row * ld + n) = u;
  };
  float4 pa[8], pb[8];
  auto prefetch_chunk = [&](int ci) {
    if (EPI != EPI_DX && EPI != EPI_UPD) return;
    const int n = n0 + ci * 32 + cg4;
    if (n + 3 >= N) return;
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
#pragma unroll 1
  for (int ci = chalf; ci < MG_BN / 32; ci += STEP) {
    if (n0 + ci * 32 >= N) break;
    prefetch_chunk(ci);
    const int n = n0 + ci * 32 + cg4;
    const bool vec = n + 3 < N;
    float4 bv = make_float4(0, 0, 0, 0);
    if ((EPI == EPI_FWD || EPI == EPI_GENERIC) && ep.bias && vec) bv = *(const float4 *)(ep.bias + n);
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(ci * 32), v);
    // 32 x 32 transpose tile, 16-byte units XOR-swizzled by the row (unit j of row r lives at r*8 + (j ^ (r & 7))): both the row-wise
    // writes (lane = row) and the reads (8 lanes per row, 4 rows per instruction) touch 8 distinct bank groups per quarter warp
    float4 *stile = (float4 *)scratch;
#pragma unroll
    for (int j = 0; j < 8; j++)
      stile[lane * 8 + (j ^ (lane & 7))] =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    auto acc_at = [&](int k) { const int rr = r8 + 4 * k; return stile[rr * 8 + ((lane & 7) ^ (rr & 7))]; };
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
            o = a4;  // same operation order as epi_one + the generic update (alpha == 1 on this path)
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
#pragma unroll 2
        for (int k = 0; k < 8; k++) {
          const int row = row0 + 4 * k;
          const float4 a4 = acc_at(k);
          if (row < M) {
            const size_t crow = (size_t)row * (size_t)ep.ldc;
            float4 cold = make_float4(0, 0, 0, 0), yv = make_float4(0, 0, 0, 0);
            if (ep.beta != 0.0f) cold = *(const float4 *)(ep.C + crow + n);
            if (ep.mulY) yv = *(const float4 *)(ep.mulY + (size_t)row * ep.ldy + n);
            float4 o;
            o.x = epi_one(ep, a4.x, cold.x, bv.x, yv.x);
            o.y = epi_one(ep, a4.y, cold.y, bv.y, yv.y);
            o.z = epi_one(ep, a4.z, cold.z, bv.z, yv.z);
            o.w = epi_one(ep, a4.w, cold.w, bv.w, yv.w);
            *(float4 *)(ep.C + crow + n) = o;
            if (ep.C16) st16(ep.C16, ep.ldc16, row, n, o);
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
        const size_t crow = (size_t)row * (size_t)ep.ldc;
        for (int t = 0; t < 4 && n + t < N; t++) {
          float cold = (ep.beta != 0.0f) ? ep.C[crow + n + t] : 0.0f;
          float bs = ep.bias ? ep.bias[n + t] : 0.0f;
          float yv = ep.mulY ? ep.mulY[(size_t)row * ep.ldy + n + t] : 0.0f;
          float o = epi_one(ep, acc[t], cold, bs, yv);
          ep.C[crow + n + t] = o;
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
    __syncwarp();  // the next chunk overwrites the scratch tile
  }
}

// ---------------------------------------------------------------------------------------------- kernel
template <int NTERMS>
__global__ void __launch_bounds__(MgCfg<NTERMS>::THREADS, 1) gemm_multi_kernel(const __grid_constant__ MgParams p) {
  using Cfg = MgCfg<NTERMS>;
  constexpr bool BF = Cfg::BF;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BH = Cfg::BH;
  constexpr int BKE = BF ? 64 : 32;
  constexpr int CW = BF ? 64 : 32;
  constexpr int CHUNK_BYTES = BKE * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("tnb gemm batch: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
  uint64_t *bars = (uint64_t *)(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t *full_bar = bars;
  uint64_t *conv_bar = bars + STAGES;
  uint64_t *empty_bar = bars + 2 * STAGES;
  uint64_t *tmem_full_bar = bars + 3 * STAGES;       // [2]
  uint64_t *tmem_empty_bar = bars + 3 * STAGES + 2;  // [2]
  uint32_t *tmem_ptr_smem = (uint32_t *)(bars + 3 * STAGES + 4);
  float *scratch_all = (float *)(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;  // 0 = leader of the pair
  const int4 *items = p.items + (size_t)(blockIdx.x >> 1) * p.ipp;
  if (threadIdx.x == 0) MG_TRACE(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], (NTERMS == 3 ? CONV_WARPS : 1) * 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], Cfg::EPIW * 2);  // every epilogue warp of both CTAs of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<2>(tmem_ptr_smem, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) MG_TRACE(1);

  auto stage_a = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + Cfg::B_BYTES; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };

  // 3xTF32: every role starts by handing registers it does not need to the CTA's pool (control and converter warpgroups) or by
  // taking them (epilogue warpgroups) — setmaxnreg is the first statement of each role's branch so that the code it dominates is
  // register-allocated under the new limit.
  // (all four warps of a warpgroup must execute the SAME setmaxnreg instruction: one site per warpgroup class.)
  if (warp < Cfg::CONV_WARP0 && NTERMS == 3) setmaxnreg_dec<Cfg::REGS_CTRL>();  // 3xTF32: warps 0-3 (2 and 3 have no other role)
  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own 128 rows of A, own half of B) =====================
    if (lane == 0) {
      for (int i = 0; i < MG_MAX_MAPS; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps[i]) : "memory");
      uint32_t it = 0;
      for (int i = 0; i < p.ipp; i++) {
        const int4 w = items[i];
        if (w.x < 0) break;
        const MgGemm &g = p.g[w.x];
        const CUtensorMap *tmA = &p.maps[g.tm_a], *tmB = &p.maps[g.tm_b];
        const int a_mn = g.a_mn, b_mn = g.b_mn;
        const int m0 = w.y + (int)rank * BM, nb = w.z + (int)rank * BH;
        const int nkb = (g.K + BKE - 1) / BKE;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (i < 8 && kb == 0) MG_TRACE(12 + 6 * i);
          if (i < 8 && kb == nkb - 1) MG_TRACE(13 + 6 * i);
          mbar_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
          const int k0 = kb * BKE;
          if (!a_mn) {
            tma_load_2d(stage_a(s), tmA, &full_bar[s], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / CW; j++) tma_load_2d(stage_a(s) + j * CHUNK_BYTES, tmA, &full_bar[s], m0 + CW * j, k0);
          }
          if (!b_mn) {
            tma_load_2d(stage_b(s), tmB, &full_bar[s], k0, nb);
          } else {
#pragma unroll
            for (int j = 0; j < BH / CW; j++) tma_load_2d(stage_b(s) + j * CHUNK_BYTES, tmB, &full_bar[s], nb + CW * j, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t FMT = BF ? 1u : 2u;
      constexpr uint32_t MN_SBO = BF ? 1024 : 512, MN_LT = BF ? 2 : 1, MN_KSTEP = BF ? 2048 : 1024;
      uint32_t it = 0;
      for (int i = 0; i < p.ipp; i++) {
        const int4 w = items[i];
        if (w.x < 0) break;
        const MgGemm &g = p.g[w.x];
        const uint32_t a_mn = (uint32_t)g.a_mn, b_mn = (uint32_t)g.b_mn;
        const int nkb = (g.K + BKE - 1) / BKE;
        const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(MG_BN >> 3) << 17) |
                               ((uint32_t)((BM * 2) >> 4) << 24);
        const uint32_t a_lbo = a_mn ? CHUNK_BYTES : 16, b_lbo = b_mn ? CHUNK_BYTES : 16;
        const uint32_t a_sbo = a_mn ? MN_SBO : 1024, b_sbo = b_mn ? MN_SBO : 1024;
        const uint32_t a_lt = a_mn ? MN_LT : 2, b_lt = b_mn ? MN_LT : 2;
        const uint32_t a_kstep = a_mn ? MN_KSTEP : 32, b_kstep = b_mn ? MN_KSTEP : 32;
        const int as = i & 1, use = i >> 1;
        // the epilogue warps of both CTAs must have drained this accumulator buffer (its previous tile: item i-2)
        if (use >= 1) mbar_wait(&tmem_empty_bar[as], (uint32_t)((use - 1) & 1));
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(as * MG_BN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&conv_bar[s], ph);
          if (i < 8 && kb == 0) MG_TRACE(8 + 6 * i);
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
              umma_tf32<2>(acc, dal, dbh, idesc, first);
              umma_tf32<2>(acc, dah, dbl, idesc, 1u);
              umma_tf32<2>(acc, dah, dbh, idesc, 1u);
            } else {
              umma_bf16<2>(acc, dah, dbh, idesc, first);
            }
          }
          umma_commit<2>(&empty_bar[s], (uint16_t)(3u << (crank & ~1u)));
        }
        umma_commit<2>(&tmem_full_bar[as], (uint16_t)(3u << (crank & ~1u)));
        if (i < 8) MG_TRACE(9 + 6 * i);
      }
    }
    __syncwarp();
  } else if (warp >= Cfg::CONV_WARP0 && warp < Cfg::CONV_WARP0 + Cfg::CONVW) {
    // ===================== converters (3xTF32) / stage forwarder + L2 prefetcher (bf16) =====================
    if (NTERMS == 3) setmaxnreg_dec<Cfg::REGS_CONV>();
    const int ct = threadIdx.x - 32 * Cfg::CONV_WARP0;
    uint32_t it = 0;
    for (int i = 0; i < p.ipp; i++) {
      const int4 w = items[i];
      if (w.x < 0) break;
      const MgGemm &g = p.g[w.x];
      const int nkb = (g.K + BKE - 1) / BKE;
      // pull this CTA's tiles of the arrays the fused epilogue re-reads into L2 while the mainloop runs
      if (NTERMS == 3 || warp == 3) {
        const EpiParams &ep = g.ep;
        const float *pc = ep.beta != 0.0f ? ep.C : nullptr, *pw = ep.W, *py = ep.mulY;
        if (pc || pw || py) {
          const int m0 = w.y + (int)rank * BM, n0 = w.z;
          const int t0 = (NTERMS == 3) ? ct : lane, tstep = (NTERMS == 3) ? CONV_THREADS : 32;
          for (int j = t0; j < BM * (MG_BN / 32); j += tstep) {
            const int r = m0 + j / (MG_BN / 32), cc = n0 + (j % (MG_BN / 32)) * 32;
            if (r < g.M && cc < g.N) {
              if (pc) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + (size_t)r * ep.ldc + cc));
              if (pw) asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + (size_t)r * ep.ldw + cc));
              if (py) asm volatile("prefetch.global.L2 [%0];" ::"l"(py + (size_t)r * ep.ldy + cc));
            }
          }
        }
      }
      if (NTERMS == 3 || warp == 2) {
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          if (NTERMS == 3) {
            const float4 *src = (const float4 *)stage_a(s);
            float4 *dst = (float4 *)stage_alo(s);
            constexpr int NV = (Cfg::A_BYTES + Cfg::B_BYTES) / 16;
            static_assert(NV % CONV_THREADS == 0, "tile bytes must split evenly over the converter threads");
#pragma unroll
            for (int j = 0; j < NV / CONV_THREADS; j++) {
              const float4 x = src[ct + CONV_THREADS * j];
              float4 l;
              l.x = lo_tf32(x.x); l.y = lo_tf32(x.y); l.z = lo_tf32(x.z); l.w = lo_tf32(x.w);
              dst[ct + CONV_THREADS * j] = l;
            }
            fence_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&conv_bar[s], crank & ~1u);
        }
      }
    }
  } else if (warp >= Cfg::EPI_WARP0) {
    // ===================== epilogue warps: drain accumulator buffer (i & 1) while the MMAs fill the other one =====================
    if (NTERMS == 3) setmaxnreg_inc<Cfg::REGS_EPI>();
    const int ew = warp - Cfg::EPI_WARP0;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    float *scratch = scratch_all + ew * (32 * 32);
    for (int i = 0; i < p.ipp; i++) {
      const int4 w = items[i];
      if (w.x < 0) break;
      const MgGemm &g = p.g[w.x];
      const EpiParams ep = g.ep;
      const int M = g.M, N = g.N;
      const int m0 = w.y + (int)rank * BM, n0 = w.z;
      const int as = i & 1, use = i >> 1;
      mbar_wait(&tmem_full_bar[as], (uint32_t)(use & 1));
      tc_fence_after();
      if (ew == 0 && lane == 0 && i < 8) MG_TRACE(10 + 6 * i);
      const uint32_t acc = tmem_base + (uint32_t)(as * MG_BN);
      constexpr int STEP = Cfg::EPIW / 4;
      switch (ep.mode) {
        case EPI_FWD: mg_epilogue<EPI_FWD, STEP>(ep, M, N, m0, n0, acc, scratch, q, chalf, lane); break;
        case EPI_DX: mg_epilogue<EPI_DX, STEP>(ep, M, N, m0, n0, acc, scratch, q, chalf, lane); break;
        case EPI_UPD: mg_epilogue<EPI_UPD, STEP>(ep, M, N, m0, n0, acc, scratch, q, chalf, lane); break;
        default: mg_epilogue<EPI_GENERIC, STEP>(ep, M, N, m0, n0, acc, scratch, q, chalf, lane); break;
      }
      tc_fence_before();  // this warp's tcgen05.ld of the buffer have completed (tcgen05.wait::ld inside tmem_ld32)
      __syncwarp();
      if (ew == 0 && lane == 0 && i < 8) MG_TRACE(11 + 6 * i);
      if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[as], crank & ~1u);
    }
  }


  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) MG_TRACE(2);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
// One tile of a job.  `main` = its mainloop, `epi` = its epilogue, both in cycles of the model below: a pair works through its list
// with the epilogue of tile i running under the mainloop of tile i+1, so its time is  m_1 + sum_i max(m_i, e_(i-1)) + e_last.
struct MgItem { int gemm, m0, n0; long main, epi; };

// measured with tools/dbg/batch_timeline.py (cycles at ~1.9 GHz): K block of the 256 x 256 pair tile and the three epilogues
static void item_cost(const TnbGemmJob &j, bool bf, long *main, long *epi) {
  const int nkb = (j.k + (bf ? 64 : 32) - 1) / (bf ? 64 : 32);
  *main = (long)nkb * (bf ? 640 : 1600);
  const bool upd = j.W != nullptr, rd = j.mulY != nullptr || j.beta != 0.0f;
  *epi = bf ? (upd ? 24000 : (rd ? 13000 : 8000)) : (upd ? 30000 : (rd ? 12000 : 7000));
}

// time of a pair's list in the order it will run: most expensive epilogue first, so that the cheapest one is the exposed last
static long list_time(std::vector<MgItem> &l) {
  std::stable_sort(l.begin(), l.end(), [](const MgItem &a, const MgItem &b) { return a.epi > b.epi; });
  long t = 0, prev_epi = 0;
  for (size_t i = 0; i < l.size(); i++) {
    t += (i == 0) ? l[i].main : std::max(l[i].main, prev_epi);
    prev_epi = l[i].epi;
  }
  return t + prev_epi;
}

static int mg_pairs(TnbContext *ctx) {
  static int pairs[64] = {};
  int &n = pairs[ctx->device & 63];
  if (n == 0) {
    n = ctx->sm_count / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * 74, 1);
    cfg.blockDim = dim3(MgCfg<3>::THREADS);
    cfg.dynamicSmemBytes = MgCfg<3>::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(gemm_multi_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MgCfg<3>::SMEM_BYTES);
    int c = 0;
    if (cudaOccupancyMaxActiveClusters(&c, gemm_multi_kernel<3>, &cfg) == cudaSuccess && c > 0) n = c; else cudaGetLastError();
    const char *e = getenv("TNB_BATCH_PAIRS");
    if (e && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
  }
  // a caller may keep SMs free for kernels that run next to the batch (the data-parallel exchange): tnb_gemm_batch_set_pairs
  return (ctx->mg_pairs_limit > 0 && ctx->mg_pairs_limit < n) ? ctx->mg_pairs_limit : n;
}

static bool job_ok(const TnbGemmJob &j, bool bf) {
  if (!(j.m > 0 && j.n > 0 && j.k > 0 && j.C)) return false;
  if (bf) {
    if (!j.A16 || !j.B16 || ((uintptr_t)j.A16 & 15) || ((uintptr_t)j.B16 & 15) || (j.lda16 & 7) || (j.ldb16 & 7)) return false;
    if ((j.C16 && (((uintptr_t)j.C16 & 7) || (j.ldc16 & 3))) || (j.W16 && (((uintptr_t)j.W16 & 7) || (j.ldw16 & 3)))) return false;
  } else {
    if (!j.A || !j.B || ((uintptr_t)j.A & 15) || ((uintptr_t)j.B & 15) || (j.lda & 3) || (j.ldb & 3)) return false;
  }
  if (((uintptr_t)j.C & 15) || (j.ldc & 3)) return false;
  if (j.bias && ((uintptr_t)j.bias & 15)) return false;
  if (j.mulY && (((uintptr_t)j.mulY & 15) || (j.ldy & 3))) return false;
  if (j.W && (((uintptr_t)j.W & 15) || (j.ldw & 3))) return false;
  return true;
}

static void job_to_gemm(const TnbGemmJob &j, MgGemm *g) {
  memset(g, 0, sizeof(*g));
  g->M = j.m; g->N = j.n; g->K = j.k;
  g->a_mn = j.transa ? 1 : 0;
  g->b_mn = j.transb ? 0 : 1;
  EpiParams &ep = g->ep;
  ep.C = j.C; ep.ldc = j.ldc; ep.alpha = j.alpha; ep.beta = j.beta; ep.bias = j.bias; ep.act = j.act;
  ep.mulY = j.mulY; ep.ldy = j.ldy; ep.W = j.W; ep.ldw = j.ldw; ep.w_scale = j.w_scale; ep.w_l2 = j.w_l2;
  ep.C16 = j.C16; ep.ldc16 = j.ldc16; ep.W16 = j.W16; ep.ldw16 = j.ldw16;
  // the specialised epilogues assume exactly their fused op's fields (as launch_tc_major does for the single-GEMM kernel)
  ep.mode = EPI_GENERIC;
  if (j.epilogue == TNB_EPI_FWD && !j.transa && !j.transb && j.alpha == 1.0f && j.beta == 0.0f && !j.mulY && !j.W) ep.mode = EPI_FWD;
  if (j.epilogue == TNB_EPI_DX && j.mulY && j.alpha == 1.0f && j.beta == 0.0f && !j.bias && !j.W && j.act == TNB_ACT_NONE) ep.mode = EPI_DX;
  if (j.epilogue == TNB_EPI_UPDATE && j.W && j.alpha == 1.0f && !j.bias && !j.mulY && j.act == TNB_ACT_NONE) ep.mode = EPI_UPD;
}

static int job_tiles(const TnbGemmJob &j) { return ((j.m + 255) / 256) * ((j.n + MG_BN - 1) / MG_BN); }

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_gemm_job_tiles(const TnbGemmJob *job) { return job ? job_tiles(*job) : 0; }

int tnb_gemm_batch_set_pairs(TnbContext *ctx, int pairs) {
  TNB_ARG(ctx && pairs >= 0, "pairs");
  ctx->mg_pairs_limit = pairs;
  return TNB_OK;
}

int tnb_gemm_batch_trace_read(TnbContext *ctx, long long *out /* [128][64] */) {
  TNB_ARG(ctx && out, "null");
  TNB_ARG(ctx->mg_trace != nullptr, "no trace: set TNB_BATCH_TRACE=1 before the first tnb_gemm_batch");
  TNB_CUDA(cudaStreamSynchronize(ctx->stream));
  TNB_CUDA(cudaMemcpy(out, ctx->mg_trace, sizeof(long long) * 64 * 128, cudaMemcpyDeviceToHost));
  return TNB_OK;
}

int tnb_job_affine_bwd_dx(TnbGemmJob *job, const float *E, TnbMatrixDim dE, const float *W, TnbMatrixDim dW, const float *Yprev,
                          TnbMatrixDim dYprev, float *Eprev, TnbMatrixDim dEprev) {
  TNB_ARG(job && E && W && Eprev, "null");
  TNB_ARG(dE.cols == dW.cols && dEprev.cols == dW.rows && dEprev.rows == dE.rows, "dimension mismatch");
  if (Yprev) TNB_ARG(dYprev.rows == dEprev.rows && dYprev.cols == dEprev.cols, "Yprev dims");
  memset(job, 0, sizeof(*job));
  job->transa = 0; job->transb = 1; job->m = dE.rows; job->n = dW.rows; job->k = dE.cols;
  job->A = E; job->lda = dE.stride; job->B = W; job->ldb = dW.stride;
  job->epilogue = Yprev ? TNB_EPI_DX : TNB_EPI_STORE; job->alpha = 1.0f; job->beta = 0.0f;
  job->C = Eprev; job->ldc = dEprev.stride; job->mulY = Yprev; job->ldy = dYprev.stride;
  job->tile_first = 0; job->tile_count = job_tiles(*job);
  return TNB_OK;
}

int tnb_job_affine_update(TnbGemmJob *job, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *W, TnbMatrixDim dW,
                          float *corrW, float lr, float mmt, float wc, int gdf, int n_frames_global) {
  TNB_ARG(job && X && E && W && corrW, "null");
  TNB_ARG(dX.rows == dE.rows && dW.rows == dX.cols && dW.cols == dE.cols, "dimension mismatch");
  float scale, l2;
  update_scalars(lr, mmt, wc, gdf, n_frames_global > 0 ? n_frames_global : dX.rows, &scale, &l2);
  memset(job, 0, sizeof(*job));
  job->transa = 1; job->transb = 0; job->m = dX.cols; job->n = dE.cols; job->k = dX.rows;
  job->A = X; job->lda = dX.stride; job->B = E; job->ldb = dE.stride;
  job->epilogue = TNB_EPI_UPDATE; job->alpha = 1.0f; job->beta = mmt;
  job->C = corrW; job->ldc = dW.stride; job->W = W; job->ldw = dW.stride; job->w_scale = scale; job->w_l2 = l2;
  job->tile_first = 0; job->tile_count = job_tiles(*job);
  return TNB_OK;
}

int tnb_job_affine_grad(TnbGemmJob *job, const float *X, TnbMatrixDim dX, const float *E, TnbMatrixDim dE, float *G, TnbMatrixDim dG) {
  TNB_ARG(job && X && E && G, "null");
  TNB_ARG(dX.rows == dE.rows && dG.rows == dX.cols && dG.cols == dE.cols, "dimension mismatch");
  memset(job, 0, sizeof(*job));
  job->transa = 1; job->transb = 0; job->m = dX.cols; job->n = dE.cols; job->k = dX.rows;
  job->A = X; job->lda = dX.stride; job->B = E; job->ldb = dE.stride;
  job->epilogue = TNB_EPI_STORE; job->alpha = 1.0f; job->beta = 0.0f;
  job->C = G; job->ldc = dG.stride;
  job->tile_first = 0; job->tile_count = job_tiles(*job);
  return TNB_OK;
}

int tnb_job_set_twins(TnbGemmJob *job, const uint16_t *A16, int lda16, const uint16_t *B16, int ldb16, uint16_t *C16, int ldc16,
                      uint16_t *W16, int ldw16) {
  TNB_ARG(job != nullptr, "null");
  job->A16 = A16; job->lda16 = lda16; job->B16 = B16; job->ldb16 = ldb16;
  job->C16 = C16; job->ldc16 = ldc16; job->W16 = W16; job->ldw16 = ldw16;
  return TNB_OK;
}

int tnb_gemm_batch_ok(TnbContext *ctx, const TnbGemmJob *job) {
  if (!ctx || !job) return 0;
  if (ctx->math_mode != TNB_MATH_3XTF32 && ctx->math_mode != TNB_MATH_BF16) return 0;
  if (ctx->profiling && getenv("TNB_BATCH_PROFILE_SPLIT")) return 0;
  static int off = -1;
  if (off < 0) { const char *e = getenv("TNB_GEMM_BATCH"); off = (e && atoi(e) == 0) ? 1 : 0; }
  if (off) return 0;
  // 256 x 256 pair tiles: worth it only when the output and the contraction are at least a tile or so in every direction
  if (job->m < 256 || job->n < 256 || job->k < 128) return 0;
  return job_ok(*job, ctx->math_mode == TNB_MATH_BF16) ? 1 : 0;
}

int tnb_gemm_batch(TnbContext *ctx, const TnbGemmJob *must, int n_must, TnbGemmJob *pool, int n_pool) {
  TNB_ARG(ctx && (must || n_must == 0) && (pool || n_pool == 0), "null");
  TNB_ARG(n_must >= 0 && n_pool >= 0 && n_must + n_pool >= 1, "no jobs");
  TNB_ARG(ctx->math_mode == TNB_MATH_3XTF32 || ctx->math_mode == TNB_MATH_BF16, "tnb_gemm_batch runs in the 3xTF32 and bf16 modes");
  const bool bf = ctx->math_mode == TNB_MATH_BF16;
  const int pairs = mg_pairs(ctx);
  // ---- the PLAN of a launch (which tiles, on which pair, in which order) depends only on the jobs' shapes, epilogue kinds and tile
  // ranges, and a training run repeats the same few launches for ever: plans are computed once and kept in the context
  std::vector<int> key;
  key.reserve(4 + 10 * (size_t)(n_must + n_pool));
  key.push_back(bf ? 1 : 0); key.push_back(pairs); key.push_back(n_must); key.push_back(n_pool);
  auto key_job = [&](const TnbGemmJob &j, int first, int count) {
    key.push_back(j.m); key.push_back(j.n); key.push_back(j.k); key.push_back(j.transa * 2 + j.transb);
    key.push_back((j.W ? 4 : 0) + (j.mulY ? 2 : 0) + (j.beta != 0.0f ? 1 : 0)); key.push_back(first); key.push_back(count);
  };
  for (int i = 0; i < n_must; i++) {
    TNB_ARG(job_ok(must[i], bf), "job operands must be 16-byte aligned with pitches that are multiples of 4 (fp32) / 8 (bf16) elements");
    const int total = job_tiles(must[i]);
    const int first = must[i].tile_count > 0 ? must[i].tile_first : 0, count = must[i].tile_count > 0 ? must[i].tile_count : total;
    TNB_ARG(first >= 0 && first + count <= total, "tile range");
    key_job(must[i], first, count);
  }
  for (int i = 0; i < n_pool; i++) {
    if (pool[i].tile_count > 0) {
      TNB_ARG(job_ok(pool[i], bf), "pool job operands must be 16-byte aligned with pitches that are multiples of 4 (fp32) / 8 (bf16) elements");
      TNB_ARG(pool[i].tile_first >= 0 && pool[i].tile_first + pool[i].tile_count <= job_tiles(pool[i]), "pool tile range");
    }
    key_job(pool[i], pool[i].tile_first, pool[i].tile_count > 0 ? pool[i].tile_count : 0);
  }
  auto &plans = ctx->mg_plans;
  auto itp = plans.find(key);
  if (itp == plans.end()) {
    MgPlan plan;
    std::vector<MgItem> its;
    // mandatory jobs first; a pool job gets a GEMM slot when its first tile is admitted
    for (int i = 0; i < n_must; i++) {
      const TnbGemmJob &j = must[i];
      const int nt = (j.n + MG_BN - 1) / MG_BN;
      const int first = j.tile_count > 0 ? j.tile_first : 0, count = j.tile_count > 0 ? j.tile_count : job_tiles(j);
      long mc, ec;
      item_cost(j, bf, &mc, &ec);
      plan.src.push_back(i);
      for (int t = first; t < first + count; t++) its.push_back(MgItem{(int)plan.src.size() - 1, (t / nt) * 256, (t % nt) * MG_BN, mc, ec});
    }
    TNB_ARG((int)plan.src.size() <= MG_MAX_GEMMS, "too many GEMMs in one batch");
    // longest-first packing: each item goes to the pair whose list it lengthens least (the first empty pair if there is one)
    std::vector<std::vector<MgItem>> lists((size_t)pairs);
    std::vector<long> load((size_t)pairs, 0);
    std::stable_sort(its.begin(), its.end(), [](const MgItem &a, const MgItem &b) { return a.main + a.epi > b.main + b.epi; });
    auto place = [&](const MgItem &t, long limit) -> bool {  // limit > 0: only if the pair's time stays within it
      long best_t = -1;
      size_t best = 0;
      for (size_t q = 0; q < lists.size(); q++) {
        std::vector<MgItem> trial = lists[q];
        trial.push_back(t);
        const long tt = list_time(trial);
        if (best_t < 0 || tt < best_t) { best_t = tt; best = q; }
        if (lists[q].empty()) break;
      }
      if (limit > 0 && best_t > limit) return false;
      lists[best].push_back(t);
      load[best] = list_time(lists[best]);
      return true;
    };
    for (const MgItem &t : its) place(t, 0);
    long makespan = 0;
    for (long l : load) makespan = std::max(makespan, l);
    // pool tiles (in order) wherever they keep the launch within 1.2x of what the mandatory work needs anyway (a tile on an otherwise
    // idle pair is nearly free; the tolerance lets equal-sized neighbours in); everything when there is no mandatory work
    plan.pool_taken.assign((size_t)n_pool, 0);
    for (int i = 0; i < n_pool; i++) {
      const TnbGemmJob &j = pool[i];
      if (j.tile_count <= 0) continue;
      if ((int)plan.src.size() >= MG_MAX_GEMMS) break;
      long mc, ec;
      item_cost(j, bf, &mc, &ec);
      const int nt = (j.n + MG_BN - 1) / MG_BN;
      int gi = -1;
      for (int t = j.tile_first; t < j.tile_first + j.tile_count; t++) {
        const int gi_try = gi < 0 ? (int)plan.src.size() : gi;
        if (!place(MgItem{gi_try, (t / nt) * 256, (t % nt) * MG_BN, mc, ec}, n_must > 0 ? makespan + makespan / 5 : 0)) break;
        if (gi < 0) { plan.src.push_back(n_must + i); gi = gi_try; }
        plan.pool_taken[(size_t)i]++;
      }
    }
    for (long l : load) makespan = std::max(makespan, l);
    plan.makespan = makespan;
    size_t ipp = 1;
    for (size_t q = 0; q < lists.size(); q++) { ipp = std::max(ipp, lists[q].size()); if (!lists[q].empty()) plan.used_pairs = (int)q + 1; }
    plan.ipp = (int)ipp;
    if (plan.used_pairs > 0) {
      std::vector<int4> h((size_t)plan.used_pairs * ipp, make_int4(-1, 0, 0, 0));
      for (int q = 0; q < plan.used_pairs; q++) {
        std::vector<MgItem> &l = lists[(size_t)q];
        (void)list_time(l);  // leaves the list in the order it is modelled in (most expensive epilogue first)
        for (size_t e = 0; e < l.size(); e++) {
          h[(size_t)q * ipp + e] = make_int4(l[e].gemm, l[e].m0, l[e].n0, 0);
          const int src = plan.src[(size_t)l[e].gemm];
          const TnbGemmJob &j = src < n_must ? must[src] : pool[src - n_must];
          plan.flops += 2.0 * (double)std::min(256, j.m - l[e].m0) * (double)std::min(MG_BN, j.n - l[e].n0) * (double)j.k;
        }
      }
      if (plans.size() > 256) {
        TNB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (auto &kv : plans) cudaFree(kv.second.dlist);
        plans.clear();
      }
      TNB_CUDA(cudaMalloc(&plan.dlist, h.size() * sizeof(int4)));
      TNB_CUDA(cudaMemcpy(plan.dlist, h.data(), h.size() * sizeof(int4), cudaMemcpyHostToDevice));  // synchronous, first use of a plan only
    }
    itp = plans.emplace(key, plan).first;
  }
  const MgPlan &plan = itp->second;
  for (int i = 0; i < n_pool; i++) { pool[i].tile_first += plan.pool_taken[(size_t)i]; pool[i].tile_count -= plan.pool_taken[(size_t)i]; }
  if (plan.used_pairs == 0) return TNB_OK;
  // ---- parameters: tensor maps (deduplicated), GEMM descriptors, the plan's work lists
  MgParams prm;
  memset(&prm, 0, sizeof(prm));
  int n_maps = 0;
  auto map_index = [&](const CUtensorMap &m) {
    for (int q = 0; q < n_maps; q++) if (!memcmp(&prm.maps[q], &m, sizeof(m))) return q;
    if (n_maps >= MG_MAX_MAPS) return -1;
    prm.maps[n_maps] = m;
    return n_maps++;
  };
  for (size_t gi = 0; gi < plan.src.size(); gi++) {
    const int src = plan.src[gi];
    const TnbGemmJob &j = src < n_must ? must[src] : pool[src - n_must];
    MgGemm &g = prm.g[gi];
    job_to_gemm(j, &g);
    CUtensorMap tmA, tmB;
    int rc;
    if (bf) {
      if (!g.a_mn) rc = get_tmap(ctx, j.A16, j.m, j.k, j.lda16, BM, 64, 0, &tmA, 2); else rc = get_tmap(ctx, j.A16, j.k, j.m, j.lda16, 64, 64, 0, &tmA, 2);
      if (rc != TNB_OK) return rc;
      if (!g.b_mn) rc = get_tmap(ctx, j.B16, j.n, j.k, j.ldb16, MG_BN / 2, 64, 0, &tmB, 2); else rc = get_tmap(ctx, j.B16, j.k, j.n, j.ldb16, 64, 64, 0, &tmB, 2);
      if (rc != TNB_OK) return rc;
    } else {
      if (!g.a_mn) rc = get_tmap(ctx, j.A, j.m, j.k, j.lda, BM, BK, 0, &tmA); else rc = get_tmap(ctx, j.A, j.k, j.m, j.lda, BK, 32, 1, &tmA);
      if (rc != TNB_OK) return rc;
      if (!g.b_mn) rc = get_tmap(ctx, j.B, j.n, j.k, j.ldb, MG_BN / 2, BK, 0, &tmB); else rc = get_tmap(ctx, j.B, j.k, j.n, j.ldb, BK, 32, 1, &tmB);
      if (rc != TNB_OK) return rc;
    }
    g.tm_a = map_index(tmA);
    g.tm_b = map_index(tmB);
    TNB_ARG(g.tm_a >= 0 && g.tm_b >= 0, "too many distinct operands in one batch");
  }
  for (int q = n_maps; q < MG_MAX_MAPS; q++) prm.maps[q] = prm.maps[0];  // the kernel prefetches every slot
  prm.items = (const int4 *)plan.dlist;
  prm.ipp = plan.ipp;
  static int want_trace = -1;
  if (want_trace < 0) { const char *e = getenv("TNB_BATCH_TRACE"); want_trace = (e && atoi(e) != 0) ? 1 : 0; }
  if (want_trace) {
    if (!ctx->mg_trace) { TNB_CUDA(cudaMalloc(&ctx->mg_trace, sizeof(long long) * 64 * 128)); }
    TNB_CUDA(cudaMemsetAsync(ctx->mg_trace, 0, sizeof(long long) * 64 * 128, ctx->stream));
    prm.trace = (long long *)ctx->mg_trace;
  }
  if (getenv("TNB_GEMM_DEBUG")) {
    fprintf(stderr, "[tnb] gemm batch: %d GEMMs, %d pairs (of %d), up to %d tiles per pair, modelled makespan %ld cycles\n", (int)plan.src.size(),
            plan.used_pairs, pairs, plan.ipp, plan.makespan);
    for (size_t gi = 0; gi < plan.src.size(); gi++)
      fprintf(stderr, "[tnb]   gemm %d: a_mn %d b_mn %d M=%d N=%d K=%d epilogue mode %d\n", (int)gi, prm.g[gi].a_mn, prm.g[gi].b_mn, prm.g[gi].M, prm.g[gi].N,
              prm.g[gi].K, prm.g[gi].ep.mode);
  }
  // ---- launch
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
    ctx->prof_flops += plan.flops;
    TNB_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * plan.used_pairs, 1);
  cfg.blockDim = dim3(bf ? MgCfg<16>::THREADS : MgCfg<3>::THREADS);
  cfg.dynamicSmemBytes = bf ? MgCfg<16>::SMEM_BYTES : MgCfg<3>::SMEM_BYTES;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (ctx->pdl && !ctx->profiling && !ctx->capturing) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  static bool attr_set[64][2] = {};
  if (!attr_set[ctx->device & 63][bf ? 1 : 0]) {
    if (bf) TNB_CUDA(cudaFuncSetAttribute(gemm_multi_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MgCfg<16>::SMEM_BYTES));
    else TNB_CUDA(cudaFuncSetAttribute(gemm_multi_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MgCfg<3>::SMEM_BYTES));
    attr_set[ctx->device & 63][bf ? 1 : 0] = true;
  }
  if (bf) TNB_CUDA(cudaLaunchKernelEx(&cfg, gemm_multi_kernel<16>, prm));
  else TNB_CUDA(cudaLaunchKernelEx(&cfg, gemm_multi_kernel<3>, prm));
  if (e1) TNB_CUDA(cudaEventRecord(e1, ctx->stream));
  TNB_LAUNCHED(ctx);
  return TNB_OK;
}

}  // extern "C"
