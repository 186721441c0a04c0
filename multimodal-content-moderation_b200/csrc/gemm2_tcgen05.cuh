// K2 (2-CTA): out[M,N] = epilogue(A[M,K] @ W[N,K]^T) with tcgen05.mma.cta_group::2 -- one 256 x BLOCK_N tile per
// CTA PAIR (two SMs of one TPC, launched as a 2-CTA cluster).
//
// Why: with one CTA per tile the tensor pipe reads A (4 KB) + B (8 KB) from shared memory for every 128-cycle
// UMMA while TMA writes the same 96 B/cycle back in -- more than the SM's shared-memory bandwidth, which capped the
// single-CTA kernel at ~59 % tensor-pipe activity (profiles/r01_gemm_v2_*.txt).  In pair mode each CTA stages its own
// 128 rows of A and only HALF of the B tile; the pair's UMMA (M = 256) reads both halves, so shared-memory and L2
// traffic per FLOP drop by a third and the ring holds 6 stages instead of 4.
//
// Protocol (rank 0 = leader):
//   * both producers load their halves into their own smem and complete_tx on the LEADER's full barrier
//     (cp.async.bulk.tensor ... .cta_group::2, barrier address with the peer bit cleared); the leader's producer
//     arms it with expect_tx(2 x stage bytes)
//   * the leader's MMA warp issues tcgen05.mma.cta_group::2 and releases ring slots / publishes accumulators with
//     tcgen05.commit ... multicast::cluster to the same barrier offset in BOTH CTAs
//   * each CTA's 8 epilogue warps drain that CTA's 128 accumulator rows from its own TMEM (same coalesced epilogue
//     as the single-CTA kernel) and arrive on the leader's tmem-empty barrier (remote mbarrier.arrive)
//   * tiles are assigned round-robin to the pairs (static: the pair shares no scheduler state)
#pragma once
#include "gemm_tcgen05.cuh"

namespace mmcm {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (pair leader)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the pair leader's copy of `bar`.  RELAXED: the only thing the waiter (the MMA warp) does afterwards is
// overwrite TMEM, whose reads by this thread are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; no
// generic-proxy memory is handed over.  The default .release.cluster arrive cost ~2.6 k cycles per tile here
// (tools/gemm_trace.py: acc-ready -> epilogue-done 7.5 k cycles although the two chunks took 4.9 k).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread have retired) on `bar` in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

// ---- TMA store epilogue ------------------------------------------------------------------------------
// The per-thread STG read-back of the first pair kernel cost 1.4-1.9 k cycles per 32x64 chunk per warp
// (tools/gemm_trace.py): the SM's LSU store path, not the math, bounded the K=512 GEMMs.  Here a warp writes its
// 32 rows x 128 B chunk into a 128B-swizzled staging tile (conflict free: 16-byte slot i of row r goes to slot
// i ^ (r & 7)) and ONE lane hands it to the TMA unit:
//   bf16 outputs : cp.async.bulk.tensor.2d.global.shared::cta          (plain tile store)
//   fp32 residual: cp.reduce.async.bulk.tensor.2d ... .add.f32          (x += acc + bias inside L2: the residual
//                  stream is never read into the SM, and fp32 addition is commutative so the result is identical)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
constexpr int EPI_TMA_STAGE_BYTES = 32 * 128;   // per epilogue warp: 32 rows x 128 B, 128B-swizzled, 1 KB aligned
// Staging tiles per epilogue warp for the TMA-store epilogues.  With one tile every chunk waits for the previous
// chunk's bulk store to finish READING the tile before refilling it; two tiles (chunk c fills tile c & 1,
// cp.async.bulk.wait_group.read 1) remove that wait but cost one operand-ring stage.  Measured (tools/gemm_bench_fold.py,
// bench.py): two tiles + 4 stages 56.5 k samples/s, one tile + 5 stages 57.2 k -- the store wait is not what bounds the
// K = 512 GEMMs, the deeper ring is worth more.  Kept as a build switch.
#ifndef MMCM_EPI_NSTG
#define MMCM_EPI_NSTG 1
#endif

// ---- LN fold ------------------------------------------------------------------------------------------
// LayerNorm is row-wise over the K dimension of the GEMM that consumes it:
//     LN(x) W^T + b = rstd * ((x - mean) (W*gamma)^T) + b' = rstd * (x W'^T) + b',   b' = b + W beta,
//     W' = W*gamma with every ROW centred (sum_k W'[n,k] = 0, so x W'[n] = (x - mean(x)) W'[n] for any x)
// so the separate LayerNorm pass (read x fp32, write bf16: 6 B per element, 10 % of the forward) can go:
//   * the PRODUCER of x -- the residual GEMM (out_proj / fc2), EPI_RESID_STATS -- pulls its x tile in by TMA,
//     adds acc + bias in place, stores it back by TMA together with a bf16 copy (the consumer's A operand) and leaves,
//     per row and 128-column slab, (sum, M2 about the slab mean): thread == row, so no shuffles.
//   * the CONSUMER -- qkv / fc1, EPI_LNFOLD_* -- multiplies the un-normalised bf16 rows by W' and applies rstd
//     (variance merged from the slabs with Chan's formula, fixed order) in its epilogue: one FMA per element, the
//     same instruction and shared-memory traffic as the plain bias epilogue.  (A first version kept W*gamma
//     un-centred and subtracted mean * colsum in the epilogue: the extra column-sum loads from shared memory, which
//     the main loop already saturates, cost fc1 10 % -- tools/gemm_bench_fold.py.)
// DRAM per residual element: 4 (read x) + 4 (write x) + 2 (write bf16) = 10 B instead of 8 (L2 reduce-add) + 6 (LN).
#ifndef MMCM_STATS_XBUFS
#define MMCM_STATS_XBUFS 2    /* in-place x tiles (32 rows x 32 fp32, 4 KB) per epilogue warp */
#endif
#ifndef MMCM_STATS_STAGES
#define MMCM_STATS_STAGES 4   /* operand ring of the EPI_RESID_STATS instantiation (the x tiles need the room) */
#endif
#ifndef MMCM_STATS_EARLY
#define MMCM_STATS_EARLY 0    /* 1: refill an x buffer at the top of the next chunk instead of in its middle */
#endif
constexpr int EPI_X_TILE_BYTES = 32 * 128;    // fp32 32 x 32, 128B-swizzled
constexpr int EPI_XB_TILE_BYTES = 32 * 64;    // bf16 32 x 32, 64B-swizzled

template <int BLOCK_N, int EPI = EPI_BIAS_BF16>
struct Gemm2Cfg {
  static constexpr int BLOCK_M = 256;           // per pair; 128 rows per CTA
  static constexpr int BLOCK_K = 64;
  static constexpr int UMMA_K = 16;
  static constexpr int A_BYTES = 128 * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8;
  static constexpr bool kStats = (EPI == EPI_RESID_STATS);
  static constexpr bool kFold = (EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_ACT_BF16);
  static constexpr int XBUFS = MMCM_STATS_XBUFS;
  static constexpr int EPI_BIAS_BYTES = (BLOCK_N / 2) * 4;
  static constexpr int EPI_STAGE = kStats ? XBUFS * EPI_X_TILE_BYTES : MMCM_EPI_NSTG * EPI_TMA_STAGE_BYTES;
  static constexpr int EPI_STAGE2 = kStats ? EPI_XB_TILE_BYTES : 0;
#ifndef MMCM_PAIR_STAGES
#define MMCM_PAIR_STAGES (MMCM_EPI_NSTG == 2 ? 4 : 5)   /* 4 stages already saturate the TMA->UMMA loop (tools/mainloop_probe.cu) */
#endif
#ifndef MMCM_PAIR_REG_THREADS
#define MMCM_PAIR_REG_THREADS 384   /* __launch_bounds__ thread count used only to cap registers per thread */
#endif
  static constexpr int STAGES = kStats ? MMCM_STATS_STAGES : ((BLOCK_N == 256) ? MMCM_PAIR_STAGES : MMCM_PAIR_STAGES + 2);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * (EPI_STAGE + EPI_STAGE2 + EPI_BIAS_BYTES) + 1024;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static_assert(SMEM_BYTES <= 232448, "CTA shared memory over the 227 KB limit");
};

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MMCM_PAIR_REG_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_d,
                     const EpiParams ep, const int M_host, const int N, const int K, const int tma_out) {
  using C = Gemm2Cfg<BLOCK_N, EPI>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES];    // used in the leader only
  __shared__ __align__(8) uint64_t bar_empty[C::STAGES];   // one copy per CTA (multicast commit)
  __shared__ __align__(8) uint64_t bar_tfull[2];           // one copy per CTA (multicast commit)
  __shared__ __align__(8) uint64_t bar_tempty[2];          // used in the leader only: 2 x EPI_WARPS arrivals
  __shared__ __align__(8) uint64_t bar_x[C::EPI_WARPS][4];  // EPI_RESID_STATS: x tile of epilogue warp e, buffer j, has landed
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_stamp(ep, 0);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;

  const int tiles_n = N / BLOCK_N;
  const int num_kb = K / C::BLOCK_K;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (tma_out) prefetch_tmap(&tmap_c);
    if (C::kStats) prefetch_tmap(&tmap_d);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 2 * C::EPI_WARPS);
    }
    if (C::kStats)
      for (int s = 0; s < C::EPI_WARPS * 4; ++s) mbar_init(smem_u32(&bar_x[0][0] + s), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_holder), C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of BOTH CTAs are visible before any remote traffic
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  if (threadIdx.x == 0) trace_stamp(ep, 1);
  // PDL: barrier init, TMEM allocation and the cluster handshake above overlapped the previous kernel's tail
  pdl_trigger();
  pdl_wait();
  // packed variable-length text: the live row count is produced on the device by the previous kernel
  const int M = ep.m_dev ? min(__ldg(ep.m_dev), M_host) : M_host;
  const int tiles_m = (M + C::BLOCK_M - 1) / C::BLOCK_M;
  // split-K: only the fp32 store epilogue can hold partial sums (EpiParams::ksplit)
  const int ksplit = (EPI == EPI_BIAS_RESID_F32 && ep.ksplit > 1) ? ep.ksplit : 1;
  const int kb_per = num_kb / ksplit;
  const int num_tiles = tiles_m * tiles_n * ksplit;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int mn = tile / ksplit, ks = tile - mn * ksplit;
      const int m_blk = mn / tiles_n, n_blk = mn - m_blk * tiles_n;
      const int a_row = m_blk * C::BLOCK_M + (int)rank * 128;
      const int b_row = n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2);
      for (int kb = 0; kb < kb_per; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        if (lane == 0) {
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const int kcol = (ks * kb_per + kb) * C::BLOCK_K;
          if (leader) mbar_expect_tx(full, 2 * C::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
          tma_load_2d_pair(&tmap_a, full, sa, kcol, a_row);
          tma_load_2d_pair(&tmap_b, full, sa + C::A_BYTES, kcol, b_row);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader only) =====================
    constexpr uint32_t idesc = make_idesc_bf16(C::BLOCK_M, BLOCK_N);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);   // both CTAs' epilogues have drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
      for (int kb = 0; kb < kb_per; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          if (kb == 0 && tile == pair) trace_stamp(ep, 2);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k)
            umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit_pair(smem_u32(&bar_empty[stage]));
          if (kb == kb_per - 1) {
            umma_commit_pair(smem_u32(&bar_tfull[as]));
            if (tile == pair) trace_stamp(ep, 3);
            trace_stamp(ep, 6);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): this CTA's 128 rows of the pair's tile =====================
    const int e = warp - 4;
    const int lg = e & 3, ch = e >> 2;
    const uint32_t stage_smem = epi_base + e * C::EPI_STAGE;   // 1 KB aligned (epi_base is, 4 / 8 / 12 KB per warp)
    const uint32_t stage2_smem = epi_base + C::EPI_WARPS * C::EPI_STAGE + e * C::EPI_STAGE2;   // bf16 tile (kStats)
    const uint32_t bias_smem = epi_base + C::EPI_WARPS * (C::EPI_STAGE + C::EPI_STAGE2) + e * C::EPI_BIAS_BYTES;
    const bool tma_path = tma_out != 0 && EPI != EPI_PATCH_F32;
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr bool kF32 = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_PATCH_F32);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t xph = 0;   // kStats: phase bit per x buffer
    uint32_t cc = 0;    // TMA-store epilogues: running chunk counter -> staging tile cc % MMCM_EPI_NSTG
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int mn = tile / ksplit, ks = tile - mn * ksplit;
      const int m_blk = mn / tiles_n, n_blk = mn - m_blk * tiles_n;
      const int row_base = m_blk * C::BLOCK_M + (int)rank * 128 + lg * 32;
      const int col_base = n_blk * BLOCK_N + ch * HALF_N;
      const int out_row = row_base + ks * ep.part_rows;          // split-K: plane ks of the partial-sum buffer
      float4 xa[8];
      float4 fb[HALF_N / 32];
      if (kF32 && !tma_path) {
#pragma unroll
        for (int c = 0; c < HALF_N / 32; ++c)
          fb[c] = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base + c * 32 + (lane & 7) * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_base < M) epi_load_addend<EPI>(ep, lane, row_base, col_base, M, xa);
      } else {
        // QuickGELU epilogues work on halved constants (epi_pack_bf16): exact, one instruction fewer per element
        constexpr bool kActEpi = (EPI == EPI_BIAS_ACT_BF16 || EPI == EPI_LNFOLD_ACT_BF16);
        const float hs = (kActEpi && ep.act == ACT_QUICK_GELU) ? 0.5f : 1.0f;
#pragma unroll
        for (int j = lane; j < HALF_N / 4; j += 32) {
          const float4 b = (ep.bias && ks == 0) ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base) + j)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          sts128(bias_smem + j * 16, __float_as_uint(b.x * hs), __float_as_uint(b.y * hs), __float_as_uint(b.z * hs),
                 __float_as_uint(b.w * hs));
        }
        __syncwarp();
      }
      // LN fold, consumer side: this lane's row statistics, merged from the 128-column slabs in fixed order
      float ln_rs = 0.f;   // rstd; rows beyond M keep 0 -> they store the (finite) bias
      if (C::kFold) {
        const int row = row_base + lane;
        if (row < M) {
          const int ns = ep.ln_slabs;
          float2 p[8];
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < ns) {
              p[i] = __ldg(ep.stats + (size_t)i * ep.stats_pitch + row);
              sum += p[i].x;
            }
          const float inv_d = 1.0f / (float)(ns * LN_SLAB);
          const float mean = sum * inv_d;
          float m2 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < ns) {
              const float d = p[i].x * (1.0f / LN_SLAB) - mean;
              m2 += p[i].y + (float)LN_SLAB * d * d;
            }
          ln_rs = rsqrtf(m2 * inv_d + ep.ln_eps);
          if (EPI == EPI_LNFOLD_ACT_BF16 && ep.act == ACT_QUICK_GELU) ln_rs *= 0.5f;   // the bias is staged halved
        }
      }
      // LN fold, producer side: pull the first x tiles of this warp in while the main loop runs
      if (C::kStats) {
        if (row_base < M && lane == 0) {
          bulk_wait_read0();       // the previous tile's stores have finished reading the x tiles / the bf16 tile
          fence_proxy_async();
#pragma unroll
          for (int j = 0; j < C::XBUFS; ++j) {
            const uint32_t bar = smem_u32(&bar_x[e][j]);
            mbar_expect_tx(bar, EPI_X_TILE_BYTES);
            tma_load_2d(&tmap_c, bar, stage_smem + j * EPI_X_TILE_BYTES, col_base + j * 32, row_base);
          }
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      if (e == 0 && lane == 0) { if (tile == pair) trace_stamp(ep, 4); trace_stamp(ep, 7); }
      const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BLOCK_N + ch * HALF_N);
      if (row_base < M) {
        if constexpr (C::kStats) {
          // x_new = x + acc + bias in place (thread == row), bf16 copy, running (sum, M2) of this row's 128 columns
          static_assert(!C::kStats || HALF_N == LN_SLAB, "EPI_RESID_STATS needs BLOCK_N == 256");
          constexpr int NCH = HALF_N / 32;
          float run_s = 0.f, run_m2 = 0.f;
          const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int j = c % C::XBUFS;
            const uint32_t xt = stage_smem + j * EPI_X_TILE_BYTES + lane * 128;
#if MMCM_STATS_EARLY
            if (c >= 1) {   // refill the previous chunk's x buffer before touching this chunk
              if (lane == 0) {
                bulk_wait_read0();
                const int cn = c - 1 + C::XBUFS;
                if (cn < NCH) {
                  const int jn = (c - 1) % C::XBUFS;
                  const uint32_t bar = smem_u32(&bar_x[e][jn]);
                  fence_proxy_async();
                  mbar_expect_tx(bar, EPI_X_TILE_BYTES);
                  tma_load_2d(&tmap_c, bar, stage_smem + jn * EPI_X_TILE_BYTES, col_base + cn * 32, row_base);
                }
              }
              __syncwarp();
            }
#endif
            mbar_wait(smem_u32(&bar_x[e][j]), (xph >> j) & 1u);
            xph ^= 1u << j;
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            float cs = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 xo = lds128(xt + ((i ^ sw) << 4));
              const uint4 bb = lds128(bias_smem + c * 128 + i * 16);
              const float v0 = (__uint_as_float(r[4 * i]) + __uint_as_float(bb.x)) + __uint_as_float(xo.x);
              const float v1 = (__uint_as_float(r[4 * i + 1]) + __uint_as_float(bb.y)) + __uint_as_float(xo.y);
              const float v2 = (__uint_as_float(r[4 * i + 2]) + __uint_as_float(bb.z)) + __uint_as_float(xo.z);
              const float v3 = (__uint_as_float(r[4 * i + 3]) + __uint_as_float(bb.w)) + __uint_as_float(xo.w);
              r[4 * i] = __float_as_uint(v0); r[4 * i + 1] = __float_as_uint(v1);
              r[4 * i + 2] = __float_as_uint(v2); r[4 * i + 3] = __float_as_uint(v3);
              cs += (v0 + v1) + (v2 + v3);
            }
            const float cm = cs * (1.0f / 32.0f);
            float cq = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = __uint_as_float(r[i]) - cm;
              cq = fmaf(d, d, cq);
            }
            if (c == 0) { run_s = cs; run_m2 = cq; }
            else {   // Chan: merge 32 new elements into the 32 c seen so far
              const float n0 = 32.0f * c;
              const float d = cm - run_s / n0;
              run_m2 += cq + d * d * (n0 * 32.0f / (n0 + 32.0f));
              run_s += cs;
            }
            // the previous chunk's stores have read their tiles: its x buffer can take the chunk XBUFS ahead of it
            if (!MMCM_STATS_EARLY && c >= 1) {
              if (lane == 0) {
                bulk_wait_read0();
                const int cn = c - 1 + C::XBUFS;
                if (cn < NCH) {
                  const int jn = (c - 1) % C::XBUFS;
                  const uint32_t bar = smem_u32(&bar_x[e][jn]);
                  fence_proxy_async();
                  mbar_expect_tx(bar, EPI_X_TILE_BYTES);
                  tma_load_2d(&tmap_c, bar, stage_smem + jn * EPI_X_TILE_BYTES, col_base + cn * 32, row_base);
                }
              }
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sts128(xt + ((i ^ sw) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            {   // bf16 copy: 32 rows x 64 B, 64B swizzle (16-byte slot i of row r sits at slot i ^ ((r >> 1) & 3))
              const uint32_t bt = stage2_smem + lane * 64;
              const uint32_t sw2 = (uint32_t)((lane >> 1) & 3);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                sts128(bt + ((i ^ sw2) << 4),
                       pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                       pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                       pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                       pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_c, stage_smem + j * EPI_X_TILE_BYTES, col_base + c * 32, row_base);
              tma_store_2d(&tmap_d, stage2_smem, col_base + c * 32, row_base);
              bulk_commit();
            }
          }
          const int row = row_base + lane;
          if (row < M) ep.stats[(size_t)(col_base / LN_SLAB) * ep.stats_pitch + row] = make_float2(run_s, run_m2);
        } else if (!kF32 && HALF_N == 32) {
          // 64-column tiles (single-row-block GEMMs): one 32-column chunk per warp, staged as a 32 x 64 B tile with the
          // 64B swizzle (16-byte slot i of row r sits at slot i ^ ((r >> 1) & 3)) and stored by TMA
          uint32_t r[32], w[16];
          tmem_ld32(t_row, r);
          tmem_ld_wait();
          epi_pack_bf16<EPI>(ep, bias_smem, r, w, ln_rs);
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          const uint32_t bt = stage_smem + lane * 64;
          const uint32_t sw2 = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int i = 0; i < 4; ++i) sts128(bt + ((i ^ sw2) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_c, stage_smem, col_base, row_base);
            bulk_commit();
          }
        } else if (!kF32) {
#pragma unroll 1
          for (int c = 0; c < HALF_N / 64; ++c) {
            const bool tr = (e == 0 && lane == 0 && tile == pair);
            if (tr && c == 0) trace_stamp(ep, 10);
            uint32_t w[32];
            {
              uint32_t r[32];
              tmem_ld32(t_row + (uint32_t)(c * 64), r);
              tmem_ld_wait();
              if (tr && c == 0) trace_stamp(ep, 11);
              epi_pack_bf16<EPI>(ep, bias_smem + c * 256, r, &w[0], ln_rs);
            }
            {
              uint32_t r[32];
              tmem_ld32(t_row + (uint32_t)(c * 64 + 32), r);
              tmem_ld_wait();
              epi_pack_bf16<EPI>(ep, bias_smem + c * 256 + 128, r, &w[16], ln_rs);
            }
            if (tr && c == 0) trace_stamp(ep, 14);
            if (tma_path) {
              const uint32_t sbuf = stage_smem + (cc % MMCM_EPI_NSTG) * EPI_TMA_STAGE_BYTES;
              ++cc;
              if (lane == 0) bulk_wait_read<MMCM_EPI_NSTG - 1>();   // the store that last used this tile has read it
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                sts128(sbuf + lane * 128 + ((i ^ (lane & 7)) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
              fence_proxy_async();                          // generic-proxy writes -> visible to the TMA unit
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmap_c, sbuf, col_base + c * 64, row_base);
                bulk_commit();
              }
            } else {
              epi_store_bf16(ep, stage_smem, lane, row_base, col_base + c * 64, M, w);
            }
            if (tr && c == 0) trace_stamp(ep, 12);
            if (tr && c == 1) trace_stamp(ep, 13);
          }
        } else if (tma_path) {
#pragma unroll 1
          for (int c = 0; c < HALF_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            const uint32_t sbuf = stage_smem + (cc % MMCM_EPI_NSTG) * EPI_TMA_STAGE_BYTES;
            ++cc;
            if (lane == 0) bulk_wait_read<MMCM_EPI_NSTG - 1>();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 b = lds128(bias_smem + c * 128 + i * 16);
              sts128(sbuf + lane * 128 + ((i ^ (lane & 7)) << 4),
                     __float_as_uint(__uint_as_float(r[4 * i]) + __uint_as_float(b.x)),
                     __float_as_uint(__uint_as_float(r[4 * i + 1]) + __uint_as_float(b.y)),
                     __float_as_uint(__uint_as_float(r[4 * i + 2]) + __uint_as_float(b.z)),
                     __float_as_uint(__uint_as_float(r[4 * i + 3]) + __uint_as_float(b.w)));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (ep.resid) tma_reduce_add_2d(&tmap_c, sbuf, col_base + c * 32, row_base);   // x += acc + bias
              else tma_store_2d(&tmap_c, sbuf, col_base + c * 32, out_row);
              bulk_commit();
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < HALF_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            epi_chunk_f32<EPI>(ep, stage_smem, lane, row_base, col_base + c * 32, M, r, xa, fb[c]);
            if (c + 1 < HALF_N / 32) epi_load_addend<EPI>(ep, lane, row_base, col_base + (c + 1) * 32, M, xa);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&bar_tempty[as]));
      if (e == 0 && lane == 0) { if (tile == pair) trace_stamp(ep, 5); trace_stamp(ep, 8); }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (tma_path && lane == 0) bulk_wait0();   // all bulk stores of this warp are complete (globally visible)
  }

  if (threadIdx.x == 0) trace_stamp(ep, 9);
  tc_fence_before();
  cluster_sync_all();   // nobody frees TMEM / exits while the peer may still read its smem or write its TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace mmcm
