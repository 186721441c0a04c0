// K3 (tcgen05): fused multi-head attention for sequences of up to 128 tokens on the 5th-generation tensor cores.
//
// The mma.sync kernel in attention.cuh is instruction-issue bound (~1200 SASS instructions per warp and (sample, head):
// quad-shuffle row reductions, per-thread fragment bookkeeping, four warps each re-reading all K/V fragments), it
// reaches 47 % of the HBM roofline (profiles/r01_attention_*).  Here one CTA handles a 128-row TILE of one head:
// G = 128 / SLOT consecutive samples, each in its own SLOT-row slot (SLOT = 16 / 32 / 64 / 128 >= T: vision 2 x 50 rows
// in 64-row slots, text 1 x 77, packed text 1 x len).  Slots start at multiples of 16 rows, so a sample's keys meet
// the UMMA k-steps in the same grouping whichever slot it sits in: together with exact zeros for the other slots'
// keys this keeps the result of a sample bit-identical for every batch composition.
//
//   warp 0      TMA: Q, K, V boxes (128 rows x 64 dh, 128B swizzle) of the tile straight out of the packed qkv matrix
//   warp 1      one lane issues tcgen05.mma:  S = Q K^T  (M=128, N=keys, K=64)   accumulators in TMEM cols [0,128)
//                                             O = P V    (M=128, N=64,  K=keys)  accumulators in TMEM cols [0,64)
//   warps 2-5   softmax: thread == row.  tcgen05.ld brings a whole S row to one thread, so max / sum are plain
//               in-register reductions (no shuffles); block-diagonal (sample), causal and key-padding masks are a range
//               test + one bit test per score; P is written as bf16 into the shared-memory tiles that held Q and K (the
//               A operand of the second MMA, K-major, 128B swizzle); V is consumed in place as an MN-major B operand.
//               Epilogue: O row * 1/sum -> bf16 -> per-warp smem transpose -> 128-byte-line global stores.
//
// 48 KB of shared memory and 128 TMEM columns per CTA -> 4 CTAs per SM overlap each other's load / MMA / softmax
// phases; CTAs are persistent over the tiles.  Semantics are those of attention.cuh (SDPA with safe softmax: a query
// whose keys are all masked yields exactly 0).
#pragma once
#include "attention.cuh"
#include "gemm_tcgen05.cuh"

namespace mmcm {

constexpr int ATC_THREADS = 192;
constexpr int ATC_TILE_BYTES = 128 * 128;            // 128 rows x 64 bf16
// KMAX = most keys a query row can see (= TMEM columns of S): 128 for everything up to 128 tokens (G samples per tile),
// 256 for longer sequences (SigLIP vision, 196 tokens: one sample, ceil(T/128) row tiles that each see all keys).
// smem: Q (16 KB) | K (KMAX rows) | [pad so that P = 128 x KMAX bf16 fits over Q|K|pad] | V (KMAX rows)
template <int KMAX>
struct AtcCfg {
  static constexpr int K_BYTES = KMAX * 128;
  static constexpr int P_BYTES = (KMAX / 64) * ATC_TILE_BYTES;
  static constexpr int V_OFF = P_BYTES > ATC_TILE_BYTES + K_BYTES ? P_BYTES : ATC_TILE_BYTES + K_BYTES;
  static constexpr int SMEM_BYTES = V_OFF + K_BYTES + 1024;   // + 1 KB alignment slack
  static constexpr int TMEM_COLS = KMAX;
  static constexpr int KW = KMAX / 32;                         // 32-key validity words / score chunks
  static constexpr int CTAS_PER_SM = 512 / KMAX;               // TMEM (and smem: 49 KB / 97 KB) limit
};

// instruction descriptor: D=f32, A=B=bf16, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__device__ __forceinline__ uint32_t atc_idesc(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// qkv : bf16 [rows_total, 3*D] via tensor map `tmap_qkv` (box 64 dh x box_rows);  out : bf16 [rows_total, D]
// fixed-length mode (seq_start == nullptr): sample b owns rows [b*T, (b+1)*T), box_rows = T, G samples per tile
// packed mode: sample b owns rows [seq_start[b], +seq_len[b]), box_rows = 128, one sample per tile
// key_valid: one byte per qkv row (fixed-length: [B, T] contiguous) or nullptr
template <int KMAX>
__global__ void __launch_bounds__(ATC_THREADS)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    __nv_bfloat16* __restrict__ out, const uint8_t* __restrict__ key_valid,
                    const int* __restrict__ seq_start, const int* __restrict__ seq_len, const int T_fixed, const int D,
                    const int causal, const int B, const int heads, const int q_box, const int kv_box,
                    long long* __restrict__ trace) {
  using C = AtcCfg<KMAX>;
  extern __shared__ uint8_t atc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_s, bar_p, bar_o, bar_free;
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem = (smem_u32(atc_smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem, sK = smem + ATC_TILE_BYTES, sV = smem + C::V_OFF;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), 128);
    mbar_init(smem_u32(&bar_o), 1);
    mbar_init(smem_u32(&bar_free), 128);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_imm<C::TMEM_COLS>(smem_u32(&tmem_holder));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  pdl_trigger();
  pdl_wait();

  const bool packed = seq_start != nullptr;
  const bool longseq = !packed && T_fixed > 128;             // one sample per tile, RT row tiles, all keys visible
  // slot = rows reserved per sample inside the tile (power of two >= T, at least one UMMA k-step)
  const int slot_shift = (packed || longseq) ? 7 : (T_fixed <= 16 ? 4 : (T_fixed <= 32 ? 5 : (T_fixed <= 64 ? 6 : 7)));
  const int SLOT = 1 << slot_shift;
  const int G = 128 >> slot_shift;                            // samples per tile
  const int RT = longseq ? (T_fixed + 127) >> 7 : 1;          // row tiles per sample
  const int groups = (B + G - 1) / G;
  const int total = groups * heads * RT;

  // V padding rows (slot rows >= T) are never written by TMA: zero the V tile once so that 0-probability x stale
  // shared memory cannot produce NaN.  (The Q | K tiles are rewritten with finite P values every tile.)
  for (int i = threadIdx.x; i < (C::V_OFF + C::K_BYTES) / 16; i += ATC_THREADS) sts128(smem + i * 16, 0u, 0u, 0u, 0u);
  fence_proxy_async();
  __syncthreads();

  // tile -> head, row tile, first sample, samples in the tile, rows per sample
  auto tile_info = [&](int item, int& h, int& rt, int& b0, int& ns, int& Tcur) {
    h = item % heads;
    int r = item / heads;
    rt = r % RT;
    r /= RT;
    b0 = r * G;
    ns = min(G, B - b0);
    Tcur = packed ? seq_len[b0] : T_fixed;
  };
  auto first_row = [&](int b) { return packed ? seq_start[b] : b * T_fixed; };
  // keys of the tile, rounded to the UMMA N / K step
  auto key_extent = [&](int ns, int Tcur) { return (((ns - 1) << slot_shift) + Tcur + 15) & ~15; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int h, rt, b0, ns, Tcur;
      tile_info(item, h, rt, b0, ns, Tcur);
      mbar_wait(smem_u32(&bar_free), ph ^ 1u);       // previous tile fully consumed (smem and TMEM)
      if (lane == 0) {
        const uint32_t full = smem_u32(&bar_full);
        // one Q box (64 dh x q_box rows) and one K / V box (64 dh x kv_box rows) per sample slot
        mbar_expect_tx(full, (uint32_t)(ns * (q_box + 2 * kv_box) * 128));
        for (int g = 0; g < ns; ++g) {
          const int row = first_row(b0 + g);
          const uint32_t off = (uint32_t)(g << slot_shift) * 128u;
          tma_load_2d(&tmap_q, full, sQ + off, h * ATT_DH, row + rt * 128);
          tma_load_2d(&tmap_kv, full, sK + off, D + h * ATT_DH, row);
          tma_load_2d(&tmap_kv, full, sV + off, 2 * D + h * ATT_DH, row);
        }
      }
      __syncwarp();
      ph ^= 1u;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int h, rt, b0, ns, Tcur;
      tile_info(item, h, rt, b0, ns, Tcur);
      const int nk = key_extent(ns, Tcur);
      mbar_wait(smem_u32(&bar_full), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t id1 = atc_idesc(128, nk, 0);
        const uint64_t qd = make_smem_desc_sw128(sQ), kd = make_smem_desc_sw128(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), id1, k != 0);   // S = Q K^T
        umma_commit(smem_u32(&bar_s));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar_p), ph);                // P written (and S fully read) by all 128 softmax threads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t id2 = atc_idesc(128, ATT_DH, 1);
        const uint64_t vd = make_smem_desc_sw128(sV);
        for (int k = 0; k < nk / 16; ++k) {           // O = P V; P slab k/4 (64 keys each), V advances 16 key rows = 2 KB
          const uint64_t pd = make_smem_desc_sw128(sQ + (k >> 2) * ATC_TILE_BYTES) + (uint64_t)(2 * (k & 3));
          umma_f16(tmem, pd, vd + (uint64_t)(k * 128), id2, k != 0);
        }
        umma_commit(smem_u32(&bar_o));
      }
      __syncwarp();
      ph ^= 1u;
    }
  } else {
    // ===================== softmax + epilogue: thread == tile row =====================
    const int lg = warp & 3;                          // TMEM lane group this warp may access
    const int r = lg * 32 + lane;                     // tile row of this thread
    const uint32_t t_row = tmem + ((uint32_t)(lg * 32) << 16);
    const float L2E = 1.4426950408889634f;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int h, rt, b0, ns, Tcur;
      tile_info(item, h, rt, b0, ns, Tcur);
      // keys this row may attend: its own sample's slot, cut at the diagonal when causal
      const int g = r >> slot_shift;
      const int q = (r & (SLOT - 1)) + rt * 128;                     // position of this row inside its sample
      const bool row_ok = g < ns && q < Tcur;
      const int lo = g << slot_shift;
      const int hi = row_ok ? lo + (causal ? q + 1 : Tcur) : lo;     // empty range for padding rows
      const int grow = row_ok ? first_row(b0 + g) + q : 0;           // global qkv / out row of this thread
      const int nk = key_extent(ns, Tcur);
      const int nchunks = (nk + 31) >> 5;                            // 32-column chunks covering the MMA's K extent
      uint32_t kv[C::KW];
#pragma unroll
      for (int w = 0; w < C::KW; ++w) {
        const int c = w * 32 + lane;
        const int cg = c >> slot_shift, ck = c & (SLOT - 1);
        bool ok = longseq ? c < Tcur : (cg < ns && ck < Tcur);
        if (ok && key_valid) ok = key_valid[(size_t)first_row(b0 + (longseq ? 0 : cg)) + (longseq ? c : ck)] != 0;
        kv[w] = __ballot_sync(0xffffffffu, ok);
      }

      const bool tr = trace && blockIdx.x == 0 && r == 0 && item == blockIdx.x + gridDim.x;   // 2nd tile of CTA 0
      if (tr) trace[0] = clock64();
      mbar_wait(smem_u32(&bar_s), ph);
      tc_fence_after();
      if (tr) trace[1] = clock64();
      // ---- pass 1: row maximum over the allowed keys.  A chunk is "inner" when all its 32 keys are allowed for this
      // row (then no per-score masking is needed) and "outer" when none is (then it is skipped).
      // (tcgen05.ld is warp-collective: whether a chunk is loaded at all must be a warp-uniform decision.)
      float mx = -INFINITY;
      for (int ci = 0; ci < nchunks; ++ci) {
        const int c0 = ci * 32;
        const bool outer = c0 >= hi || c0 + 32 <= lo;
        if (__all_sync(0xffffffffu, outer)) continue;
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        const uint32_t bits = kv[ci];
        if (outer) {
        } else if (bits == 0xffffffffu && c0 >= lo && c0 + 32 <= hi) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            const bool ok = ((bits >> j) & 1u) && c >= lo && c < hi;
            mx = fmaxf(mx, ok ? __uint_as_float(v[j]) : -INFINITY);
          }
        }
      }
      if (mx == -INFINITY) mx = 0.f;                  // nothing to attend: every p below is 0 -> output row 0
      const float nm = -mx * L2E;
      if (tr) trace[2] = clock64();
      // ---- pass 2: p = 2^((s - max) log2 e), row sum, bf16 P into the Q|K tiles (K-major, 128B swizzle)
      float sum = 0.f;
      for (int ci = 0; ci < nchunks; ++ci) {
        const int c0 = ci * 32;
        const bool outer = c0 >= hi || c0 + 32 <= lo;
        uint32_t pw[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pw[j] = 0u;
        if (!__all_sync(0xffffffffu, outer)) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          const uint32_t bits = kv[ci];
          if (outer) {
          } else if (bits == 0xffffffffu && c0 >= lo && c0 + 32 <= hi) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = ex2_fast(fmaf(__uint_as_float(v[j]), L2E, nm));
              const float p1 = ex2_fast(fmaf(__uint_as_float(v[j + 1]), L2E, nm));
              sum += p0 + p1;
              pw[j >> 1] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int c = c0 + j;
              const bool ok0 = ((bits >> j) & 1u) && c >= lo && c < hi;
              const bool ok1 = ((bits >> (j + 1)) & 1u) && c + 1 >= lo && c + 1 < hi;
              const float p0 = ok0 ? ex2_fast(fmaf(__uint_as_float(v[j]), L2E, nm)) : 0.f;
              const float p1 = ok1 ? ex2_fast(fmaf(__uint_as_float(v[j + 1]), L2E, nm)) : 0.f;
              sum += p0 + p1;
              pw[j >> 1] = pack_bf16x2(p0, p1);
            }
          }
        }
        // chunk ci = keys [32 ci, 32 ci + 32) = 16-byte slots 4*(ci&1) .. +3 of row r in P slab ci>>1
        const uint32_t slab = sQ + (ci >> 1) * ATC_TILE_BYTES + r * 128;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int slot = (ci & 1) * 4 + s4;
          sts128(slab + ((slot ^ (r & 7)) << 4), pw[4 * s4], pw[4 * s4 + 1], pw[4 * s4 + 2], pw[4 * s4 + 3]);
        }
      }
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      if (tr) trace[3] = clock64();
      tc_fence_before();
      fence_proxy_async();                            // P (generic proxy) -> visible to the tensor core's async proxy
      mbar_arrive(smem_u32(&bar_p));

      // ---- epilogue: O row / sum -> bf16 -> transpose through this warp's 32 rows of the (now free) Q tile
      mbar_wait(smem_u32(&bar_o), ph);
      tc_fence_after();
      if (tr) trace[4] = clock64();
      {
        uint32_t o0[32], o1[32];
        tmem_ld32(t_row, o0);
        tmem_ld32(t_row + 32u, o1);
        tmem_ld_wait();
        const uint32_t dst = sQ + r * 128;
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = s8 * 8 + 2 * j;
            const float a = __uint_as_float(e < 32 ? o0[e] : o1[e - 32]) * inv;
            const float b = __uint_as_float(e + 1 < 32 ? o0[e + 1] : o1[e + 1 - 32]) * inv;
            w[j] = pack_bf16x2(a, b);
          }
          sts128(dst + ((s8 ^ (r & 7)) << 4), w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      {
        const int sub = lane >> 3, c16 = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = lg * 32 + it * 4 + sub;      // tile row stored by this lane
          const uint4 v = lds128(sQ + rl * 128 + ((c16 ^ (rl & 7)) << 4));
          const int gro = __shfl_sync(0xffffffffu, row_ok ? grow : -1, it * 4 + sub);   // that row's global row (or -1)
          if (gro >= 0) *reinterpret_cast<uint4*>(out + (size_t)gro * D + h * ATT_DH + c16 * 8) = v;
        }
      }
      __syncwarp();
      if (tr) trace[5] = clock64();
      mbar_arrive(smem_u32(&bar_free));               // smem tiles and TMEM may be reused for the next tile
      ph ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_imm<C::TMEM_COLS>(tmem);
  }
}

}  // namespace mmcm
