// K3 (ring): the mma.sync attention of attention.cuh fed by a TMA ring instead of per-thread cp.async.
//
// Why: attention_kernel<80, 5> (77-token causal CLIP text) sat at 48 % of the HBM roofline with 43 % of the issue slots
// busy (profiles/r02_other_kernels_ncu_full.txt).  Its source counters show where the time goes: a warp spends ~980
// SASS instructions per (sample, head), 180 of them computing cp.async addresses (whose 16-byte requests cost 24 L2
// sectors per instruction instead of 16), every item is bracketed by two block barriers (16 % of the stall samples),
// the key-validity bytes are a synchronous global load on the critical path (17 % long-scoreboard), the output leaves as
// 4-byte stores (two L2 write sectors per useful one) and all ten key tiles are multiplied although the causal mask
// hides 40 % of them.  Here:
//
//   warp NG*QW   producer: per item one header (validity bit words, length, first row, head) and three TMA boxes
//                (Q | K | V: 64 dh x T rows, or 16*ceil(T/16) rows for packed text; 128B swizzle) into the next stage of
//                an NSTAGES-deep ring;
//                the metadata of the following items is loaded two and three iterations ahead
//   NG groups    of QW = TPAD/16 consumer warps; group g takes every NG-th item of the CTA.  Warp w of a group owns
//                query rows 16w .. 16w+15 exactly as in attention.cuh (same fragments, same order of operations: the
//                results are bit-identical to attention_kernel), but 16-key steps that the causal mask or the sequence
//                length hides from ALL 16 rows are skipped (their probabilities are exact zeros), the normalised O
//                tile is transposed through the warp's own (dead) Q rows and stored as 128-byte lines, and no block
//                barrier exists: a stage is handed over through one full / one empty mbarrier.
//
// One persistent CTA per SM.  Text tower <80, 3, 7>: 7 x 30 KB stages, four items in flight while three are computed,
// 16 warps -> 128 registers.  Measured (ncu, 1024 x 77 x 8 heads, profiles/r02_attention_ring_ncu.txt): 96.7 -> 60 us =
// 4.95 TB/s of DRAM traffic = 76 % of the measured HBM peak (80 us first version, 72 with interleaved chains, 66 with
// the rotating row blocks, 59-60 with three consumer groups instead of four).  Vision tower <64, 5, 8>: 79 us.
// Tried and dropped: a contiguous run of items per CTA instead of round-robin (same 80 us at that stage: DRAM page
// locality is not the limit), mbarrier.try_wait with a suspend-time hint (same duration, same instruction count).
// Rows T .. 16*ceil(T/16)-1 of a stage hold the next sample's rows (packed mode: the box is a multiple of 16 rows) or
// whatever an earlier item left there (fixed-length mode: the box is exactly T rows); their keys are masked by a
// select, their V rows are zeroed in shared memory before P V so that a non-finite neighbour can never leak into this
// sample (0 * NaN) -- the semantics stay those of attention.cuh / SDPA safe softmax.
#pragma once
#include "attention.cuh"
#include "gemm_tcgen05.cuh"

namespace mmcm {

constexpr int ATR_MAXBOX = 8;
struct AttRingMaps {
  CUtensorMap m[ATR_MAXBOX];   // m[k]: box of 64 dh x 16*(k+1) rows over qkv [rows, 3*D] (fixed-length mode: T rows)
};

template <int TPAD, int NG, int NSTAGES>
struct AttRingCfg {
  static_assert(TPAD % 16 == 0 && TPAD <= 16 * ATR_MAXBOX, "TPAD: multiple of 16, at most 128");
  static constexpr int QW = TPAD / 16;
  static constexpr int CONSUMERS = NG * QW;                 // consumer warps
  static constexpr int THREADS = (CONSUMERS + 1) * 32;
  static constexpr int TILE_BYTES = TPAD * 128;             // Q, K or V: TPAD rows x 64 bf16
  static constexpr int STAGE_BYTES = 3 * TILE_BYTES;
  static constexpr int HDR_WORDS = 8;                       // validity words [4], T, first row, head, -
  static constexpr int SMEM_BYTES = 1024 + NSTAGES * STAGE_BYTES + NSTAGES * (HDR_WORDS * 4 + 16);
};

__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// One warp, one item: query rows r0 .. r0+15 against the first `kmax` key steps (16 keys = two 8-key tiles each) of the
// staged K / V tiles.  Fragments and order of operations are those of attention_kernel (attention.cuh).  A skipped
// step is one warp-uniform branch around two interleaved accumulator chains.  (One straight-line instantiation per
// step count -- a switch over kmax -- was tried: 3.8 k SASS instructions, 80 -> 130 us: the five warps of a group sit
// in five different bodies and the instruction cache thrashes.)
template <int KM>
__device__ __forceinline__ void atr_item(const uint32_t Qs, const uint32_t Ks, const uint32_t Vs,
                                         const uint32_t (&vmask)[4], const int T, const int r0, const int causal,
                                         const int lane, __nv_bfloat16* __restrict__ orow, const int D, const int kmax) {
  constexpr int NTK = 2 * KM;
  const int g = lane >> 2, tq = lane & 3;
  const int m = lane >> 3, rr = lane & 7;

  // ---- S = Q K^T
  uint32_t qa[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int qrow = r0 + rr + ((m & 1) << 3);
    const int chunk = kk * 2 + (m >> 1);
    ldmatrix_x4(qa[kk], Qs + qrow * 128 + ((chunk ^ rr) << 4));
  }
  float s[NTK][4];
#pragma unroll
  for (int kk = 0; kk < KM; ++kk) {
    const int j0 = 2 * kk, j1 = 2 * kk + 1;
    s[j0][0] = s[j0][1] = s[j0][2] = s[j0][3] = 0.f;
    s[j1][0] = s[j1][1] = s[j1][2] = s[j1][3] = 0.f;
    if (kk < kmax) {
      uint32_t ka0[4], ka1[4], kb0[4], kb1[4];
      ldmatrix_x4(ka0, Ks + (j0 * 8 + rr) * 128 + ((m ^ rr) << 4));          // tile j0, dh 0..31
      ldmatrix_x4(ka1, Ks + (j0 * 8 + rr) * 128 + (((4 + m) ^ rr) << 4));    //          dh 32..63
      ldmatrix_x4(kb0, Ks + (j1 * 8 + rr) * 128 + ((m ^ rr) << 4));          // tile j1
      ldmatrix_x4(kb1, Ks + (j1 * 8 + rr) * 128 + (((4 + m) ^ rr) << 4));
      mma_bf16_16816(s[j0], qa[0], ka0[0], ka0[1]);
      mma_bf16_16816(s[j1], qa[0], kb0[0], kb0[1]);
      mma_bf16_16816(s[j0], qa[1], ka0[2], ka0[3]);
      mma_bf16_16816(s[j1], qa[1], kb0[2], kb0[3]);
      mma_bf16_16816(s[j0], qa[2], ka1[0], ka1[1]);
      mma_bf16_16816(s[j1], qa[2], kb1[0], kb1[1]);
      mma_bf16_16816(s[j0], qa[3], ka1[2], ka1[3]);
      mma_bf16_16816(s[j1], qa[3], kb1[2], kb1[3]);
    }
  }

  // ---- mask + fp32 softmax (as attention.cuh)
  const int qr0 = r0 + g, qr1 = r0 + g + 8;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NTK; ++j) {
    if (j >= 2 * kmax) continue;      // warp-uniform
    const uint32_t tile_bits = (vmask[(j * 8) / 32] >> ((j * 8) & 31)) & 0xffu;
    const bool full = tile_bits == 0xffu && (!causal || j * 8 + 7 <= r0);
    if (full) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    } else {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = j * 8 + 2 * tq + e;
        const bool kok = (tile_bits >> (2 * tq + e)) & 1u;
        const bool ok0 = kok && (!causal || key <= qr0);
        const bool ok1 = kok && (!causal || key <= qr1);
        s[j][e] = ok0 ? s[j][e] : -INFINITY;
        s[j][2 + e] = ok1 ? s[j][2 + e] : -INFINITY;
        mx0 = fmaxf(mx0, s[j][e]);
        mx1 = fmaxf(mx1, s[j][2 + e]);
      }
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  if (mx0 == -INFINITY) mx0 = 0.f;  // fully masked row: exp(-inf) = 0 everywhere, sum = 0 -> output 0
  if (mx1 == -INFINITY) mx1 = 0.f;
  const float L2E = 1.4426950408889634f;
  const float nm0 = -mx0 * L2E, nm1 = -mx1 * L2E;
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t p[NTK][2];           // probabilities as bf16 pairs (the A fragments of P V): half the registers of s[][]
#pragma unroll
  for (int kk = 0; kk < KM; ++kk) {
    p[2 * kk][0] = p[2 * kk][1] = p[2 * kk + 1][0] = p[2 * kk + 1][1] = 0u;
    if (kk < kmax) {
#pragma unroll
      for (int j = 2 * kk; j < 2 * kk + 2; ++j) {
        const float e0 = ex2_fast(fmaf(s[j][0], L2E, nm0));
        const float e1 = ex2_fast(fmaf(s[j][1], L2E, nm0));
        const float e2 = ex2_fast(fmaf(s[j][2], L2E, nm1));
        const float e3 = ex2_fast(fmaf(s[j][3], L2E, nm1));
        sum0 += e0 + e1;
        sum1 += e2 + e3;
        p[j][0] = pack_bf16x2(e0, e1);
        p[j][1] = pack_bf16x2(e2, e3);
      }
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f;
  const float inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;

  // ---- O = P V
  float o[8][4];
#pragma unroll
  for (int jd = 0; jd < 8; ++jd) o[jd][0] = o[jd][1] = o[jd][2] = o[jd][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KM; ++kk) {
    if (kk >= kmax) continue;         // warp-uniform
    uint32_t pa[4];
    pa[0] = p[2 * kk][0];
    pa[1] = p[2 * kk][1];
    pa[2] = p[2 * kk + 1][0];
    pa[3] = p[2 * kk + 1][1];
    const int vrow = kk * 16 + rr + ((m & 1) << 3);
#pragma unroll
    for (int jd = 0; jd < 8; jd += 2) {
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, Vs + vrow * 128 + (((jd + (m >> 1)) ^ rr) << 4));
      mma_bf16_16816(o[jd], pa, vb[0], vb[1]);
      mma_bf16_16816(o[jd + 1], pa, vb[2], vb[3]);
    }
  }

  // ---- normalise, transpose through this warp's own (dead) Q rows, store 128-byte lines
#pragma unroll
  for (int jd = 0; jd < 8; ++jd) {
    const uint32_t a = Qs + (r0 + g) * 128 + ((jd ^ g) << 4) + tq * 4;
    sts32(a, pack_bf16x2(o[jd][0] * inv0, o[jd][1] * inv0));
    sts32(a + 8 * 128, pack_bf16x2(o[jd][2] * inv1, o[jd][3] * inv1));
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = r0 + m + 4 * i;
    const uint4 v = lds128(Qs + row * 128 + ((rr ^ (row & 7)) << 4));
    if (row < T) stg128(orow + (size_t)row * D + rr * 8, v);
  }
}

template <int TPAD, int NG, int NSTAGES>
__global__ void __launch_bounds__(AttRingCfg<TPAD, NG, NSTAGES>::THREADS, 1)
attention_ring_kernel(const __grid_constant__ AttRingMaps maps, __nv_bfloat16* __restrict__ out,
                      const uint8_t* __restrict__ key_valid, const int* __restrict__ seq_start,
                      const int* __restrict__ seq_len, const int T_fixed, const int D, const int causal,
                      const int kv_stride, const int B, const int heads) {
  using C = AttRingCfg<TPAD, NG, NSTAGES>;
  constexpr int QW = C::QW;
  constexpr int KT = TPAD / 16;
  constexpr int VW = (TPAD + 31) / 32;
  extern __shared__ __align__(16) uint8_t atr_smem_raw[];

  const uint32_t base = (smem_u32(atr_smem_raw) + 1023u) & ~1023u;        // 128B swizzle atom = 1024 bytes
  uint8_t* gen = atr_smem_raw + (base - smem_u32(atr_smem_raw));
  int* hdr_all = reinterpret_cast<int*>(gen + NSTAGES * C::STAGE_BYTES);
  const uint32_t bars = base + NSTAGES * C::STAGE_BYTES + NSTAGES * C::HDR_WORDS * 4;   // full[s] at +16 s, empty[s] at +16 s + 8

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int total = B * heads;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(bars + 16 * s, 1);
      mbar_init(bars + 16 * s + 8, QW);
    }
    fence_barrier_init();
  }
  if (warp == C::CONSUMERS && lane < QW) prefetch_tmap(&maps.m[lane]);
  pdl_trigger();
  pdl_wait();
  __syncthreads();

  if (warp == C::CONSUMERS) {
    // ===================================== producer =====================================
    struct Meta { int T, row0, kvoff; };
    auto meta = [&](int it) {
      Meta mt{0, 0, 0};
      if (it < total) {
        const int b = it / heads;
        mt.T = seq_len ? seq_len[b] : T_fixed;
        mt.row0 = seq_start ? seq_start[b] : b * T_fixed;
        mt.kvoff = seq_start ? mt.row0 : b * kv_stride;
      }
      return mt;
    };
    auto bytes = [&](const Meta& mt, uint32_t (&ok)[VW]) {
#pragma unroll
      for (int w = 0; w < VW; ++w) {
        const int key = w * 32 + lane;
        uint32_t v = (key < mt.T && key < TPAD) ? 1u : 0u;
        if (v && key_valid) v = key_valid[(size_t)mt.kvoff + key];
        ok[w] = v;
      }
    };
    const int stride = gridDim.x;
    int item = blockIdx.x;
    // Software pipeline of the global loads: lengths / first rows are read three items ahead of the TMA issue, the
    // key-validity bytes (whose address depends on them in packed mode) two items ahead; the ballot that consumes the
    // bytes runs on values loaded two iterations earlier, so no iteration waits for a load it issued itself.
    Meta m0 = meta(item), m1 = meta(item + stride), m2 = meta(item + 2 * stride);
    uint32_t b0[VW], b1[VW];
    bytes(m0, b0);
    bytes(m1, b1);
    for (int n = 0; item < total; item += stride, ++n) {
      const Meta m3 = meta(item + 3 * stride);
      uint32_t b2[VW];
      bytes(m2, b2);
      uint32_t vm[VW];
#pragma unroll
      for (int w = 0; w < VW; ++w) vm[w] = __ballot_sync(0xffffffffu, b0[w] != 0u);
      const int stage = n % NSTAGES, use = n / NSTAGES;
      if (use > 0) mbar_wait(bars + 16 * stage + 8, (uint32_t)(use - 1) & 1u);
      if (lane == 0) {
        int* hdr = hdr_all + stage * C::HDR_WORDS;
#pragma unroll
        for (int w = 0; w < VW; ++w) hdr[w] = (int)vm[w];
        hdr[4] = m0.T;
        hdr[5] = m0.row0;
        hdr[6] = item % heads;
        int nb = (min(m0.T, TPAD) + 15) >> 4;
        if (nb < 1) nb = 1;
        const uint32_t full = bars + 16 * stage;
        const uint32_t dst = base + stage * C::STAGE_BYTES;
        const int c = (item % heads) * ATT_DH;
        // fixed-length mode: the box is exactly T rows (nothing of the next sample is fetched); packed: 16 nb rows
        const int box_rows = seq_start ? nb * 16 : T_fixed;
        mbar_expect_tx(full, (uint32_t)(3 * box_rows * 128));
        tma_load_2d(&maps.m[nb - 1], full, dst, c, m0.row0);
        tma_load_2d(&maps.m[nb - 1], full, dst + C::TILE_BYTES, D + c, m0.row0);
        tma_load_2d(&maps.m[nb - 1], full, dst + 2 * C::TILE_BYTES, 2 * D + c, m0.row0);
      }
      __syncwarp();
      m0 = m1; m1 = m2; m2 = m3;
#pragma unroll
      for (int w = 0; w < VW; ++w) { b0[w] = b1[w]; b1[w] = b2[w]; }
    }
    return;
  }

  // ======================================= consumers =======================================
  const int grp = warp / QW, wlocal = warp - grp * QW;
  int n = grp;
  int rot = 0;                   // rounds of this group, modulo QW
  for (int item = blockIdx.x + grp * gridDim.x; item < total; item += NG * gridDim.x, n += NG) {
    // Causal: row block w costs w + 1 key steps, so the block a warp owns rotates from item to item -- every warp
    // sees the same work over QW items, and since the warps of a group only meet at the stage's empty barrier the
    // warp that drew a light block already works on the group's next item while the heavy block finishes.
    const int wq = causal ? (wlocal + rot >= QW ? wlocal + rot - QW : wlocal + rot) : wlocal;
    rot = rot + 1 == QW ? 0 : rot + 1;
    const int r0 = wq * 16;
    const int stage = n % NSTAGES;
    mbar_wait(bars + 16 * stage, (uint32_t)(n / NSTAGES) & 1u);
    const int* hdr = hdr_all + stage * C::HDR_WORDS;
    const int T = min(hdr[4], TPAD), row0 = hdr[5], h = hdr[6];
    const uint32_t Qs = base + stage * C::STAGE_BYTES;
    const uint32_t Ks = Qs + C::TILE_BYTES;
    const uint32_t Vs = Ks + C::TILE_BYTES;
    const int nb = (T + 15) >> 4;
    uint32_t vmask[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) vmask[w] = w < VW ? (uint32_t)hdr[w] : 0u;

    // V rows of the box behind the sequence: zero them before anybody multiplies them by a zero probability
    if ((T & 15) != 0) {
      const bool mine = causal ? (wq == nb - 1) : (wq == 0);
      if (mine) {
        const int rows = nb * 16 - T;                 // 1..15 rows of 8 chunks
        for (int i = lane; i < rows * 8; i += 32) {
          const int row = T + (i >> 3);
          sts128(Vs + row * 128 + (((i & 7) ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
        }
      }
      if (causal) __syncwarp();
      else named_bar_sync(1 + grp, QW * 32);
    }

    if (r0 < T) {
      // key steps (16 keys) any of this warp's rows can see
      int kmax = nb;
      if (causal) kmax = min(kmax, wq + 1);
      __nv_bfloat16* orow = out + (size_t)row0 * D + h * ATT_DH;
      atr_item<KT>(Qs, Ks, Vs, vmask, T, r0, causal, lane, orow, D, kmax);
    }
    fence_proxy_async();          // generic-proxy writes (O tile, V padding) before the next TMA box lands here
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 16 * stage + 8);
  }
}

}  // namespace mmcm
