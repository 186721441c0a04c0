// K3: fused multi-head attention for the short CLIP / SigLIP sequences (50 / 77 / 64 / 196 tokens).
//
// Semantics = eager_attention_forward (HF/models/clip/modeling_clip.py:261-279) as dispatched through
// sdpa_attention_forward (HF/integrations/sdpa_attention.py:40-104): softmax_fp32(q k^T * dh^-1/2 + mask) v,
// mask = causal (CLIP text, HF clip :546-551) and/or key padding (HF/masking_utils.py:882,1001).
// A query whose keys are all masked produces exactly 0 (SDPA safe-softmax).
//
// One CTA per (sample, head).  The whole K/V of one head fits in shared memory, so there is no online
// softmax: S = Q K^T is held in registers (mma.sync m16n8k16 bf16, fp32 accumulate), the row softmax is done
// in fp32 with quad shuffles, P is re-used in registers as the A operand of P V.  The attention core is
// 1.6 % (CLIP) / 3.4 % (SigLIP) of the model FLOPs, so the legacy warp-level MMA is sufficient here; the
// kernel is bound by the QKV read + O write (HBM/L2), which is 128-bit vectorised.
#pragma once
#include "common.cuh"

namespace mmcm {

constexpr int ATT_DH = 64;       // head dim of every supported tower
constexpr int ATT_LD = 72;       // smem row pitch in bf16 (144 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// qkv  : bf16 [rows, 3*D]  (Q | K | V, Q already multiplied by dh^-1/2 -- folded into the weights)
// out  : bf16 [rows, D]
// seq_start/seq_len : per-sample first row and length (packed variable-length text) or nullptr => b*T, T
// key_valid : uint8 [B, Tstride] (1 = key may be attended; packed chunks: one byte per packed row) or nullptr
//
// Persistent and software-pipelined: a CTA (QW warps) walks work items (sample b, head h, query block z) with stride
// gridDim.x; while it computes item i out of one shared-memory buffer, cp.async (LDGSTS, 16 B, zero-fill for padding
// rows) is already filling the other buffer with item i+1.  The first version staged, synchronised, computed and
// stored one item per CTA and reached 47 % of the HBM roofline (profiles/r01_attention_layernorm_ncu_full.txt):
// latency bound, nothing in flight during the math.  Warp w of query block z owns query rows 16*(z*QW+w) .. +15.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 2^x as ONE MUFU: exp2f() wraps ex2.approx in denormal range fix-ups (2 FMUL + compare + select per call), which
// the softmax does not need -- arguments are <= 0 and a flushed 2^-127 contributes nothing to a sum that contains 1
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int TPAD, int QW>
struct AttCfg {
  static constexpr int QROWS = QW * 16;
  static constexpr int BUF_BYTES = ((2 * TPAD + QROWS) * ATT_LD * 2 + TPAD + 15) / 16 * 16;
  // Occupancy beats software pipelining here.  The 77-token text kernel <80, 5> double-buffered its 34 KB item
  // (69 KB, 3 CTAs = 15 warps per SM, 82 registers) and sat at 47 % of the HBM roofline with 41 % of the issue slots
  // busy: latency bound, too few warps (profiles/r01_attention_layernorm_final_ncu_full.txt).  With ONE buffer and the
  // register allocator held to 5 CTAs per SM (25 warps, 72 registers) the other CTAs' math covers a CTA's load phase:
  // bench.py 56.5 k -> 58.0 k samples/s (profiles/r02_attention_occupancy.txt).  Items above the limit below are single
  // buffered (SigLIP vision, 196 tokens: 76 KB per item, two CTAs per SM); smaller ones keep the double buffer.
#ifndef MMCM_ATT_DB_LIMIT
#define MMCM_ATT_DB_LIMIT (30 * 1024)
#endif
  static constexpr int NBUF = (BUF_BYTES > MMCM_ATT_DB_LIMIT) ? 1 : 2;
  static constexpr int SMEM_BYTES = NBUF * BUF_BYTES;
  static constexpr int QBLOCKS = (TPAD / 16 + QW - 1) / QW;
  // min resident CTAs per SM the register allocator must allow (the text kernel: 5 x 160 threads x 80 registers)
  static constexpr int MIN_CTAS = (TPAD == 80 && QW == 5) ? 5 : 1;
};

template <int TPAD, int QW>
__global__ void __launch_bounds__(QW * 32, AttCfg<TPAD, QW>::MIN_CTAS)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                 const uint8_t* __restrict__ key_valid, const int* __restrict__ seq_start,
                 const int* __restrict__ seq_len, const int T_fixed, const int D, const int causal,
                 const int kv_stride, const int B, const int heads) {
  static_assert(TPAD % 16 == 0, "TPAD must be a multiple of 16");
  using C = AttCfg<TPAD, QW>;
  constexpr int NT = TPAD / 8;   // key tiles of 8
  constexpr int KT = TPAD / 16;  // key steps of 16 for P V
  constexpr int NTHR = QW * 32;
  extern __shared__ __align__(16) uint8_t att_smem[];

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int ld_qkv = 3 * D;
  const int total = B * heads * C::QBLOCKS;

  // item -> (b, h, z): h fastest so that neighbouring CTAs read neighbouring 128-byte column blocks of the same rows
  auto decode = [&](int item, int& b, int& h, int& z) {
    h = item % heads;
    const int r = item / heads;
    z = r % C::QBLOCKS;
    b = r / C::QBLOCKS;
  };
  // issue the asynchronous copies of one item into buffer `buf` (every thread, no barrier)
  auto prefetch = [&](int item, int buf) {
    int b, h, z;
    decode(item, b, h, z);
    const int T = seq_len ? seq_len[b] : T_fixed;
    const int row0 = seq_start ? seq_start[b] : b * T_fixed;
    const int q0 = z * C::QROWS;
    uint8_t* base = att_smem + buf * C::BUF_BYTES;
    const uint32_t Qs = smem_u32(base);
    const uint32_t Ks = Qs + C::QROWS * ATT_LD * 2;
    const uint32_t Vs = Ks + TPAD * ATT_LD * 2;
    uint8_t* kvs = base + (2 * TPAD + C::QROWS) * ATT_LD * 2;
    if (q0 < T) {   // (a query block that is pure padding loads nothing and is skipped by compute)
      for (int idx = tid; idx < TPAD * 16; idx += NTHR) {
        const int t = idx >> 4, c = idx & 15;
        const bool in = t < T;
        const __nv_bfloat16* src = qkv + (size_t)(row0 + (in ? t : 0)) * ld_qkv + (1 + (c >> 3)) * D + h * ATT_DH + (c & 7) * 8;
        cp_async16(((c >> 3) == 0 ? Ks : Vs) + (t * ATT_LD + (c & 7) * 8) * 2, src, in ? 16u : 0u);
      }
      for (int idx = tid; idx < C::QROWS * 8; idx += NTHR) {
        const int tl = idx >> 3, t = q0 + tl;
        const bool in = t < T;
        const __nv_bfloat16* src = qkv + (size_t)(row0 + (in ? t : 0)) * ld_qkv + h * ATT_DH + (idx & 7) * 8;
        cp_async16(Qs + (tl * ATT_LD + (idx & 7) * 8) * 2, src, in ? 16u : 0u);
      }
      for (int t = tid; t < TPAD; t += NTHR) {
        uint8_t ok = (t < T) ? 1 : 0;
        if (ok && key_valid) ok = key_valid[seq_start ? (size_t)(row0 + t) : (size_t)b * kv_stride + t] ? 1 : 0;
        kvs[t] = ok;
      }
    }
  };

  int item = blockIdx.x;
  int buf = 0;
  if (C::NBUF == 2) {
    if (item < total) prefetch(item, 0);
    cp_async_commit();
  }
  for (; item < total; item += gridDim.x) {
    if (C::NBUF == 2) {
      const int next = item + gridDim.x;
      if (next < total) prefetch(next, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();        // everything but the newest group (= the next item) has landed
    } else {
      prefetch(item, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();

    int b, h, z;
    decode(item, b, h, z);
    const int T = seq_len ? seq_len[b] : T_fixed;
    const int row0 = seq_start ? seq_start[b] : b * T_fixed;
    const int q0 = z * C::QROWS;
    const int r0 = q0 + warp * 16;
    if (r0 < T) {                // warp-uniform: warps whose 16 rows are padding skip the math, not the barriers
      uint8_t* base = att_smem + buf * C::BUF_BYTES;
      const __nv_bfloat16* Qs = reinterpret_cast<const __nv_bfloat16*>(base);
      const __nv_bfloat16* Ks = Qs + C::QROWS * ATT_LD;
      const __nv_bfloat16* Vs = Ks + TPAD * ATT_LD;
      const uint8_t* kvs = base + (2 * TPAD + C::QROWS) * ATT_LD * 2;

      // ---- S = Q K^T
      uint32_t qa[4][4];
      {
        const int m = lane >> 3, rr = lane & 7;
        const int qrow = warp * 16 + rr + ((m & 1) ? 8 : 0);  // local row inside Qs
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int col = kk * 16 + ((m & 2) ? 8 : 0);
          ldmatrix_x4(qa[kk], smem_u32(Qs + qrow * ATT_LD + col));
        }
      }
      float s[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const int m = lane >> 3, rr = lane & 7;
        const int krow = j * 8 + rr;
        uint32_t kb0[4], kb1[4];
        ldmatrix_x4(kb0, smem_u32(Ks + krow * ATT_LD + m * 8));        // dh 0..31 : (b0,b1) of k-steps 0,1
        ldmatrix_x4(kb1, smem_u32(Ks + krow * ATT_LD + 32 + m * 8));   // dh 32..63: (b0,b1) of k-steps 2,3
        mma_bf16_16816(s[j], qa[0], kb0[0], kb0[1]);
        mma_bf16_16816(s[j], qa[1], kb0[2], kb0[3]);
        mma_bf16_16816(s[j], qa[2], kb1[0], kb1[1]);
        mma_bf16_16816(s[j], qa[3], kb1[2], kb1[3]);
      }

      // ---- mask + fp32 softmax (rows g and g+8 of this warp's 16-row slab live in one quad)
      // key validity as bit masks (one 32-bit word per 32 keys, built once per item from the staged bytes) instead
      // of a shared-memory byte load + compare per score: the kernel is instruction-issue bound (profiles/), so the
      // per-element work is kept to: select, max | FFMA, EX2, add.
      const int qr0 = r0 + g, qr1 = r0 + g + 8;
      uint32_t vmask[(TPAD + 31) / 32];
#pragma unroll
      for (int w = 0; w < (TPAD + 31) / 32; ++w) {
        const int key = w * 32 + lane;
        vmask[w] = __ballot_sync(0xffffffffu, key < TPAD && kvs[key] != 0);
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        // warp-uniform fast path: all 8 keys of the tile are valid and (if causal) not after this warp's first row
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
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = ex2_fast(fmaf(s[j][0], L2E, nm0));
        s[j][1] = ex2_fast(fmaf(s[j][1], L2E, nm0));
        s[j][2] = ex2_fast(fmaf(s[j][2], L2E, nm1));
        s[j][3] = ex2_fast(fmaf(s[j][3], L2E, nm1));
        sum0 += s[j][0] + s[j][1];
        sum1 += s[j][2] + s[j][3];
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f;
      const float inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;

      // ---- O = P V   (P: accumulator layout of two adjacent key tiles == A fragment of one 16-key step)
      float o[8][4];
#pragma unroll
      for (int jd = 0; jd < 8; ++jd) o[jd][0] = o[jd][1] = o[jd][2] = o[jd][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int m = lane >> 3, rr = lane & 7;
        const int vrow = kk * 16 + rr + ((m & 1) ? 8 : 0);
#pragma unroll
        for (int jd = 0; jd < 8; jd += 2) {
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, smem_u32(Vs + vrow * ATT_LD + (jd + (m >> 1)) * 8));
          mma_bf16_16816(o[jd], pa, vb[0], vb[1]);
          mma_bf16_16816(o[jd + 1], pa, vb[2], vb[3]);
        }
      }

      // ---- normalise and store (bf16x2 per thread per 8-wide dh tile)
#pragma unroll
      for (int jd = 0; jd < 8; ++jd) {
        const int col = h * ATT_DH + jd * 8 + 2 * tq;
        if (qr0 < T)
          *reinterpret_cast<uint32_t*>(out + (size_t)(row0 + qr0) * D + col) = pack_bf16x2(o[jd][0] * inv0, o[jd][1] * inv0);
        if (qr1 < T)
          *reinterpret_cast<uint32_t*>(out + (size_t)(row0 + qr1) * D + col) = pack_bf16x2(o[jd][2] * inv1, o[jd][3] * inv1);
      }
    }
    __syncthreads();             // everyone is done with `buf` before the prefetch of the iteration after next refills it
    if (C::NBUF == 2) buf ^= 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// K9: SigLIP multi-head attention pooling (SiglipMultiheadAttentionPoolingHead,
// HF/models/siglip/modeling_siglip.py:628-649): ONE learned query (the probe, projected once at weight-load
// time and pre-scaled by dh^-1/2) attends over the T tokens of each sample.
//   kv  : bf16 [B*T, 2*D]  (K | V from the packed in_proj rows D..3D)
//   q   : fp32 [D]
//   out : bf16 [B, D]
// grid = (heads, B), block = 128: warps stride over keys for the scores, fp32 softmax in shared memory, then each
// warp accumulates a quarter of the keys for P V (lane owns 2 of the 64 head dims) and the quarters are summed.
// ------------------------------------------------------------------------------------------------
constexpr int MAP_MAXT = 1024;
__global__ void __launch_bounds__(128)
map_attention_kernel(const __nv_bfloat16* __restrict__ kv, const float* __restrict__ q,
                     __nv_bfloat16* __restrict__ out, const int T, const int D) {
  __shared__ float sc[MAP_MAXT];
  __shared__ float red[4];
  __shared__ float acc[4][ATT_DH];
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = kv + (size_t)b * T * 2 * D + h * ATT_DH;
  const float q0 = q[h * ATT_DH + 2 * lane], q1 = q[h * ATT_DH + 2 * lane + 1];
  float mx = -INFINITY;
  for (int t = warp; t < T; t += 4) {
    const __nv_bfloat162 k2 = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)t * 2 * D + 2 * lane);
    float s = warp_sum(q0 * __bfloat162float(k2.x) + q1 * __bfloat162float(k2.y));
    if (lane == 0) sc[t] = s;
    mx = fmaxf(mx, s);
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += 128) {
    const float e = __expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
  float o0 = 0.f, o1 = 0.f;
  for (int t = warp; t < T; t += 4) {
    const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)t * 2 * D + D + 2 * lane);
    const float p = sc[t];
    o0 = fmaf(p, __bfloat162float(v2.x), o0);
    o1 = fmaf(p, __bfloat162float(v2.y), o1);
  }
  acc[warp][2 * lane] = o0;
  acc[warp][2 * lane + 1] = o1;
  __syncthreads();
  if (threadIdx.x < ATT_DH) {
    const int d = threadIdx.x;
    const float v = (acc[0][d] + acc[1][d] + acc[2][d] + acc[3][d]) * inv;
    out[(size_t)b * D + h * ATT_DH + d] = __float2bfloat16_rn(v);
  }
}

// q[n] = scale * (W[n,:] . probe + bias[n])  -- the probe's query projection, done once when weights are finalized
__global__ void probe_query_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                   const float* __restrict__ probe, float* __restrict__ q, const int D,
                                   const float scale) {
  pdl_wait();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= D) return;
  float a = 0.f;
  for (int k = lane; k < D; k += 32) a = fmaf(W[(size_t)n * D + k], probe[k], a);
  a = warp_sum(a);
  if (lane == 0) q[n] = (a + bias[n]) * scale;
}

}  // namespace mmcm
