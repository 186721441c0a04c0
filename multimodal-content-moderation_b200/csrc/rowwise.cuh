// K4 / K5 / K6 and the small HBM-bound helpers: LayerNorm, embeddings, im2col, pooling-row selection.
// All of them are one-pass, 128-bit vectorised, one warp per row with shuffle reductions.
#pragma once
#include "common.cuh"

namespace mmcm {

// ------------------------------------------------------------------------------------------------
// K4 LayerNorm over D (nn.LayerNorm semantics: biased variance, eps inside the sqrt).
// Replaces layer_norm1/2, pre_layrnorm, post_layernorm, final_layer_norm
// (HF/models/clip/modeling_clip.py:359-384,522,562,659-661,677,686).
//   x        fp32 [*, D] residual stream
//   gather   optional row indices (pooling: only the EOS / CLS rows are normalised) else row i
//   out_bf16 optional bf16 [rows, D]  (operand of the next GEMM)
//   out_f32  optional fp32 [rows, D]  (may alias x when gather == nullptr: pre_layrnorm in place)
//   part     optional split-K partial sums of the residual GEMM in front of this LayerNorm (gemm2_tcgen05.cuh,
//            EpiParams::ksplit): x_new = ((x + part[0]) + part[1]) + ... in this fixed order, written back to x_rw
//            before it is normalised -- the reduction of a split-K GEMM without a reduction kernel or inter-CTA waits
// One warp per row; D/128 float4 per lane held in registers (two-pass mean/variance, no re-read).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float eps, const int rows_host, const int* __restrict__ gather,
                 __nv_bfloat16* __restrict__ out_bf16, float* out_f32, const int* __restrict__ rows_dev,
                 const float* __restrict__ part, const int ksplit, const int part_rows, float* x_rw) {
  constexpr int V = D / 128;  // float4 per lane
  pdl_trigger();
  pdl_wait();
  const int rows = rows_dev ? min(__ldg(rows_dev), rows_host) : rows_host;   // packed text: live rows known on device
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int src = gather ? gather[warp] : warp;
  // (with partial sums the row is read through the pointer it is rewritten through, not the __restrict__ one)
  const float4* xr = part ? reinterpret_cast<const float4*>(x_rw + (size_t)src * D)
                          : reinterpret_cast<const float4*>(x + (size_t)src * D);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xr[lane + 32 * i];
  if (part) {
    for (int k = 0; k < ksplit; ++k) {
      const float4* pr = reinterpret_cast<const float4*>(part + ((size_t)k * part_rows + src) * D);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 p = pr[lane + 32 * i];
        v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w;
      }
    }
    float4* xw = reinterpret_cast<float4*>(x_rw + (size_t)src * D);
#pragma unroll
    for (int i = 0; i < V; ++i) xw[lane + 32 * i] = v[i];
  }
#pragma unroll
  for (int i = 0; i < V; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c4 = lane + 32 * i;
    // gamma == nullptr: plain normalisation -- the affine part is folded into the consuming GEMM's weights (fold_ln_kernel)
    const float4 g = gamma ? __ldg(g4 + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bb = gamma ? __ldg(b4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + bb.x;
    y.y = (v[i].y - mean) * rstd * g.y + bb.y;
    y.z = (v[i].z - mean) * rstd * g.z + bb.z;
    y.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)warp * D)[c4] = y;
    if (out_bf16) {
      uint2 p;
      p.x = pack_bf16x2(y.x, y.y);
      p.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + (size_t)warp * D)[c4] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LN fold, entry of a tower (gemm2_tcgen05.cuh "LN fold"): the first layer's LayerNorm has no residual GEMM in front
// of it, so this kernel leaves what EPI_RESID_STATS leaves -- a bf16 copy of the rows and, per 128-column slab,
// (sum, M2 about the slab mean).  Lane l of the warp holds columns 4l..4l+3 of every slab, so slab i is v[i] across
// the warp.  gamma != nullptr: the rows are first LayerNorm-ed in place (CLIP's pre_layrnorm, HF clip :677).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
prep_rows_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                 const int rows_host, const int* __restrict__ rows_dev, __nv_bfloat16* __restrict__ xb,
                 float2* __restrict__ stats, const int stats_pitch) {
  constexpr int V = D / 128;
  pdl_trigger();
  pdl_wait();
  const int rows = rows_dev ? min(__ldg(rows_dev), rows_host) : rows_host;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + (size_t)warp * D);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xr[lane + 32 * i];
  if (gamma) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c4 = lane + 32 * i;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      v[i].x = (v[i].x - mean) * rstd * g.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * g.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * g.w + bb.w;
      xr[c4] = v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float s = warp_sum(v[i].x + v[i].y + v[i].z + v[i].w);
    const float m = s * (1.0f / 128.0f);
    const float a = v[i].x - m, b = v[i].y - m, c = v[i].z - m, d = v[i].w - m;
    const float q = warp_sum(a * a + b * b + c * c + d * d);
    if (lane == 0) stats[(size_t)i * stats_pitch + warp] = make_float2(s, q);
    uint2 p;
    p.x = pack_bf16x2(v[i].x, v[i].y);
    p.y = pack_bf16x2(v[i].z, v[i].w);
    reinterpret_cast<uint2*>(xb + (size_t)warp * D)[lane + 32 * i] = p;
  }
}

// Weight side of the LN fold, run once per weight load (mmcm_finalize_weights): for output row n of a Linear that
// consumes LayerNorm(x)
//   g[n,k]      = scale_n * W[n,k] * gamma[k]              scale_n = dh^-1/2 for the q rows of the fused QKV matrix
//   Wout[n,k]   = bf16(g[n,k] - mean_k g[n,:])             CENTRED rows: x . Wout[n] = (x - mean(x)) . g[n] for any x,
//                                                          so the consumer never has to subtract mean(x) * colsum
//   bias_out[n] = scale_n * (b[n] + sum_k W[n,k] * beta[k])
//   resid[n]    = sum_k float(Wout[n,k])                   what the bf16 rounding leaves of the zero row sum (optional;
//                                                          the tests bound it: it multiplies mean(x) * rstd)
// One warp per n; the row (K <= 1024) stays in registers between the two passes.
__global__ void __launch_bounds__(256)
fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
               const float* __restrict__ beta, const int N, const int K, const int q_rows, const float q_scale,
               __nv_bfloat16* __restrict__ Wout, float* __restrict__ resid, float* __restrict__ bias_out) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float sc = n < q_rows ? q_scale : 1.0f;
  float g[32];                      // K <= 1024
  float gs = 0.f, bs = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * i;
    g[i] = 0.f;
    if (k < K) {
      const float w = W[(size_t)n * K + k];
      g[i] = sc * w * gamma[k];
      gs += g[i];
      bs = fmaf(w, beta[k], bs);
    }
  }
  const float mean = warp_sum(gs) / (float)K;
  bs = warp_sum(bs);
  float rs = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      const __nv_bfloat16 wg = __float2bfloat16_rn(g[i] - mean);
      Wout[(size_t)n * K + k] = wg;
      rs += __bfloat162float(wg);
    }
  }
  rs = warp_sum(rs);
  if (lane == 0) {
    if (resid) resid[n] = rs;
    bias_out[n] = sc * (b[n] + bs);
  }
}

// ------------------------------------------------------------------------------------------------
// K5 text embeddings + pooling-row selection.
//   x[b*S+s, :] = token_embedding[ids[b,s]] + position_embedding[s]      (HF clip :234-258)
//   pool_row[b] = b*S + (first s with ids==eos_id, else 0)               (HF clip :575-584)
//                 b*S + argmax(ids)   when eos_id == 2 (legacy branch)    (HF clip :564-574)
//                 b*S + S-1           when eos_id < 0 (SigLIP last token)  (HF siglip :520)
//   key_valid[b,s] = attention_mask[b,s] != 0 (all ones when mask == nullptr)
// One warp per token row; lane 0 of the first warp of each sample also does the pooling scan.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
text_embed_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                  const float* __restrict__ tok, const float* __restrict__ pos, const int B, const int S,
                  const int vocab, const int eos_id, float* __restrict__ x, int* __restrict__ pool_row,
                  uint8_t* __restrict__ key_valid) {
  constexpr int V = D / 128;
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * S) return;
  const int b = warp / S, s = warp - b * S;
  long long id = ids[warp];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;  // torch would raise an index error; clamp instead of reading out of bounds
  const float4* tr = reinterpret_cast<const float4*>(tok + (size_t)id * D);
  const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)s * D);
  float4* xo = reinterpret_cast<float4*>(x + (size_t)warp * D);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 a = __ldg(tr + lane + 32 * i), p = __ldg(pr + lane + 32 * i);
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    xo[lane + 32 * i] = a;
  }
  if (lane == 0) key_valid[warp] = mask ? (mask[warp] != 0 ? 1 : 0) : 1;
  if (s == 0) {
    // pooling scan: S <= 77, one warp strided over the row
    int best = 0;
    if (eos_id < 0) {
      best = S - 1;
    } else if (eos_id == 2) {
      long long bv = -0x7fffffffffffffffLL - 1;
      int bi = 0;
      for (int t = lane; t < S; t += 32) {
        long long v = (long long)(int)ids[(size_t)b * S + t];  // HF casts to int32 before argmax
        if (v > bv) { bv = v; bi = t; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        long long ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      best = bi;
    } else {
      int first = S;
      for (int t = lane; t < S; t += 32)
        if ((int)ids[(size_t)b * S + t] == eos_id) { first = t; break; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      best = (first == S) ? 0 : first;
    }
    if (lane == 0) pool_row[b] = b * S + best;
  }
}

// ------------------------------------------------------------------------------------------------
// K5' packed (variable-length) CLIP text chunks.  The CLIP text tower is causal and only the EOS row is pooled
// (HF clip :546-551, :575-584), so rows after the pooled position never influence the consumed row: measured on the
// reference, truncating at EOS changes the pooled output by 2.4e-6 = fp32 rounding.  text_plan_kernel
// computes, per sample, the pooled index p_b (same rules as text_embed_kernel), keeps L_b = p_b + 1 rows and
// prefix-sums them; text_embed_packed_kernel writes only those rows, back to back.  One block, B <= 4096.
//   seq_start[b], seq_len[b]        packed coordinates of sample b
//   pool_row[b] = seq_start[b] + p_b
//   rows_total[0] = sum_b L_b        (device-side M of every GEMM / LayerNorm of the chunk)
// Absent text (skip_mode != 0): a sample whose text feature cannot reach the logits keeps ONE row (its BOS token) --
//   fusion head (skip_mode 1): text_present < 0.5 zeroes the normalised text feature (R/src/models/fusion.py:188-189)
//   MTL head    (skip_mode 2): text_present < 0.5 AND image_present >= 0.5 selects the image branch and the gate is
//                              unused (R/src/models/multitask.py:194-197); with both absent the text branch wins, so
//                              those samples keep all their rows (measured on the reference)
// The pooled row of such a sample is a finite value nobody reads through: the logits are bit-identical.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
text_plan_kernel(const int64_t* __restrict__ ids, const int B, const int S, const int eos_id,
                 int* __restrict__ seq_start, int* __restrict__ seq_len, int* __restrict__ pool_row,
                 int* __restrict__ rows_total, const float* __restrict__ text_present,
                 const float* __restrict__ image_present, const int skip_mode) {
  __shared__ int warp_tot[32], warp_excl[32];
  __shared__ int carry, slab_total;
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int b = base + tid;
    int len = 0;
    if (b < B) {
      const int64_t* row = ids + (size_t)b * S;
      int best = 0;
      if (eos_id == 2) {            // legacy branch: argmax(ids), first maximum (HF clip :564-574)
        int bv = (int)row[0];
        for (int t = 1; t < S; ++t) { const int v = (int)row[t]; if (v > bv) { bv = v; best = t; } }
      } else {                      // first position equal to eos, 0 if none (HF clip :575-584)
        for (int t = 0; t < S; ++t) if ((int)row[t] == eos_id) { best = t; break; }
      }
      len = best + 1;
      if (skip_mode && text_present[b] < 0.5f && (skip_mode == 1 || image_present[b] >= 0.5f)) len = 1;
    }
    int incl = len;                 // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {                // scan of the 32 warp totals
      const int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
      warp_excl[lane] = wi - w;
      if (lane == 31) slab_total = wi;
    }
    __syncthreads();
    const int excl = carry + warp_excl[warp] + incl - len;
    if (b < B) {
      seq_start[b] = excl;
      seq_len[b] = len;
      pool_row[b] = excl + len - 1;
    }
    __syncthreads();
    if (tid == 0) carry += slab_total;
    __syncthreads();
  }
  if (tid == 0) rows_total[0] = carry;
}

template <int D>
__global__ void __launch_bounds__(256)
text_embed_packed_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask,
                         const float* __restrict__ tok, const float* __restrict__ pos, const int B, const int S,
                         const int vocab, const int* __restrict__ seq_start, const int* __restrict__ seq_len,
                         float* __restrict__ x, uint8_t* __restrict__ key_valid) {
  constexpr int V = D / 128;
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * S) return;
  const int b = warp / S, s = warp - b * S;
  if (s >= seq_len[b]) return;
  const int orow = seq_start[b] + s;
  long long id = ids[warp];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const float4* tr = reinterpret_cast<const float4*>(tok + (size_t)id * D);
  const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)s * D);
  float4* xo = reinterpret_cast<float4*>(x + (size_t)orow * D);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 a = __ldg(tr + lane + 32 * i), p = __ldg(pr + lane + 32 * i);
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    xo[lane + 32 * i] = a;
  }
  if (lane == 0) key_valid[orow] = mask ? (mask[warp] != 0 ? 1 : 0) : 1;
}

// ------------------------------------------------------------------------------------------------
// K1a im2col + cast for the patch-embedding convolution (stride == kernel == patch):
//   A[b*P + py*G + px, c*p*p + ky*p + kx] = pixel_values[b, c, py*p + ky, px*p + kx]   (bf16)
// which makes Conv2d(3, D, p, p) (HF clip :148-154,209-210) the GEMM A @ W.view(D, 3*p*p)^T.
// Each thread converts 8 consecutive kx (two float4 loads -> one 16-byte store).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ A, const int B, const int img,
              const int p) {
  pdl_trigger();
  pdl_wait();
  const int G = img / p;
  const int K = 3 * p * p;
  const int chunks_per_row = K / 8;
  const size_t total = (size_t)B * G * G * chunks_per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks_per_row);
    const size_t row = i / chunks_per_row;
    const int k = ch * 8;
    const int c = k / (p * p), rem = k - c * p * p;
    const int ky = rem / p, kx = rem - ky * p;
    const int b = (int)(row / (G * G)), pr = (int)(row - (size_t)b * G * G);
    const int py = pr / G, pxi = pr - py * G;
    const float* src = px + (((size_t)b * 3 + c) * img + (py * p + ky)) * img + pxi * p + kx;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(d.x, d.y); o.w = pack_bf16x2(d.z, d.w);
    *reinterpret_cast<uint4*>(A + row * K + k) = o;
  }
}

// K1a', uint8 source: the same patch matrix straight from raw HWC uint8 images, with ToTensor +
// Normalize (R/src/data/dataset.py:106-111) applied in registers:
//   A[b*P + py*G + px, c*p*p + ky*p + kx] = bf16( (u8[b, py*p+ky, px*p+kx, c] / 255 - mean[c]) / std[c] )
// fp32 operation order of torchvision (div, sub, div; no contraction), then the same round-to-nearest bf16 cast as
// im2col_kernel -> bit-identical to preprocess_u8_kernel followed by im2col_kernel, without the fp32 image in HBM
// (150 KB read + 301 KB written per 224 px sample instead of 752 KB + 903 KB) and with 4x less to ship from the host.
// One thread = 8 consecutive kx of one (patch, ky) for all three channels: 24 interleaved bytes -> three 16-byte stores.
__global__ void __launch_bounds__(256)
im2col_u8_kernel(const uint8_t* __restrict__ hwc, __nv_bfloat16* __restrict__ A, const int B, const int img, const int p,
                 const float m0, const float m1, const float m2, const float s0, const float s1, const float s2,
                 const int aligned8) {
  // the transform has 3 x 256 possible results: tabulate them once per CTA with the exact fp32 expression
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    lut[c][i & 255] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), mean), sd);
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  const int G = img / p;
  const int K = 3 * p * p;
  const int p8 = p / 8;
  const size_t total = (size_t)B * G * G * p * p8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kx = (int)(i % p8) * 8;
    size_t r = i / p8;
    const int ky = (int)(r % p);
    const size_t row = r / p;
    const int b = (int)(row / (G * G)), pr = (int)(row - (size_t)b * G * G);
    const int py = pr / G, pxi = pr - py * G;
    const uint8_t* src = hwc + (((size_t)b * img + (py * p + ky)) * img + (pxi * p + kx)) * 3;
    uint8_t v[24];
    if (aligned8) {
      const uint2* s2p = reinterpret_cast<const uint2*>(src);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint2 w = __ldg(s2p + j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[j * 8 + k] = (uint8_t)(w.x >> (8 * k));
          v[j * 8 + 4 + k] = (uint8_t)(w.y >> (8 * k));
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 24; ++j) v[j] = __ldg(src + j);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = lut[c][v[k * 3 + c]];
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(A + row * K + (size_t)c * p * p + ky * p + kx) = o;
    }
  }
}

// CLIP class-token rows: x[b*T + 0, :] = class_embedding + position_embedding[0]   (HF clip :212-217)
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                const int B, const int T, const int D) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[(size_t)b * T * D + d] = cls[d] + pos[d];
}

// pool_row[b] = b*T + t0  (vision CLS row)
__global__ void fill_pool_rows_kernel(int* __restrict__ pool_row, const int B, const int T, const int t0) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) pool_row[b] = b * T + t0;
}

// Last encoder layer, pooled rows only: only row pool_row[b] of each sample is read after the last layer (CLS / EOS /
// last token, HF clip :575-584, :688-689), and everything after the attention core is row-wise (out_proj, residual,
// LN2, MLP, final LN).  Gather those rows of the attention output (bf16) and of the residual stream (fp32) into
// dense [B, D] buffers; the remaining GEMMs of the layer then run with M = B instead of M = B*T.  One warp per row.
__global__ void gather_pool_rows_kernel(const __nv_bfloat16* __restrict__ att, const float* __restrict__ x,
                                        const int* __restrict__ pool_row, const int B, const int D,
                                        __nv_bfloat16* __restrict__ att_p, float* __restrict__ x_p) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const size_t src = (size_t)pool_row[b] * D, dst = (size_t)b * D;
  for (int c = lane * 8; c < D; c += 256)
    *reinterpret_cast<uint4*>(att_p + dst + c) = *reinterpret_cast<const uint4*>(att + src + c);
  for (int c = lane * 4; c < D; c += 128)
    *reinterpret_cast<float4*>(x_p + dst + c) = *reinterpret_cast<const float4*>(x + src + c);
}

// fp32 -> bf16 (weights repack; `scale` folds dh^-1/2 into the Q projection -- exact, 1/8 is a power of two)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, const size_t n,
                                 const float scale) {
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i] * scale);
}
__global__ void scale_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, const size_t n,
                                 const float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale;
}

}  // namespace mmcm
