// K6 (projection) + K7 / K8: the late-fusion head and the multi-task head, each as ONE kernel, fp32.
//
//   fusion : R/src/models/fusion.py:187-216   (+ CLIP text_projection / visual_projection,
//            HF/models/clip/modeling_clip.py:822-823,860-861, which only ever see the pooled rows)
//   mtl    : R/src/models/multitask.py:183-207
//
// The heads are 0.04 % of the model FLOPs but ~15 tiny launches in the reference.  Here one CTA owns
// HEAD_SB samples, keeps every intermediate in shared memory and streams the (L2-resident, 11.6 MB fp32)
// weights once per CTA with 128-bit loads: one warp per output column, lanes split K, shuffle reduction.
// Everything stays fp32 (the head is where bf16 error would be amplified into the logits).
#pragma once
#include "common.cuh"

namespace mmcm {

constexpr int HEAD_SB = 8;        // samples per CTA
constexpr int HEAD_THREADS = 512;
constexpr int HEAD_MAXD = 768;    // widest tower feature

enum HeadAct : int { HA_NONE = 0, HA_TANH = 1, HA_SIGMOID = 2, HA_GELU = 3 };

struct HeadWeights {
  int head;        // 0 fusion, 1 mtl
  int backend;     // 0 clip, 1 siglip
  int dt, dv;      // pooled feature widths coming out of the towers
  int dp;          // width after the CLIP projections (fusion-clip) else == dt/dv
  int fd;          // fusion_dim
  int n_out;       // labels / tasks
  int hh;          // mtl head_hidden_dim (0 = single Linear per task)
  const float *text_proj, *vis_proj;                // [dp,dt], [dp,dv] no bias (fusion + clip only)
  const float *text_head_w, *text_head_b;           // SigLIP text_model.head Linear(dt,dt)+bias (HF siglip :520-522)
  const float *proj_t_w, *proj_t_b, *proj_i_w, *proj_i_b;
  const float *g_t_w, *g_t_b, *g_i_w, *g_i_b;
  const float *gate_w, *gate_wp, *gate_b;           // gate_w [fd, 2fd] (repacked), gate_wp [fd,2] presence columns
  const float *ln_fused_g, *ln_fused_b, *cls0_g, *cls0_b, *cls1_w, *cls1_b, *cls4_w, *cls4_b;  // fusion
  const float *shared_w, *shared_b;                 // mtl shared_head.1
  const float *h0_w, *h0_b, *h3_w, *h3_b;           // mtl: [T,hh,fd],[T,hh],[T,hh],[T]  (hh>0) | h3_w [T,fd], h3_b [T]
};

__device__ __forceinline__ float head_act(float v, int act) {
  if (act == HA_TANH) return tanhf(v);
  if (act == HA_SIGMOID) return 1.0f / (1.0f + expf(-v));
  if (act == HA_GELU) return gelu_erf(v);
  return v;
}

// ys[s][n] = act(sum_k xs[s][k] * W[n][k] + bias[n]) for s < HEAD_SB, n < N.  K % 4 == 0, W rows 16-B aligned.
// A warp produces HEAD_NB = 4 output columns per pass: every 16-byte slice of x that is read from shared memory is
// used for 4 weight rows (the first version re-read x for every row and was shared-memory bound, 1.1 ms per 1024
// samples), the weight rows are streamed with coalesced 128-bit loads, and the 8 x 4 = 32 partial sums per lane are
// reduced across the warp with a transposing butterfly (31 shuffles instead of 160) that leaves (sample, column) =
// (lane / 4, lane % 4) in lane `lane`.
constexpr int HEAD_NB = 4;
static_assert(HEAD_SB * HEAD_NB == 32, "the butterfly below assumes 32 partial sums per lane");

// ---- small batches: one sample group per CLUSTER --------------------------------------------------------------
// With B <= 8 a single CTA streamed all 11.6 MB of fp32 head weights through one SM: 341 us, a quarter of the B = 1
// latency.  Launched as a thread-block cluster (HEAD_CLUSTER CTAs per group of HEAD_SB samples), every CTA computes
// 1 / HEAD_CLUSTER of the output columns of each Linear and stores them into the activation buffers of ALL CTAs of
// the cluster through distributed shared memory; the cheap element-wise steps between the Linears are done
// redundantly by everyone on its full local copy.  cluster size 1 (large batches) is the same code.
constexpr int HEAD_CLUSTER = 8;
__device__ __forceinline__ uint32_t head_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t head_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// CTA barrier + cluster barrier with release / acquire: DSMEM stores before it are visible to every CTA after it
__device__ __forceinline__ void head_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void head_store_all(float* local, const float v, const uint32_t csize) {
  if (csize == 1) { *local = v; return; }
  const uint32_t la = smem_u32(local);
  for (uint32_t r = 0; r < csize; ++r) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}

__device__ __forceinline__ void block_linear(const float* __restrict__ W, const float* __restrict__ bias, const int N,
                                             const int K, const float* xs, const int ldx, float* ys,
                                             const int ldy, const int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int K4 = K >> 2;
  const uint32_t crank = head_cluster_rank(), csize = head_cluster_size();
  for (int n0 = ((int)crank * nw + warp) * HEAD_NB; n0 < N; n0 += (int)csize * nw * HEAD_NB) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    const float4* wr[HEAD_NB];
#pragma unroll
    for (int j = 0; j < HEAD_NB; ++j) wr[j] = reinterpret_cast<const float4*>(W + (size_t)min(n0 + j, N - 1) * K);
    // two k-slices per iteration: 8 independent 128-bit weight loads in flight per lane (the kernel is latency bound:
    // one CTA per SM, so bytes in flight = warps x loads per lane x 512 B)
    for (int k4 = lane; k4 < K4; k4 += 64) {
      float4 w[2][HEAD_NB];
      const bool second = (k4 + 32) < K4;
#pragma unroll
      for (int j = 0; j < HEAD_NB; ++j) {
        w[0][j] = __ldg(wr[j] + k4);
        w[1][j] = second ? __ldg(wr[j] + k4 + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !second) break;
#pragma unroll
        for (int s = 0; s < HEAD_SB; ++s) {
          const float4 x = *reinterpret_cast<const float4*>(xs + s * ldx + 4 * (k4 + 32 * u));
#pragma unroll
          for (int j = 0; j < HEAD_NB; ++j)
            v[s * HEAD_NB + j] = fmaf(w[u][j].x, x.x, fmaf(w[u][j].y, x.y, fmaf(w[u][j].z, x.z, fmaf(w[u][j].w, x.w, v[s * HEAD_NB + j]))));
        }
      }
    }
    // transposing butterfly: after the step with offset `off`, a lane keeps the half of its values selected by its bit
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = upper ? v[i] : v[i + off];
        const float keep = upper ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    const int s = lane / HEAD_NB, j = lane % HEAD_NB;
    if (n0 + j < N) head_store_all(ys + s * ldy + n0 + j, head_act(v[0] + (bias ? __ldg(bias + n0 + j) : 0.f), act), csize);
  }
}

// in-place LayerNorm of HEAD_SB smem rows of width D (one warp per row), eps = 1e-5 (nn.LayerNorm default)
__device__ __forceinline__ void block_layernorm(float* xs, const int ld, const int D, const float* __restrict__ g,
                                                const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int s = warp; s < HEAD_SB; s += nw) {
    float* r = xs + s * ld;
    float sum = 0.f;
    for (int k = lane; k < D; k += 32) sum += r[k];
    const float mean = warp_sum(sum) / D;
    float q = 0.f;
    for (int k = lane; k < D; k += 32) { const float d = r[k] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
    for (int k = lane; k < D; k += 32) r[k] = (r[k] - mean) * rstd * __ldg(g + k) + __ldg(b + k);
  }
}

__global__ void __launch_bounds__(HEAD_THREADS, 1)
head_kernel(const HeadWeights w, const float* __restrict__ pooled_t, const float* __restrict__ pooled_v,
            const float* __restrict__ tpres, const float* __restrict__ ipres, const int B,
            float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ feat_t_out,
            float* __restrict__ feat_v_out) {
  extern __shared__ __align__(16) float hs[];
  pdl_trigger();
  pdl_wait();
  const int fd = w.fd;
  const int ldf = 5 * fd;
  float* bufA = hs;                              // [SB][HEAD_MAXD]
  float* bufB = bufA + HEAD_SB * HEAD_MAXD;      // [SB][HEAD_MAXD]
  float* z0 = bufB + HEAD_SB * HEAD_MAXD;        // [SB][fd] x3
  float* z1 = z0 + HEAD_SB * fd;
  float* z2 = z1 + HEAD_SB * fd;
  float* feat = z2 + HEAD_SB * fd;               // [SB][5 fd] = [fused | tp | vp | |tp-vp| | tp*vp]
  __shared__ float s_tp[HEAD_SB], s_ip[HEAD_SB];

  const int s0 = (blockIdx.x / head_cluster_size()) * HEAD_SB;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;

  for (int i = tid; i < HEAD_SB * w.dt; i += nthr) {
    const int s = i / w.dt, k = i - s * w.dt;
    bufA[s * HEAD_MAXD + k] = (s0 + s < B) ? pooled_t[(size_t)(s0 + s) * w.dt + k] : 0.f;
  }
  for (int i = tid; i < HEAD_SB * w.dv; i += nthr) {
    const int s = i / w.dv, k = i - s * w.dv;
    bufB[s * HEAD_MAXD + k] = (s0 + s < B) ? pooled_v[(size_t)(s0 + s) * w.dv + k] : 0.f;
  }
  if (tid < HEAD_SB) {
    s_tp[tid] = (s0 + tid < B) ? tpres[s0 + tid] : 0.f;
    s_ip[tid] = (s0 + tid < B) ? ipres[s0 + tid] : 0.f;
  }
  head_sync();

  const float* tin = bufA;   // text feature rows, pitch ldt
  const float* vin = bufB;
  int ldt = HEAD_MAXD, ldv = HEAD_MAXD, dt = w.dt, dv = w.dv;
  if (w.text_head_w) {  // SigLIP: pooler_output = head(final_LN(x)[:, -1]); staged through `feat` (free until later)
    block_linear(w.text_head_w, w.text_head_b, w.dt, w.dt, bufA, HEAD_MAXD, feat, HEAD_MAXD, HA_NONE);
    head_sync();
    for (int i = tid; i < HEAD_SB * w.dt; i += nthr) {
      const int s = i / w.dt, k = i - s * w.dt;
      bufA[s * HEAD_MAXD + k] = feat[s * HEAD_MAXD + k];
    }
    head_sync();
  }
  if (w.head == 0) {
    if (w.text_proj) {  // CLIP: pooled -> projection_dim, no bias
      block_linear(w.text_proj, nullptr, w.dp, w.dt, bufA, HEAD_MAXD, z0, fd, HA_NONE);
      block_linear(w.vis_proj, nullptr, w.dp, w.dv, bufB, HEAD_MAXD, z1, fd, HA_NONE);
      head_sync();
      tin = z0; vin = z1; ldt = ldv = fd; dt = dv = w.dp;
    }
    if (feat_t_out) {  // introspection: get_text_features / get_image_features outputs (before normalisation)
      for (int i = tid; i < HEAD_SB * dt; i += nthr) {
        const int s = i / dt, k = i - s * dt;
        if (s0 + s < B) feat_t_out[(size_t)(s0 + s) * dt + k] = tin[s * ldt + k];
      }
      for (int i = tid; i < HEAD_SB * dv; i += nthr) {
        const int s = i / dv, k = i - s * dv;
        if (s0 + s < B) feat_v_out[(size_t)(s0 + s) * dv + k] = vin[s * ldv + k];
      }
      head_sync();
    }
    // F.normalize(dim=-1, eps=1e-12) * presence   (fusion.py:188-189)
    for (int s = warp; s < 2 * HEAD_SB; s += nw) {
      const bool is_t = s < HEAD_SB;
      const int ss = is_t ? s : s - HEAD_SB;
      float* r = const_cast<float*>(is_t ? tin + ss * ldt : vin + ss * ldv);
      const int d = is_t ? dt : dv;
      float q = 0.f;
      for (int k = lane; k < d; k += 32) q += r[k] * r[k];
      const float nrm = fmaxf(sqrtf(warp_sum(q)), 1e-12f);
      const float sc = (is_t ? s_tp[ss] : s_ip[ss]) / nrm;
      for (int k = lane; k < d; k += 32) r[k] *= sc;
    }
    head_sync();
  }
  // proj_t / proj_i -> feat[:, fd:2fd], feat[:, 2fd:3fd]
  block_linear(w.proj_t_w, w.proj_t_b, fd, dt, tin, ldt, feat + fd, ldf, HA_NONE);
  block_linear(w.proj_i_w, w.proj_i_b, fd, dv, vin, ldv, feat + 2 * fd, ldf, HA_NONE);
  head_sync();
  // zt, zi, gate(cat[tp, vp, presence])
  block_linear(w.g_t_w, w.g_t_b, fd, fd, feat + fd, ldf, z0, fd, HA_TANH);
  block_linear(w.g_i_w, w.g_i_b, fd, fd, feat + 2 * fd, ldf, z1, fd, HA_TANH);
  block_linear(w.gate_w, nullptr, fd, 2 * fd, feat + fd, ldf, z2, fd, HA_NONE);
  head_sync();
  for (int i = tid; i < HEAD_SB * fd; i += nthr) {
    const int s = i / fd, n = i - s * fd;
    const float tpv = s_tp[s], ipv = s_ip[s];
    const float gpre = z2[i] + __ldg(w.gate_wp + 2 * n) * tpv + __ldg(w.gate_wp + 2 * n + 1) * ipv + __ldg(w.gate_b + n);
    const float g = 1.0f / (1.0f + expf(-gpre));
    const float zt = z0[i], zi = z1[i];
    const float fused = (ipv < 0.5f) ? zt : ((tpv < 0.5f) ? zi : g * zt + (1.0f - g) * zi);
    feat[s * ldf + n] = fused;
    if (w.head == 0) {
      const float a = feat[s * ldf + fd + n], b = feat[s * ldf + 2 * fd + n];
      feat[s * ldf + 3 * fd + n] = fabsf(a - b);
      feat[s * ldf + 4 * fd + n] = a * b;
    }
  }
  head_sync();

  if (w.head == 0) {
    block_layernorm(feat, ldf, fd, w.ln_fused_g, w.ln_fused_b);      // ln_fused on the fused slice only
    head_sync();
    block_layernorm(feat, ldf, 5 * fd, w.cls0_g, w.cls0_b);          // cls.0
    head_sync();
    block_linear(w.cls1_w, w.cls1_b, fd, 5 * fd, feat, ldf, z0, fd, HA_GELU);   // cls.1 + GELU (Dropout: eval no-op)
    head_sync();
    block_linear(w.cls4_w, w.cls4_b, w.n_out, fd, z0, fd, z1, fd, HA_NONE);     // cls.4
    head_sync();
  } else {
    block_linear(w.shared_w, w.shared_b, fd, fd, feat, ldf, z0, fd, HA_GELU);   // shared_head
    head_sync();
    if (w.hh > 0) {
      for (int j = 0; j < w.n_out; ++j) {
        block_linear(w.h0_w + (size_t)j * w.hh * fd, w.h0_b + (size_t)j * w.hh, w.hh, fd, z0, fd, z2, fd, HA_GELU);
        head_sync();
        // Linear(hh, 1): one warp per sample
        for (int s = warp; s < HEAD_SB; s += nw) {
          float a = 0.f;
          for (int k = lane; k < w.hh; k += 32) a = fmaf(z2[s * fd + k], __ldg(w.h3_w + (size_t)j * w.hh + k), a);
          a = warp_sum(a);
          if (lane == 0) z1[s * fd + j] = a + __ldg(w.h3_b + j);
        }
        head_sync();
      }
    } else {
      block_linear(w.h3_w, w.h3_b, w.n_out, fd, z0, fd, z1, fd, HA_NONE);
      head_sync();
    }
  }
  if (head_cluster_rank() != 0) return;   // every CTA of the cluster holds the same logits: one of them stores
  for (int i = tid; i < HEAD_SB * w.n_out; i += nthr) {
    const int s = i / w.n_out, n = i - s * w.n_out;
    if (s0 + s < B) {
      const float l = z1[s * fd + n];
      logits[(size_t)(s0 + s) * w.n_out + n] = l;
      if (probs) probs[(size_t)(s0 + s) * w.n_out + n] = 1.0f / (1.0f + expf(-l));
    }
  }
}

inline int head_smem_bytes(int fd) { return (2 * HEAD_SB * HEAD_MAXD + 3 * HEAD_SB * fd + HEAD_SB * 5 * fd) * 4; }

}  // namespace mmcm
