// K2: out[M,N] = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear of the encoder blocks (q/k/v/out_proj, fc1, fc2:
// HF/models/clip/modeling_clip.py:295-298,344-345) and the patch-embedding convolution
// (:148-154, as an im2col GEMM), with their bias / activation / residual fused into the epilogue.
//
// Design (sm_100a):
//   * persistent CTAs (grid = min(#tiles, 148)), 256 threads, one CTA per SM (>=192 KB smem ring)
//   * warp 0   : TMA producer  -- cp.async.bulk.tensor.2d, 128B-swizzled 64-wide K slabs of A (128 rows)
//                                 and W (BLOCK_N rows) into a STAGES-deep shared-memory ring
//   * warp 1   : MMA issuer    -- one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16),
//                                 accumulators live in TMEM; tcgen05.commit frees ring slots / publishes tiles
//   * warp 2   : TMEM allocator (2 x BLOCK_N columns: the accumulator is double buffered so the epilogue
//                                 of tile i overlaps the main loop of tile i+1)
//   * warps 4-7: epilogue      -- tcgen05.ld 32x32b.x32 (lane == output row), fused bias/act/residual,
//                                 128-bit global stores
//   Both operands are K-major, which is exactly nn.Linear's [out,in] weight layout: no transposes anywhere.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mmcm {

enum Epi : int { EPI_BIAS_BF16 = 0, EPI_BIAS_ACT_BF16 = 1, EPI_BIAS_RESID_F32 = 2, EPI_PATCH_F32 = 3 };

struct EpiParams {
  const float* bias;   // [N] or nullptr
  void* out;           // bf16 [M,ldo] or fp32 [*,ldo]
  const float* resid;  // fp32 [M,ldo] (EPI_BIAS_RESID_F32; may alias out)
  const float* pos;    // fp32 [T,N]   (EPI_PATCH_F32)
  int ldo;             // row pitch of out / resid in elements
  int P, T;            // EPI_PATCH_F32: patches per sample, tokens per sample (T-P leading class rows)
  int act;             // Act for EPI_BIAS_ACT_BF16
};

// ------------------------------------------------------------------------------------------------
// epilogue math, shared by the tcgen05 kernel (32-column chunks) and the SIMT validation kernel
// ------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epi_store1(const EpiParams& ep, int row, int col, float acc) {
  float v = acc + (ep.bias ? __ldg(ep.bias + col) : 0.0f);
  if (EPI == EPI_BIAS_BF16) {
    reinterpret_cast<__nv_bfloat16*>(ep.out)[(size_t)row * ep.ldo + col] = __float2bfloat16_rn(v);
  } else if (EPI == EPI_BIAS_ACT_BF16) {
    reinterpret_cast<__nv_bfloat16*>(ep.out)[(size_t)row * ep.ldo + col] = __float2bfloat16_rn(apply_act(v, ep.act));
  } else if (EPI == EPI_BIAS_RESID_F32) {
    size_t o = (size_t)row * ep.ldo + col;
    reinterpret_cast<float*>(ep.out)[o] = v + (ep.resid ? ep.resid[o] : 0.0f);
  } else {
    int b = row / ep.P, p = row - b * ep.P, off = ep.T - ep.P;
    size_t o = ((size_t)b * ep.T + off + p) * ep.ldo + col;
    reinterpret_cast<float*>(ep.out)[o] = v + __ldg(ep.pos + (size_t)(off + p) * ep.ldo + col);
  }
}

// one output row, 32 consecutive columns starting at col0 (col0 % 32 == 0)
template <int EPI>
__device__ __forceinline__ void epi_store32(const EpiParams& ep, int row, int col0, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_ACT_BF16) {
    if (EPI == EPI_BIAS_ACT_BF16) {
      if (ep.act == ACT_QUICK_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
      } else if (ep.act == ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      o[i] = w;
    }
  } else if (EPI == EPI_BIAS_RESID_F32) {
    size_t off = (size_t)row * ep.ldo + col0;
    const float4* rs = reinterpret_cast<const float4*>(ep.resid + off);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off);
    const bool has_resid = ep.resid != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 x = has_resid ? rs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      x.x += v[4 * i + 0]; x.y += v[4 * i + 1]; x.z += v[4 * i + 2]; x.w += v[4 * i + 3];
      o[i] = x;
    }
  } else {
    int b = row / ep.P, p = row - b * ep.P, poff = ep.T - ep.P;
    const float4* ps = reinterpret_cast<const float4*>(ep.pos + (size_t)(poff + p) * ep.ldo + col0);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) +
                                          ((size_t)b * ep.T + poff + p) * ep.ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 x = __ldg(ps + i);
      x.x += v[4 * i + 0]; x.y += v[4 * i + 1]; x.z += v[4 * i + 2]; x.w += v[4 * i + 3];
      o[i] = x;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier / TMA / tcgen05).  sm_100a only.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on the phase with a wall-clock bound (4 s): a protocol bug must trap (launch failure), never hang the GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (it & 0x3FFu) == 0x3FFu) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load: coordinates are (innermost = K element index, outer = row index)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile whose rows are 128 bytes (64 elements) wide and
// 128B-swizzled by TMA: 8-row groups are 1024 B apart (SBO), LBO is unused for swizzled K-major layouts,
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  Field layout: cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (>>4), bits [16,30) -- ignored
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset (>>4), bits [32,46)
  d |= (uint64_t)1 << 46;                        // version = 1, bits [46,48)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int BLOCK_K = 64;  // one 128-byte swizzle row of bf16
  static constexpr int UMMA_K = 16;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // +1024: manual 1 KB alignment
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                   // double-buffered accumulator
  static constexpr int THREADS = 256;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const EpiParams ep, const int M, const int N, const int K) {
  using C = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES];
  __shared__ __align__(8) uint64_t bar_empty[C::STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B aligned tiles

  const int tiles_n = N / BLOCK_N;
  const int tiles_m = (M + C::BLOCK_M - 1) / C::BLOCK_M;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / C::BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 128);  // every epilogue thread arrives
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_holder), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        if (lane == 0) {
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          mbar_expect_tx(full, C::STAGE_BYTES);
          tma_load_2d(&tmap_a, full, sa, kb * C::BLOCK_K, m_blk * C::BLOCK_M);
          tma_load_2d(&tmap_b, full, sa + C::A_BYTES, kb * C::BLOCK_K, n_blk * BLOCK_N);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(C::BLOCK_M, BLOCK_N);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);      // TMA bytes have landed
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-byte units
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(&bar_empty[stage]));                       // ring slot reusable when MMAs retire
          if (kb == num_kb - 1) umma_commit(smem_u32(&bar_tfull[as]));   // accumulator complete
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (4 warps; warp w owns TMEM lanes 32*(w%4)..+31) =====================
    const int ew = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      const int row = m_blk * C::BLOCK_M + ew * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (row < M) epi_store32<EPI>(ep, row, n_blk * BLOCK_N + c * 32, r);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT validation kernel: same operands, same epilogue, no tensor cores.  Used only by tests
// (`gemm_impl = 1`) to separate "pipeline wrong" from "tcgen05 descriptor wrong".
// ------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __nv_bfloat16* __restrict__ A,
                                                        const __nv_bfloat16* __restrict__ W, const EpiParams ep,
                                                        const int M, const int N, const int K) {
  __shared__ float As[32][33];
  __shared__ float Ws[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = ty + 8 * i;
      int gm = m0 + r;
      As[r][tx] = (gm < M) ? __bfloat162float(A[(size_t)gm * K + k0 + tx]) : 0.f;
      Ws[r][tx] = __bfloat162float(W[(size_t)(n0 + r) * K + k0 + tx]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      float w = Ws[tx][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][kk], w, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty + 8 * i;
    if (gm < M) epi_store1<EPI>(ep, gm, n0 + tx, acc[i]);
  }
}

}  // namespace mmcm
