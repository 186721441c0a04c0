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
//   * warps 4-11: epilogue     -- tcgen05.ld 32x32b.x32 (lane == output row), fused bias/act in registers, then a
//                                 per-warp shared-memory transpose so that every global access (bf16 / fp32 store,
//                                 fp32 residual or position-embedding read) is a full 128-byte line per row:
//                                 warps e and e+4 share TMEM lanes 32*(e%4).. and split the tile's columns
//   * tiles are handed out by a global atomic counter (dynamic persistent scheduler): when the text and the vision
//     tower run on two streams, late-starting CTAs simply take fewer tiles
//   Both operands are K-major, which is exactly nn.Linear's [out,in] weight layout: no transposes anywhere.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mmcm {

// 0-3: every GEMM kernel.  4-6: CTA-pair kernel only -- LayerNorm folded into the GEMMs (gemm2_tcgen05.cuh, "LN fold"):
//   EPI_RESID_STATS      x += acc + bias (fp32, in place), plus a bf16 copy of the updated rows and per-row
//                        (sum, M2) partials over every 128-column slab  -> replaces the residual GEMM + half the LN pass
//   EPI_LNFOLD_BF16      out = rstd * acc + bias'   with acc = bf16(x) @ W'^T, W' = bf16(W * gamma, rows centred)
//   EPI_LNFOLD_ACT_BF16  same, then the activation                    -> the other half: LN applied in the consumer
enum Epi : int { EPI_BIAS_BF16 = 0, EPI_BIAS_ACT_BF16 = 1, EPI_BIAS_RESID_F32 = 2, EPI_PATCH_F32 = 3,
                 EPI_RESID_STATS = 4, EPI_LNFOLD_BF16 = 5, EPI_LNFOLD_ACT_BF16 = 6 };
constexpr int LN_SLAB = 128;   // columns per row-statistics partial = the column span of one epilogue warp (BLOCK_N / 2)

struct EpiParams {
  const float* bias;   // [N] or nullptr
  void* out;           // bf16 [M,ldo] or fp32 [*,ldo]
  const float* resid;  // fp32 [M,ldo] (EPI_BIAS_RESID_F32; may alias out)
  const float* pos;    // fp32 [T,N]   (EPI_PATCH_F32)
  int ldo;             // row pitch of out / resid in elements
  int P, T;            // EPI_PATCH_F32: patches per sample, tokens per sample (T-P leading class rows)
  int act;             // Act for EPI_BIAS_ACT_BF16
  const int* m_dev;    // optional device-side row count (<= M): packed variable-length text chunks
  // LN fold (EPI_RESID_STATS writes, EPI_LNFOLD_* read): float2 stats[slab][stats_pitch] = (sum, M2) of the row's
  // 128-column slab; xb = bf16 copy of the updated residual rows
  float2* stats;
  void* xb;
  int stats_pitch, ln_slabs;
  float ln_eps;
  // split-K (CTA-pair kernel, fp32 store epilogue only): tile t covers k-blocks [ks, ks + 1) * K / (64 ksplit) of output
  // tile t / ksplit and stores its partial sums into plane ks of `out` (planes part_rows rows apart); the bias goes into
  // plane 0.  Whoever reads the result adds the planes in order (layernorm_kernel) -- deterministic, no inter-CTA waits.
  int ksplit, part_rows;
  long long* trace;    // dev tool (mmcm_debug_set_gemm_trace): per-CTA clock64 stamps, 16 slots per CTA; else nullptr
};

// trace slots: 0 entry, 1 prologue done, 2 first operands landed, 3 MMAs of tile 0 issued, 4 accumulator 0 ready,
// 5 epilogue of tile 0 done, 6 MMAs of the last tile issued, 7 last accumulator ready, 8 last epilogue done, 9 exit
__device__ __forceinline__ void trace_stamp(const EpiParams& ep, int slot) {
  if (ep.trace) ep.trace[(size_t)blockIdx.x * 16 + slot] = clock64();
}

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

// ---- coalesced epilogue -------------------------------------------------------------------------
// A warp owns 32 accumulator rows (lane == row).  A chunk is 128 bytes of output per row (32 fp32 or 64 bf16
// columns).  Half a warp at a time writes its rows into a padded 16-row staging tile (pitch 144 B: conflict-free
// for 16-byte accesses), then the whole warp re-reads it so that 8 lanes cover one row: every global instruction
// touches 4 complete 128-byte lines instead of 32 partial ones.
constexpr int EPI_PITCH = 144;                   // bytes per staged row
constexpr int EPI_STAGE_BYTES = 16 * EPI_PITCH;  // per epilogue warp

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// bf16 outputs: 64 columns starting at col0 (two TMEM loads), rows row_base .. row_base+31
// `bias_smem` = shared-memory address of this chunk's 64 bias values (staged per tile by the warp itself: with
// ~220 KB of the SM's 228 KB used as shared memory the L1 is a few KB, and per-chunk __ldg's of the bias were L2
// round trips sitting directly in front of the FADDs -- the top stall of the first profile of this epilogue).
template <int EPI>
__device__ __forceinline__ void epi_chunk_bf16(const EpiParams& ep, const uint32_t stage, const uint32_t bias_smem,
                                               const int lane, const int row_base, const int col0, const int M,
                                               const uint32_t (&r0)[32], const uint32_t (&r1)[32]) {
  uint32_t w[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(i < 4 ? r0[8 * i + j] : r1[8 * (i - 4) + j]);
    {
      const uint4 ba = lds128(bias_smem + i * 32), bb = lds128(bias_smem + i * 32 + 16);
      v[0] += __uint_as_float(ba.x); v[1] += __uint_as_float(ba.y); v[2] += __uint_as_float(ba.z); v[3] += __uint_as_float(ba.w);
      v[4] += __uint_as_float(bb.x); v[5] += __uint_as_float(bb.y); v[6] += __uint_as_float(bb.z); v[7] += __uint_as_float(bb.w);
    }
    if (EPI == EPI_BIAS_ACT_BF16) {
      if (ep.act == ACT_QUICK_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = quick_gelu_fast(v[j]);
      } else if (ep.act == ACT_GELU_TANH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_tanh_fast(v[j]);
      }
    }
    w[4 * i + 0] = pack_bf16x2(v[0], v[1]);
    w[4 * i + 1] = pack_bf16x2(v[2], v[3]);
    w[4 * i + 2] = pack_bf16x2(v[4], v[5]);
    w[4 * i + 3] = pack_bf16x2(v[6], v[7]);
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
  const int sub = lane >> 3, c16 = lane & 7;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if ((lane >> 4) == pass) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        sts128(stage + (lane & 15) * EPI_PITCH + i * 16, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rl = it * 4 + sub;
      const uint4 v = lds128(stage + rl * EPI_PITCH + c16 * 16);
      const int row = row_base + pass * 16 + rl;
      if (row < M) *reinterpret_cast<uint4*>(out + (size_t)row * ep.ldo + col0 + c16 * 8) = v;
    }
    __syncwarp();
  }
}

// the same epilogue in two steps, for callers that want only 32 accumulator registers live at a time:
// pack 32 fp32 accumulators (+ bias from smem, + activation) into 16 bf16x2 words ...
// QuickGELU epilogues run on HALVED constants: the caller stages 0.5 * bias and passes 0.5 * rstd,
// so the pre-activation arrives as h = x / 2 (exact: scaling by a power of two commutes with every rounding) and
// x * sigmoid(1.702 x) = h + h * tanh(1.702 h) costs FMUL + MUFU + FFMA -- the LN-fold epilogue then issues exactly as
// many instructions per element as the plain bias + QuickGELU one did (tools/gemm_bench_fold.py: fc1 +10 % before).
template <int EPI>
__device__ __forceinline__ void epi_pack_bf16(const EpiParams& ep, const uint32_t bias_smem, const uint32_t (&r)[32],
                                              uint32_t* w, const float ln_rs = 0.f) {
  constexpr bool kFold = (EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_ACT_BF16);
  constexpr bool kAct = (EPI == EPI_BIAS_ACT_BF16 || EPI == EPI_LNFOLD_ACT_BF16);
  const bool halved = kAct && ep.act == ACT_QUICK_GELU;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * i + j]);
    const uint4 ba = lds128(bias_smem + i * 32), bb = lds128(bias_smem + i * 32 + 16);
    if (kFold) {   // LN fold: rstd * acc + bias'  (the weight rows are centred, so acc already is (x - mean) . W*gamma)
      v[0] = fmaf(v[0], ln_rs, __uint_as_float(ba.x)); v[1] = fmaf(v[1], ln_rs, __uint_as_float(ba.y));
      v[2] = fmaf(v[2], ln_rs, __uint_as_float(ba.z)); v[3] = fmaf(v[3], ln_rs, __uint_as_float(ba.w));
      v[4] = fmaf(v[4], ln_rs, __uint_as_float(bb.x)); v[5] = fmaf(v[5], ln_rs, __uint_as_float(bb.y));
      v[6] = fmaf(v[6], ln_rs, __uint_as_float(bb.z)); v[7] = fmaf(v[7], ln_rs, __uint_as_float(bb.w));
    } else if (halved) {   // h = 0.5 * (acc + bias) = fma(acc, 0.5, 0.5 * bias)
      v[0] = fmaf(v[0], 0.5f, __uint_as_float(ba.x)); v[1] = fmaf(v[1], 0.5f, __uint_as_float(ba.y));
      v[2] = fmaf(v[2], 0.5f, __uint_as_float(ba.z)); v[3] = fmaf(v[3], 0.5f, __uint_as_float(ba.w));
      v[4] = fmaf(v[4], 0.5f, __uint_as_float(bb.x)); v[5] = fmaf(v[5], 0.5f, __uint_as_float(bb.y));
      v[6] = fmaf(v[6], 0.5f, __uint_as_float(bb.z)); v[7] = fmaf(v[7], 0.5f, __uint_as_float(bb.w));
    } else {
      v[0] += __uint_as_float(ba.x); v[1] += __uint_as_float(ba.y); v[2] += __uint_as_float(ba.z); v[3] += __uint_as_float(ba.w);
      v[4] += __uint_as_float(bb.x); v[5] += __uint_as_float(bb.y); v[6] += __uint_as_float(bb.z); v[7] += __uint_as_float(bb.w);
    }
    if (kAct) {
      if (ep.act == ACT_QUICK_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], tanh_approx(1.702f * v[j]), v[j]);   // v holds h = x / 2
      } else if (ep.act == ACT_GELU_TANH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_tanh_fast(v[j]);
      }
    }
    w[4 * i + 0] = pack_bf16x2(v[0], v[1]);
    w[4 * i + 1] = pack_bf16x2(v[2], v[3]);
    w[4 * i + 2] = pack_bf16x2(v[4], v[5]);
    w[4 * i + 3] = pack_bf16x2(v[6], v[7]);
  }
}
// ... then transpose 32 rows x 64 bf16 columns (w[32] per lane) through the staging tile and store full lines
__device__ __forceinline__ void epi_store_bf16(const EpiParams& ep, const uint32_t stage, const int lane,
                                               const int row_base, const int col0, const int M, const uint32_t (&w)[32]) {
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
  const int sub = lane >> 3, c16 = lane & 7;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if ((lane >> 4) == pass) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        sts128(stage + (lane & 15) * EPI_PITCH + i * 16, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rl = it * 4 + sub;
      const uint4 v = lds128(stage + rl * EPI_PITCH + c16 * 16);
      const int row = row_base + pass * 16 + rl;
      if (row < M) *reinterpret_cast<uint4*>(out + (size_t)row * ep.ldo + col0 + c16 * 8) = v;
    }
    __syncwarp();
  }
}

// fp32 outputs: 32 columns starting at col0.  The addend (residual stream or position embedding) of a chunk is 8
// float4 per lane; it is loaded by `epi_load_addend` BEFORE the accumulator is waited for (the residual may alias
// the output, so the compiler cannot hoist these loads above earlier stores by itself -- a load->store->load chain
// of L2 round trips is what made the first version of this epilogue 5x slower than its MMAs).
template <int EPI>
__device__ __forceinline__ void epi_load_addend(const EpiParams& ep, const int lane, const int row_base,
                                                const int col0, const int M, float4 (&x)[8]) {
  const int sub = lane >> 3, c16 = lane & 7;
  const int col = col0 + c16 * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + i * 4 + sub;   // i = pass * 4 + it
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < M) {
      if (EPI == EPI_BIAS_RESID_F32) {
        if (ep.resid) x[i] = *reinterpret_cast<const float4*>(ep.resid + (size_t)row * ep.ldo + col);
      } else {
        const int b = row / ep.P, p = row - b * ep.P;
        x[i] = __ldg(reinterpret_cast<const float4*>(ep.pos + (size_t)(ep.T - ep.P + p) * ep.ldo + col));
      }
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epi_chunk_f32(const EpiParams& ep, const uint32_t stage, const int lane,
                                              const int row_base, const int col0, const int M,
                                              const uint32_t (&r)[32], const float4 (&x)[8], const float4 bias) {
  float* out = reinterpret_cast<float*>(ep.out);
  const int sub = lane >> 3, c16 = lane & 7;
  const int col = col0 + c16 * 4;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if ((lane >> 4) == pass) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        sts128(stage + (lane & 15) * EPI_PITCH + i * 16, r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rl = it * 4 + sub;
      const uint4 u = lds128(stage + rl * EPI_PITCH + c16 * 16);
      const float4 a = x[pass * 4 + it];
      const float4 v = make_float4(__uint_as_float(u.x) + bias.x + a.x, __uint_as_float(u.y) + bias.y + a.y,
                                   __uint_as_float(u.z) + bias.z + a.z, __uint_as_float(u.w) + bias.w + a.w);
      const int row = row_base + pass * 16 + rl;
      if (row < M) {
        if (EPI == EPI_BIAS_RESID_F32) {
          *reinterpret_cast<float4*>(out + (size_t)row * ep.ldo + col) = v;
        } else {  // EPI_PATCH_F32: GEMM row (sample b, patch p) -> token row b*T + (T-P) + p
          const int b = row / ep.P, p = row - b * ep.P;
          *reinterpret_cast<float4*>(out + ((size_t)b * ep.T + ep.T - ep.P + p) * ep.ldo + col) = v;
        }
      }
    }
    __syncwarp();
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
// immediate column count: ptxas records the kernel's TMEM footprint from it -- with a register operand it assumes all
// 512 columns and the occupancy calculator (and the block scheduler) allow only ONE such CTA per SM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_imm(uint32_t holder_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_imm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
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
  static constexpr int EPI_WARPS = 8;
  static constexpr int EPI_BIAS_BYTES = (BLOCK_N / 2) * 4;  // per epilogue warp: the bias of its column half
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + EPI_WARPS * (EPI_STAGE_BYTES + EPI_BIAS_BYTES) + 1024;  // +1024: 1 KB alignment
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                   // double-buffered accumulator
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int SCHED_SLOTS = 4;
};

// sched[0] = next tile, sched[1] = CTAs finished (the last one re-arms both for the next launch on this stream)
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GemmCfg<BLOCK_N>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const EpiParams ep, const int M_host, const int N, const int K, int* __restrict__ sched) {
  using C = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES];
  __shared__ __align__(8) uint64_t bar_empty[C::STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ __align__(8) uint64_t bar_sfull[C::SCHED_SLOTS];
  __shared__ __align__(8) uint64_t bar_sempty[C::SCHED_SLOTS];
  __shared__ int sched_tile[C::SCHED_SLOTS];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B aligned tiles
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;

  const int tiles_n = N / BLOCK_N;
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
      mbar_init(smem_u32(&bar_tempty[s]), C::EPI_WARPS);  // lane 0 of every epilogue warp arrives
    }
    for (int s = 0; s < C::SCHED_SLOTS; ++s) {
      mbar_init(smem_u32(&bar_sfull[s]), 1);
      mbar_init(smem_u32(&bar_sempty[s]), 1 + C::EPI_WARPS);  // MMA warp + every epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_holder), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  // PDL: everything above overlapped the previous kernel's tail; its outputs (A, the residual) are safe to read now
  pdl_trigger();
  pdl_wait();
  // packed variable-length text: the live row count is produced on the device by the previous kernel
  const int M = ep.m_dev ? min(__ldg(ep.m_dev), M_host) : M_host;
  const int tiles_m = (M + C::BLOCK_M - 1) / C::BLOCK_M;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0) {
    // ===================== tile scheduler + TMA producer =====================
    int stage = 0, slot = 0;
    uint32_t phase = 0, sphase = 0;
    while (true) {
      mbar_wait(smem_u32(&bar_sempty[slot]), sphase ^ 1u);
      int tile = 0;
      if (lane == 0) {
        tile = atomicAdd(&sched[0], 1);
        if (tile >= num_tiles) tile = -1;
        sched_tile[slot] = tile;
        mbar_arrive(smem_u32(&bar_sfull[slot]));
      }
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (++slot == C::SCHED_SLOTS) { slot = 0; sphase ^= 1u; }
      if (tile < 0) break;
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
    if (lane == 0) {
      __threadfence();
      const int done = atomicAdd(&sched[1], 1);
      if (done == (int)gridDim.x - 1) {  // every CTA has drawn its last tile: re-arm for the next launch
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(C::BLOCK_M, BLOCK_N);
    int stage = 0, as = 0, slot = 0;
    uint32_t phase = 0, aphase = 0, sphase = 0;
    while (true) {
      mbar_wait(smem_u32(&bar_sfull[slot]), sphase);
      const int tile = sched_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_sempty[slot]));
      if (++slot == C::SCHED_SLOTS) { slot = 0; sphase ^= 1u; }
      if (tile < 0) break;
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
    // ===================== epilogue: warp e owns TMEM lanes 32*(e%4).., columns half (e/4) =====================
    const int e = warp - 4;
    const int lg = e & 3, ch = e >> 2;
    const uint32_t stage_smem = epi_base + e * EPI_STAGE_BYTES;
    const uint32_t bias_smem = epi_base + C::EPI_WARPS * EPI_STAGE_BYTES + e * C::EPI_BIAS_BYTES;
    constexpr int HALF_N = BLOCK_N / 2;
    int as = 0, slot = 0;
    uint32_t aphase = 0, sphase = 0;
    while (true) {
      mbar_wait(smem_u32(&bar_sfull[slot]), sphase);
      const int tile = sched_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_sempty[slot]));
      if (++slot == C::SCHED_SLOTS) { slot = 0; sphase ^= 1u; }
      if (tile < 0) break;
      const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
      const int row_base = m_blk * C::BLOCK_M + lg * 32;
      const int col_base = n_blk * BLOCK_N + ch * HALF_N;
      constexpr bool kF32 = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_PATCH_F32);
      float4 xa[8];
      float4 fb[HALF_N / 32];   // fp32 path: this lane's bias for each 32-column chunk
      // everything the epilogue needs from global memory is requested BEFORE waiting for the accumulator
      if (kF32) {
#pragma unroll
        for (int c = 0; c < HALF_N / 32; ++c)
          fb[c] = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base + c * 32 + (lane & 7) * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_base < M) epi_load_addend<EPI>(ep, lane, row_base, col_base, M, xa);
      } else {
#pragma unroll
        for (int j = lane; j < HALF_N / 4; j += 32) {
          const float4 b = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base) + j)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          sts128(bias_smem + j * 16, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BLOCK_N + ch * HALF_N);
      if (row_base < M) {
        if (!kF32) {
#pragma unroll 1
          for (int c = 0; c < HALF_N / 64; ++c) {
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row + (uint32_t)(c * 64), r0);
            tmem_ld32(t_row + (uint32_t)(c * 64 + 32), r1);
            tmem_ld_wait();
            epi_chunk_bf16<EPI>(ep, stage_smem, bias_smem + c * 256, lane, row_base, col_base + c * 64, M, r0, r1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < HALF_N / 32; ++c) {
            uint32_t r[32];
            float4 xn[8];
            tmem_ld32(t_row + (uint32_t)(c * 32), r);
            if (c + 1 < HALF_N / 32) epi_load_addend<EPI>(ep, lane, row_base, col_base + (c + 1) * 32, M, xn);
            tmem_ld_wait();
            epi_chunk_f32<EPI>(ep, stage_smem, lane, row_base, col_base + c * 32, M, r, xa, fb[c]);
            if (c + 1 < HALF_N / 32) {
#pragma unroll
              for (int i = 0; i < 8; ++i) xa[i] = xn[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));
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
                                                        const int M_host, const int N, const int K) {
  __shared__ float As[32][33];
  __shared__ float Ws[32][33];
  pdl_trigger();
  pdl_wait();
  const int M = ep.m_dev ? min(__ldg(ep.m_dev), M_host) : M_host;
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
