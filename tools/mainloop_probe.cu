// Dev probe: GEMM main loop only (TMA ring -> tcgen05.mma, accumulators discarded) to separate operand-feed limits
// from epilogue effects.  Each CTA (or CTA pair) streams `kblocks` k-blocks per tile over `tiles` tiles.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "gemm2_tcgen05.cuh"
using namespace mmcm;

template <int BN, int STAGES, bool PAIR>
__global__ void __launch_bounds__(128, 1)
mainloop(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int tiles_m, int tiles_n,
         int num_kb, long long* out) {
  constexpr int A_BYTES = 128 * 64 * 2;
  constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 64 * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_done;
  __shared__ uint32_t holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init(); fence_proxy_async();
  }
  if (warp == 2) { if (PAIR) tmem_alloc_pair(smem_u32(&holder), 512); else tmem_alloc(smem_u32(&holder), 512); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  const int unit = PAIR ? blockIdx.x >> 1 : blockIdx.x, units = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int num_tiles = tiles_m * tiles_n;
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = unit; tile < num_tiles; tile += units) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        if (lane == 0) {
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = base + stage * STAGE_BYTES;
          if (PAIR) {
            if (rank == 0) mbar_expect_tx(full, 2 * STAGE_BYTES);
            tma_load_2d_pair(&ta, full, sa, kb * 64, m_blk * 256 + rank * 128);
            tma_load_2d_pair(&tb, full, sa + A_BYTES, kb * 64, n_blk * BN + rank * (BN / 2));
          } else {
            mbar_expect_tx(full, STAGE_BYTES);
            tma_load_2d(&ta, full, sa, kb * 64, m_blk * 128);
            tma_load_2d(&tb, full, sa + A_BYTES, kb * 64, n_blk * BN);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN);
    int stage = 0; uint32_t phase = 0;
    for (int tile = unit; tile < num_tiles; tile += units) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t ad = make_smem_desc_sw128(sa), bd = make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (PAIR) umma_f16_pair(tm, ad + 2 * k, bd + 2 * k, idesc, 1);
            else umma_f16(tm, ad + 2 * k, bd + 2 * k, idesc, 1);
          }
          if (PAIR) umma_commit_pair(smem_u32(&bar_empty[stage])); else umma_commit(smem_u32(&bar_empty[stage]));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (lane == 0) {
      if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
      else umma_commit(smem_u32(&bar_done));
    }
    mbar_wait(smem_u32(&bar_done), 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tm, 512); else tmem_dealloc(tm, 512); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn enc;
static CUtensorMap mk(void* p, long rows, long K, int box_rows, CUtensorMapL2promotion prom) {
  CUtensorMap m; cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
  cuuint32_t bx[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int BN, int STAGES, bool PAIR>
void run(const char* name, void* A, void* W, long M, int N, int K, CUtensorMapL2promotion prom) {
  CUtensorMap ta = mk(A, M, K, 128, prom), tb = mk(W, N, K, PAIR ? BN / 2 : BN, prom);
  const int bm = PAIR ? 256 : 128;
  const int tiles_m = (int)((M + bm - 1) / bm), tiles_n = N / BN;
  constexpr int smem = STAGES * (128 * 128 + (PAIR ? BN / 2 : BN) * 128) + 1024;
  auto k = mainloop<BN, STAGES, PAIR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 40; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, ta, tb, tiles_m, tiles_n, K / 64, d);
    cudaEventRecord(e1);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
  const double units = PAIR ? 74 : 148;
  const double kb_per_unit = (double)tiles_m * tiles_n * (K / 64) / units;
  printf("%-44s M=%6ld N=%5d K=%5d  %7.1f us  %7.1f TFLOP/s  %6.0f cycles/k-block (ideal 512)\n", name, M, N, K,
         best * 1e3, 2.0 * M * N * K / best / 1e9, mx / kb_per_unit);
  cudaFree(d);
}

int main() {
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q); enc = (EncodeFn)fn;
  const long M = 12288; const int N = 3072, K = 3072;
  void *A, *W; cudaMalloc(&A, M * K * 2); cudaMalloc(&W, (long)N * K * 2);
  cudaMemset(A, 0, M * K * 2); cudaMemset(W, 0, (long)N * K * 2);
  if (getenv("RANDOM_DATA")) {   // bf16 noise in [-1, 1): realistic switching activity / power
    size_t na = (size_t)M * K, nw = (size_t)N * K;
    unsigned short* h = (unsigned short*)malloc((na > nw ? na : nw) * 2);
    unsigned s = 12345u;
    for (size_t i = 0; i < na; ++i) { s = s * 1664525u + 1013904223u; h[i] = (unsigned short)(((s >> 16) & 0x807F) | 0x3F00 | ((s >> 9) & 0x80)); }
    cudaMemcpy(A, h, na * 2, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < nw; ++i) { s = s * 1664525u + 1013904223u; h[i] = (unsigned short)(((s >> 16) & 0x807F) | 0x3F00 | ((s >> 9) & 0x80)); }
    cudaMemcpy(W, h, nw * 2, cudaMemcpyHostToDevice);
    free(h);
  }
  auto P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; auto P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  run<256, 4, false>("single 128x256 4 stages", A, W, M, N, K, P128);
  run<256, 4, false>("single 128x256 4 stages promo256", A, W, M, N, K, P256);
  run<256, 3, false>("single 128x256 3 stages", A, W, M, N, K, P128);
  run<128, 6, false>("single 128x128 6 stages", A, W, M, N, K, P128);
  run<256, 6, true>("pair 256x256 6 stages", A, W, M, N, K, P128);
  run<256, 6, true>("pair 256x256 6 stages promo256", A, W, M, N, K, P256);
  run<256, 4, true>("pair 256x256 4 stages", A, W, M, N, K, P128);
  run<256, 6, true>("pair 256x256 6 stages K=768", A, W, M, N, 768, P128);
  run<256, 6, true>("pair 256x256 6 stages small(L2) M=2048", A, W, 2048, N, K, P128);
  run<256, 4, false>("single 128x256 4 stages small(L2) M=2048", A, W, 2048, N, K, P128);
  return 0;
}
