// Dev probe: raw tcgen05.mma issue throughput (no TMA, operands = whatever is in smem).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal-content-moderation_b200/csrc tools/mma_probe.cu -o /tmp/mma_probe
#include <cstdio>
#include "gemm2_tcgen05.cuh"
using namespace mmcm;

template <int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out, int a_stride_k) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 2) { if (PAIR) tmem_alloc_pair(smem_u32(&holder), 512); else tmem_alloc(smem_u32(&holder), 512); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  const bool issuer = warp == 1 && lane == 0 && (!PAIR || cluster_ctarank() == 0);
  if (issuer) {
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N);
    const uint64_t ad = make_smem_desc_sw128(base), bd = make_smem_desc_sw128(base + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t off = (uint64_t)(2 * k) + (uint64_t)((i & 3) * a_stride_k);
        if (PAIR) umma_f16_pair(tm, ad + off, bd + off, idesc, 1);
        else umma_f16(tm, ad + off, bd + off, idesc, 1);
      }
    }
    if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tm, 512); else tmem_dealloc(tm, 512); }
}

template <int N, bool PAIR>
void run(const char* name, int grid, int stride) {
  long long* d; cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8);
  auto k = probe<N, PAIR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, iters, d, stride);
    if (e != cudaSuccess) { printf("%s launch: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s sync: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[1024]; cudaMemcpy(h, d, 1024 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  printf("%-34s grid=%3d  %.1f cycles / MMA (K=16)\n", name, grid, (double)mx / (iters * 4));
  cudaFree(d);
}

int main() {
  run<256, false>("cg1 M=128 N=256", 1, 0);
  run<256, false>("cg1 M=128 N=256", 148, 0);
  run<256, false>("cg1 M=128 N=256 (4 k-slabs)", 148, 3072);   // walk 4 different 48 KB stage-like regions
  run<128, false>("cg1 M=128 N=128", 148, 0);
  run<256, true>("cg2 M=256 N=256", 2, 0);
  run<256, true>("cg2 M=256 N=256", 148, 0);
  run<128, true>("cg2 M=256 N=128", 148, 0);
  return 0;
}
