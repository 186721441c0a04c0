// Shared device helpers for the sm_100a scoring kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmcm {

enum Act : int { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_GELU_TANH = 2 };

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch (PDL): every kernel of the forward is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may become resident while the previous kernel of
// the stream is still draining.  pdl_wait() blocks until all prerequisite grids have completed and their memory is
// visible -- it must precede the first global access that depends on them; pdl_trigger() lets the NEXT kernel start
// launching.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

// x * sigmoid(1.702 x)   (HF/activations.py:122-123)
__device__ __forceinline__ float quick_gelu(float x) {
  return __fdividef(x, 1.0f + __expf(-1.702f * x));
}

// 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))   (HF/activations.py:45, torch gelu(approximate="tanh"))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  // tanh(u) = 1 - 2/(exp(2u)+1); stable for large |u| with __expf saturating to inf / 0
  float t = 1.0f - __fdividef(2.0f, __expf(2.0f * u) + 1.0f);
  return 0.5f * x * (1.0f + t);
}

// MUFU-light variants for the GEMM epilogue (one tanh.approx per element instead of ex2 + rcp; max relative error of
// tanh.approx.f32 is 2^-11, below the 2^-9 rounding of the bf16 store that follows):
//   x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x))
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quick_gelu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(0.851f * x), h);
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float u = x * fmaf(k0k1, x * x, k0);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(u), h);
}

// exact GELU (erf), nn.GELU() default -- used by the fp32 heads (fusion.py:143, multitask.py:101)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_QUICK_GELU) return quick_gelu(x);
  if (act == ACT_GELU_TANH) return gelu_tanh(x);
  return x;
}

}  // namespace mmcm
