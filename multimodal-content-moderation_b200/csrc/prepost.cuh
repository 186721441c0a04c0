// SURVEY 8(f) rows either side of the forward: the callers' pre- and post-processing.
//
//   preprocess_u8_kernel : ToTensor + Normalize of the eval transform (R/src/data/dataset.py:106-111) for an already
//                          resized / cropped uint8 HWC image: out[b,c,y,x] = (u8/255 - mean[c]) / std[c]  (fp32 CHW).
//                          Same operation order as torchvision (div by 255, sub, div) -> bit-identical pixel_values,
//                          with 4x less host->device traffic than shipping fp32 pixels.
//   postprocess_kernel   : probs = 1/(1+exp(-logits)), label = prob >= thresholds[c], any_harmful = any(label)
//                          (R/scripts/inference.py:218-232, R/sagemaker/inference.py:281-296), and per-class confusion
//                          counts TP/FP/FN/TN against labels, from which F1 / precision / recall of
//                          R/src/training/metrics.py:180-205 follow without a per-batch D2H of the logits.
#pragma once
#include "common.cuh"

namespace mmcm {

// one thread per 4 consecutive x of one (b, c, y): reads 12 interleaved bytes, writes one float4
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ hwc, float* __restrict__ chw, const int B, const int H, const int W,
                     const float m0, const float m1, const float m2, const float s0, const float s1, const float s2) {
  pdl_wait();
  const int W4 = W >> 2;
  const size_t total = (size_t)B * 3 * H * W4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % W4);
    size_t r = i / W4;
    const int y = (int)(r % H);
    r /= H;
    const int c = (int)(r % 3);
    const int b = (int)(r / 3);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const uint8_t* src = hwc + (((size_t)b * H + y) * W + (size_t)x4 * 4) * 3 + c;
    float4 o;
    o.x = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[0], 255.0f), mean), sd);
    o.y = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[3], 255.0f), mean), sd);
    o.z = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[6], 255.0f), mean), sd);
    o.w = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[9], 255.0f), mean), sd);
    *reinterpret_cast<float4*>(chw + (((size_t)b * 3 + c) * H + y) * W + (size_t)x4 * 4) = o;
  }
}

constexpr int POST_MAXC = 64;
// one thread per sample; per-class confusion counts are reduced in shared memory, then one atomic per class and CTA
__global__ void __launch_bounds__(256)
postprocess_kernel(const float* __restrict__ logits, const float* __restrict__ thresholds,
                   const float* __restrict__ labels, const int B, const int C, float* __restrict__ probs,
                   uint8_t* __restrict__ decisions, uint8_t* __restrict__ any_harmful,
                   unsigned long long* __restrict__ confusion) {
  __shared__ unsigned int cnt[POST_MAXC][4];
  pdl_wait();
  for (int i = threadIdx.x; i < C * 4; i += blockDim.x) cnt[i >> 2][i & 3] = 0u;
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    bool any = false;
    for (int c = 0; c < C; ++c) {
      const float l = logits[(size_t)b * C + c];
      const float p = __fdiv_rn(1.0f, 1.0f + expf(-l));     // numpy: 1 / (1 + np.exp(-logits)) in float32
      const bool d = p >= thresholds[c];
      any |= d;
      if (probs) probs[(size_t)b * C + c] = p;
      if (decisions) decisions[(size_t)b * C + c] = d ? 1 : 0;
      if (confusion && labels) {
        const bool y = labels[(size_t)b * C + c] >= 0.5f;
        atomicAdd(&cnt[c][d ? (y ? 0 : 1) : (y ? 2 : 3)], 1u);   // TP, FP, FN, TN
      }
    }
    if (any_harmful) any_harmful[b] = any ? 1 : 0;
  }
  __syncthreads();
  if (confusion && labels)
    for (int i = threadIdx.x; i < C * 4; i += blockDim.x)
      if (cnt[i >> 2][i & 3]) atomicAdd(&confusion[i], (unsigned long long)cnt[i >> 2][i & 3]);
}

}  // namespace mmcm
