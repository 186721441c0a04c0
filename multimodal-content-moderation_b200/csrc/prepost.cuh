// SURVEY 8(f) rows either side of the forward: the callers' pre- and post-processing.
//
//   preprocess_u8_kernel : ToTensor + Normalize of the eval transform (R/src/data/dataset.py:106-111) for an already
//                          resized / cropped uint8 HWC image: out[b,c,y,x] = (u8/255 - mean[c]) / std[c]  (fp32 CHW).
//                          Same operation order as torchvision (div by 255, sub, div) -> bit-identical pixel_values,
//                          with 4x less host->device traffic than shipping fp32 pixels.
//   resize_crop_u8_kernel: Resize(size, antialias=True) + CenterCrop(size) of the same transform on decoded uint8 RGB
//                          images of different sizes -- Pillow's two-pass fixed-point bilinear resampler
//                          (Image.resize as torchvision calls it), bit-identical, only the cropped region computed.
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

// ------------------------------------------------------------------------------------------------
// Resize + CenterCrop (R/src/data/dataset.py:106-108).  torchvision resizes a PIL image with Image.resize(BILINEAR):
// Pillow's resampler (src/libImaging/Resample.c) computes, per output index, a tap window and double-precision
// triangle weights (precompute_coeffs), turns them into 22-bit fixed point (normalize_coeffs_8bpc) and runs a
// horizontal pass into a uint8 image followed by a vertical pass, each pixel = clip8((2^21 + sum p*k) >> 22).
// The kernel below evaluates the same expressions in the same order (IEEE double ops without contraction, integer
// accumulation), so the crop equals Pillow's byte for byte; oracle/prepost_oracle.py restates it in numpy and the
// tests pin both against Pillow itself.
// ------------------------------------------------------------------------------------------------
constexpr int RESIZE_PRECISION_BITS = 32 - 8 - 2;

struct ResizeGeom {
  int new_h, new_w;   // size after T.Resize(size): shorter side -> size, longer = int(size * long / short)
  int top, left;      // T.CenterCrop origin: int(round((new - size) / 2.0)), Python's round-half-to-even
  int ksize_h, ksize_v;
  double scale_h, scale_v;   // source pixels per resized pixel along the width / the height
};

__host__ __device__ inline int resize_half_even(int d) {   // round(d / 2.0) for d >= 0
  const int q = d >> 1;
  return (d & 1) ? ((q & 1) ? q + 1 : q) : q;
}
__host__ __device__ inline int resize_ksize(double scale) {
  const double support = scale < 1.0 ? 1.0 : scale;      // bilinear support 1.0 * filterscale
  return (int)ceil(support) * 2 + 1;
}
__host__ __device__ inline void resize_geometry(int h, int w, int size, ResizeGeom& g) {
  const int shrt = w <= h ? w : h, lng = w <= h ? h : w;
  const int new_long = (int)((double)((long long)size * lng) / (double)shrt);
  g.new_w = w <= h ? size : new_long;
  g.new_h = w <= h ? new_long : size;
  g.top = resize_half_even(g.new_h - size);
  g.left = resize_half_even(g.new_w - size);
  g.scale_h = (double)w / (double)g.new_w;
  g.scale_v = (double)h / (double)g.new_h;
  g.ksize_h = resize_ksize(g.scale_h);
  g.ksize_v = resize_ksize(g.scale_v);
}
// source rows one strip of `rows` output rows can touch (upper bound used to size the shared-memory tile)
__host__ __device__ inline int resize_strip_rows(double scale, int rows) {
  const double support = scale < 1.0 ? 1.0 : scale;
  return (int)ceil((rows - 1) * scale + 2.0 * support) + 2;
}

// precompute_coeffs + normalize_coeffs_8bpc for output index xx of an axis in_size -> out_size
__device__ inline void resize_taps(const int in_size, const double scale, const int xx, int& first, int& count,
                                   int* __restrict__ k, const int kpitch) {
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    const double v = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    ww = __dadd_rn(ww, v < 1.0 ? __dsub_rn(1.0, v) : 0.0);
  }
  for (int x = 0; x < kpitch; ++x) {
    int c = 0;
    if (x < n) {
      const double v = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
      double w = v < 1.0 ? __dsub_rn(1.0, v) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      c = __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, (double)(1 << RESIZE_PRECISION_BITS))));
    }
    k[x] = c;
  }
  first = xmin;
  count = n;
}

__device__ __forceinline__ uint8_t resize_clip8(const int acc) {
  const int v = acc >> RESIZE_PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid (strips, B); one CTA = `rows` output rows of one image.  Shared memory: tap tables of the `size` cropped
// columns and of the strip's rows, then the horizontally resampled uint8 tile of the source rows the strip needs.
__global__ void __launch_bounds__(256)
resize_crop_u8_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ offsets,
                      const int* __restrict__ heights, const int* __restrict__ widths, const int size, const int rows,
                      const int kpitch_h, const int kpitch_v, const int max_src_rows, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  int* hb = reinterpret_cast<int*>(rs_smem);             // [size][2]   first source column, tap count
  int* kh = hb + 2 * size;                               // [size][kpitch_h]
  int* vb = kh + size * kpitch_h;                        // [rows][2]
  int* kv = vb + 2 * rows;                               // [rows][kpitch_v]
  uint8_t* tile = reinterpret_cast<uint8_t*>(kv + rows * kpitch_v);   // [max_src_rows][size][3]
  pdl_wait();
  const int b = blockIdx.y;
  const int h = heights[b], w = widths[b];
  const uint8_t* img = src + offsets[b];
  ResizeGeom g;
  resize_geometry(h, w, size, g);
  const int y0 = blockIdx.x * rows;
  const int ny = min(rows, size - y0);
  for (int t = threadIdx.x; t < size + ny; t += blockDim.x) {
    if (t < size) resize_taps(w, g.scale_h, g.left + t, hb[2 * t], hb[2 * t + 1], kh + t * kpitch_h, kpitch_h);
    else {
      const int r = t - size;
      resize_taps(h, g.scale_v, g.top + y0 + r, vb[2 * r], vb[2 * r + 1], kv + r * kpitch_v, kpitch_v);
    }
  }
  __syncthreads();
  const int r0 = vb[0];
  const int nrows = vb[2 * (ny - 1)] + vb[2 * (ny - 1) + 1] - r0;
  if (nrows > max_src_rows) __trap();                    // the host sized the tile from the same geometry
  // horizontal pass (ImagingResampleHorizontal_8bpc) for the source rows [r0, r0 + nrows) and the cropped columns
  for (int i = threadIdx.x; i < nrows * size; i += blockDim.x) {
    const int r = i / size, x = i - r * size;
    const int first = hb[2 * x], n = hb[2 * x + 1];
    const int* k = kh + x * kpitch_h;
    const uint8_t* p = img + ((size_t)(r0 + r) * w + first) * 3;
    int a0 = 1 << (RESIZE_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < n; ++j) {
      const int c = k[j];
      a0 += (int)p[3 * j] * c;
      a1 += (int)p[3 * j + 1] * c;
      a2 += (int)p[3 * j + 2] * c;
    }
    uint8_t* q = tile + (size_t)i * 3;
    q[0] = resize_clip8(a0); q[1] = resize_clip8(a1); q[2] = resize_clip8(a2);
  }
  __syncthreads();
  // vertical pass (ImagingResampleVertical_8bpc) over the uint8 tile
  for (int i = threadIdx.x; i < ny * size; i += blockDim.x) {
    const int r = i / size, x = i - r * size;
    const int first = vb[2 * r] - r0, n = vb[2 * r + 1];
    const int* k = kv + r * kpitch_v;
    const uint8_t* p = tile + ((size_t)first * size + x) * 3;
    int a0 = 1 << (RESIZE_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < n; ++j) {
      const int c = k[j];
      const uint8_t* pj = p + (size_t)j * size * 3;
      a0 += (int)pj[0] * c;
      a1 += (int)pj[1] * c;
      a2 += (int)pj[2] * c;
    }
    uint8_t* q = out + (((size_t)b * size + (y0 + r)) * size + x) * 3;
    q[0] = resize_clip8(a0); q[1] = resize_clip8(a1); q[2] = resize_clip8(a2);
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
