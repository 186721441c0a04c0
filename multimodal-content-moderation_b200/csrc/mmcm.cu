// libmmcm.so -- the C ABI declared in include/mmcm.h and the host-side engine behind it.
//
// The engine owns (a) a repacked copy of the reference's state dict (bf16 GEMM operands with Q|K|V concatenated
// and dh^-1/2 folded into Q; fp32 norms, biases, embeddings and heads) and (b) a per-tower activation arena sized for
// one micro-batch.  mmcm_forward enqueues, per micro-batch, the text tower on one stream and the vision tower on a
// second one, joins them on the caller's stream and runs the fused head kernel over the whole batch.
//
// Reference call stack this replaces: R/src/models/fusion.py:157-216 / R/src/models/multitask.py:156-207 and, below
// them, HF/models/clip/modeling_clip.py (CLIPTextTransformer :531-589, CLIPVisionTransformer :667-691, encoder layer
// :363-384) and HF/models/siglip/modeling_siglip.py (:489-527, :604-649).
//
// There is NO CPU path in this file: without a CUDA device every entry point fails with MMCM_ECUDA.
#include "../../include/mmcm.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "attention.cuh"
#include "attention_ring.cuh"
#include "attention_tc.cuh"
#include "common.cuh"
#include "gemm2_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "heads.cuh"
#include "prepost.cuh"
#include "rowwise.cuh"
#include "tokenizer.h"

using namespace mmcm;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(expr)                                                                                            \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return fail(MMCM_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CKR(expr)                 \
  do {                            \
    int _r = (expr);              \
    if (_r != MMCM_OK) return _r; \
  } while (0)

static int g_num_sms = kNumSMs;

// Launch options live in the handle (mmcm_set_option); every entry point that launches kernels installs its handle's
// set for the calling thread, so two handles driven from two threads never see each other's settings.  The
// stand-alone kernels (mmcm_gemm_bf16, mmcm_attention, ...) use the defaults, which mmcm_set_option(NULL, ...) edits.
struct LaunchOpts {
  bool pdl = true;            // programmatic dependent launch on every kernel
  bool tma_epilogue = true;   // TMA tile-store / reduce-add epilogue of the pair GEMM
  int attention_impl = 0;     // 0 auto, 1 mma.sync, 2 tcgen05 wherever T <= 256, 3 TMA-ring mma.sync for 32 < T <= 80
  int attention_ring = 1;     // auto: 77-token text on the TMA-ring kernel (0 = the cp.async kernel of round 1)
  int pair_limit = 0;         // > 0: cap the CTA pairs a GEMM launch may occupy (SM partitioning between the towers)
  bool narrow_tiles = true;   // GEMMs with one row block (M <= 256) use 64 / 128-column tiles
};
static LaunchOpts g_default_opts;
static thread_local const LaunchOpts* t_opts = &g_default_opts;
struct OptsScope {
  const LaunchOpts* prev;
  explicit OptsScope(const LaunchOpts* o) : prev(t_opts) { t_opts = o; }
  ~OptsScope() { t_opts = prev; }
};

// cudaFuncSetAttribute is per device: remember which devices a given kernel instantiation was configured on
struct AttrOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// ------------------------------------------------------------------------------------------------ launches
// Every kernel goes through launch_k: cudaLaunchKernelEx with programmatic stream serialization (PDL), see common.cuh.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = t_opts->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------ TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_mu;

static int ensure_driver() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_encode) return MMCM_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
    return fail(MMCM_ECUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) g_num_sms = n;
  }
  return MMCM_OK;
}

// K-major bf16 matrix [rows, K] -> 2D map with a (64 x box_rows) box, 128-byte swizzle; out-of-range rows read as 0.
static int make_tmap(CUtensorMap* m, const void* ptr, int64_t rows, int64_t K, int box_rows, bool weight) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        weight ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMCM_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld", (int)r,
                                     (long long)rows, (long long)K);
  return MMCM_OK;
}

struct TmapKey {
  const void* p;
  int64_t rows, K;
  int box;   // operand maps: box rows; output maps: -1 (bf16 64x32 box) / -2 (fp32 32x32 box) / -3 (bf16 32x32 box, 64B swizzle)
  bool operator<(const TmapKey& o) const { return std::tie(p, rows, K, box) < std::tie(o.p, o.rows, o.K, o.box); }
};

// output tile map for the TMA-store epilogue: row-major [rows, ld] bf16 or fp32, box = 128 bytes x 32 rows, 128B swizzle
static int make_out_tmap(CUtensorMap* m, const void* ptr, int64_t rows, int64_t ld, int kind) {
  const bool f32 = kind == -2;
  const int esz = f32 ? 4 : 2;
  const int inner_bytes = kind == -3 ? 64 : 128;
  cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(inner_bytes / esz), 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        kind == -3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMCM_ECUDA, "cuTensorMapEncodeTiled(out) failed (%d) rows=%lld ld=%lld", (int)r,
                                     (long long)rows, (long long)ld);
  return MMCM_OK;
}
static std::map<TmapKey, CUtensorMap> g_tmaps;

static int get_tmap(CUtensorMap* out, const void* ptr, int64_t rows, int64_t K, int box, bool weight) {
  std::lock_guard<std::mutex> lk(g_mu);
  TmapKey k{ptr, rows, K, box};
  auto it = g_tmaps.find(k);
  if (it == g_tmaps.end()) {
    CUtensorMap m;
    if (box < 0) CKR(make_out_tmap(&m, ptr, rows, K, box));
    else CKR(make_tmap(&m, ptr, rows, K, box, weight));
    if (g_tmaps.size() > 8192) g_tmaps.clear();
    it = g_tmaps.emplace(k, m).first;
  }
  *out = it->second;
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ tile scheduler state
// One {next tile, finished CTAs} pair per stream: launches on a stream are serialised and every launch leaves the
// pair at {0, 0} (the last CTA re-arms it), so no memset is needed between GEMMs.
static std::map<std::pair<int, cudaStream_t>, int*> g_sched;
static int get_sched(cudaStream_t st, int** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto key = std::make_pair(dev, st);
  auto it = g_sched.find(key);
  if (it == g_sched.end()) {
    int* p = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&p), 2 * sizeof(int)));
    CK(cudaMemset(p, 0, 2 * sizeof(int)));
    it = g_sched.emplace(key, p).first;
  }
  *out = it->second;
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ launch counting
struct LaunchStats {
  int64_t launches = 0;
  bool time_gemms = false;
  struct GemmRec { cudaEvent_t e0, e1; double flops_per_row; int m_host; const int* m_dev; int epi, N, K; };
  std::vector<GemmRec> gemm_events;      // per-GEMM CUDA events (time_gemms) + what is needed to count EXECUTED FLOPs
  int text_chunk = 0;                    // packed text chunks of this forward (each owns one device row counter)
};

// ------------------------------------------------------------------------------------------------ GEMM launcher
template <int BN, int EPI>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& ep, int M, int N, int K,
                     cudaStream_t st) {
  using C = GemmCfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, EPI>;
  static AttrOnce once;
  if (once.need()) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = ((M + C::BLOCK_M - 1) / C::BLOCK_M) * (N / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  int* sched = nullptr;
  CKR(get_sched(st, &sched));
  CK(launch_k(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, ep, M, N, K, sched));
  CK(cudaGetLastError());
  return MMCM_OK;
}

template <int BN, int EPI>
static int launch_tc_pair(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& ep, int M, int N, int K,
                          cudaStream_t st) {
  using C = Gemm2Cfg<BN, EPI>;
  auto kern = gemm2_tcgen05_kernel<BN, EPI>;
  static AttrOnce once;
  if (once.need()) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = ((M + C::BLOCK_M - 1) / C::BLOCK_M) * (N / BN) * (ep.ksplit > 1 ? ep.ksplit : 1);
  int pairs = g_num_sms / 2;
  if (t_opts->pair_limit > 0 && t_opts->pair_limit < pairs) pairs = t_opts->pair_limit;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  // TMA-store epilogue: bf16 tile stores; fp32 residual GEMMs as an L2 reduce-add (needs resid == out or no resid);
  // rows / pitches must keep the 16-byte global alignment TMA wants.  Otherwise the per-thread store path is used.
  constexpr bool f32 = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_RESID_STATS);
  const bool aligned = (ep.ldo * (f32 ? 4 : 2)) % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0;
  int tma_out = t_opts->tma_epilogue && EPI != EPI_PATCH_F32 && aligned;
  if (EPI == EPI_BIAS_RESID_F32 && ep.resid && ep.resid != ep.out) tma_out = 0;
  CUtensorMap tc = ta, td = ta;
  if (EPI == EPI_RESID_STATS) {   // the x tile travels in and out by TMA, so does its bf16 copy: no per-thread path
    if (!aligned || !ep.xb || !ep.stats || (ep.ldo * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(ep.xb) & 15) != 0)
      return fail(MMCM_EINVAL, "gemm: EPI_RESID_STATS needs 16-byte aligned out / xb rows and a stats buffer");
    tma_out = 1;
    CKR(get_tmap(&td, ep.xb, M, ep.ldo, -3, false));
  }
  if ((EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_ACT_BF16) &&
      (!ep.stats || !ep.bias || ep.ln_slabs < 1 || ep.ln_slabs > 8 || ep.ln_slabs * LN_SLAB != K))
    return fail(MMCM_EINVAL, "gemm: EPI_LNFOLD needs stats, bias and K == 128 * slabs <= 1024");
  if (ep.ksplit > 1) {   // partial sums go to planes of a scratch buffer: plain TMA stores, no per-thread path
    if (EPI != EPI_BIAS_RESID_F32 || !tma_out || ep.resid || (K / C::BLOCK_K) % ep.ksplit != 0 || ep.part_rows < M)
      return fail(MMCM_EINVAL, "gemm: split-K needs the fp32 TMA-store epilogue, no residual and K / 64 %% ksplit == 0");
    CKR(get_tmap(&tc, ep.out, (int64_t)ep.ksplit * ep.part_rows, ep.ldo, -2, false));
  } else if (tma_out) CKR(get_tmap(&tc, ep.out, M, ep.ldo, f32 ? -2 : (BN == 64 ? -3 : -1), false));
  if (BN == 64 && !f32 && EPI != EPI_PATCH_F32 && !tma_out)
    return fail(MMCM_EINVAL, "gemm: 64-column bf16 tiles need the TMA-store epilogue");
  CK(launch_k(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, tc, td, ep, M, N, K, tma_out));   // __cluster_dims__(2,1,1)
  CK(cudaGetLastError());
  return MMCM_OK;
}

template <int EPI>
static int launch_gemm_epi(const bf16* A, const bf16* W, int M, int N, int K, const EpiParams& ep, int impl,
                           cudaStream_t st) {
  constexpr bool kPairOnly = EPI >= EPI_RESID_STATS;   // LN-fold epilogues exist in the CTA-pair kernel only
  if constexpr (kPairOnly) {
    if (impl != 0 || N % 256 != 0)
      return fail(MMCM_EINVAL, "gemm: epilogue %d needs gemm_impl 0 and N %% 256 == 0 (got impl %d, N %d)", EPI, impl, N);
  } else {
    if (impl == 1) {
      dim3 grid(N / 32, (M + 31) / 32);
      CK(launch_k(gemm_simt_kernel<EPI>, dim3(grid), dim3(256), 0, st, A, W, ep, M, N, K));
      CK(cudaGetLastError());
      return MMCM_OK;
    }
  }
  CKR(ensure_driver());
  int BN = (N % 256 == 0) ? 256 : 128;
  // One row block (M <= 256: the online B = 1 path, the pooled-rows last layer): a 256-wide tile leaves N / 256 = 2-3
  // CTA pairs streaming the whole weight matrix while 140 SMs idle -- fc2 took 20 us at B = 1, bound by what two SMs
  // pull from L2 (tools/b1_launches.py).  Narrow tiles spread the weight stream: 64 columns (one 32-column chunk per
  // epilogue warp; 128 for bf16 outputs without the TMA-store epilogue).  The per-element k order is unchanged.
  if (impl == 0 && M <= 256 && t_opts->narrow_tiles) {
    if constexpr (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_PATCH_F32) { if (N % 64 == 0) BN = 64; }
    else if constexpr (!kPairOnly) {
      const bool tma_ok = t_opts->tma_epilogue && (ep.ldo * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0;
      if (N % 64 == 0 && tma_ok) BN = 64;          // the 64-column bf16 epilogue exists as a TMA store only
      else if (N % 128 == 0) BN = 128;
    }
  }
  CUtensorMap ta, tb;
  CKR(get_tmap(&ta, A, M, K, 128, false));
  if (impl == 0) {  // CTA-pair kernel: every CTA stages half of the B tile
    CKR(get_tmap(&tb, W, N, K, BN / 2, true));
    if (BN == 256) return launch_tc_pair<256, EPI>(ta, tb, ep, M, N, K, st);
    if constexpr (!kPairOnly)
      if (BN == 64) return launch_tc_pair<64, EPI>(ta, tb, ep, M, N, K, st);
    if constexpr (!kPairOnly) return launch_tc_pair<128, EPI>(ta, tb, ep, M, N, K, st);
  }
  if constexpr (!kPairOnly) {
    CKR(get_tmap(&tb, W, N, K, BN, true));
    if (BN == 256) return launch_tc<256, EPI>(ta, tb, ep, M, N, K, st);
    return launch_tc<128, EPI>(ta, tb, ep, M, N, K, st);
  }
  return fail(MMCM_EINVAL, "gemm: unsupported epilogue / implementation combination");
}

static int launch_gemm(const bf16* A, const bf16* W, int M, int N, int K, int epi, const EpiParams& ep, int impl,
                       cudaStream_t st, LaunchStats* stats) {
  if (M <= 0) return MMCM_OK;
  if (N % 128 != 0 || K % 64 != 0 || N <= 0 || K <= 0)
    return fail(MMCM_EINVAL, "gemm: need N %% 128 == 0 and K %% 64 == 0 (got M=%d N=%d K=%d)", M, N, K);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (stats && stats->time_gemms) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
  }
  int r;
  switch (epi) {
    case EPI_BIAS_BF16: r = launch_gemm_epi<EPI_BIAS_BF16>(A, W, M, N, K, ep, impl, st); break;
    case EPI_BIAS_ACT_BF16: r = launch_gemm_epi<EPI_BIAS_ACT_BF16>(A, W, M, N, K, ep, impl, st); break;
    case EPI_BIAS_RESID_F32: r = launch_gemm_epi<EPI_BIAS_RESID_F32>(A, W, M, N, K, ep, impl, st); break;
    case EPI_PATCH_F32: r = launch_gemm_epi<EPI_PATCH_F32>(A, W, M, N, K, ep, impl, st); break;
    case EPI_RESID_STATS: r = launch_gemm_epi<EPI_RESID_STATS>(A, W, M, N, K, ep, impl, st); break;
    case EPI_LNFOLD_BF16: r = launch_gemm_epi<EPI_LNFOLD_BF16>(A, W, M, N, K, ep, impl, st); break;
    case EPI_LNFOLD_ACT_BF16: r = launch_gemm_epi<EPI_LNFOLD_ACT_BF16>(A, W, M, N, K, ep, impl, st); break;
    default: return fail(MMCM_EINVAL, "gemm: unknown epilogue %d", epi);
  }
  CKR(r);
  if (stats) {
    stats->launches++;
    if (stats->time_gemms) {
      CK(cudaEventRecord(e1, st));
      stats->gemm_events.push_back(LaunchStats::GemmRec{e0, e1, 2.0 * (double)N * K, M, ep.m_dev, epi, N, K});
    }
  }
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ other launchers
// split-K partial sums waiting to be folded into a residual buffer by the next LayerNorm that reads it
struct PendingParts {
  const float* part = nullptr;
  int ksplit = 0, part_rows = 0;
};

static int launch_layernorm(const float* x, const float* g, const float* b, float eps, int rows, int D,
                            const int* gather, bf16* out_bf16, float* out_f32, cudaStream_t st, LaunchStats* stats,
                            const int* rows_dev = nullptr, PendingParts* pend = nullptr) {
  if (rows <= 0) return MMCM_OK;
  const int blocks = (rows + 7) / 8;  // 8 warps (rows) per 256-thread block
  const float* part = (pend && pend->part) ? pend->part : nullptr;
  const int ks = part ? pend->ksplit : 0, pr = part ? pend->part_rows : 0;
  float* xrw = part ? const_cast<float*>(x) : nullptr;
  if (pend) *pend = PendingParts();
  if (D == 512) CK(launch_k(layernorm_kernel<512>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, gather, out_bf16, out_f32, rows_dev, part, ks, pr, xrw));
  else if (D == 768) CK(launch_k(layernorm_kernel<768>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, gather, out_bf16, out_f32, rows_dev, part, ks, pr, xrw));
  else if (D == 1024) CK(launch_k(layernorm_kernel<1024>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, gather, out_bf16, out_f32, rows_dev, part, ks, pr, xrw));
  else return fail(MMCM_EINVAL, "layernorm: unsupported width %d (512, 768, 1024)", D);
  CK(cudaGetLastError());
  if (stats) stats->launches++;
  return MMCM_OK;
}

// LN fold entry of a tower: optional in-place LayerNorm (CLIP pre_layrnorm), then bf16 copy + slab statistics
static int launch_prep_rows(float* x, const float* g, const float* b, float eps, int rows, int D, bf16* xb, float2* stats,
                            int pitch, cudaStream_t st, LaunchStats* S, const int* rows_dev = nullptr) {
  if (rows <= 0) return MMCM_OK;
  const int blocks = (rows + 7) / 8;
  if (D == 512) CK(launch_k(prep_rows_kernel<512>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, rows_dev, xb, stats, pitch));
  else if (D == 768) CK(launch_k(prep_rows_kernel<768>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, rows_dev, xb, stats, pitch));
  else if (D == 1024) CK(launch_k(prep_rows_kernel<1024>, dim3(blocks), dim3(256), 0, st, x, g, b, eps, rows, rows_dev, xb, stats, pitch));
  else return fail(MMCM_EINVAL, "prep_rows: unsupported width %d (512, 768, 1024)", D);
  CK(cudaGetLastError());
  if (S) S->launches++;
  return MMCM_OK;
}

template <int TPAD, int QW>
static int launch_att(const bf16* qkv, bf16* out, const uint8_t* kvalid, int B, int T, int heads, int causal,
                      cudaStream_t st, const int* seq_start, const int* seq_len) {
  using C = AttCfg<TPAD, QW>;
  auto kern = attention_kernel<TPAD, QW>;
  static AttrOnce once;
  static int ctas_per_sm = 1;
  if (once.need()) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int n = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, QW * 32, C::SMEM_BYTES));
    ctas_per_sm = n > 0 ? n : 1;
    if (getenv("MMCM_DEBUG")) fprintf(stderr, "[mmcm] attention<%d,%d>: occupancy %d CTAs/SM\n", TPAD, QW, n);
  }
  const int total = B * heads * C::QBLOCKS;
  int grid = g_num_sms * ctas_per_sm;
  if (grid > total) grid = total;
  CK(launch_k(kern, dim3(grid), dim3(QW * 32), C::SMEM_BYTES, st, qkv, out, kvalid, seq_start, seq_len, T, heads * ATT_DH,
              causal, T, B, heads));
  return MMCM_OK;
}

// TMA-ring variant of the mma.sync kernel (attention_ring.cuh): one persistent CTA per SM
template <int TPAD, int NG, int NSTAGES>
static int launch_att_ring(const bf16* qkv, bf16* out, const uint8_t* kvalid, int B, int T, int heads, int causal,
                           cudaStream_t st, const int* seq_start, const int* seq_len) {
  using C = AttRingCfg<TPAD, NG, NSTAGES>;
  CKR(ensure_driver());
  auto kern = attention_ring_kernel<TPAD, NG, NSTAGES>;
  static AttrOnce once;
  if (once.need()) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const int D = heads * ATT_DH;
  AttRingMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int k = 0; k < C::QW; ++k)
    CKR(get_tmap(&maps.m[k], qkv, (int64_t)B * T, 3 * D, seq_start ? 16 * (k + 1) : std::min(T, 16 * (k + 1)), false));
  const int total = B * heads;
  int grid = g_num_sms;
  if (grid > total) grid = total;
  CK(launch_k(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, maps, out, kvalid, seq_start, seq_len, T, D, causal, T, B,
              heads));
  return MMCM_OK;
}

static long long* g_gemm_trace = nullptr;   // dev tool, see mmcm_debug_set_gemm_trace
// 0 = auto: tcgen05 kernel when at least two samples share a 128-row tile (T <= 64: measured 80 vs 87 us for the
// 50-token vision tower) and for 128 < T <= 256 (SigLIP vision), mma.sync kernel otherwise (77-token text: 90 vs
// 98 us; packed text stays on it too although tcgen05 would be 2 % faster end to end, so that packed and dense text
// run the SAME attention kernel and stay bit-identical); 1 = always mma.sync; 2 = tcgen05 whenever T <= 256
// (LaunchOpts::attention_impl)

template <int KMAX>
static int launch_attention_tc(const bf16* qkv, const uint8_t* kvalid, int B, int T, int heads, int causal, bf16* out,
                               cudaStream_t st, const int* seq_start, const int* seq_len) {
  using C = AtcCfg<KMAX>;
  CKR(ensure_driver());
  auto kern = attention_tc_kernel<KMAX>;
  static AttrOnce once;
  if (once.need()) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  // KMAX = 128: 49 KB of shared memory and 128 of the SM's 512 TMEM columns per CTA -> 4 resident CTAs per SM
  // (KMAX = 256: 97 KB, 256 columns -> 2).  The occupancy calculator answers 1 for a kernel that allocates TMEM; the
  // hardware does co-schedule them: 247 -> 142 -> 98 us for 1 / 2 / 4 CTAs per SM (tools/attn_trace.py).
  int ctas_per_sm = C::CTAS_PER_SM;
  if (getenv("MMCM_ATC_CTAS")) ctas_per_sm = atoi(getenv("MMCM_ATC_CTAS"));
  const int D = heads * ATT_DH;
  const bool longseq = !seq_start && T > 128;
  const int q_box = (seq_start || longseq) ? 128 : T;     // one TMA box per sample slot
  const int kv_box = seq_start ? 128 : T;
  CUtensorMap tq, tkv;
  CKR(get_tmap(&tq, qkv, (int64_t)B * T, 3 * D, q_box, false));
  CKR(get_tmap(&tkv, qkv, (int64_t)B * T, 3 * D, kv_box, false));
  const int slot = (seq_start || longseq) ? 128 : (T <= 16 ? 16 : (T <= 32 ? 32 : (T <= 64 ? 64 : 128)));
  const int G = 128 / slot;
  const int RT = longseq ? (T + 127) / 128 : 1;
  const int total = ((B + G - 1) / G) * heads * RT;
  int grid = g_num_sms * ctas_per_sm;
  if (grid > total) grid = total;
  CK(launch_k(kern, dim3(grid), dim3(ATC_THREADS), C::SMEM_BYTES, st, tq, tkv, out, kvalid, seq_start, seq_len, T, D, causal, B,
              heads, q_box, kv_box, g_gemm_trace));
  return MMCM_OK;
}

static int launch_attention(const bf16* qkv, const uint8_t* kvalid, int B, int T, int heads, int causal, bf16* out,
                            cudaStream_t st, LaunchStats* stats, const int* seq_start = nullptr,
                            const int* seq_len = nullptr) {
  if (B <= 0) return MMCM_OK;
  const bool tc128 = T <= 128 && (t_opts->attention_impl == 2 || (t_opts->attention_impl == 0 && T <= 64 && !seq_start));
  const bool tc256 = T > 128 && T <= 256 && !seq_start && t_opts->attention_impl != 1;
  // TMA-ring mma.sync kernel (32 < T <= 80): auto (with attention_ring = 1) and attention_impl 3; the 77-token text
  // tower runs it dense and packed alike (so the two stay bit-identical), the 50-token vision tower dense
  const bool ring_ok = T > 32 && T <= 80 &&
                       (t_opts->attention_impl == 3 || (t_opts->attention_impl == 0 && t_opts->attention_ring));
  const int ring = ring_ok ? (T <= 64 ? 64 : 80) : 0;
  if (ring) {
    // consumer groups x ring stages, measured per 1024 samples (ncu, profiles/r02_attention_ring_ncu.txt):
    //   77-token causal text  <80, NG, 7>: NG = 2 / 3 / 4 -> 72.5 / 59.1 / 64.0 us (3 groups compute, 4 stages in flight,
    //                                      16 warps -> 128 registers, no spills)
    //   50-token vision       <64, NG, S>: <3, 9> / <4, 9> / <5, 8> / <6, 9> -> 90.5 / 84.3 / 78.6 / 83 us (every warp multiplies
    //                                      all keys: more warps help until 25 warps cap the registers at 72 and leave
    //                                      three stages in flight; the tcgen05 kernel: 82.8 us)
    if (ring == 64) CKR((launch_att_ring<64, 5, 8>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len)));
    else CKR((launch_att_ring<80, 3, 7>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len)));
    if (stats) stats->launches++;
    return MMCM_OK;
  }
  if (tc128 || tc256) {
    if (tc128) CKR(launch_attention_tc<128>(qkv, kvalid, B, T, heads, causal, out, st, seq_start, seq_len));
    else CKR(launch_attention_tc<256>(qkv, kvalid, B, T, heads, causal, out, st, seq_start, seq_len));
    if (stats) stats->launches++;
    return MMCM_OK;
  }
  int r;
  if (T <= 16) r = launch_att<16, 1>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 32) r = launch_att<32, 2>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 64) r = launch_att<64, 4>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 80) r = launch_att<80, 5>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 128) r = launch_att<128, 8>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 208) r = launch_att<208, 4>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else if (T <= 256) r = launch_att<256, 8>(qkv, out, kvalid, B, T, heads, causal, st, seq_start, seq_len);
  else return fail(MMCM_EINVAL, "attention: sequence length %d > 256 is not supported", T);
  CKR(r);
  if (stats) stats->launches++;
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ weights
struct Part {          // one rectangular piece of a state-dict tensor and where it goes
  int64_t src_off;     // element offset into the source tensor
  int64_t rows, cols;  // piece shape
  int64_t src_pitch;   // elements between consecutive source rows
  void* dst;
  int64_t dst_pitch;
  int kind;            // 0 = fp32, 1 = bf16, 2 = fp32 into the fold staging block (dst = float offset into it)
  float scale;
};
struct Slot {
  int64_t numel = 0;
  std::vector<Part> parts;
  bool loaded = false;
  bool fold = false;   // input of the LN fold (a staged Linear or a layer norm): loading it re-arms the fold
};

__global__ void copy2d_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t rows, int64_t cols,
                              int64_t src_pitch, int64_t dst_pitch, int kind, float scale) {
  pdl_wait();
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float v = src[r * src_pitch + c] * scale;
    if (kind == 1) reinterpret_cast<bf16*>(dst)[r * dst_pitch + c] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[r * dst_pitch + c] = v;
  }
}

// The layer norms are folded into the Linears that consume them (fold_ln_kernel): wqkv / w1 hold bf16(W * gamma) with
// centred rows, bqkv / b1 the folded biases b + W beta.
struct LayerW {
  bf16 *wqkv, *wo, *w1, *w2;
  float *bqkv, *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b;
};
struct FoldJob {   // one fold_ln_kernel launch: staged fp32 W [N,K] and b [N] -> bf16 W*gamma, colsum, folded bias
  int64_t w_off, b_off;
  const float *gamma, *beta;
  int N, K, q_rows;
  float q_scale;
  bf16* wout;
  float* bias_out;
};
struct TowerW {
  int D, H, L, F, act;
  float eps;
  std::vector<LayerW> layers;
};

struct Arena {  // activations of one tower for one micro-batch
  float* x = nullptr;    // fp32 residual stream [rows, D]
  bf16* h = nullptr;     // LayerNorm output     [rows, D]
  bf16* qkv = nullptr;   // [rows, 3D]
  bf16* att = nullptr;   // [rows, D]
  bf16* ff = nullptr;    // [rows, F]
  int* pool_row = nullptr;
  int *seq_start = nullptr, *seq_len = nullptr, *rows_dev = nullptr;   // packed variable-length text
  // last layer on the pooled rows only (one row per sample): residual, attention output, LN2 output, MLP hidden
  float* xp = nullptr;
  bf16 *attp = nullptr, *hp = nullptr, *ffp = nullptr;
  // LN fold: per-row (sum, M2) of every 128-column slab of x / xp, written by the producer of the rows
  float2 *stats = nullptr, *statsp = nullptr;
  // small forwards (B < 16): split-K partial sums of the residual GEMMs, kSplitPlanes planes of kSplitRows rows, and
  // which of x / xp still has to absorb them (the next LayerNorm on that buffer does)
  float* part = nullptr;
  PendingParts pend_x, pend_xp;
  int64_t rows = 0;
  int mb = 0;
};
constexpr int kSplitPlanes = 4, kSplitRows = 3072;   // >= 15 samples x 197 tokens (the largest forward that splits)

struct mmcm_handle_s {
  mmcm_config cfg;
  int device = 0;
  std::vector<void*> allocs;
  std::unordered_map<std::string, Slot> slots;
  bool finalized = false;
  // LN fold inputs: fp32 copies of the q/k/v/fc1 weights and biases, alive only between the first mmcm_load_weight
  // of such a tensor and the next mmcm_finalize_weights
  float* stage32 = nullptr;
  int64_t stage32_floats = 0;
  std::vector<FoldJob> fold_jobs;
  bool fold_pending = false;
  float* upload_buf = nullptr;   // host -> device bounce buffer of mmcm_load_weight (freed by finalize)
  int64_t upload_floats = 0;
  LaunchOpts opts;
  // every buffer of the repacked weight set, in allocation order (a pure function of cfg): the packed weight file
  struct WBuf { void* ptr; size_t bytes; };
  std::vector<WBuf> wbufs;
  bool recording_weights = false;

  TowerW text, vis;
  // text extras
  float *tok_emb = nullptr, *tpos_emb = nullptr, *tfin_g = nullptr, *tfin_b = nullptr;
  float *thead_w = nullptr, *thead_b = nullptr;  // siglip text head
  // vision extras
  bf16* wpatch = nullptr;
  float *bpatch = nullptr, *cls_emb = nullptr, *vpos_emb = nullptr, *pre_g = nullptr, *pre_b = nullptr,
        *post_g = nullptr, *post_b = nullptr;
  // siglip MAP head
  float *map_probe = nullptr, *map_inw = nullptr, *map_inb = nullptr, *map_q = nullptr;
  bf16 *map_wkv = nullptr, *map_wo = nullptr, *map_w1 = nullptr, *map_w2 = nullptr;
  float *map_bkv = nullptr, *map_bo = nullptr, *map_lng = nullptr, *map_lnb = nullptr, *map_b1 = nullptr,
        *map_b2 = nullptr;
  HeadWeights hw;

  // activations
  int mb_text = 0, mb_vis = 0;  // micro-batch capacities the arenas are sized for
  Arena at, av;
  bf16* im2col = nullptr;
  uint8_t* key_valid = nullptr;
  bf16 *map_kv = nullptr, *map_att = nullptr, *map_h = nullptr, *map_ff = nullptr;
  float* map_y = nullptr;
  int64_t cap_B = 0;
  float *pooled_t = nullptr, *pooled_v = nullptr, *feat_t = nullptr, *feat_v = nullptr;
  int last_B = 0, last_text_rows = 0, last_vis_rows = 0;
  // host-call staging
  int64_t host_cap = 0;
  int64_t* d_ids = nullptr;
  int64_t* d_mask = nullptr;
  float *d_px = nullptr, *d_tp = nullptr, *d_ip = nullptr, *d_logits = nullptr, *d_probs = nullptr;
  int host_S = 0;

  cudaStream_t s_text = nullptr, s_vis = nullptr, s_copy = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_text = nullptr, ev_vis = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_copy_end = nullptr, ev_end = nullptr;   // timed: what bounded the last host call
  bool h2d_bound = false;
  float last_h2d_share = 0.f;
  std::vector<cudaEvent_t> ev_chunk;
  // second input set + what mmcm_prefetch_host* put into it (double-buffered input pipeline across calls)
  int64_t alt_cap = 0;
  int64_t* a_ids = nullptr;
  int64_t* a_mask = nullptr;
  float *a_px = nullptr, *a_tp = nullptr, *a_ip = nullptr;
  std::vector<cudaEvent_t> ev_chunk_alt;
  cudaEvent_t ev_small = nullptr, ev_small_alt = nullptr;
  struct HostPlan { int ct = 0, cv = 0; std::vector<int> stages; };
  struct Prefetched {
    bool valid = false, u8 = false, need_vis = true;
    const void *ids = nullptr, *mask = nullptr, *px = nullptr, *tp = nullptr, *ip = nullptr;
    int B = 0, S = 0;
    HostPlan plan;
  };
  Prefetched pf;      // what the second set holds (copies possibly still in flight on s_copy)
  Prefetched ready;   // what the CURRENT set holds after a promotion (prefetched, not consumed yet)
  // options
  // CUDA graphs of whole forwards for small batches (launch-bound regime): key = (B, S, mask?, probs?)
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int warm = 0; };
  std::map<std::tuple<int, int, int, int>, GraphEntry> graphs;
  int opt_graph_max_batch = 0;   // off by default: small batches are bound by the GPU-side kernel chain, not the host
  int opt_pairs_text = 0, opt_pairs_vis = 0;   // > 0: CTA pairs the text / vision GEMMs may occupy (two-stream SM split)
  int opt_host_chunk = 0;     // mmcm_forward_host*: samples per H2D pipeline stage of the vision tower (0 = ~200 MB)
  int opt_pooled_last = 1;   // last layer: out_proj / MLP / final LN only for the one row per sample that is pooled (exact)
  int opt_varlen_text = 1;   // CLIP text: keep only the rows up to the pooled (EOS) position -- exact, see rowwise.cuh
  int opt_streams = 2, opt_gemm_impl = 0, opt_micro_batch = 1024, opt_debug_feats = 0, opt_auto_chunk = 1;
  bool fold_forward = true;   // this forward runs the LN fold (opt_ln_fold and B >= kLnFoldMinBatch)
  bool split_forward = false; // this forward splits K in its residual GEMMs (B < kLnFoldMinBatch, see run_layers)
  int opt_skip_absent = 1;    // packed text: samples whose text cannot reach the logits keep one row (text_plan_kernel)
  int opt_split_k = 1;        // small forwards: split-K residual GEMMs, partials absorbed by the following LayerNorm
  int opt_head_cluster = 1;   // B <= 144: the head kernel runs as clusters of 8 CTAs per 8 samples (heads.cuh)
  int opt_ln_fold = 1;        // LayerNorm folded into the residual / consumer GEMMs (gemm_impl 0 only), else a separate pass
  int last_chunk_text = 0, last_chunk_vis = 0;
  LaunchStats stats;
};
typedef mmcm_handle_s Eng;

template <typename T>
static int dalloc(Eng* e, T** out, int64_t count) {
  void* p = nullptr;
  if (count <= 0) count = 1;
  cudaError_t err = cudaMalloc(&p, (size_t)count * sizeof(T));
  if (err != cudaSuccess) return fail(MMCM_ECUDA, "cudaMalloc(%lld bytes) failed: %s", (long long)(count * sizeof(T)),
                                      cudaGetErrorString(err));
  e->allocs.push_back(p);
  if (e->recording_weights) e->wbufs.push_back({p, (size_t)count * sizeof(T)});
  *out = reinterpret_cast<T*>(p);
  return MMCM_OK;
}
static void dfree(Eng* e, void* p) {
  if (!p) return;
  for (size_t i = 0; i < e->allocs.size(); ++i)
    if (e->allocs[i] == p) {
      e->allocs[i] = e->allocs.back();
      e->allocs.pop_back();
      break;
    }
  cudaFree(p);
}

static void reg(Eng* e, const std::string& key, int64_t numel, void* dst, int kind, float scale = 1.0f) {
  Slot& s = e->slots[key];
  s.numel = numel;
  s.parts.push_back(Part{0, 1, numel, numel, dst, numel, kind, scale});
}
static void reg2d(Eng* e, const std::string& key, int64_t numel, int64_t src_off, int64_t rows, int64_t cols,
                  int64_t src_pitch, void* dst, int64_t dst_pitch, int kind, float scale = 1.0f) {
  Slot& s = e->slots[key];
  s.numel = numel;
  s.parts.push_back(Part{src_off, rows, cols, src_pitch, dst, dst_pitch, kind, scale});
}

// fp32 staging of a fold input: `off` floats into e->stage32 (allocated lazily by mmcm_load_weight)
static void reg_staged(Eng* e, const std::string& key, int64_t numel, int64_t off) {
  Slot& s = e->slots[key];
  s.numel = numel;
  s.fold = true;
  s.parts.push_back(Part{0, 1, numel, numel, reinterpret_cast<void*>(static_cast<intptr_t>(off)), numel, 2, 1.0f});
}

static int setup_tower(Eng* e, TowerW& t, const std::string& prefix, int D, int H, int L, int F, int act, float eps) {
  t.D = D; t.H = H; t.L = L; t.F = F; t.act = act; t.eps = eps;
  if (D != H * ATT_DH) return fail(MMCM_EINVAL, "tower %s: hidden %d != heads %d * 64", prefix.c_str(), D, H);
  if (D % 128 != 0 || F % 128 != 0) return fail(MMCM_EINVAL, "tower %s: hidden/ffn must be multiples of 128", prefix.c_str());
  t.layers.resize(L);
  const float qs = 1.0f / sqrtf((float)ATT_DH);  // 0.125: exact power of two, folding it into Wq/bq is lossless
  for (int i = 0; i < L; ++i) {
    LayerW& w = t.layers[i];
    CKR(dalloc(e, &w.wqkv, (int64_t)3 * D * D));
    CKR(dalloc(e, &w.bqkv, 3 * D));
    CKR(dalloc(e, &w.wo, (int64_t)D * D));
    CKR(dalloc(e, &w.bo, D));
    CKR(dalloc(e, &w.w1, (int64_t)F * D));
    CKR(dalloc(e, &w.b1, F));
    CKR(dalloc(e, &w.w2, (int64_t)D * F));
    CKR(dalloc(e, &w.b2, D));
    CKR(dalloc(e, &w.ln1g, D)); CKR(dalloc(e, &w.ln1b, D));
    CKR(dalloc(e, &w.ln2g, D)); CKR(dalloc(e, &w.ln2b, D));
    const std::string p = prefix + "encoder.layers." + std::to_string(i) + ".";
    const int64_t dd = (int64_t)D * D;
    // q/k/v_proj and fc1 consume a LayerNorm: staged in fp32, folded with its gamma / beta at finalize
    const int64_t o_wqkv = e->stage32_floats, o_bqkv = o_wqkv + 3 * dd, o_w1 = o_bqkv + 3 * D, o_b1 = o_w1 + (int64_t)F * D;
    e->stage32_floats = o_b1 + F;
    reg_staged(e, p + "self_attn.q_proj.weight", dd, o_wqkv);
    reg_staged(e, p + "self_attn.k_proj.weight", dd, o_wqkv + dd);
    reg_staged(e, p + "self_attn.v_proj.weight", dd, o_wqkv + 2 * dd);
    reg_staged(e, p + "self_attn.q_proj.bias", D, o_bqkv);
    reg_staged(e, p + "self_attn.k_proj.bias", D, o_bqkv + D);
    reg_staged(e, p + "self_attn.v_proj.bias", D, o_bqkv + 2 * D);
    reg_staged(e, p + "mlp.fc1.weight", (int64_t)F * D, o_w1);
    reg_staged(e, p + "mlp.fc1.bias", F, o_b1);
    reg(e, p + "self_attn.out_proj.weight", dd, w.wo, 1);
    reg(e, p + "self_attn.out_proj.bias", D, w.bo, 0);
    reg(e, p + "layer_norm1.weight", D, w.ln1g, 0);
    reg(e, p + "layer_norm1.bias", D, w.ln1b, 0);
    reg(e, p + "layer_norm2.weight", D, w.ln2g, 0);
    reg(e, p + "layer_norm2.bias", D, w.ln2b, 0);
    for (const char* k : {"layer_norm1.weight", "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias"})
      e->slots[p + k].fold = true;
    reg(e, p + "mlp.fc2.weight", (int64_t)D * F, w.w2, 1);
    reg(e, p + "mlp.fc2.bias", D, w.b2, 0);
    // dh^-1/2 = 0.125 goes into the q rows (exact: a power of two)
    e->fold_jobs.push_back(FoldJob{o_wqkv, o_bqkv, w.ln1g, w.ln1b, 3 * D, D, D, qs, w.wqkv, w.bqkv});
    e->fold_jobs.push_back(FoldJob{o_w1, o_b1, w.ln2g, w.ln2b, F, D, 0, 1.0f, w.w1, w.b1});
  }
  return MMCM_OK;
}

static int setup_weights(Eng* e) {
  const mmcm_config& c = e->cfg;
  const bool clip = c.backend == MMCM_BACKEND_CLIP;
  const bool fusion = c.head == MMCM_HEAD_FUSION;
  const std::string tp = fusion ? "backbone.text_model." : "tower_txt.text_model.";
  const std::string vp = fusion ? "backbone.vision_model." : "tower_img.vision_model.";
  CKR(setup_tower(e, e->text, tp, c.text_hidden, c.text_heads, c.text_layers, c.text_ffn, c.text_act, c.text_eps));
  CKR(setup_tower(e, e->vis, vp, c.vis_hidden, c.vis_heads, c.vis_layers, c.vis_ffn, c.vis_act, c.vis_eps));
  const int Dt = c.text_hidden, Dv = c.vis_hidden;
  // text embeddings + final norm
  CKR(dalloc(e, &e->tok_emb, (int64_t)c.vocab * Dt));
  CKR(dalloc(e, &e->tpos_emb, (int64_t)c.max_pos * Dt));
  CKR(dalloc(e, &e->tfin_g, Dt)); CKR(dalloc(e, &e->tfin_b, Dt));
  reg(e, tp + "embeddings.token_embedding.weight", (int64_t)c.vocab * Dt, e->tok_emb, 0);
  reg(e, tp + "embeddings.position_embedding.weight", (int64_t)c.max_pos * Dt, e->tpos_emb, 0);
  reg(e, tp + "final_layer_norm.weight", Dt, e->tfin_g, 0);
  reg(e, tp + "final_layer_norm.bias", Dt, e->tfin_b, 0);
  // vision embeddings + norms
  const int G = c.image / c.patch, P = G * G, Kp = 3 * c.patch * c.patch;
  const int Tv = P + (clip ? 1 : 0);
  if (Kp % 64 != 0) return fail(MMCM_EINVAL, "patch %d: 3*patch^2 must be a multiple of 64", c.patch);
  CKR(dalloc(e, &e->wpatch, (int64_t)Dv * Kp));
  CKR(dalloc(e, &e->vpos_emb, (int64_t)Tv * Dv));
  CKR(dalloc(e, &e->post_g, Dv)); CKR(dalloc(e, &e->post_b, Dv));
  reg(e, vp + "embeddings.patch_embedding.weight", (int64_t)Dv * Kp, e->wpatch, 1);
  reg(e, vp + "embeddings.position_embedding.weight", (int64_t)Tv * Dv, e->vpos_emb, 0);
  reg(e, vp + "post_layernorm.weight", Dv, e->post_g, 0);
  reg(e, vp + "post_layernorm.bias", Dv, e->post_b, 0);
  if (clip) {
    CKR(dalloc(e, &e->cls_emb, Dv));
    CKR(dalloc(e, &e->pre_g, Dv)); CKR(dalloc(e, &e->pre_b, Dv));
    reg(e, vp + "embeddings.class_embedding", Dv, e->cls_emb, 0);
    reg(e, vp + "pre_layrnorm.weight", Dv, e->pre_g, 0);
    reg(e, vp + "pre_layrnorm.bias", Dv, e->pre_b, 0);
  } else {
    const int F = c.vis_ffn;
    const int64_t dd = (int64_t)Dv * Dv;
    CKR(dalloc(e, &e->bpatch, Dv));
    reg(e, vp + "embeddings.patch_embedding.bias", Dv, e->bpatch, 0);
    CKR(dalloc(e, &e->thead_w, (int64_t)c.proj_dim * Dt)); CKR(dalloc(e, &e->thead_b, c.proj_dim));
    reg(e, tp + "head.weight", (int64_t)c.proj_dim * Dt, e->thead_w, 0);
    reg(e, tp + "head.bias", c.proj_dim, e->thead_b, 0);
    // MAP head (HF siglip :586-649): packed in_proj -> fp32 q rows (probe projection) + bf16 K|V rows
    CKR(dalloc(e, &e->map_probe, Dv));
    CKR(dalloc(e, &e->map_inw, dd)); CKR(dalloc(e, &e->map_inb, Dv)); CKR(dalloc(e, &e->map_q, Dv));
    CKR(dalloc(e, &e->map_wkv, 2 * dd)); CKR(dalloc(e, &e->map_bkv, 2 * Dv));
    CKR(dalloc(e, &e->map_wo, dd)); CKR(dalloc(e, &e->map_bo, Dv));
    CKR(dalloc(e, &e->map_lng, Dv)); CKR(dalloc(e, &e->map_lnb, Dv));
    CKR(dalloc(e, &e->map_w1, (int64_t)F * Dv)); CKR(dalloc(e, &e->map_b1, F));
    CKR(dalloc(e, &e->map_w2, (int64_t)Dv * F)); CKR(dalloc(e, &e->map_b2, Dv));
    const std::string h = vp + "head.";
    reg(e, h + "probe", Dv, e->map_probe, 0);
    reg2d(e, h + "attention.in_proj_weight", 3 * dd, 0, 1, dd, dd, e->map_inw, dd, 0);
    reg2d(e, h + "attention.in_proj_weight", 3 * dd, dd, 1, 2 * dd, 2 * dd, e->map_wkv, 2 * dd, 1);
    reg2d(e, h + "attention.in_proj_bias", 3 * Dv, 0, 1, Dv, Dv, e->map_inb, Dv, 0);
    reg2d(e, h + "attention.in_proj_bias", 3 * Dv, Dv, 1, 2 * Dv, 2 * Dv, e->map_bkv, 2 * Dv, 0);
    reg(e, h + "attention.out_proj.weight", dd, e->map_wo, 1);
    reg(e, h + "attention.out_proj.bias", Dv, e->map_bo, 0);
    reg(e, h + "layernorm.weight", Dv, e->map_lng, 0);
    reg(e, h + "layernorm.bias", Dv, e->map_lnb, 0);
    reg(e, h + "mlp.fc1.weight", (int64_t)F * Dv, e->map_w1, 1);
    reg(e, h + "mlp.fc1.bias", F, e->map_b1, 0);
    reg(e, h + "mlp.fc2.weight", (int64_t)Dv * F, e->map_w2, 1);
    reg(e, h + "mlp.fc2.bias", Dv, e->map_b2, 0);
  }
  // ---- heads (fp32)
  HeadWeights& w = e->hw;
  memset(&w, 0, sizeof(w));
  const int fd = c.fusion_dim, C = c.num_outputs;
  w.head = c.head; w.backend = c.backend; w.fd = fd; w.n_out = C; w.hh = c.head_hidden_dim;
  w.dt = clip ? Dt : c.proj_dim;  // siglip: the text head maps Dt -> proj_dim (== Dt for every released SigLIP)
  w.dv = Dv;
  if (!clip && c.proj_dim != Dt) return fail(MMCM_EINVAL, "siglip: projection_size %d != text hidden %d", c.proj_dim, Dt);
  auto f32 = [&](const std::string& key, int64_t n, const float** field) -> int {
    float* p = nullptr;
    CKR(dalloc(e, &p, n));
    reg(e, key, n, p, 0);
    *field = p;
    return MMCM_OK;
  };
  int din_t, din_v;
  if (fusion) {
    if (clip) {
      w.dp = c.proj_dim;
      CKR(f32("backbone.text_projection.weight", (int64_t)c.proj_dim * Dt, &w.text_proj));
      CKR(f32("backbone.visual_projection.weight", (int64_t)c.proj_dim * Dv, &w.vis_proj));
      din_t = din_v = c.proj_dim;
    } else {
      w.dp = c.proj_dim;
      w.text_head_w = e->thead_w; w.text_head_b = e->thead_b;
      din_t = c.proj_dim; din_v = Dv;
      if (din_t != din_v) return fail(MMCM_EINVAL, "siglip fusion: text %d and vision %d feature widths differ", din_t, din_v);
    }
  } else {
    if (!clip) return fail(MMCM_EINVAL, "MultiTaskClassifier supports the clip backend only "
                                        "(the reference asserts for AutoModel backends, multitask.py:81-88)");
    w.dp = 0;
    din_t = Dt; din_v = Dv;
  }
  if (din_t > HEAD_MAXD || din_v > HEAD_MAXD || fd > HEAD_MAXD || fd % 4 != 0 || din_t % 4 != 0 || din_v % 4 != 0 ||
      (clip && fusion && c.proj_dim > fd))
    return fail(MMCM_EINVAL, "head: unsupported widths (text %d vision %d fusion %d)", din_t, din_v, fd);
  if (C > fd) return fail(MMCM_EINVAL, "head: num_outputs %d > fusion_dim %d", C, fd);
  // the SigLIP text head is staged through the 5*fd-wide interaction buffer with the tower's row pitch (heads.cuh)
  if (!clip && 5 * fd < HEAD_MAXD)
    return fail(MMCM_EINVAL, "head: siglip backend needs fusion_dim >= %d (got %d)", (HEAD_MAXD + 4) / 5, fd);
  CKR(f32("proj_t.weight", (int64_t)fd * din_t, &w.proj_t_w)); CKR(f32("proj_t.bias", fd, &w.proj_t_b));
  CKR(f32("proj_i.weight", (int64_t)fd * din_v, &w.proj_i_w)); CKR(f32("proj_i.bias", fd, &w.proj_i_b));
  CKR(f32("g_t.weight", (int64_t)fd * fd, &w.g_t_w)); CKR(f32("g_t.bias", fd, &w.g_t_b));
  CKR(f32("g_i.weight", (int64_t)fd * fd, &w.g_i_w)); CKR(f32("g_i.bias", fd, &w.g_i_b));
  {  // gate.weight [fd, 2fd+2] -> dense [fd, 2fd] + the two presence columns [fd, 2]
    float *gw = nullptr, *gp = nullptr;
    CKR(dalloc(e, &gw, (int64_t)fd * 2 * fd));
    CKR(dalloc(e, &gp, (int64_t)fd * 2));
    const int64_t n = (int64_t)fd * (2 * fd + 2);
    reg2d(e, "gate.weight", n, 0, fd, 2 * fd, 2 * fd + 2, gw, 2 * fd, 0);
    reg2d(e, "gate.weight", n, 2 * fd, fd, 2, 2 * fd + 2, gp, 2, 0);
    w.gate_w = gw; w.gate_wp = gp;
    CKR(f32("gate.bias", fd, &w.gate_b));
  }
  if (fusion) {
    CKR(f32("ln_fused.weight", fd, &w.ln_fused_g)); CKR(f32("ln_fused.bias", fd, &w.ln_fused_b));
    CKR(f32("cls.0.weight", 5 * fd, &w.cls0_g)); CKR(f32("cls.0.bias", 5 * fd, &w.cls0_b));
    CKR(f32("cls.1.weight", (int64_t)fd * 5 * fd, &w.cls1_w)); CKR(f32("cls.1.bias", fd, &w.cls1_b));
    CKR(f32("cls.4.weight", (int64_t)C * fd, &w.cls4_w)); CKR(f32("cls.4.bias", C, &w.cls4_b));
  } else {
    CKR(f32("shared_head.1.weight", (int64_t)fd * fd, &w.shared_w)); CKR(f32("shared_head.1.bias", fd, &w.shared_b));
    const int hh = c.head_hidden_dim;
    if (hh > 0) {
      if (hh > fd || hh % 4 != 0) return fail(MMCM_EINVAL, "mtl: head_hidden_dim %d must be <= fusion_dim and %% 4 == 0", hh);
      float *h0w, *h0b, *h3w, *h3b;
      CKR(dalloc(e, &h0w, (int64_t)C * hh * fd)); CKR(dalloc(e, &h0b, (int64_t)C * hh));
      CKR(dalloc(e, &h3w, (int64_t)C * hh)); CKR(dalloc(e, &h3b, C));
      for (int j = 0; j < C; ++j) {
        const std::string p = "heads." + std::to_string(j) + ".";
        reg(e, p + "0.weight", (int64_t)hh * fd, h0w + (int64_t)j * hh * fd, 0);
        reg(e, p + "0.bias", hh, h0b + (int64_t)j * hh, 0);
        reg(e, p + "3.weight", hh, h3w + (int64_t)j * hh, 0);
        reg(e, p + "3.bias", 1, h3b + j, 0);
      }
      w.h0_w = h0w; w.h0_b = h0b; w.h3_w = h3w; w.h3_b = h3b;
    } else {
      float *h3w, *h3b;
      CKR(dalloc(e, &h3w, (int64_t)C * fd)); CKR(dalloc(e, &h3b, C));
      for (int j = 0; j < C; ++j) {
        const std::string p = "heads." + std::to_string(j) + ".";
        reg(e, p + "weight", fd, h3w + (int64_t)j * fd, 0);
        reg(e, p + "bias", 1, h3b + j, 0);
      }
      w.h3_w = h3w; w.h3_b = h3b;
    }
  }
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ arenas
static void free_arena(Eng* e, Arena& a) {
  dfree(e, a.x); dfree(e, a.h); dfree(e, a.qkv); dfree(e, a.att); dfree(e, a.ff); dfree(e, a.pool_row);
  dfree(e, a.seq_start); dfree(e, a.seq_len); dfree(e, a.rows_dev);
  dfree(e, a.xp); dfree(e, a.attp); dfree(e, a.hp); dfree(e, a.ffp);
  dfree(e, a.stats); dfree(e, a.statsp); dfree(e, a.part);
  a = Arena();
}
static int alloc_arena(Eng* e, Arena& a, const TowerW& t, int64_t rows, int mb) {
  a.rows = rows;
  a.mb = mb;
  CKR(dalloc(e, &a.stats, rows * (t.D / LN_SLAB)));
  CKR(dalloc(e, &a.statsp, (int64_t)mb * (t.D / LN_SLAB)));
  CK(cudaMemset(a.stats, 0, rows * (t.D / LN_SLAB) * sizeof(float2)));
  CK(cudaMemset(a.statsp, 0, (int64_t)mb * (t.D / LN_SLAB) * sizeof(float2)));
  CKR(dalloc(e, &a.part, (int64_t)kSplitPlanes * kSplitRows * t.D));
  CK(cudaMemset(a.part, 0, (int64_t)kSplitPlanes * kSplitRows * t.D * sizeof(float)));
  CKR(dalloc(e, &a.x, rows * t.D));
  CKR(dalloc(e, &a.h, rows * t.D));
  CKR(dalloc(e, &a.qkv, rows * 3 * t.D));
  CKR(dalloc(e, &a.att, rows * t.D));
  CKR(dalloc(e, &a.ff, rows * t.F));
  CKR(dalloc(e, &a.pool_row, mb));
  CKR(dalloc(e, &a.seq_start, mb));
  CKR(dalloc(e, &a.seq_len, mb));
  CKR(dalloc(e, &a.rows_dev, 256));   // one live-row counter per packed chunk of a forward (accounting reads them back)
  CKR(dalloc(e, &a.xp, (int64_t)mb * t.D));
  CKR(dalloc(e, &a.attp, (int64_t)mb * t.D));
  CKR(dalloc(e, &a.hp, (int64_t)mb * t.D));
  CKR(dalloc(e, &a.ffp, (int64_t)mb * t.F));
  CK(cudaMemset(a.xp, 0, (int64_t)mb * t.D * sizeof(float)));
  CK(cudaMemset(a.attp, 0, (int64_t)mb * t.D * sizeof(bf16)));
  CK(cudaMemset(a.hp, 0, (int64_t)mb * t.D * sizeof(bf16)));
  CK(cudaMemset(a.ffp, 0, (int64_t)mb * t.F * sizeof(bf16)));
  // TMA tiles of partially filled chunks also cover rows nobody wrote in this forward: they must hold finite numbers
  // (a masked probability of 0 times a stale NaN in V would still be NaN)
  CK(cudaMemset(a.x, 0, rows * t.D * sizeof(float)));
  CK(cudaMemset(a.h, 0, rows * t.D * sizeof(bf16)));
  CK(cudaMemset(a.qkv, 0, rows * 3 * t.D * sizeof(bf16)));
  CK(cudaMemset(a.att, 0, rows * t.D * sizeof(bf16)));
  CK(cudaMemset(a.ff, 0, rows * t.F * sizeof(bf16)));
  return MMCM_OK;
}

static int vis_tokens(const mmcm_config& c) {
  const int G = c.image / c.patch;
  return G * G + (c.backend == MMCM_BACKEND_CLIP ? 1 : 0);
}

// Captured graphs bake device pointers: whenever a buffer they may reference is reallocated they must go.
static void invalidate_graphs(Eng* e) {
  for (auto& kv : e->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
}

static int ensure_arenas(Eng* e, int mbt, int mbv) {
  const mmcm_config& c = e->cfg;
  if (mbt > e->mb_text || mbv > e->mb_vis) invalidate_graphs(e);
  if (mbt > e->mb_text) {
    CK(cudaDeviceSynchronize());
    free_arena(e, e->at);
    dfree(e, e->key_valid);
    e->key_valid = nullptr;
    CKR(alloc_arena(e, e->at, e->text, (int64_t)mbt * c.max_pos, mbt));
    CKR(dalloc(e, &e->key_valid, (int64_t)mbt * c.max_pos));
    e->mb_text = mbt;
  }
  if (mbv > e->mb_vis) {
    CK(cudaDeviceSynchronize());
    free_arena(e, e->av);
    dfree(e, e->im2col);
    dfree(e, e->map_kv); dfree(e, e->map_att); dfree(e, e->map_h); dfree(e, e->map_ff); dfree(e, e->map_y);
    e->im2col = nullptr;
    e->map_kv = e->map_att = e->map_h = e->map_ff = nullptr; e->map_y = nullptr;
    const int Tv = vis_tokens(c);
    const int G = c.image / c.patch, P = G * G, Kp = 3 * c.patch * c.patch;
    CKR(alloc_arena(e, e->av, e->vis, (int64_t)mbv * Tv, mbv));
    CKR(dalloc(e, &e->im2col, (int64_t)mbv * P * Kp));
    if (c.backend == MMCM_BACKEND_SIGLIP) {
      const int D = c.vis_hidden;
      CKR(dalloc(e, &e->map_kv, (int64_t)mbv * Tv * 2 * D));
      CKR(dalloc(e, &e->map_att, (int64_t)mbv * D));
      CKR(dalloc(e, &e->map_h, (int64_t)mbv * D));
      CKR(dalloc(e, &e->map_ff, (int64_t)mbv * c.vis_ffn));
      CKR(dalloc(e, &e->map_y, (int64_t)mbv * D));
    }
    e->mb_vis = mbv;
  }
  return MMCM_OK;
}

// ---- micro-batch selection ------------------------------------------------------------------------
// The persistent GEMMs hand out 128 x 256 tiles to 148 CTAs, so a launch costs ceil(tiles / 148) "rounds": a chunk
// of 128 samples of the CLIP text tower is 77 row blocks x 2 column blocks = 154 tiles = TWO rounds for out_proj/fc2,
// a chunk of 123 samples is 148 tiles = one.  Pick, per tower, the chunk size (<= cap) that minimises the modelled
// cycles of the whole batch: rounds * (k-blocks * 512 + epilogue) per GEMM + a fixed cost per launch.
static double chunk_cost(const TowerW& t, int T, int n, int sms) {
  if (n <= 0) return 0.0;
  const double mt = std::ceil((double)n * T / 256.0);   // CTA-pair tiles: 256 rows
  const double units = sms / 2;                         // pairs
  auto gemm = [&](int N, int K) {
    const int bn = (N % 256 == 0) ? 256 : 128;
    const double rounds = std::ceil(mt * (N / bn) / units);
    return rounds * ((K / 64) * 620.0 * bn / 256.0 + 5000.0) + 12000.0;
  };
  const double layer = gemm(3 * t.D, t.D) + gemm(t.D, t.D) + gemm(t.F, t.D) + gemm(t.D, t.F) +
                       3 * 8000.0 /* LN, LN, attention launches */ + 3.0 * n * T * t.D / 148.0 / 32.0;
  return layer * t.L;
}
static int choose_chunk(const TowerW& t, int T, int B, int cap, int sms) {
  if (cap > B) cap = B;
  // the search is ~1000 cost evaluations: remember the answer per (tower shape, T, B, cap)
  static std::map<std::tuple<int, int, int, int, int, int>, int> memo;
  const auto key = std::make_tuple(t.D, t.F, t.L, T, B, cap);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = memo.find(key);
    if (it != memo.end()) return it->second;
  }
  int best = cap;
  double best_cost = 1e300;
  for (int c = cap; c >= 16 || c == cap; --c) {
    if (c < 1) break;
    const int full = B / c, rem = B - full * c;
    const double cost = full * chunk_cost(t, T, c, sms) + chunk_cost(t, T, rem, sms);
    if (cost < best_cost * 0.999) { best_cost = cost; best = c; }
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (memo.size() > 4096) memo.clear();
  memo[key] = best;
  return best;
}

static int ensure_batch(Eng* e, int64_t B) {
  if (B <= e->cap_B) return MMCM_OK;
  CK(cudaDeviceSynchronize());
  invalidate_graphs(e);
  dfree(e, e->pooled_t); dfree(e, e->pooled_v); dfree(e, e->feat_t); dfree(e, e->feat_v);
  int64_t cap = e->cap_B ? e->cap_B : 64;
  while (cap < B) cap *= 2;
  CKR(dalloc(e, &e->pooled_t, cap * e->cfg.text_hidden));
  CKR(dalloc(e, &e->pooled_v, cap * e->cfg.vis_hidden));
  CKR(dalloc(e, &e->feat_t, cap * HEAD_MAXD));
  CKR(dalloc(e, &e->feat_v, cap * HEAD_MAXD));
  e->cap_B = cap;
  return MMCM_OK;
}

// ------------------------------------------------------------------------------------------------ towers
// `pooled_last`: a.pool_row holds the one row per sample the caller reads after the last layer; the last layer then
// runs out_proj / LN2 / MLP on those B rows only and leaves their residual in a.xp [B, D] (a.x keeps the last layer's
// INPUT).  Row-wise ops on gathered rows give the same bits as on the full matrix.
// Small batches (B < 16): the residual GEMM is a handful of tiles and the x tiles the EPI_RESID_STATS epilogue pulls in
// are pure exposed latency (B = 1: 1.39 ms with the fold, 1.16 ms with the separate pass; B = 32: 1.30 vs 1.41 ms --
// tools/latency.py), so they keep the separate normalisation kernel.  The choice is made once per forward from B, not
// per micro-batch: every chunk of a forward runs the same arithmetic, so logits do not depend on the chunking.
constexpr int kLnFoldMinBatch = 16;
static bool use_ln_fold(const Eng* e, const TowerW& t) {
  return e->fold_forward && e->opt_gemm_impl == 0 && e->opts.tma_epilogue && t.D % 256 == 0 && t.D <= 1024;
}

// On entry with the LN fold on, a.h holds bf16(a.x) and a.stats the slab statistics of a.x (launch_prep_rows or the
// previous layer's fc2); every layer keeps that invariant.  With it off a.h is scratch for the LayerNorm output.
static int run_layers(Eng* e, const TowerW& t, Arena& a, int rows, int B, int T, const uint8_t* kvalid, int causal,
                      cudaStream_t st, const int* packed_rows = nullptr, bool pooled_last = false) {
  LaunchStats* S = &e->stats;
  const int D = t.D, F = t.F, impl = e->opt_gemm_impl;
  const bool packed = packed_rows != nullptr;
  const bool fold = use_ln_fold(e, t);
  const int* rdev = packed_rows;                             // live row count of a packed chunk (device side)
  const int* sstart = packed ? a.seq_start : nullptr;
  const int* slen = packed ? a.seq_len : nullptr;
  // h @ W^T with the LayerNorm in front of it: normalised rows x folded weights, or raw bf16 rows + LN-fold epilogue
  auto ln_linear = [&](const float* x, bf16* h, float2* stats, int pitch, const bf16* W, const float* bias,
                       int M, int N, bf16* out, int act, const int* mdev) -> int {
    EpiParams ep{};
    ep.bias = bias; ep.out = out; ep.ldo = N; ep.act = act; ep.m_dev = mdev;
    if (fold) {
      ep.stats = stats; ep.stats_pitch = pitch; ep.ln_slabs = D / LN_SLAB; ep.ln_eps = t.eps;
      return launch_gemm(h, W, M, N, D, act ? EPI_LNFOLD_ACT_BF16 : EPI_LNFOLD_BF16, ep, impl, st, S);
    }
    // (affine part lives in W / bias; split-K partials of the residual GEMM in front are absorbed here)
    CKR(launch_layernorm(x, nullptr, nullptr, t.eps, M, D, nullptr, h, nullptr, st, S, mdev, x == a.x ? &a.pend_x : &a.pend_xp));
    return launch_gemm(h, W, M, N, D, act ? EPI_BIAS_ACT_BF16 : EPI_BIAS_BF16, ep, impl, st, S);
  };
  // x += A @ W^T + b; with the fold also the bf16 copy and the slab statistics the next LN-consuming GEMM needs
  auto resid_linear = [&](const bf16* A, const bf16* W, const float* bias, float* x, bf16* xb, float2* stats, int pitch,
                          int M, int K, const int* mdev, bool want_stats, bool ln_follows = true) -> int {
    EpiParams ep{};
    ep.bias = bias; ep.out = x; ep.resid = x; ep.ldo = D; ep.m_dev = mdev;
    // Small forwards (B < 16, fold off): a handful of tiles, each CTA pair bound by the few KB it keeps in flight.
    // Split K over idle pairs; the partial sums go to planes of a.part and the LayerNorm that follows adds them to x
    // in a fixed order (deterministic, no reduction kernel, no inter-CTA waits).
    // The split count depends on K only, never on M: every row of a forward -- whatever micro-batch it travels in, and
    // in the pooled-rows last layer as much as in the all-rows one -- sees the same summation order.
    if (!fold && ln_follows && impl == 0 && e->split_forward && M <= kSplitRows) {
      const int num_kb = K / 64;
      int ks = std::min(kSplitPlanes, num_kb / 4);
      while (ks > 1 && num_kb % ks != 0) --ks;
      if (ks > 1) {
        ep.out = a.part; ep.resid = nullptr; ep.ksplit = ks; ep.part_rows = kSplitRows;
        CKR(launch_gemm(A, W, M, D, K, EPI_BIAS_RESID_F32, ep, impl, st, S));
        PendingParts& pend = (x == a.x) ? a.pend_x : a.pend_xp;
        pend.part = a.part; pend.ksplit = ks; pend.part_rows = kSplitRows;
        return MMCM_OK;
      }
    }
    if (fold && want_stats) {
      ep.xb = xb; ep.stats = stats; ep.stats_pitch = pitch;
      return launch_gemm(A, W, M, D, K, EPI_RESID_STATS, ep, impl, st, S);
    }
    return launch_gemm(A, W, M, D, K, EPI_BIAS_RESID_F32, ep, impl, st, S);
  };
  for (int i = 0; i < t.L; ++i) {
    const LayerW& w = t.layers[i];
    const bool last = i == t.L - 1;
    // qkv = LN1(x) @ [Wq*s | Wk | Wv]^T + [bq*s | bk | bv]     HF clip :372, :313-319
    CKR(ln_linear(a.x, a.h, a.stats, (int)a.rows, w.wqkv, w.bqkv, rows, 3 * D, a.qkv, ACT_NONE, rdev));
    // att = softmax(q k^T + mask) v                            HF clip :321-332
    CKR(launch_attention(a.qkv, kvalid, B, T, t.H, causal, a.att, st, S, sstart, slen));
    if (pooled_last && last) {
      CK(launch_k(gather_pool_rows_kernel, dim3((B + 7) / 8), dim3(256), 0, st, (const bf16*)a.att, (const float*)a.x,
                  (const int*)a.pool_row, B, D, a.attp, a.xp));
      CK(cudaGetLastError());
      S->launches++;
      CKR(resid_linear(a.attp, w.wo, w.bo, a.xp, a.hp, a.statsp, a.mb, B, D, nullptr, true));
      CKR(ln_linear(a.xp, a.hp, a.statsp, a.mb, w.w1, w.b1, B, F, a.ffp, t.act, nullptr));
      // (the final LN reads x itself; no split-K here: the all-rows path keeps its last fc2 unsplit so that x stays
      // complete for every row, and the two paths must add in the same order to stay bit-identical)
      CKR(resid_linear(a.ffp, w.w2, w.b2, a.xp, nullptr, nullptr, 0, B, F, nullptr, false, false));
      break;
    }
    // x = x + att @ Wo^T + bo                                  HF clip :334, :379
    CKR(resid_linear(a.att, w.wo, w.bo, a.x, a.h, a.stats, (int)a.rows, rows, D, rdev, true));
    // ff = act(LN2(x) @ W1^T + b1); x = x + ff @ W2^T + b2     HF clip :381-384, :347-351
    CKR(ln_linear(a.x, a.h, a.stats, (int)a.rows, w.w1, w.b1, rows, F, a.ff, t.act, rdev));
    // (after the last layer the final LayerNorm may read only the pooled rows: keep x itself complete there)
    CKR(resid_linear(a.ff, w.w2, w.b2, a.x, a.h, a.stats, (int)a.rows, rows, F, rdev, !last, !last));
  }
  return MMCM_OK;
}

static int run_text(Eng* e, const int64_t* ids, const int64_t* mask, int n, int S, float* pooled, cudaStream_t st,
                    const float* tp = nullptr, const float* ip = nullptr) {
  const mmcm_config& c = e->cfg;
  const TowerW& t = e->text;
  Arena& a = e->at;
  const int rows = n * S;
  const bool clip = c.backend == MMCM_BACKEND_CLIP;
  const int blocks = (rows + 7) / 8;
  const int eos = clip ? c.eos_id : -1;
  const bool packed = clip && e->opt_varlen_text;   // exact for the causal CLIP tower only (see text_plan_kernel)
  int* const rows_slot = a.rows_dev + (e->stats.text_chunk++ & 255);
  if (packed) {
    // absent-text shortcut (exact, see text_plan_kernel): only with the presence flags at hand and the option on
    const int skip_mode = (e->opt_skip_absent && tp && ip) ? (c.head == MMCM_HEAD_FUSION ? 1 : 2) : 0;
    CK(launch_k(text_plan_kernel, dim3(1), dim3(1024), 0, st, ids, n, S, eos, a.seq_start, a.seq_len, a.pool_row, rows_slot,
                tp, ip, skip_mode));
    if (t.D == 512)
      CK(launch_k(text_embed_packed_kernel<512>, dim3(blocks), dim3(256), 0, st, ids, mask, e->tok_emb, e->tpos_emb, n, S,
                  c.vocab, a.seq_start, a.seq_len, a.x, e->key_valid));
    else if (t.D == 768)
      CK(launch_k(text_embed_packed_kernel<768>, dim3(blocks), dim3(256), 0, st, ids, mask, e->tok_emb, e->tpos_emb, n, S,
                  c.vocab, a.seq_start, a.seq_len, a.x, e->key_valid));
    else return fail(MMCM_EINVAL, "text hidden %d unsupported (512, 768)", t.D);
    e->stats.launches += 2;
  } else {
    if (t.D == 512)
      CK(launch_k(text_embed_kernel<512>, dim3(blocks), dim3(256), 0, st, ids, mask, e->tok_emb, e->tpos_emb, n, S, c.vocab, eos, a.x,
                  a.pool_row, e->key_valid));
    else if (t.D == 768)
      CK(launch_k(text_embed_kernel<768>, dim3(blocks), dim3(256), 0, st, ids, mask, e->tok_emb, e->tpos_emb, n, S, c.vocab, eos, a.x,
                  a.pool_row, e->key_valid));
    else return fail(MMCM_EINVAL, "text hidden %d unsupported (512, 768)", t.D);
    e->stats.launches++;
  }
  if (use_ln_fold(e, t))   // bf16 copy + slab statistics of the embedded rows: what layer 0's QKV GEMM consumes
    CKR(launch_prep_rows(a.x, nullptr, nullptr, t.eps, rows, t.D, a.h, a.stats, (int)a.rows, st, &e->stats,
                         packed ? rows_slot : nullptr));
  e->opts.pair_limit = e->opt_streams >= 2 ? e->opt_pairs_text : 0;
  const bool pl = e->opt_pooled_last && S > 1;
  const int rl = run_layers(e, t, a, rows, n, S, e->key_valid, clip ? 1 : 0, st, packed ? rows_slot : nullptr, pl);
  e->opts.pair_limit = 0;
  CKR(rl);
  // pooled = final_layer_norm(x)[pool_row]   (LayerNorm is row-wise, so only the pooled rows are normalised)
  if (pl) CKR(launch_layernorm(a.xp, e->tfin_g, e->tfin_b, t.eps, n, t.D, nullptr, nullptr, pooled, st, &e->stats, nullptr, &a.pend_xp));
  else CKR(launch_layernorm(a.x, e->tfin_g, e->tfin_b, t.eps, n, t.D, a.pool_row, nullptr, pooled, st, &e->stats, nullptr, &a.pend_x));
  if (a.pend_x.part || a.pend_xp.part) return fail(MMCM_ESTATE, "internal: split-K partial sums of the text tower were not absorbed");
  e->last_text_rows = rows;
  return MMCM_OK;
}

// Pixel source of a forward: the reference's normalised fp32 CHW `pixel_values`, or raw uint8 HWC images with the
// Normalize constants (ToTensor + Normalize then happen inside the im2col, see im2col_u8_kernel).
struct Pixels {
  const float* f32 = nullptr;
  const uint8_t* u8 = nullptr;
  float mean[3] = {0.f, 0.f, 0.f}, std[3] = {1.f, 1.f, 1.f};
  size_t bytes_per_sample(const mmcm_config& c) const { return (size_t)3 * c.image * c.image * (u8 ? 1 : 4); }
  Pixels at(int64_t b, const mmcm_config& c) const {        // the same source advanced by b samples
    Pixels r = *this;
    const int64_t per = (int64_t)3 * c.image * c.image;
    if (u8) r.u8 = u8 + b * per; else r.f32 = f32 + b * per;
    return r;
  }
};

static int run_vision(Eng* e, const Pixels& px, int n, float* pooled, cudaStream_t st) {
  const mmcm_config& c = e->cfg;
  const TowerW& t = e->vis;
  Arena& a = e->av;
  LaunchStats* S = &e->stats;
  const bool clip = c.backend == MMCM_BACKEND_CLIP;
  const int G = c.image / c.patch, P = G * G, Kp = 3 * c.patch * c.patch, T = vis_tokens(c), D = t.D;
  const int rows = n * T;
  // patch embedding = im2col + GEMM, position embedding (and conv bias) fused into the epilogue   HF clip :202-218
  {
    const int64_t chunks = (int64_t)n * P * (Kp / 8);
    const int blocks = (int)((chunks + 255) / 256 < 148 * 16 ? (chunks + 255) / 256 : 148 * 16);
    if (px.u8)
      CK(launch_k(im2col_u8_kernel, dim3(blocks), dim3(256), 0, st, px.u8, e->im2col, n, c.image, c.patch, px.mean[0],
                  px.mean[1], px.mean[2], px.std[0], px.std[1], px.std[2],
                  (int)((reinterpret_cast<uintptr_t>(px.u8) & 7) == 0)));
    else
      CK(launch_k(im2col_kernel, dim3(blocks), dim3(256), 0, st, px.f32, e->im2col, n, c.image, c.patch));
    CK(cudaGetLastError());
    S->launches++;
  }
  EpiParams ep{};
  ep.bias = e->bpatch; ep.out = a.x; ep.pos = e->vpos_emb; ep.ldo = D; ep.P = P; ep.T = T;
  CKR(launch_gemm(e->im2col, e->wpatch, n * P, D, Kp, EPI_PATCH_F32, ep, e->opt_gemm_impl, st, S));
  if (clip) {
    CK(launch_k(cls_rows_kernel, dim3((n * D + 255) / 256), dim3(256), 0, st, e->cls_emb, e->vpos_emb, a.x, n, T, D));
    CK(cudaGetLastError());
    S->launches++;
  }
  if (use_ln_fold(e, t))   // pre_layrnorm (CLIP, in place) fused with the bf16 copy + slab statistics layer 0 consumes
    CKR(launch_prep_rows(a.x, clip ? e->pre_g : nullptr, clip ? e->pre_b : nullptr, t.eps, rows, D, a.h, a.stats,
                         (int)a.rows, st, S));
  else if (clip)
    CKR(launch_layernorm(a.x, e->pre_g, e->pre_b, t.eps, rows, D, nullptr, nullptr, a.x, st, S));  // pre_layrnorm
  if (clip) {   // pooled row = CLS (row 0 of each sample)
    CK(launch_k(fill_pool_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, st, a.pool_row, n, T, 0));
    CK(cudaGetLastError());
    S->launches++;
  }
  const bool pl = clip && e->opt_pooled_last;   // the SigLIP MAP head reads every token of the last layer
  e->opts.pair_limit = e->opt_streams >= 2 ? e->opt_pairs_vis : 0;
  const int rl = run_layers(e, t, a, rows, n, T, nullptr, 0, st, nullptr, pl);
  e->opts.pair_limit = 0;
  CKR(rl);
  if (clip) {
    if (pl) CKR(launch_layernorm(a.xp, e->post_g, e->post_b, t.eps, n, D, nullptr, nullptr, pooled, st, S, nullptr, &a.pend_xp));
    else CKR(launch_layernorm(a.x, e->post_g, e->post_b, t.eps, n, D, a.pool_row, nullptr, pooled, st, S, nullptr, &a.pend_x));
  } else {
    // post_layernorm over all tokens, then the MAP head   HF siglip :617-649
    const int impl = e->opt_gemm_impl;
    CKR(launch_layernorm(a.x, e->post_g, e->post_b, t.eps, rows, D, nullptr, a.h, nullptr, st, S, nullptr, &a.pend_x));
    ep = EpiParams{};
    ep.bias = e->map_bkv; ep.out = e->map_kv; ep.ldo = 2 * D;
    CKR(launch_gemm(a.h, e->map_wkv, rows, 2 * D, D, EPI_BIAS_BF16, ep, impl, st, S));
    if (T > MAP_MAXT) return fail(MMCM_EINVAL, "MAP head: %d tokens > %d", T, MAP_MAXT);
    CK(launch_k(map_attention_kernel, dim3(dim3(t.H, n)), dim3(128), 0, st, e->map_kv, e->map_q, e->map_att, T, D));
    CK(cudaGetLastError());
    S->launches++;
    ep = EpiParams{};
    ep.bias = e->map_bo; ep.out = e->map_y; ep.ldo = D;
    CKR(launch_gemm(e->map_att, e->map_wo, n, D, D, EPI_BIAS_RESID_F32, ep, impl, st, S));
    CKR(launch_layernorm(e->map_y, e->map_lng, e->map_lnb, t.eps, n, D, nullptr, e->map_h, nullptr, st, S));
    ep = EpiParams{};
    ep.bias = e->map_b1; ep.out = e->map_ff; ep.ldo = t.F; ep.act = t.act;
    CKR(launch_gemm(e->map_h, e->map_w1, n, t.F, D, EPI_BIAS_ACT_BF16, ep, impl, st, S));
    ep = EpiParams{};
    ep.bias = e->map_b2; ep.out = pooled; ep.resid = e->map_y; ep.ldo = D;
    CKR(launch_gemm(e->map_ff, e->map_w2, n, D, t.F, EPI_BIAS_RESID_F32, ep, impl, st, S));
  }
  if (a.pend_x.part || a.pend_xp.part) return fail(MMCM_ESTATE, "internal: split-K partial sums of the vision tower were not absorbed");
  e->last_vis_rows = rows;
  return MMCM_OK;
}

static int run_head(Eng* e, const float* tp, const float* ip, int B, float* logits, float* probs, cudaStream_t st) {
  static AttrOnce once;
  const int smem = head_smem_bytes(e->cfg.fusion_dim);
  if (once.need()) CK(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (smem > 200 * 1024) return fail(MMCM_EINVAL, "head: fusion_dim %d needs too much shared memory", e->cfg.fusion_dim);
  const bool dbg = e->opt_debug_feats && e->cfg.head == MMCM_HEAD_FUSION;
  // few sample groups: a cluster of HEAD_CLUSTER CTAs per group splits every Linear's columns (heads.cuh)
  const int groups = (B + HEAD_SB - 1) / HEAD_SB;
  const int csize = (e->opt_head_cluster && groups * HEAD_CLUSTER <= g_num_sms) ? HEAD_CLUSTER : 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * csize);
  cfg.blockDim = dim3(HEAD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (t_opts->pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (csize > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = csize;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  CK(cudaLaunchKernelEx(&cfg, head_kernel, e->hw, (const float*)e->pooled_t, (const float*)e->pooled_v, tp, ip, B, logits,
                        probs, dbg ? e->feat_t : (float*)nullptr, dbg ? e->feat_v : (float*)nullptr));
  CK(cudaGetLastError());
  e->stats.launches++;
  return MMCM_OK;
}

static void clear_gemm_events(LaunchStats& s) {
  for (auto& p : s.gemm_events) {
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  s.gemm_events.clear();
  s.text_chunk = 0;
}

static int check_forward_args(Eng* e, const void* ids, const void* px, const void* tp, const void* ip, int B, int S,
                              const void* logits) {
  if (!e) return fail(MMCM_EINVAL, "null handle");
  if (!e->finalized) return fail(MMCM_ESTATE, "weights are not finalized (call mmcm_finalize_weights)");
  if (B < 0 || S <= 0) return fail(MMCM_EINVAL, "bad batch %d or sequence length %d", B, S);
  if (S > e->cfg.max_pos)  // mirrors HF/models/clip/modeling_clip.py:243-247 (siglip :211-215)
    return fail(MMCM_EINVAL,
                "Sequence length must be less than max_position_embeddings (got `sequence length`: %d and "
                "max_position_embeddings: %d", S, e->cfg.max_pos);
  if (B > 0 && (!ids || !px || !tp || !ip || !logits)) return fail(MMCM_EINVAL, "null input/output pointer");
  return MMCM_OK;
}

static int ensure_static_io(Eng* e, int64_t B) {
  const mmcm_config& c = e->cfg;
  if (B <= e->host_cap) return MMCM_OK;
  const int64_t px_per = (int64_t)3 * c.image * c.image;
  CK(cudaDeviceSynchronize());
  dfree(e, e->d_ids); dfree(e, e->d_mask); dfree(e, e->d_px); dfree(e, e->d_tp); dfree(e, e->d_ip);
  dfree(e, e->d_logits); dfree(e, e->d_probs);
  int64_t cap = e->host_cap ? e->host_cap : 64;
  while (cap < B) cap *= 2;
  CKR(dalloc(e, &e->d_ids, cap * c.max_pos)); CKR(dalloc(e, &e->d_mask, cap * c.max_pos));
  CKR(dalloc(e, &e->d_px, cap * px_per));
  CKR(dalloc(e, &e->d_tp, cap)); CKR(dalloc(e, &e->d_ip, cap));
  CKR(dalloc(e, &e->d_logits, cap * c.num_outputs)); CKR(dalloc(e, &e->d_probs, cap * c.num_outputs));
  e->host_cap = cap; e->host_S = c.max_pos;
  invalidate_graphs(e);   // graphs captured earlier point into the freed buffers
  return MMCM_OK;
}

// Whole-tower skip (exact): when NO sample of a forward has an image, the vision tower's output cannot reach
// a logit (fusion: feature x image_present = 0, fusion.py:188-189; MTL: image_present < 0.5 selects the text branch,
// multitask.py:194-197), and likewise the text tower when no sample's text can (fusion: text_present < 0.5; MTL: text
// absent AND image present).  The pooled buffer is zero-filled instead; the logits are bit-identical.  The online
// callers build such batches all the time (a text-only or image-only request is B = 1, inference.py:201-211).
// `htp` / `hip`: HOST copies of the presence flags.
static void towers_needed(const Eng* e, const float* htp, const float* hip, int B, bool* text, bool* vision) {
  bool t = false, v = false;
  for (int i = 0; i < B; ++i) {
    v = v || hip[i] >= 0.5f;
    t = t || htp[i] >= 0.5f || (e->cfg.head == MMCM_HEAD_MTL && hip[i] < 0.5f);
  }
  *text = t;
  *vision = v;
}
constexpr int kFlagPeekBatch = 16;   // device-pointer forwards of at most this many samples read their flags back first

static int forward_device(Eng* e, const int64_t* ids, const int64_t* mask, const Pixels& px, const float* tp,
                          const float* ip, int B, int S, float* logits, float* probs, cudaStream_t st) {
  const mmcm_config& c = e->cfg;
  e->stats.launches = 0;
  clear_gemm_events(e->stats);
  if (B == 0) return MMCM_OK;
  bool need_text = true, need_vis = true;
  if (e->opt_skip_absent && B <= kFlagPeekBatch) {
    // small forwards are latency bound and skipping a tower halves them; one 2 x B-float read-back (the callers
    // synchronise for the logits right afterwards anyway).  Not while the stream is being captured into a graph.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      float flags[2 * kFlagPeekBatch];
      CK(cudaMemcpyAsync(flags, tp, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(flags + kFlagPeekBatch, ip, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      towers_needed(e, flags, flags + kFlagPeekBatch, B, &need_text, &need_vis);
    }
  }
  e->fold_forward = e->opt_ln_fold && B >= kLnFoldMinBatch;
  e->split_forward = e->opt_split_k && e->opts.tma_epilogue && B < kLnFoldMinBatch &&
                     (int64_t)B * std::max(e->cfg.max_pos, vis_tokens(e->cfg)) <= kSplitRows;
  const int ct = e->opt_auto_chunk ? choose_chunk(e->text, S, B, e->opt_micro_batch, g_num_sms) : std::min(B, e->opt_micro_batch);
  const int cv = e->opt_auto_chunk ? choose_chunk(e->vis, vis_tokens(c), B, e->opt_micro_batch, g_num_sms)
                                   : std::min(B, e->opt_micro_batch);
  e->last_chunk_text = ct; e->last_chunk_vis = cv;
  CKR(ensure_arenas(e, ct, cv));
  CKR(ensure_batch(e, B));
  const bool two = e->opt_streams >= 2;
  cudaStream_t stx = two ? e->s_text : st, svx = two ? e->s_vis : st;
  if (two) {
    CK(cudaEventRecord(e->ev_fork, st));
    CK(cudaStreamWaitEvent(stx, e->ev_fork, 0));
    CK(cudaStreamWaitEvent(svx, e->ev_fork, 0));
  }
  // enqueue the two towers' chunks alternately so that neither stream starves while the host is still launching
  if (!need_text) CK(cudaMemsetAsync(e->pooled_t, 0, (size_t)B * c.text_hidden * sizeof(float), stx));
  if (!need_vis) CK(cudaMemsetAsync(e->pooled_v, 0, (size_t)B * c.vis_hidden * sizeof(float), svx));
  for (int bt = need_text ? 0 : B, bv = need_vis ? 0 : B; bt < B || bv < B;) {
    if (bt < B) {
      const int n = std::min(ct, B - bt);
      CKR(run_text(e, ids + (int64_t)bt * S, mask ? mask + (int64_t)bt * S : nullptr, n, S,
                   e->pooled_t + (int64_t)bt * c.text_hidden, stx, tp + bt, ip + bt));
      bt += n;
    }
    if (bv < B) {
      const int n = std::min(cv, B - bv);
      CKR(run_vision(e, px.at(bv, c), n, e->pooled_v + (int64_t)bv * c.vis_hidden, svx));
      bv += n;
    }
  }
  if (two) {
    CK(cudaEventRecord(e->ev_text, stx));
    CK(cudaEventRecord(e->ev_vis, svx));
    CK(cudaStreamWaitEvent(st, e->ev_text, 0));
    CK(cudaStreamWaitEvent(st, e->ev_vis, 0));
  }
  CKR(run_head(e, tp, ip, B, logits, probs, st));
  e->last_B = B;
  return MMCM_OK;
}

// ================================================================================================ C ABI
extern "C" {

const char* mmcm_last_error(void) { return g_err; }
const char* mmcm_version(void) { return "mmcm-b200 0.2 (sm_100a; tcgen05/TMEM/TMA)"; }

int mmcm_create(const mmcm_config* cfg, int device, mmcm_handle* out) {
  if (!cfg || !out) return fail(MMCM_EINVAL, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return fail(MMCM_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
  if (device < 0 || device >= ndev) return fail(MMCM_EINVAL, "device %d out of range [0,%d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MMCM_ECUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  if (cfg->backend != MMCM_BACKEND_CLIP && cfg->backend != MMCM_BACKEND_SIGLIP) return fail(MMCM_EINVAL, "bad backend");
  if (cfg->head != MMCM_HEAD_FUSION && cfg->head != MMCM_HEAD_MTL) return fail(MMCM_EINVAL, "bad head");
  if (cfg->image <= 0 || cfg->patch <= 0 || cfg->image % cfg->patch != 0 || cfg->patch % 8 != 0)
    return fail(MMCM_EINVAL, "bad image/patch %d/%d", cfg->image, cfg->patch);
  if (cfg->num_outputs <= 0 || cfg->max_pos <= 0 || cfg->vocab <= 0) return fail(MMCM_EINVAL, "bad sizes");
  if (cfg->max_pos > 256 || vis_tokens(*cfg) > 256) return fail(MMCM_EINVAL, "sequences longer than 256 tokens are unsupported");
  CKR(ensure_driver());
  Eng* e = new Eng();
  e->cfg = *cfg;
  e->device = device;
  e->recording_weights = true;
  int r = setup_weights(e);
  e->recording_weights = false;
  if (r == MMCM_OK) {
    cudaError_t ce = cudaSuccess;
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_text, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_vis, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_text, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_vis, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev_t0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev_copy_end);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_small, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_small_alt, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev_end);
    if (ce != cudaSuccess) r = fail(MMCM_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(ce));
  }
  if (r != MMCM_OK) {
    mmcm_destroy(e);
    return r;
  }
  *out = e;
  return MMCM_OK;
}

int mmcm_destroy(mmcm_handle h) {
  if (!h) return MMCM_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  clear_gemm_events(h->stats);
  for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : h->allocs) cudaFree(p);
  if (h->stage32) cudaFree(h->stage32);
  if (h->upload_buf) cudaFree(h->upload_buf);
  {  // cached tensor maps may point into freed memory that a later allocation re-uses with another shape
    std::lock_guard<std::mutex> lk(g_mu);
    g_tmaps.clear();
  }
  if (h->s_text) cudaStreamDestroy(h->s_text);
  if (h->s_vis) cudaStreamDestroy(h->s_vis);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_text) cudaEventDestroy(h->ev_text);
  if (h->ev_vis) cudaEventDestroy(h->ev_vis);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_copy_end) cudaEventDestroy(h->ev_copy_end);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  for (cudaEvent_t ev : h->ev_chunk) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->ev_chunk_alt) cudaEventDestroy(ev);
  if (h->ev_small) cudaEventDestroy(h->ev_small);
  if (h->ev_small_alt) cudaEventDestroy(h->ev_small_alt);
  delete h;
  return MMCM_OK;
}

static bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

int mmcm_load_weight(mmcm_handle h, const char* key, const float* src, int64_t numel) {
  if (!h || !key || !src) return fail(MMCM_EINVAL, "null argument");
  CK(cudaSetDevice(h->device));
  const std::string k(key);
  auto it = h->slots.find(k);
  if (it == h->slots.end()) {
    // tensors of the reference's state dict that the scoring path never reads
    if (ends_with(k, "logit_scale") || ends_with(k, "logit_bias") || k == "pos_weight" || k == "log_vars" ||
        ends_with(k, "position_ids") || k.rfind("criterion.", 0) == 0)
      return MMCM_OK;
    return fail(MMCM_EINVAL, "unexpected key '%s' for this model configuration", key);
  }
  Slot& s = it->second;
  if (numel != s.numel)
    return fail(MMCM_EINVAL, "size mismatch for '%s': got %lld elements, expected %lld", key, (long long)numel,
                (long long)s.numel);
  // LN-fold inputs go to an fp32 staging block that lives until the next finalize.  If it is gone (a tensor is being
  // replaced after a finalize) it comes back empty, and every staged tensor has to be pushed again before the fold
  // can be redone -- mmcm_finalize_weights reports what is missing.
  if (s.fold) {
    if (!h->stage32) {
      cudaError_t ce = cudaMalloc(reinterpret_cast<void**>(&h->stage32), (size_t)h->stage32_floats * 4);
      if (ce != cudaSuccess) return fail(MMCM_ECUDA, "cudaMalloc of the fold staging block failed: %s", cudaGetErrorString(ce));
      for (auto& kv : h->slots)
        for (const Part& p : kv.second.parts)
          if (p.kind == 2) { kv.second.loaded = false; break; }
    }
    h->fold_pending = true;
  }
  // the source may be host or device memory: host tensors pass through a grow-only device buffer of the handle
  const float* dsrc = src;
  cudaPointerAttributes attr;
  cudaError_t pe = cudaPointerGetAttributes(&attr, src);
  const bool on_device = (pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged));
  if (pe != cudaSuccess) cudaGetLastError();
  if (!on_device) {
    if (numel > h->upload_floats) {
      CK(cudaDeviceSynchronize());
      if (h->upload_buf) cudaFree(h->upload_buf);
      h->upload_buf = nullptr;
      h->upload_floats = 0;
      cudaError_t ce = cudaMalloc(reinterpret_cast<void**>(&h->upload_buf), (size_t)numel * 4);
      if (ce != cudaSuccess) return fail(MMCM_ECUDA, "cudaMalloc staging for '%s' failed: %s", key, cudaGetErrorString(ce));
      h->upload_floats = numel;
    }
    // stream-ordered on the legacy stream behind the previous tensor's repack kernels, which read the same buffer
    cudaError_t ce = cudaMemcpyAsync(h->upload_buf, src, (size_t)numel * 4, cudaMemcpyHostToDevice, nullptr);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(nullptr);   // pageable sources may be re-used by the caller at once
    if (ce != cudaSuccess) return fail(MMCM_ECUDA, "H2D copy of '%s' failed: %s", key, cudaGetErrorString(ce));
    dsrc = h->upload_buf;
  }
  OptsScope scope(&h->opts);
  for (const Part& p : s.parts) {
    const int64_t total = p.rows * p.cols;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    void* dst = p.kind == 2 ? static_cast<void*>(h->stage32 + reinterpret_cast<intptr_t>(p.dst)) : p.dst;
    CK(launch_k(copy2d_kernel, dim3(blocks), dim3(256), 0, nullptr, dsrc + p.src_off, dst, p.rows, p.cols, p.src_pitch,
                p.dst_pitch, p.kind == 1 ? 1 : 0, p.scale));
  }
  CK(cudaGetLastError());   // no device sync per tensor: the repack kernels are stream ordered, finalize waits once
  s.loaded = true;
  h->finalized = false;
  return MMCM_OK;
}

int mmcm_finalize_weights(mmcm_handle h) {
  if (!h) return fail(MMCM_EINVAL, "null handle");
  CK(cudaSetDevice(h->device));
  int missing = 0;
  std::string first;
  for (auto& kv : h->slots)
    if (!kv.second.loaded) {
      if (!missing || kv.first < first) first = kv.first;
      ++missing;
    }
  if (missing) return fail(MMCM_ESTATE, "%d weight tensor(s) missing, e.g. '%s'", missing, first.c_str());
  OptsScope scope(&h->opts);
  if (h->fold_pending) {
    // LayerNorm gamma / beta into the Linears that consume them (rowwise.cuh fold_ln_kernel)
    if (!h->stage32) return fail(MMCM_ESTATE, "internal: LN-fold inputs are gone");
    for (const FoldJob& j : h->fold_jobs)
      CK(launch_k(fold_ln_kernel, dim3((j.N + 7) / 8), dim3(256), 0, nullptr, (const float*)(h->stage32 + j.w_off),
                  (const float*)(h->stage32 + j.b_off), j.gamma, j.beta, j.N, j.K, j.q_rows,
                  j.q_scale, j.wout, (float*)nullptr, j.bias_out));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaFree(h->stage32);
    h->stage32 = nullptr;
    h->fold_pending = false;
  }
  if (h->cfg.backend == MMCM_BACKEND_SIGLIP) {
    const int D = h->cfg.vis_hidden;
    CK(launch_k(probe_query_kernel, dim3((D + 7) / 8), dim3(256), 0, nullptr, h->map_inw, h->map_inb, h->map_probe, h->map_q, D, 1.0f / sqrtf((float)ATT_DH)));
    CK(cudaGetLastError());
  }
  CK(cudaDeviceSynchronize());
  if (h->upload_buf) {
    cudaFree(h->upload_buf);
    h->upload_buf = nullptr;
    h->upload_floats = 0;
  }
  h->finalized = true;
  return MMCM_OK;
}

// ---- packed weight file -----------------------------------------------------------------------
// The repacked weight set (bf16 GEMM operands, Q|K|V fused, dh^-1/2 folded, fp32 everything else) as one blob that a
// later process maps and copies straight into the handle's buffers: no fp32 checkpoint read, no repack kernels, no
// fp32 master copy.  Layout: PackedHeader | sizes[n_bufs] | pad to 4096 | buffers, each 256-byte aligned, in the
// handle's allocation order (a pure function of mmcm_config, which the header carries and the loader compares).
struct PackedHeader {
  char magic[8];            // "MMCMPK02" (02: layer norms folded into the qkv / fc1 operands, column sums stored)
  uint32_t header_bytes;    // sizeof(PackedHeader)
  uint32_t cfg_bytes;       // sizeof(mmcm_config)
  uint64_t n_bufs;
  uint64_t payload_offset;  // from the start of the file
  uint64_t payload_bytes;
  uint64_t checksum;        // packed_checksum over the payload
  mmcm_config cfg;
};
static const char kPackedMagic[8] = {'M', 'M', 'C', 'M', 'P', 'K', '0', '2'};
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static uint64_t packed_checksum(const unsigned char* p, size_t n, uint64_t h = 0x9E3779B97F4A7C15ull) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
  }
  for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001B3ull;
  return h;
}

int mmcm_save_packed(mmcm_handle h, const char* path) {
  if (!h || !path) return fail(MMCM_EINVAL, "null argument");
  if (!h->finalized) return fail(MMCM_ESTATE, "weights are not finalized (nothing to save)");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  PackedHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, kPackedMagic, 8);
  hd.header_bytes = sizeof(PackedHeader);
  hd.cfg_bytes = sizeof(mmcm_config);
  hd.n_bufs = h->wbufs.size();
  hd.cfg = h->cfg;
  std::vector<uint64_t> sizes;
  size_t payload = 0;
  for (const auto& b : h->wbufs) {
    sizes.push_back(b.bytes);
    payload += align_up(b.bytes, 256);
  }
  hd.payload_offset = align_up(sizeof(PackedHeader) + sizes.size() * 8, 4096);
  hd.payload_bytes = payload;
  std::vector<unsigned char> host(payload, 0);
  size_t off = 0;
  for (const auto& b : h->wbufs) {
    CK(cudaMemcpy(host.data() + off, b.ptr, b.bytes, cudaMemcpyDeviceToHost));
    off += align_up(b.bytes, 256);
  }
  hd.checksum = packed_checksum(host.data(), payload);
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(MMCM_EINVAL, "cannot open '%s' for writing: %s", tmp.c_str(), strerror(errno));
  std::vector<unsigned char> head(hd.payload_offset, 0);
  memcpy(head.data(), &hd, sizeof(hd));
  memcpy(head.data() + sizeof(hd), sizes.data(), sizes.size() * 8);
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() && fwrite(host.data(), 1, payload, f) == payload;
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(MMCM_EINVAL, "writing '%s' failed: %s", path, strerror(errno));
  }
  return MMCM_OK;
}

// maps the file read-only; *base / *len describe the mapping the caller must munmap
static int map_packed(const char* path, const unsigned char** base, size_t* len, const PackedHeader** hd) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(MMCM_EINVAL, "cannot open packed weight file '%s': %s", path, strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(PackedHeader)) {
    close(fd);
    return fail(MMCM_EINVAL, "'%s' is not a packed weight file (too short)", path);
  }
  void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail(MMCM_EINVAL, "mmap of '%s' failed: %s", path, strerror(errno));
  const PackedHeader* h = reinterpret_cast<const PackedHeader*>(m);
  const size_t n = (size_t)st.st_size;
  if (memcmp(h->magic, kPackedMagic, 8) != 0 || h->header_bytes != sizeof(PackedHeader) ||
      h->cfg_bytes != sizeof(mmcm_config) || h->n_bufs > (1u << 20) ||
      sizeof(PackedHeader) + h->n_bufs * 8 > h->payload_offset || h->payload_offset > n ||
      h->payload_bytes != n - h->payload_offset) {
    munmap(m, n);
    return fail(MMCM_EINVAL, "'%s' is not a packed weight file of this library version (bad header)", path);
  }
  *base = reinterpret_cast<const unsigned char*>(m);
  *len = n;
  *hd = h;
  return MMCM_OK;
}

int mmcm_packed_config(const char* path, mmcm_config* cfg_out) {
  if (!path || !cfg_out) return fail(MMCM_EINVAL, "null argument");
  const unsigned char* base = nullptr;
  size_t len = 0;
  const PackedHeader* hd = nullptr;
  CKR(map_packed(path, &base, &len, &hd));
  *cfg_out = hd->cfg;
  munmap(const_cast<unsigned char*>(base), len);
  return MMCM_OK;
}

int mmcm_load_packed(mmcm_handle h, const char* path) {
  if (!h || !path) return fail(MMCM_EINVAL, "null argument");
  CK(cudaSetDevice(h->device));
  const unsigned char* base = nullptr;
  size_t len = 0;
  const PackedHeader* hd = nullptr;
  CKR(map_packed(path, &base, &len, &hd));
  int rc = MMCM_OK;
  const uint64_t* sizes = reinterpret_cast<const uint64_t*>(base + sizeof(PackedHeader));
  if (memcmp(&hd->cfg, &h->cfg, sizeof(mmcm_config)) != 0)
    rc = fail(MMCM_EINVAL, "'%s' was packed for a different model configuration than this handle", path);
  else if (hd->n_bufs != h->wbufs.size())
    rc = fail(MMCM_EINVAL, "'%s' holds %llu buffers, this library version lays out %zu", path,
              (unsigned long long)hd->n_bufs, h->wbufs.size());
  size_t total = 0;
  for (size_t i = 0; rc == MMCM_OK && i < h->wbufs.size(); ++i) {
    if (sizes[i] != h->wbufs[i].bytes) rc = fail(MMCM_EINVAL, "'%s': buffer %zu has another size than this library lays out", path, i);
    total += align_up(h->wbufs[i].bytes, 256);
  }
  if (rc == MMCM_OK && total != hd->payload_bytes) rc = fail(MMCM_EINVAL, "'%s': payload size mismatch", path);
  if (rc == MMCM_OK && packed_checksum(base + hd->payload_offset, hd->payload_bytes) != hd->checksum)
    rc = fail(MMCM_EINVAL, "'%s': checksum mismatch (file is corrupt)", path);
  if (rc == MMCM_OK) {
    cudaError_t ce = cudaDeviceSynchronize();
    size_t off = hd->payload_offset;
    for (size_t i = 0; ce == cudaSuccess && i < h->wbufs.size(); ++i) {
      ce = cudaMemcpy(h->wbufs[i].ptr, base + off, h->wbufs[i].bytes, cudaMemcpyHostToDevice);
      off += align_up(h->wbufs[i].bytes, 256);
    }
    if (ce != cudaSuccess) rc = fail(MMCM_ECUDA, "H2D copy of packed weights failed: %s", cudaGetErrorString(ce));
  }
  munmap(const_cast<unsigned char*>(base), len);
  CKR(rc);
  for (auto& kv : h->slots) kv.second.loaded = true;
  if (h->stage32) {   // a half-pushed state dict is superseded by the file (which holds the folded weights)
    cudaFree(h->stage32);
    h->stage32 = nullptr;
  }
  h->fold_pending = false;
  h->finalized = true;
  return MMCM_OK;
}

// Small batches are launch bound (~180 launches, 1.5 ms of host time for < 0.3 ms of GPU work): replay the whole
// forward as one CUDA graph.  Inputs are copied into handle-owned buffers (graphs bake pointers), the first two calls
// of a shape run eagerly (allocations, function attributes, TMA maps), the third is captured.
static int forward_graphed(Eng* e, const int64_t* ids, const int64_t* mask, const float* px, const float* tp,
                           const float* ip, int B, int S, float* logits, float* probs, cudaStream_t st, bool* handled) {
  *handled = false;
  const mmcm_config& c = e->cfg;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return MMCM_OK;  // caller captures
  auto key = std::make_tuple(B, S, mask ? 1 : 0, probs ? 1 : 0);
  Eng::GraphEntry& g = e->graphs[key];
  if (!g.exec && g.warm < 2) { g.warm++; return MMCM_OK; }   // eager warm-up calls
  CKR(ensure_static_io(e, B));
  Eng::GraphEntry& ge = e->graphs[key];                      // ensure_static_io may have cleared the map
  const int64_t px_per = (int64_t)3 * c.image * c.image;
  const int C = c.num_outputs;
  CK(cudaMemcpyAsync(e->d_ids, ids, (size_t)B * S * 8, cudaMemcpyDeviceToDevice, st));
  if (mask) CK(cudaMemcpyAsync(e->d_mask, mask, (size_t)B * S * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(e->d_px, px, (size_t)B * px_per * 4, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(e->d_tp, tp, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(e->d_ip, ip, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  if (!ge.exec) {
    if (ge.warm >= 1000) return MMCM_OK;                     // capture failed before: stay eager for this shape
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int r = MMCM_OK;
    if (ce == cudaSuccess) {
      Pixels spx;
      spx.f32 = e->d_px;
      r = forward_device(e, e->d_ids, mask ? e->d_mask : nullptr, spx, e->d_tp, e->d_ip, B, S, e->d_logits,
                         probs ? e->d_probs : nullptr, st);
      ce = cudaStreamEndCapture(st, &graph);
    }
    if (r == MMCM_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&ge.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (r != MMCM_OK || ce != cudaSuccess || !ge.exec) {
      cudaGetLastError();
      ge.exec = nullptr;
      ge.warm = 1000;
      return MMCM_OK;                                        // eager path below handles this call
    }
    ge.launches = e->stats.launches;
  }
  CK(cudaGraphLaunch(ge.exec, st));
  e->stats.launches = ge.launches;
  e->last_B = B;
  CK(cudaMemcpyAsync(logits, e->d_logits, (size_t)B * C * 4, cudaMemcpyDeviceToDevice, st));
  if (probs) CK(cudaMemcpyAsync(probs, e->d_probs, (size_t)B * C * 4, cudaMemcpyDeviceToDevice, st));
  *handled = true;
  return MMCM_OK;
}

int mmcm_forward(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask, const float* pixel_values,
                 const float* text_present, const float* image_present, int32_t B, int32_t S, float* logits_out,
                 float* probs_out, void* stream) {
  CKR(check_forward_args(h, input_ids, pixel_values, text_present, image_present, B, S, logits_out));
  CK(cudaSetDevice(h->device));
  OptsScope scope(&h->opts);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B > 0 && B <= h->opt_graph_max_batch && !h->stats.time_gemms) {
    bool handled = false;
    CKR(forward_graphed(h, input_ids, attention_mask, pixel_values, text_present, image_present, B, S, logits_out,
                        probs_out, st, &handled));
    if (handled) return MMCM_OK;
  }
  Pixels px;
  px.f32 = pixel_values;
  return forward_device(h, input_ids, attention_mask, px, text_present, image_present, B, S, logits_out, probs_out, st);
}

static int make_u8_pixels(Pixels* px, const uint8_t* pixels_u8, const float* mean3, const float* std3) {
  if (!mean3 || !std3) return fail(MMCM_EINVAL, "null mean / std");
  px->u8 = pixels_u8;
  for (int i = 0; i < 3; ++i) {
    if (!(std3[i] > 0.f)) return fail(MMCM_EINVAL, "std[%d] must be positive", i);
    px->mean[i] = mean3[i];
    px->std[i] = std3[i];
  }
  return MMCM_OK;
}

int mmcm_forward_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask, const uint8_t* pixels_u8,
                    const float* mean3, const float* std3, const float* text_present, const float* image_present,
                    int32_t B, int32_t S, float* logits_out, float* probs_out, void* stream) {
  CKR(check_forward_args(h, input_ids, pixels_u8, text_present, image_present, B, S, logits_out));
  Pixels px;
  CKR(make_u8_pixels(&px, pixels_u8, mean3, std3));
  CK(cudaSetDevice(h->device));
  OptsScope scope(&h->opts);
  return forward_device(h, input_ids, attention_mask, px, text_present, image_present, B, S, logits_out, probs_out,
                        reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// host buffers in, host logits out; `hpx` points at HOST pixels (fp32 CHW or uint8 HWC)
// Chunks of the text tower and H2D pipeline stages (= vision chunks) of one host call
static void plan_host_call(Eng* e, const Pixels& hpx, int B, int S, Eng::HostPlan* plan, bool prefetched = false) {
  const mmcm_config& c = e->cfg;
  const size_t px_bytes = hpx.bytes_per_sample(c);
  const int ct = e->opt_auto_chunk ? choose_chunk(e->text, S, B, e->opt_micro_batch, g_num_sms) : std::min((int)B, e->opt_micro_batch);
  // the vision chunks double as the H2D pipeline stages (fp32 pixels are 602 KB/sample): the copy of chunk i+1 overlaps
  // the towers of chunk i; the text tower does not wait for pixels at all
  // stage size: ~200 MB of pixels (measured on B=1024: fp32 pixels best at 342 samples per stage, 51.6 k vs 50.3 k at
  // 256; uint8 pixels best unsplit, 53.4 k vs 49.9 k -- tools/e2e_sweep.py); option "host_chunk" overrides
  int auto_chunk = (int)std::max<int64_t>(64, ((int64_t)200 << 20) / (int64_t)px_bytes);
  // a batch that would be ONE stage gets two (2/3 + 1/3) once it is big enough for the second copy to be worth hiding:
  // batch 256, fp32 pixels: 43.7 k -> 46.7 k samples/s; below ~192 samples one stage is better (tools/e2e_sweep.py,
  // profiles/r02_e2e_stage_sweep.txt)
  if (!hpx.u8 && B >= 192 && B <= auto_chunk) auto_chunk = (2 * B + 2) / 3;
  // a prefetched batch was shipped during the PREVIOUS forward: nothing to hide behind the towers, so the vision tower
  // runs in the chunks a device-resident forward would use (fuller GEMM rounds) unless "host_chunk" says otherwise
  if (prefetched) auto_chunk = e->opt_micro_batch;
  const int vcap = std::min(e->opt_micro_batch, e->opt_host_chunk > 0 ? e->opt_host_chunk : auto_chunk);
  const int cv = e->opt_auto_chunk ? choose_chunk(e->vis, vis_tokens(c), B, vcap, g_num_sms) : std::min((int)B, vcap);
  // Stage schedule.  When the copy stream is the bottleneck (N ranks sharing a host that cannot feed N x 32 GB/s: the
  // copies of the previous call took > 85 % of its time) the step is "all copies, then the towers of the LAST stage":
  // taper the last stages (cv, ..., cv/2, cv/4, cv/4) so that little work is left when the last bytes land.  When the
  // towers are the bottleneck equal stages are better (fewer, fuller GEMM rounds), so the schedule follows what the
  // previous call measured (8 GPUs, 23 GB/s per rank: 287 k -> 295 k samples/s, profiles/r02_bench_8gpu*.json).  A small
  // FIRST stage, so that the vision tower starts sooner, measured slower on one GPU (55.2 k -> 52.1-53.3 k for 48-171
  // samples, profiles/r02_e2e_stage_sweep.txt): not done.
  std::vector<int>& stages = plan->stages;
  stages.clear();
  {
    int left = B;
    const bool taper = e->h2d_bound && e->opt_host_chunk == 0 && B >= 2 * cv;
    while (left > 0) {
      int n = std::min(cv, left);
      if (taper && left <= cv + cv / 2) n = (left > cv / 3) ? std::max(cv / 4, (left + 1) / 2) : left;
      n = std::min(n, left);
      stages.push_back(n);
      left -= n;
    }
  }
  plan->ct = ct;
  plan->cv = cv;
}

static int ensure_alt_io(Eng* e, int64_t B) {
  const mmcm_config& c = e->cfg;
  CKR(ensure_static_io(e, B));
  if (e->alt_cap >= e->host_cap) return MMCM_OK;
  const int64_t px_per = (int64_t)3 * c.image * c.image;
  CK(cudaDeviceSynchronize());
  dfree(e, e->a_ids); dfree(e, e->a_mask); dfree(e, e->a_px); dfree(e, e->a_tp); dfree(e, e->a_ip);
  const int64_t cap = e->host_cap;
  CKR(dalloc(e, &e->a_ids, cap * c.max_pos)); CKR(dalloc(e, &e->a_mask, cap * c.max_pos));
  CKR(dalloc(e, &e->a_px, cap * px_per));
  CKR(dalloc(e, &e->a_tp, cap)); CKR(dalloc(e, &e->a_ip, cap));
  e->alt_cap = cap;
  return MMCM_OK;
}

static bool prefetch_match(const Eng::Prefetched& f, const void* ids, const void* mask, const Pixels& hpx, const void* tp,
                           const void* ip, int B, int S) {
  const void* px = hpx.u8 ? static_cast<const void*>(hpx.u8) : static_cast<const void*>(hpx.f32);
  return f.valid && f.ids == ids && f.mask == mask && f.px == px && f.tp == tp && f.ip == ip && f.B == B && f.S == S &&
         f.u8 == (hpx.u8 != nullptr);
}
// The prefetched batch becomes the current input set (its copies may still be in flight: the forward waits on the
// set's stage events); the old current set -- consumed by a forward that has returned -- becomes the free second set.
static void promote_prefetched(Eng* e) {
  std::swap(e->d_ids, e->a_ids); std::swap(e->d_mask, e->a_mask); std::swap(e->d_px, e->a_px);
  std::swap(e->d_tp, e->a_tp); std::swap(e->d_ip, e->a_ip); std::swap(e->host_cap, e->alt_cap);
  std::swap(e->ev_chunk, e->ev_chunk_alt);
  std::swap(e->ev_small, e->ev_small_alt);
  e->ready = std::move(e->pf);
  e->pf.valid = false;
  invalidate_graphs(e);   // captured graphs point into what is now the second set
}

// mmcm_prefetch_host*: ship the NEXT batch into the second input set on the copy stream, with the stage events the
// forward that consumes it will wait on.  The caller's forward of the CURRENT batch is issued afterwards and overlaps.
static int prefetch_impl(Eng* e, const int64_t* input_ids, const int64_t* attention_mask, const Pixels& hpx,
                         const float* text_present, const float* image_present, int32_t B, int32_t S) {
  CK(cudaSetDevice(e->device));
  if (B == 0) return MMCM_OK;
  // a batch prefetched earlier and not consumed yet moves into the current set; the second set is then free
  if (e->pf.valid) promote_prefetched(e);
  const mmcm_config& c = e->cfg;
  const size_t px_bytes = hpx.bytes_per_sample(c);
  const char* hsrc = hpx.u8 ? reinterpret_cast<const char*>(hpx.u8) : reinterpret_cast<const char*>(hpx.f32);
  if (B > e->host_cap || e->alt_cap < e->host_cap) {   // (re)allocation synchronises the device and moves the buffers
    e->ready.valid = false;
    CKR(ensure_alt_io(e, B));
  }
  Eng::Prefetched& f = e->pf;
  plan_host_call(e, hpx, B, S, &f.plan, true);
  bool need_text = true, need_vis = true;
  if (e->opt_skip_absent) towers_needed(e, text_present, image_present, B, &need_text, &need_vis);
  CK(cudaMemcpyAsync(e->a_ids, input_ids, (size_t)B * S * 8, cudaMemcpyHostToDevice, e->s_copy));
  if (attention_mask) CK(cudaMemcpyAsync(e->a_mask, attention_mask, (size_t)B * S * 8, cudaMemcpyHostToDevice, e->s_copy));
  CK(cudaMemcpyAsync(e->a_tp, text_present, (size_t)B * 4, cudaMemcpyHostToDevice, e->s_copy));
  CK(cudaMemcpyAsync(e->a_ip, image_present, (size_t)B * 4, cudaMemcpyHostToDevice, e->s_copy));
  CK(cudaEventRecord(e->ev_small_alt, e->s_copy));
  const int nchunks = (int)f.plan.stages.size();
  while ((int)e->ev_chunk_alt.size() < nchunks) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev_chunk_alt.push_back(ev);
  }
  for (int ci = 0, b0 = 0; ci < nchunks; b0 += f.plan.stages[ci], ++ci) {
    const int n = f.plan.stages[ci];
    if (need_vis)
      CK(cudaMemcpyAsync(reinterpret_cast<char*>(e->a_px) + b0 * px_bytes, hsrc + b0 * px_bytes, (size_t)n * px_bytes,
                         cudaMemcpyHostToDevice, e->s_copy));
    CK(cudaEventRecord(e->ev_chunk_alt[ci], e->s_copy));
  }
  f.ids = input_ids; f.mask = attention_mask; f.tp = text_present; f.ip = image_present;
  f.px = hpx.u8 ? static_cast<const void*>(hpx.u8) : static_cast<const void*>(hpx.f32);
  f.B = B; f.S = S; f.u8 = hpx.u8 != nullptr; f.need_vis = need_vis;
  f.valid = true;
  return MMCM_OK;
}

static int forward_host_impl(Eng* e, const int64_t* input_ids, const int64_t* attention_mask, const Pixels& hpx,
                             const float* text_present, const float* image_present, int32_t B, int32_t S,
                             float* logits_out, float* probs_out, cudaStream_t st) {
  CK(cudaSetDevice(e->device));
  if (B == 0) return MMCM_OK;
  OptsScope scope(&e->opts);
  e->fold_forward = e->opt_ln_fold && B >= kLnFoldMinBatch;
  e->split_forward = e->opt_split_k && e->opts.tma_epilogue && B < kLnFoldMinBatch &&
                     (int64_t)B * std::max(e->cfg.max_pos, vis_tokens(e->cfg)) <= kSplitRows;
  const mmcm_config& c = e->cfg;
  const size_t px_bytes = hpx.bytes_per_sample(c);
  const char* hsrc = hpx.u8 ? reinterpret_cast<const char*>(hpx.u8) : reinterpret_cast<const char*>(hpx.f32);
  const int C = c.num_outputs;
  CKR(ensure_static_io(e, B));
  Pixels dpx = hpx;                                          // device-side staging: the fp32 buffer doubles for uint8
  if (hpx.u8) dpx.u8 = reinterpret_cast<const uint8_t*>(e->d_px); else dpx.f32 = e->d_px;
  // small inputs first, then the pixels in micro-batch chunks on the copy stream so that the H2D transfer of
  // chunk i+1 overlaps the towers of chunk i
  // prefetched (mmcm_prefetch_host*)?  Either already promoted into the current set, or still in the second one
  if (!prefetch_match(e->ready, input_ids, attention_mask, hpx, text_present, image_present, B, S) &&
      prefetch_match(e->pf, input_ids, attention_mask, hpx, text_present, image_present, B, S))
    promote_prefetched(e);
  const bool hit = prefetch_match(e->ready, input_ids, attention_mask, hpx, text_present, image_present, B, S) &&
                   B <= e->host_cap;
  Eng::HostPlan plan;
  if (hit) {   // the inputs are on the device or on their way: no copies, wait for the set's events
    plan = std::move(e->ready.plan);
    if (hpx.u8) dpx.u8 = reinterpret_cast<const uint8_t*>(e->d_px); else dpx.f32 = e->d_px;
  } else {
    plan_host_call(e, hpx, B, S, &plan);
  }
  e->ready.valid = false;   // consumed (hit) or overwritten below (miss)
  const int ct = plan.ct, cv = plan.cv;
  const std::vector<int>& stages = plan.stages;
  e->last_chunk_text = ct; e->last_chunk_vis = cv;
  CKR(ensure_arenas(e, ct, cv));
  CKR(ensure_batch(e, B));
  if (hit) CK(cudaStreamWaitEvent(st, e->ev_small, 0));
  else {
    CK(cudaMemcpyAsync(e->d_ids, input_ids, (size_t)B * S * 8, cudaMemcpyHostToDevice, st));
    if (attention_mask) CK(cudaMemcpyAsync(e->d_mask, attention_mask, (size_t)B * S * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_tp, text_present, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_ip, image_present, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  }
  const int nchunks = (int)stages.size();
  while ((int)e->ev_chunk.size() < nchunks) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev_chunk.push_back(ev);
  }
  bool need_text = true, need_vis = true;
  if (e->opt_skip_absent) towers_needed(e, text_present, image_present, B, &need_text, &need_vis);   // host flags: no read-back
  CK(cudaEventRecord(e->ev_fork, st));
  CK(cudaEventRecord(e->ev_t0, st));
  CK(cudaStreamWaitEvent(e->s_copy, e->ev_fork, 0));
  for (int ci = 0, b0 = 0; !hit && need_vis && ci < nchunks; b0 += stages[ci], ++ci) {   // no images at all: nothing to ship
    const int n = stages[ci];
    CK(cudaMemcpyAsync(reinterpret_cast<char*>(e->d_px) + b0 * px_bytes, hsrc + b0 * px_bytes, (size_t)n * px_bytes,
                       cudaMemcpyHostToDevice, e->s_copy));
    CK(cudaEventRecord(e->ev_chunk[ci], e->s_copy));
  }
  CK(cudaEventRecord(e->ev_copy_end, e->s_copy));
  e->stats.launches = 0;
  clear_gemm_events(e->stats);
  CK(cudaStreamWaitEvent(e->s_text, e->ev_fork, 0));
  CK(cudaStreamWaitEvent(e->s_vis, e->ev_fork, 0));
  const int64_t* dmask = attention_mask ? e->d_mask : nullptr;
  if (!need_text) CK(cudaMemsetAsync(e->pooled_t, 0, (size_t)B * c.text_hidden * sizeof(float), e->s_text));
  if (!need_vis) CK(cudaMemsetAsync(e->pooled_v, 0, (size_t)B * c.vis_hidden * sizeof(float), e->s_vis));
  for (int bt = need_text ? 0 : B, bv = need_vis ? 0 : B, ci = 0; bt < B || bv < B;) {
    if (bt < B) {
      const int n = std::min(ct, (int)B - bt);
      CKR(run_text(e, e->d_ids + (int64_t)bt * S, dmask ? dmask + (int64_t)bt * S : nullptr, n, S,
                   e->pooled_t + (int64_t)bt * c.text_hidden, e->s_text, e->d_tp + bt, e->d_ip + bt));
      bt += n;
    }
    if (bv < B) {
      const int n = stages[ci];
      CK(cudaStreamWaitEvent(e->s_vis, e->ev_chunk[ci], 0));   // (prefetched: recorded by mmcm_prefetch_host*)
      CKR(run_vision(e, dpx.at(bv, c), n, e->pooled_v + (int64_t)bv * c.vis_hidden, e->s_vis));
      bv += n;
      ++ci;
    }
  }
  CK(cudaEventRecord(e->ev_text, e->s_text));
  CK(cudaEventRecord(e->ev_vis, e->s_vis));
  CK(cudaStreamWaitEvent(st, e->ev_text, 0));
  CK(cudaStreamWaitEvent(st, e->ev_vis, 0));
  CKR(run_head(e, e->d_tp, e->d_ip, B, e->d_logits, probs_out ? e->d_probs : nullptr, st));
  e->last_B = B;
  CK(cudaMemcpyAsync(logits_out, e->d_logits, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
  if (probs_out) CK(cudaMemcpyAsync(probs_out, e->d_probs, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(e->ev_end, st));
  CK(cudaStreamSynchronize(st));
  if (!hit) {  // what bounded this call: the copy stream's share of the wall time (hysteresis: on above 85 %, off below 70 %)
    float copy_ms = 0.f, total_ms = 0.f;
    if (cudaEventElapsedTime(&copy_ms, e->ev_t0, e->ev_copy_end) == cudaSuccess &&
        cudaEventElapsedTime(&total_ms, e->ev_t0, e->ev_end) == cudaSuccess && total_ms > 0.f) {
      e->last_h2d_share = copy_ms / total_ms;
      if (e->last_h2d_share > 0.85f) e->h2d_bound = true;
      else if (e->last_h2d_share < 0.70f) e->h2d_bound = false;
    } else {
      cudaGetLastError();
    }
  }
  return MMCM_OK;
}

extern "C" {

int mmcm_forward_host(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                      const float* pixel_values, const float* text_present, const float* image_present, int32_t B,
                      int32_t S, float* logits_out, float* probs_out, void* stream) {
  CKR(check_forward_args(h, input_ids, pixel_values, text_present, image_present, B, S, logits_out));
  Pixels px;
  px.f32 = pixel_values;
  return forward_host_impl(h, input_ids, attention_mask, px, text_present, image_present, B, S, logits_out, probs_out,
                           reinterpret_cast<cudaStream_t>(stream));
}

int mmcm_forward_host_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                         const uint8_t* pixels_u8, const float* mean3, const float* std3, const float* text_present,
                         const float* image_present, int32_t B, int32_t S, float* logits_out, float* probs_out,
                         void* stream) {
  CKR(check_forward_args(h, input_ids, pixels_u8, text_present, image_present, B, S, logits_out));
  Pixels px;
  CKR(make_u8_pixels(&px, pixels_u8, mean3, std3));
  return forward_host_impl(h, input_ids, attention_mask, px, text_present, image_present, B, S, logits_out, probs_out,
                           reinterpret_cast<cudaStream_t>(stream));
}

int mmcm_prefetch_host(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask, const float* pixel_values,
                       const float* text_present, const float* image_present, int32_t B, int32_t S) {
  CKR(check_forward_args(h, input_ids, pixel_values, text_present, image_present, B, S, h));
  Pixels px;
  px.f32 = pixel_values;
  return prefetch_impl(h, input_ids, attention_mask, px, text_present, image_present, B, S);
}

int mmcm_prefetch_host_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                          const uint8_t* pixels_u8, const float* text_present, const float* image_present, int32_t B,
                          int32_t S) {
  CKR(check_forward_args(h, input_ids, pixels_u8, text_present, image_present, B, S, h));
  Pixels px;
  px.u8 = pixels_u8;
  return prefetch_impl(h, input_ids, attention_mask, px, text_present, image_present, B, S);
}

int mmcm_get_stage(mmcm_handle h, const char* name, float* dst, int64_t capacity, int64_t* numel_out, void* stream) {
  if (!h || !name || !numel_out) return fail(MMCM_EINVAL, "null argument");
  CK(cudaSetDevice(h->device));
  const std::string n(name);
  const float* src = nullptr;
  int64_t count = 0;
  if (n == "text_pooled") { src = h->pooled_t; count = (int64_t)h->last_B * h->cfg.text_hidden; }
  else if (n == "vision_pooled") { src = h->pooled_v; count = (int64_t)h->last_B * h->cfg.vis_hidden; }
  else if (n == "text_hidden") { src = h->at.x; count = (int64_t)h->last_text_rows * h->cfg.text_hidden; }
  else if (n == "vision_hidden") { src = h->av.x; count = (int64_t)h->last_vis_rows * h->cfg.vis_hidden; }
  else if (n == "text_feat" || n == "vision_feat") {
    if (!h->opt_debug_feats || h->cfg.head != MMCM_HEAD_FUSION)
      return fail(MMCM_ESTATE, "stage '%s' needs option debug_feats=1 and a fusion head", name);
    const bool t = n == "text_feat";
    src = t ? h->feat_t : h->feat_v;
    const int w = (h->cfg.backend == MMCM_BACKEND_CLIP) ? h->cfg.proj_dim : (t ? h->cfg.proj_dim : h->cfg.vis_hidden);
    count = (int64_t)h->last_B * w;
  } else return fail(MMCM_EINVAL, "unknown stage '%s'", name);
  *numel_out = count;
  if (!dst) return MMCM_OK;  // size query
  if (capacity < count) return fail(MMCM_EINVAL, "stage '%s' needs %lld elements, buffer has %lld", name,
                                    (long long)count, (long long)capacity);
  if (count > 0 && !src) return fail(MMCM_ESTATE, "no forward has run yet");
  if (count > 0)
    CK(cudaMemcpyAsync(dst, src, (size_t)count * 4, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return MMCM_OK;
}

int64_t mmcm_last_launch_count(mmcm_handle h) { return h ? h->stats.launches : 0; }

double mmcm_last_host_copy_share(mmcm_handle h) { return h ? (double)h->last_h2d_share : 0.0; }

int mmcm_last_chunks(mmcm_handle h, int32_t* text_chunk_out, int32_t* vision_chunk_out) {
  if (!h) return fail(MMCM_EINVAL, "null handle");
  if (text_chunk_out) *text_chunk_out = h->last_chunk_text;
  if (vision_chunk_out) *vision_chunk_out = h->last_chunk_vis;
  return MMCM_OK;
}

// algorithmic DRAM bytes of one GEMM launch: A (bf16), W (bf16) and what the epilogue reads / writes per output element
static double gemm_algorithmic_bytes(int epi, double m, int N, int K) {
  static const double out_bytes[] = {2, 2, 8 /* fp32 read + write */, 4, 10 /* fp32 read + write, bf16 copy */, 2, 2};
  return m * K * 2.0 + (double)N * K * 2.0 + m * N * out_bytes[epi];
}

static int gemm_time_impl(mmcm_handle h, int epi, double* ms_out, double* flops_out, int64_t* launches_out,
                          double* bytes_out = nullptr, int N = 0, int K = 0) {
  if (!h) return fail(MMCM_EINVAL, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  double ms = 0, flops = 0, bytes = 0;
  int64_t n = 0;
  for (auto& p : h->stats.gemm_events) {
    if (epi >= 0 && p.epi != epi) continue;
    if ((N > 0 && p.N != N) || (K > 0 && p.K != K)) continue;
    float t = 0;
    CK(cudaEventElapsedTime(&t, p.e0, p.e1));
    ms += t;
    int m = p.m_host;
    if (p.m_dev) {   // packed text chunk: count the rows that were actually multiplied
      CK(cudaMemcpy(&m, p.m_dev, sizeof(int), cudaMemcpyDeviceToHost));
      if (m > p.m_host) m = p.m_host;
    }
    flops += p.flops_per_row * m;
    bytes += gemm_algorithmic_bytes(p.epi, m, p.N, p.K);
    ++n;
  }
  if (bytes_out) *bytes_out = bytes;
  if (ms_out) *ms_out = ms;
  if (flops_out) *flops_out = flops;
  if (launches_out) *launches_out = n;
  return MMCM_OK;
}

int mmcm_gemm_time(mmcm_handle h, double* ms_out, double* flops_out, int64_t* launches_out) {
  return gemm_time_impl(h, -1, ms_out, flops_out, launches_out);
}

int mmcm_gemm_time_epi(mmcm_handle h, int32_t epilogue, double* ms_out, double* flops_out, double* bytes_out,
                       int64_t* launches_out) {
  if (epilogue < 0 || epilogue > EPI_LNFOLD_ACT_BF16) return fail(MMCM_EINVAL, "unknown epilogue %d", epilogue);
  return gemm_time_impl(h, epilogue, ms_out, flops_out, launches_out, bytes_out);
}

int mmcm_gemm_time_shape(mmcm_handle h, int32_t epilogue, int32_t N, int32_t K, double* ms_out, double* flops_out,
                         double* bytes_out, int64_t* launches_out) {
  if (epilogue < -1 || epilogue > EPI_LNFOLD_ACT_BF16) return fail(MMCM_EINVAL, "unknown epilogue %d", epilogue);
  return gemm_time_impl(h, epilogue, ms_out, flops_out, launches_out, bytes_out, N, K);
}

int mmcm_set_option(mmcm_handle h, const char* name, int64_t value) {
  if (!name) return fail(MMCM_EINVAL, "null argument");
  if (!h) {   // defaults of the stand-alone kernels (mmcm_gemm_bf16, mmcm_attention, ...): no handle to carry them
    std::lock_guard<std::mutex> lk(g_mu);
    const std::string d(name);
    if (d == "pdl") g_default_opts.pdl = value != 0;
    else if (d == "tma_epilogue") g_default_opts.tma_epilogue = value != 0;
    else if (d == "attention_impl" && value >= 0 && value <= 3) g_default_opts.attention_impl = (int)value;
    else if (d == "attention_ring") g_default_opts.attention_ring = value != 0;
    else if (d == "narrow_tiles") g_default_opts.narrow_tiles = value != 0;
    else return fail(MMCM_EINVAL, "option '%s' = %lld cannot be set without a handle", name, (long long)value);
    return MMCM_OK;
  }
  const std::string n(name);
  if (n == "time_gemms") h->stats.time_gemms = value != 0;
  else if (n == "gemm_impl") {
    if (value < 0 || value > 2)
      return fail(MMCM_EINVAL, "gemm_impl must be 0 (tcgen05 CTA pair), 1 (SIMT validation) or 2 (tcgen05 single CTA)");
    h->opt_gemm_impl = (int)value;
  } else if (n == "micro_batch") {
    if (value < 1 || value > 65536) return fail(MMCM_EINVAL, "micro_batch out of range");
    h->opt_micro_batch = (int)value;
  } else if (n == "streams") {
    if (value != 1 && value != 2) return fail(MMCM_EINVAL, "streams must be 1 or 2");
    h->opt_streams = (int)value;
  } else if (n == "pdl") h->opts.pdl = value != 0;
  else if (n == "tma_epilogue") h->opts.tma_epilogue = value != 0;
  else if (n == "attention_impl") {
    if (value < 0 || value > 3)
      return fail(MMCM_EINVAL, "attention_impl must be 0 (auto), 1 (mma.sync), 2 (tcgen05 for T <= 256) or 3 (TMA-ring mma.sync)");
    h->opts.attention_impl = (int)value;
  } else if (n == "attention_ring") h->opts.attention_ring = value != 0;
  else if (n == "ln_fold") h->opt_ln_fold = value != 0;
  else if (n == "head_cluster") h->opt_head_cluster = value != 0;
  else if (n == "split_k") h->opt_split_k = value != 0;
  else if (n == "skip_absent_text") h->opt_skip_absent = value != 0;
  else if (n == "narrow_tiles") h->opts.narrow_tiles = value != 0;
  else if (n == "debug_feats") h->opt_debug_feats = value != 0;
  else if (n == "auto_chunk") h->opt_auto_chunk = value != 0;
  else if (n == "varlen_text") h->opt_varlen_text = value != 0;
  else if (n == "pooled_last_layer") h->opt_pooled_last = value != 0;
  else if (n == "host_chunk") {
    if (value < 0 || value > 65536) return fail(MMCM_EINVAL, "host_chunk out of range");
    h->opt_host_chunk = (int)value;
  }
  else if (n == "pairs_text" || n == "pairs_vision") {
    if (value < 0 || value > 74) return fail(MMCM_EINVAL, "%s must be in [0, 74]", name);
    (n == "pairs_text" ? h->opt_pairs_text : h->opt_pairs_vis) = (int)value;
  }
  else if (n == "graph_max_batch") {
    if (value < 0 || value > 4096) return fail(MMCM_EINVAL, "graph_max_batch out of range");
    h->opt_graph_max_batch = (int)value;
  }
  else return fail(MMCM_EINVAL, "unknown option '%s'", name);
  invalidate_graphs(h);   // captured graphs bake the launch sequence of the options they were recorded under
  return MMCM_OK;
}

// dev tool: per-CTA clock64 stamps of the CTA-pair GEMM (16 slots per CTA; see trace_stamp in gemm_tcgen05.cuh)
int mmcm_debug_set_gemm_trace(void* device_buffer) {
  g_gemm_trace = reinterpret_cast<long long*>(device_buffer);
  return MMCM_OK;
}

// ---------------------------------------------------------------------------------- stand-alone kernels
int mmcm_gemm_bf16(const void* A, const void* W, const float* bias, int32_t M, int32_t N, int32_t K, int32_t epilogue,
                   int32_t act, void* out, const float* resid, const float* pos, int32_t P, int32_t T, int32_t impl,
                   void* stream) {
  if (!A || !W || !out) return fail(MMCM_EINVAL, "null pointer");
  if (epilogue == EPI_PATCH_F32 && (!pos || P <= 0 || T < P)) return fail(MMCM_EINVAL, "patch epilogue needs pos, P, T");
  EpiParams ep{};
  ep.bias = bias; ep.out = out; ep.resid = resid; ep.pos = pos; ep.ldo = N; ep.P = P; ep.T = T; ep.act = act;
  ep.trace = g_gemm_trace;
  return launch_gemm(reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(W), M, N, K, epilogue, ep, impl,
                     reinterpret_cast<cudaStream_t>(stream), nullptr);
}

// ---- LN fold, stand-alone (parity tests drive the same launchers the towers use) ----------------------------------
int mmcm_fold_ln(const float* W, const float* b, const float* gamma, const float* beta, int32_t N, int32_t K,
                 int32_t q_rows, float q_scale, void* w_out, float* resid_out, float* bias_out, void* stream) {
  if (!W || !b || !gamma || !beta || !w_out || !bias_out) return fail(MMCM_EINVAL, "null pointer");
  if (N <= 0 || K <= 0 || K > 1024) return fail(MMCM_EINVAL, "fold_ln: need 0 < K <= 1024");
  CK(launch_k(fold_ln_kernel, dim3((N + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), W, b, gamma, beta, N, K,
              q_rows, q_scale, reinterpret_cast<bf16*>(w_out), resid_out, bias_out));
  CK(cudaGetLastError());
  return MMCM_OK;
}

int mmcm_prep_rows(float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D, void* xb_out,
                   float* stats_out, void* stream) {
  if (!x || !xb_out || !stats_out || (gamma && !beta)) return fail(MMCM_EINVAL, "null pointer");
  return launch_prep_rows(x, gamma, beta, eps, rows, D, reinterpret_cast<bf16*>(xb_out),
                          reinterpret_cast<float2*>(stats_out), rows, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int mmcm_gemm_resid_stats(const void* A, const void* W, const float* bias, int32_t M, int32_t N, int32_t K, float* x,
                          void* xb_out, float* stats_out, void* stream) {
  if (!A || !W || !x || !xb_out || !stats_out) return fail(MMCM_EINVAL, "null pointer");
  EpiParams ep{};
  ep.bias = bias; ep.out = x; ep.resid = x; ep.ldo = N; ep.xb = xb_out;
  ep.stats = reinterpret_cast<float2*>(stats_out); ep.stats_pitch = M;
  ep.trace = g_gemm_trace;
  return launch_gemm(reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(W), M, N, K, EPI_RESID_STATS, ep, 0,
                     reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int mmcm_gemm_lnfold(const void* xb, const void* w_folded, const float* bias_folded, const float* stats, int32_t M,
                     int32_t N, int32_t K, float eps, int32_t act, void* out, void* stream) {
  if (!xb || !w_folded || !bias_folded || !stats || !out) return fail(MMCM_EINVAL, "null pointer");
  if (K % LN_SLAB != 0) return fail(MMCM_EINVAL, "gemm_lnfold: K must be a multiple of %d", LN_SLAB);
  EpiParams ep{};
  ep.bias = bias_folded; ep.out = out; ep.ldo = N; ep.act = act;
  ep.stats = reinterpret_cast<float2*>(const_cast<float*>(stats)); ep.stats_pitch = M; ep.ln_slabs = K / LN_SLAB;
  ep.ln_eps = eps;
  ep.trace = g_gemm_trace;
  return launch_gemm(reinterpret_cast<const bf16*>(xb), reinterpret_cast<const bf16*>(w_folded), M, N, K,
                     act ? EPI_LNFOLD_ACT_BF16 : EPI_LNFOLD_BF16, ep, 0, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int mmcm_layernorm(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D,
                   void* out_bf16, float* out_f32, void* stream) {
  if (!x || !gamma || !beta) return fail(MMCM_EINVAL, "null pointer");
  return launch_layernorm(x, gamma, beta, eps, rows, D, nullptr, reinterpret_cast<bf16*>(out_bf16), out_f32,
                          reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int mmcm_attention(const void* qkv, const uint8_t* key_valid, int32_t B, int32_t T, int32_t heads, int32_t causal,
                   void* out, void* stream) {
  if (!qkv || !out) return fail(MMCM_EINVAL, "null pointer");
  if (T <= 0 || heads <= 0) return fail(MMCM_EINVAL, "bad T/heads");
  return launch_attention(reinterpret_cast<const bf16*>(qkv), key_valid, B, T, heads, causal,
                          reinterpret_cast<bf16*>(out), reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int mmcm_preprocess_u8(const uint8_t* hwc, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3,
                       float* chw_out, void* stream) {
  if (!hwc || !chw_out || !mean3 || !std3) return fail(MMCM_EINVAL, "null pointer");
  if (B < 0 || H <= 0 || W <= 0 || (W & 3)) return fail(MMCM_EINVAL, "preprocess_u8: need W %% 4 == 0 (got %dx%d)", H, W);
  if (B == 0) return MMCM_OK;
  const int64_t total = (int64_t)B * 3 * H * (W / 4);
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  CK(launch_k(preprocess_u8_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), hwc, chw_out, B, H, W,
              mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]));
  return MMCM_OK;
}

int mmcm_resize_crop_u8(const uint8_t* src, const int64_t* offsets, const int32_t* heights, const int32_t* widths,
                        int32_t B, int32_t size, uint8_t* out, void* stream) {
  if (!src || !offsets || !heights || !widths || !out) return fail(MMCM_EINVAL, "null pointer");
  if (B < 0 || B > 65535 || size <= 0 || size > 1024)
    return fail(MMCM_EINVAL, "resize_crop_u8: bad batch %d (<= 65535 per call) or size %d (<= 1024)", B, size);
  if (B == 0) return MMCM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int kh = 3, kv = 3;
  double max_scale_v = 1.0;
  for (int b = 0; b < B; ++b) {
    if (heights[b] <= 0 || widths[b] <= 0 || heights[b] > 32768 || widths[b] > 32768 || offsets[b] < 0)
      return fail(MMCM_EINVAL, "resize_crop_u8: image %d has bad geometry %dx%d", b, heights[b], widths[b]);
    ResizeGeom g;
    resize_geometry(heights[b], widths[b], size, g);
    kh = std::max(kh, g.ksize_h);
    kv = std::max(kv, g.ksize_v);
    max_scale_v = std::max(max_scale_v, g.scale_v);
  }
  // rows per CTA: as many as the shared-memory tile allows (16 for every realistic photo)
  int rows = 16, max_src_rows = 0;
  size_t smem = 0;
  for (;; rows >>= 1) {
    max_src_rows = resize_strip_rows(max_scale_v, rows);
    smem = sizeof(int) * ((size_t)2 * size + (size_t)size * kh + 2 * rows + (size_t)rows * kv) +
           (size_t)max_src_rows * size * 3;
    if (smem <= 200 * 1024 || rows == 1) break;
  }
  if (smem > 200 * 1024)
    return fail(MMCM_EINVAL, "resize_crop_u8: a %.1fx downscale needs %zu bytes of shared memory per CTA", max_scale_v, smem);
  static AttrOnce once;
  if (once.need()) CK(cudaFuncSetAttribute(resize_crop_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // per-image metadata: one stream-ordered scratch allocation (offsets | heights | widths)
  char* meta = nullptr;
  const size_t mbytes = (size_t)B * (8 + 4 + 4);
  CK(cudaMallocAsync(reinterpret_cast<void**>(&meta), mbytes, st));
  long long* d_off = reinterpret_cast<long long*>(meta);
  int* d_h = reinterpret_cast<int*>(meta + (size_t)B * 8);
  int* d_w = d_h + B;
  cudaError_t ce = cudaMemcpyAsync(d_off, offsets, (size_t)B * 8, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_h, heights, (size_t)B * 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_w, widths, (size_t)B * 4, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess)
    ce = launch_k(resize_crop_u8_kernel, dim3((size + rows - 1) / rows, B), dim3(256), smem, st, src,
                  (const long long*)d_off, (const int*)d_h, (const int*)d_w, (int)size, rows, kh, kv, max_src_rows, out);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  cudaFreeAsync(meta, st);
  if (ce != cudaSuccess) return fail(MMCM_ECUDA, "resize_crop_u8 failed: %s", cudaGetErrorString(ce));
  return MMCM_OK;
}

int mmcm_postprocess(const float* logits, const float* thresholds, const float* labels, int32_t B, int32_t C,
                     float* probs_out, uint8_t* decisions_out, uint8_t* any_out, uint64_t* confusion_accum,
                     void* stream) {
  if (!logits || !thresholds) return fail(MMCM_EINVAL, "null pointer");
  if (B < 0 || C <= 0 || C > POST_MAXC) return fail(MMCM_EINVAL, "postprocess: need 0 < C <= %d", POST_MAXC);
  if (B == 0) return MMCM_OK;
  CK(launch_k(postprocess_kernel, dim3((B + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), logits,
              thresholds, labels, B, C, probs_out, decisions_out, any_out,
              reinterpret_cast<unsigned long long*>(confusion_accum)));
  return MMCM_OK;
}

// ---------------------------------------------------------------------------------- CLIP tokenizer (host)
struct mmcm_tokenizer_s {
  mmcm_tok::ClipTokenizer tok;
};

int mmcm_tokenizer_create(const char* vocab_json_path, const char* merges_txt_path, mmcm_tokenizer* out) {
  if (!vocab_json_path || !merges_txt_path || !out) return fail(MMCM_EINVAL, "null argument");
  *out = nullptr;
  std::unique_ptr<mmcm_tokenizer_s> t(new mmcm_tokenizer_s());
  const std::string err = t->tok.load(vocab_json_path, merges_txt_path);
  if (!err.empty()) return fail(MMCM_EINVAL, "tokenizer: %s", err.c_str());
  *out = t.release();
  return MMCM_OK;
}

int mmcm_tokenizer_destroy(mmcm_tokenizer t) {
  delete t;
  return MMCM_OK;
}

int mmcm_tokenizer_info(mmcm_tokenizer t, int32_t* vocab_size, int32_t* bos_id, int32_t* eos_id, int32_t* pad_id) {
  if (!t) return fail(MMCM_EINVAL, "null tokenizer");
  if (vocab_size) *vocab_size = t->tok.vocab_size();
  if (bos_id) *bos_id = t->tok.bos();
  if (eos_id) *eos_id = t->tok.eos();
  if (pad_id) *pad_id = t->tok.eos();     // CLIP pads with "<|endoftext|>"
  return MMCM_OK;
}

int mmcm_tokenizer_encode(mmcm_tokenizer t, const char* const* texts, const int64_t* lengths, int32_t n, int32_t max_len,
                          int64_t* input_ids_out, int64_t* attention_mask_out, int32_t n_threads) {
  if (!t || (n > 0 && (!texts || !lengths || !input_ids_out || !attention_mask_out))) return fail(MMCM_EINVAL, "null argument");
  if (n < 0 || max_len < 2) return fail(MMCM_EINVAL, "tokenizer: need n >= 0 and max_len >= 2");
  for (int i = 0; i < n; ++i)
    if (!texts[i] || lengths[i] < 0) return fail(MMCM_EINVAL, "tokenizer: text %d is null or has a negative length", i);
  mmcm_tok::encode_batch(t->tok, texts, lengths, n, max_len, input_ids_out, attention_mask_out, n_threads);
  return MMCM_OK;
}

int mmcm_cast_bf16(const float* src, void* dst, int64_t n, float scale, void* stream) {
  if (!src || !dst) return fail(MMCM_EINVAL, "null pointer");
  if (n <= 0) return MMCM_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CK(launch_k(cast_bf16_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), src, reinterpret_cast<bf16*>(dst),
                                                                              (size_t)n, scale));
  CK(cudaGetLastError());
  return MMCM_OK;
}

}  // extern "C"
