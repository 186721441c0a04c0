/*
 * mmcm.h -- C ABI of the B200-native scoring path (libmmcm.so).
 *
 * Drop-in boundary: the reference's scoring hot path is
 *     outputs = model(**batch); logits = outputs["logits"]
 * (R/scripts/evaluate.py:168-178, R/scripts/inference.py:214-216, R/sagemaker/inference.py:277-279),
 * i.e. `MultiModalFusionClassifier.forward` (R/src/models/fusion.py:157-229) and
 * `MultiTaskClassifier.forward` (R/src/models/multitask.py:156-227), whose encoder arithmetic is
 * Hugging Face `transformers` CLIP / SigLIP (HF/models/clip/modeling_clip.py,
 * HF/models/siglip/modeling_siglip.py).  The reference has no FFI of its own (pure Python); these
 * entry points are what a ctypes binding on the reference side calls -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success, non-zero on failure; mmcm_last_error() gives the message
 *     (thread-local).  Codes: 1 = invalid argument (Python side raises ValueError, mirroring
 *     HF/models/clip/modeling_clip.py:204-207,243-247), 2 = CUDA error / no device (RuntimeError),
 *     3 = state error (weights missing, handle not finalized).
 *   - device pointers are BORROWED: the caller keeps them alive until the work enqueued on `stream`
 *     has completed.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - a handle is bound to one device and is not thread-safe; use one handle per process/GPU.
 *   - there is no CPU fallback: without a CUDA device mmcm_create fails with code 2.
 */
#ifndef MMCM_H_
#define MMCM_H_

#include <stdint.h>

#if defined(__GNUC__)
#define MMCM_API __attribute__((visibility("default")))
#else
#define MMCM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MMCM_OK 0
#define MMCM_EINVAL 1
#define MMCM_ECUDA 2
#define MMCM_ESTATE 3

#define MMCM_BACKEND_CLIP 0   /* fusion.py:100-108  CLIPModel                    */
#define MMCM_BACKEND_SIGLIP 1 /* fusion.py:109-127  AutoModel (SigLIP / SigLIP2)   */
#define MMCM_HEAD_FUSION 0    /* R/src/models/fusion.py  MultiModalFusionClassifier */
#define MMCM_HEAD_MTL 1       /* R/src/models/multitask.py MultiTaskClassifier      */
#define MMCM_ACT_QUICK_GELU 1 /* HF/activations.py:122-123 */
#define MMCM_ACT_GELU_TANH 2  /* HF/activations.py:45      */

typedef struct mmcm_handle_s* mmcm_handle;

/* Shapes of one model instance.  Mirrors CLIPConfig / SiglipConfig plus the head ctor kwargs
 * (fusion.py:83-95, multitask.py:40-52). */
typedef struct mmcm_config {
  int32_t backend;          /* MMCM_BACKEND_*                               */
  int32_t head;             /* MMCM_HEAD_*                                  */
  int32_t text_hidden, text_heads, text_layers, text_ffn, text_act;
  int32_t vis_hidden, vis_heads, vis_layers, vis_ffn, vis_act;
  float text_eps, vis_eps;
  int32_t vocab, max_pos;   /* text vocabulary and positions (77 / 64)      */
  int32_t eos_id;           /* CLIP pooling id; 2 = legacy argmax branch    */
  int32_t image, patch;     /* 224, 32 | 16                                 */
  int32_t proj_dim;         /* CLIP projection_dim / SigLIP projection_size */
  int32_t fusion_dim;       /* 512                                          */
  int32_t num_outputs;      /* num_labels (fusion) or number of tasks (mtl) */
  int32_t head_hidden_dim;  /* mtl: 0 => Linear(fusion_dim,1) heads         */
} mmcm_config;

/* Lifetime ------------------------------------------------------------------------------------ */
MMCM_API int mmcm_create(const mmcm_config* cfg, int device, mmcm_handle* out);
MMCM_API int mmcm_destroy(mmcm_handle h);

/* Weights: one call per state-dict entry, `key` is the reference's own key
 * ("backbone.text_model.encoder.layers.3.self_attn.q_proj.weight", "cls.1.bias", ... -- the names
 * `model.load_state_dict` consumes at R/scripts/evaluate.py:139-151).  `src` is fp32, host or device
 * memory, `numel` elements; the library keeps its own repacked copy (bf16 GEMM operands with Q/K/V
 * concatenated and 1/sqrt(dh) folded into Q; fp32 for norms, biases, embeddings' positions and heads).
 * Unknown keys that the path does not consume (logit_scale, logit_bias, pos_weight, log_vars) are
 * accepted and ignored; anything else returns MMCM_EINVAL. */
MMCM_API int mmcm_load_weight(mmcm_handle h, const char* key, const float* src, int64_t numel);
/* Verifies that every tensor of the configured model was loaded and builds the TMA descriptors. */
MMCM_API int mmcm_finalize_weights(mmcm_handle h);

/* Packed weight file (checkpoint -> repacked bf16 blob).  mmcm_save_packed writes the handle's
 * finalized, repacked weight set (bf16 GEMM operands, fp32 norms / biases / embeddings / heads) with the handle's
 * mmcm_config and a checksum into one file; mmcm_load_packed maps such a file and copies it straight into a handle
 * created with the SAME config (replaces the mmcm_load_weight loop + mmcm_finalize_weights: no fp32 checkpoint read,
 * no repack); mmcm_packed_config reads the config a file was packed for, so a caller can mmcm_create from the file
 * alone.  Replaces the model.safetensors load of R/scripts/evaluate.py:139-151 on every start after the first.
 * Errors: MMCM_EINVAL for a missing / foreign / corrupt file or a config mismatch. */
MMCM_API int mmcm_save_packed(mmcm_handle h, const char* path);
MMCM_API int mmcm_load_packed(mmcm_handle h, const char* path);
MMCM_API int mmcm_packed_config(const char* path, mmcm_config* cfg_out);

/* The hot path.  Replaces MultiModalFusionClassifier.forward / MultiTaskClassifier.forward
 * (fusion.py:157-216, multitask.py:156-207): all pointers are DEVICE pointers.
 *   input_ids      int64 [B,S]            attention_mask int64 [B,S] or NULL (== all ones)
 *   pixel_values   fp32  [B,3,image,image]
 *   text_present   fp32  [B]              image_present  fp32 [B]
 *   logits_out     fp32  [B,num_outputs]  probs_out      fp32 [B,num_outputs] or NULL (sigmoid(logits),
 *                                          the callers' post-processing, inference.py:218)
 * S must be <= max_pos (else MMCM_EINVAL, like HF's ValueError). Work is enqueued on `stream`. */
MMCM_API int mmcm_forward(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                 const float* pixel_values, const float* text_present, const float* image_present,
                 int32_t B, int32_t S, float* logits_out, float* probs_out, void* stream);

/* Same call with HOST buffers (pinned memory recommended): copies inputs H2D, runs the path, copies
 * logits (and probs) back and synchronises `stream` before returning.  This is the end-to-end call a
 * CPU-side caller such as sagemaker/inference.py:predict_fn would bind. */
MMCM_API int mmcm_forward_host(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                      const float* pixel_values, const float* text_present, const float* image_present,
                      int32_t B, int32_t S, float* logits_out, float* probs_out, void* stream);

/* Same two calls with RAW uint8 pixels (the step immediately before the path).  pixels_u8 is
 * [B, image, image, 3] (HWC, RGB), already resized / centre-cropped by the caller; ToTensor + Normalize of the eval
 * transform (R/src/data/dataset.py:106-111) are applied in registers inside the patch im2col:
 *     pixel_values[b,c,y,x] = (u8 / 255 - mean3[c]) / std3[c]      (fp32, torchvision's operation order)
 * so the logits are bit-identical to mmcm_forward on the fp32 pixel_values that transform produces, the fp32 image
 * never exists in HBM and the host ships 147 KB instead of 588 KB per 224 px sample.  mean3 / std3: 3 HOST floats.
 * mmcm_forward_u8 takes device pointers (e.g. images decoded on the GPU), mmcm_forward_host_u8 host pointers. */
MMCM_API int mmcm_forward_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                    const uint8_t* pixels_u8, const float* mean3, const float* std3, const float* text_present,
                    const float* image_present, int32_t B, int32_t S, float* logits_out, float* probs_out,
                    void* stream);
MMCM_API int mmcm_forward_host_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                         const uint8_t* pixels_u8, const float* mean3, const float* std3, const float* text_present,
                         const float* image_present, int32_t B, int32_t S, float* logits_out, float* probs_out,
                         void* stream);

/* Double-buffered input pipeline for callers that score batch after batch (the reference's evaluate() loop over a
 * DataLoader with pin_memory, R/scripts/evaluate.py:163-183): enqueue the host->device copies of the NEXT batch now, on
 * the handle's copy stream and into a second set of input buffers, then call mmcm_forward_host* for the CURRENT batch --
 * the copies run on the copy engine while the towers compute.  The following mmcm_forward_host* call that is given the
 * SAME host pointers, B and S finds its inputs on the device (it waits for the per-stage copy events only) and issues no
 * copies of its own; any other call ignores and drops the prefetched batch.  The host buffers must stay untouched until
 * that forward returns.  Logits are bit-identical with and without the prefetch.  One batch can be in flight. */
MMCM_API int mmcm_prefetch_host(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                                const float* pixel_values, const float* text_present, const float* image_present,
                                int32_t B, int32_t S);
MMCM_API int mmcm_prefetch_host_u8(mmcm_handle h, const int64_t* input_ids, const int64_t* attention_mask,
                                   const uint8_t* pixels_u8, const float* text_present, const float* image_present,
                                   int32_t B, int32_t S);

/* Introspection ------------------------------------------------------------------------------ */
/* Copies an intermediate of the LAST forward into dst (device fp32).  (With "skip_absent_text" the text rows of samples
 * whose text cannot reach the logits hold the BOS row's output, not the reference's pooler output.)  Names: "text_pooled",
 * "vision_pooled" (tower pooler_output, fp32 [B,D]), "text_hidden", "vision_hidden" (residual stream
 * fp32 [B*T,D]: after the last layer with option "pooled_last_layer" = 0; with the default 1 the last layer only
 * advances the pooled rows, so these hold the last layer's INPUT).  *numel_out receives the element count. */
MMCM_API int mmcm_get_stage(mmcm_handle h, const char* name, float* dst, int64_t capacity, int64_t* numel_out,
                   void* stream);
/* Number of kernels this library launched during the last mmcm_forward on this handle. */
MMCM_API int64_t mmcm_last_launch_count(mmcm_handle h);
/* Share of the last mmcm_forward_host* call's wall time during which its pixel copies were still running (CUDA events).
 * Above 0.85 the call was bound by host->device bandwidth; the next call then tapers its last H2D stages (cv, ..., cv/2,
 * cv/4, cv/4) so that little tower work is left when the last bytes land. */
MMCM_API double mmcm_last_host_copy_share(mmcm_handle h);
/* Samples per internal pass (micro-batch) the last forward used for the text and the vision tower. */
MMCM_API int mmcm_last_chunks(mmcm_handle h, int32_t* text_chunk_out, int32_t* vision_chunk_out);
/* CUDA-event time (ms) of the GEMM launches of the last forward when profiling was enabled with
 * mmcm_set_option(h, "time_gemms", 1); also returns the FLOPs they EXECUTED (2*M*N*K with the live row count of
 * packed text chunks). Synchronises the device. */
MMCM_API int mmcm_gemm_time(mmcm_handle h, double* ms_out, double* flops_out, int64_t* launches_out);
/* The same, restricted to the launches with one epilogue (MMCM_EPI_*), plus their ALGORITHMIC DRAM bytes (A and W in
 * bf16 once, the epilogue's reads / writes per output element): the residual GEMMs (EPI_RESID_STATS) are bound by the
 * fp32 residual stream in HBM, the others by the tensor pipe -- bench.py reports them apart. */
MMCM_API int mmcm_gemm_time_epi(mmcm_handle h, int32_t epilogue, double* ms_out, double* flops_out, double* bytes_out,
                                int64_t* launches_out);
/* The same again, restricted to one weight shape as well (N, K; 0 = any; epilogue -1 = any): out_proj (K = N, bound by
 * the residual stream in HBM) and fc2 (K = 4 N, bound by the tensor pipe) share EPI_RESID_STATS and are reported apart. */
MMCM_API int mmcm_gemm_time_shape(mmcm_handle h, int32_t epilogue, int32_t N, int32_t K, double* ms_out,
                                  double* flops_out, double* bytes_out, int64_t* launches_out);
/* Options (name -> meaning).  Every option is per handle; h == NULL edits the defaults that the stand-alone kernels
 * below use ("pdl", "tma_epilogue", "attention_impl", "attention_ring", "narrow_tiles" only):
 *   "streams"          1 or 2: text / vision towers on separate internal streams (default 2)
 *   "micro_batch"      upper bound on the samples per internal pass of a tower (default 1024)
 *   "auto_chunk"       1 = pick, per tower, the chunk <= micro_batch whose GEMM tile counts fill whole rounds of the
 *                      74 CTA pairs; 0 = use micro_batch as is
 *   "varlen_text"      1 = default: the causal CLIP text tower keeps only the rows up to each sample's pooled EOS
 *                      position, packed back to back -- bit-identical logits, fewer rows; 0 = all S rows like the reference
 *   "pooled_last_layer" 1 = default: after the last layer's attention only the one row per sample that is pooled
 *                      (CLIP vision CLS, CLIP text EOS, SigLIP text last token) goes through out_proj / LN2 / MLP /
 *                      final LN -- those ops are row-wise, so the logits are bit-identical and 6 % of the GEMM work
 *                      is skipped; 0 = all rows like the reference.  (SigLIP vision always runs all rows: its MAP
 *                      head reads every token.)
 *   "host_chunk"       mmcm_forward_host / _host_u8: samples per H2D pipeline stage of the vision tower
 *                      (0 = default: about 200 MB of pixels per stage)
 *   "ln_fold"          1 = default: no LayerNorm pass inside the encoder layers -- the residual GEMMs (out_proj, fc2)
 *                      leave a bf16 copy and per-row statistics of the updated rows, the qkv / fc1 GEMMs apply mean /
 *                      rstd in their epilogue on folded weights (see mmcm_gemm_lnfold); 0 = separate normalisation pass
 *                      in front of qkv / fc1 (always the case for gemm_impl 1 and 2)
 *   "head_cluster"     1 = default: for B <= 144 the fused head kernel runs as one thread-block cluster of 8 CTAs per
 *                      8 samples, each CTA computing an eighth of every Linear's columns and sharing them through
 *                      distributed shared memory (the B = 1 head: 341 -> ~45 us); bit-identical logits; 0 = one CTA
 *   "narrow_tiles"     1 = default: GEMMs with a single 256-row block (B = 1 requests, the pooled-rows last layer)
 *                      use 64-column (fp32 epilogues) / 128-column (bf16 epilogues) tiles so that 4x / 2x as many CTA
 *                      pairs share the weight stream; 0 = 256-column tiles always.  Same per-element k order: identical bits.
 *   "skip_absent_text" 1 = default (with varlen_text): a sample whose text feature cannot reach the logits -- fusion head:
 *                      text_present < 0.5 (fusion.py:188-189); MTL head: text_present < 0.5 and image_present >= 0.5
 *                      (multitask.py:194-197) -- keeps one row of the packed text tower instead of up to 77; and a
 *                      forward in which NO sample has an image (or usable text) skips that tower altogether (flags are
 *                      read on the host: directly in mmcm_forward_host*, by one small read-back in mmcm_forward* when
 *                      B <= 16 and the stream is not being captured).  Bit-identical logits; the "text_pooled" /
 *                      "vision_pooled" stages of skipped samples hold placeholders.  0 = every sample runs both towers
 *   "split_k"          1 = default: in forwards with B < 16 the residual GEMMs (out_proj, fc2) split their K loop over
 *                      up to 4 CTA pairs per tile; the partial sums are added to the residual stream, in a fixed order,
 *                      by the LayerNorm kernel that follows (deterministic; no reduction launch); 0 = one pair per tile
 *   "gemm_impl"        0 = tcgen05 CTA-pair kernel, 1 = SIMT validation kernel, 2 = tcgen05 single-CTA kernel
 *   "tma_epilogue"     1 = TMA tile-store / reduce-add epilogue of the pair GEMM, 0 = per-thread stores
 *   "attention_impl"   0 = auto: the TMA-ring mma.sync kernel for 32 < T <= 80 (the 77-token CLIP text tower, dense and
 *                      packed; the 50-token vision tower; 64-token SigLIP text), the tcgen05 attention kernel for
 *                      T <= 32 and 128 < T <= 256 (SigLIP vision), the cp.async mma.sync kernel otherwise;
 *                      1 = cp.async mma.sync always; 2 = tcgen05 whenever T <= 256; 3 = TMA-ring wherever it exists
 *   "attention_ring"   1 = default; 0 = auto never picks the TMA-ring kernel (tcgen05 for T <= 64, cp.async mma.sync for
 *                      the text tower: the round-1 choice, for A/B runs).  The ring kernel is bit-identical to the
 *                      cp.async one (same fragments and order of operations)
 *   "pdl"              1 = every kernel is launched with programmatic stream serialization (prologues overlap the
 *                      previous kernel's tail)
 *   "graph_max_batch"  forwards with B <= this are replayed as one CUDA graph from the third call of a shape on
 *                      (default 0 = off: a B=1 forward is bound by its 177-kernel dependency chain on the GPU, 1.45 ms
 *                      with or without replay)
 *   "pairs_text", "pairs_vision"  > 0: cap the CTA pairs the tower's GEMMs may occupy (SM partitioning experiment;
 *                      measured slower than letting every GEMM use all 74 pairs)
 *   "time_gemms"       1 = record CUDA events around every GEMM launch (see mmcm_gemm_time)
 *   "debug_feats"      1 = keep the projected features of the fusion head for mmcm_get_stage "text_feat"/"vision_feat" */
MMCM_API int mmcm_set_option(mmcm_handle h, const char* name, int64_t value);
MMCM_API const char* mmcm_last_error(void);
MMCM_API const char* mmcm_version(void);

/* Stand-alone kernels (unit parity tests call these through the same ABI) ----------------------- */
#define MMCM_EPI_BIAS_BF16 0       /* out bf16 = acc + bias                                  */
#define MMCM_EPI_BIAS_ACT_BF16 1   /* out bf16 = act(acc + bias)                             */
#define MMCM_EPI_BIAS_RESID_F32 2  /* out fp32 = acc + bias + resid (resid may alias out)    */
#define MMCM_EPI_PATCH_F32 3       /* out fp32[(r/P)*T+off+r%P] = acc + bias? + pos[off+r%P] */
#define MMCM_EPI_RESID_STATS 4     /* mmcm_gemm_resid_stats: x += acc + bias, bf16 copy, row statistics */
#define MMCM_EPI_LNFOLD_BF16 5     /* mmcm_gemm_lnfold, act = 0                                  */
#define MMCM_EPI_LNFOLD_ACT_BF16 6 /* mmcm_gemm_lnfold, act != 0                                 */

/* out[M,N] = epilogue(A[M,K] @ W[N,K]^T): A, W bf16 row-major device pointers (K % 64 == 0, N % 128 == 0).
 * impl 0 = tcgen05/TMEM/TMA kernel on CTA pairs (cta_group::2, 256 x BLOCK_N tiles), 1 = SIMT validation kernel,
 * 2 = tcgen05 single-CTA kernel (128 x BLOCK_N tiles, dynamic tile scheduler). act = MMCM_ACT_* for EPI_BIAS_ACT.
 * For EPI_PATCH: pos fp32 [T,N], P = patches per sample, T = tokens per sample, off = T - P. */
MMCM_API int mmcm_gemm_bf16(const void* A, const void* W, const float* bias, int32_t M, int32_t N, int32_t K,
                   int32_t epilogue, int32_t act, void* out, const float* resid, const float* pos,
                   int32_t P, int32_t T, int32_t impl, void* stream);
/* --- LayerNorm folded into the GEMMs (what the towers run; stand-alone for the parity tests) -----------------------
 * The encoder's pre-LN blocks (HF/models/clip/modeling_clip.py:372-384) compute Linear(LayerNorm(x)).  LayerNorm is
 * row-wise over the Linear's K dimension, so
 *     LayerNorm(x) W^T + b = rstd * ((x - mean) (W*gamma)^T) + b' = rstd * (x W'^T) + b',   b' = b + W beta,
 * where W' = W*gamma with every row centred (sum_k W'[n,k] = 0), so that x W'[n] = (x - mean(x)) W'[n] for any x.
 * mmcm_fold_ln        weights, once per load: w_out bf16 [N,K] = bf16(scale_n * W * gamma - its row mean), bias_out fp32
 *                     [N] = scale_n * (b + W beta); resid_out (may be NULL) = sum_k w_out[n,k], what bf16 rounding
 *                     leaves of the zero row sum (scale_n = q_scale for the first q_rows rows: dh^-1/2 of the fused
 *                     Q|K|V matrix).  K <= 1024.
 * mmcm_gemm_resid_stats   x[M,N] += A[M,K] @ W[N,K]^T + bias in place (fp32), xb_out = bf16(x), stats_out =
 *                     float2 [N/128][M]: (sum, sum of squares about the slab mean) of every 128-column slab of the
 *                     updated row.  N % 256 == 0.  Replaces out_proj / fc2 + the read half of the LayerNorm pass.
 * mmcm_prep_rows      the same by-products for rows no GEMM produced (embeddings); gamma != NULL first applies
 *                     LayerNorm(gamma, beta) to x in place (CLIP pre_layrnorm).  D in {512, 768, 1024}.
 * mmcm_gemm_lnfold    out bf16 [M,N] = act(rstd * (xb @ w_folded^T) + bias_folded), rstd from `stats`
 *                     (K = 128 * slabs <= 1024).  Replaces LayerNorm + qkv / fc1.  act = 0 or MMCM_ACT_*. */
MMCM_API int mmcm_fold_ln(const float* W, const float* b, const float* gamma, const float* beta, int32_t N, int32_t K,
                          int32_t q_rows, float q_scale, void* w_out, float* resid_out, float* bias_out, void* stream);
MMCM_API int mmcm_prep_rows(float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D,
                            void* xb_out, float* stats_out, void* stream);
MMCM_API int mmcm_gemm_resid_stats(const void* A, const void* W, const float* bias, int32_t M, int32_t N, int32_t K,
                                   float* x, void* xb_out, float* stats_out, void* stream);
MMCM_API int mmcm_gemm_lnfold(const void* xb, const void* w_folded, const float* bias_folded, const float* stats,
                              int32_t M, int32_t N, int32_t K, float eps, int32_t act, void* out, void* stream);
/* y = LayerNorm(x) over the last dim D (512 or 768); x fp32 [rows,D]; out_bf16 and/or out_f32 may be NULL. */
MMCM_API int mmcm_layernorm(const float* x, const float* gamma, const float* beta, float eps, int32_t rows, int32_t D,
                   void* out_bf16, float* out_f32, void* stream);
/* Multi-head attention over a packed QKV buffer bf16 [B*T, 3*D] (Q pre-scaled by 1/sqrt(64)); head dim 64.
 * key_valid uint8 [B,T] or NULL; causal 0/1; out bf16 [B*T, D]. A query with no admissible key yields 0. */
MMCM_API int mmcm_attention(const void* qkv, const uint8_t* key_valid, int32_t B, int32_t T, int32_t heads,
                   int32_t causal, void* out, void* stream);
/* --- the callers' pre- and post-processing, device pointers ----------------------------------------
 * ToTensor + Normalize of the eval transform (R/src/data/dataset.py:106-111) for an already resized/cropped uint8
 * HWC image batch [B,H,W,3]: chw_out[b,c,y,x] = (u8/255 - mean[c]) / std[c], fp32 [B,3,H,W], bit-identical to
 * torchvision.  mean3 / std3 are HOST pointers to 3 floats. W % 4 == 0. */
MMCM_API int mmcm_preprocess_u8(const uint8_t* hwc, int32_t B, int32_t H, int32_t W, const float* mean3,
                                const float* std3, float* chw_out, void* stream);
/* T.Resize(size, antialias=True) + T.CenterCrop((size, size)) of the eval transform (R/src/data/dataset.py:106-108)
 * for B decoded uint8 RGB images of DIFFERENT sizes: src = device buffer holding the HWC images back to back,
 * offsets[b] = byte offset of image b, heights / widths = its size (three HOST arrays of length B); out = device
 * [B, size, size, 3] uint8, ready for mmcm_forward_u8.  The arithmetic is Pillow's Image.resize(BILINEAR) -- tap windows
 * and weights in double precision, 22-bit fixed-point coefficients, horizontal pass rounded to uint8, then the vertical
 * pass -- so the crops are byte-identical to what torchvision produces from a PIL image.  Only the cropped region is
 * computed.  Grid dimension y carries B: B <= 65535 per call. */
MMCM_API int mmcm_resize_crop_u8(const uint8_t* src, const int64_t* offsets, const int32_t* heights, const int32_t* widths,
                        int32_t B, int32_t size, uint8_t* out, void* stream);
/* probs = 1/(1+exp(-logits)); decisions[b,c] = probs >= thresholds[c]; any[b] = OR_c decisions
 * (R/scripts/inference.py:218-232); when labels != NULL, confusion_accum[c*4 + {0,1,2,3}] += {TP,FP,FN,TN}
 * (the counts behind f1/precision/recall of R/src/training/metrics.py:180-205).  Outputs may be NULL. C <= 64. */
MMCM_API int mmcm_postprocess(const float* logits, const float* thresholds, const float* labels, int32_t B, int32_t C,
                              float* probs_out, uint8_t* decisions_out, uint8_t* any_out, uint64_t* confusion_accum,
                              void* stream);
/* --- CLIP tokenizer, host side: the step that produces input_ids / attention_mask --------------
 * Replaces `tokenizer(text, padding="max_length", truncation=True, max_length=77, return_attention_mask=True)` of
 * R/src/data/dataset.py:148-165 / R/scripts/inference.py:168-180 for the CLIP backends, i.e. the pipeline Hugging Face's
 * CLIPTokenizer configures (HF/models/clip/tokenization_clip.py:68-118): special tokens cut out of the raw text; NFC,
 * white-space runs -> ' ', Unicode lower-casing; Split on 's|'t|'re|'ve|'m|'ll|'d|\p{L}+|\p{N}|[^\s\p{L}\p{N}]+;
 * byte-level BPE with "</w>"; [bos] ... [eos], truncation, padding with "<|endoftext|>".  vocab.json / merges.txt are the
 * checkpoint's own files (none ship here).  texts[i] is UTF-8, lengths[i] its byte length (no terminator needed);
 * outputs are HOST int64 [n, max_len].  n_threads <= 0: all hardware threads.  Thread-safe for concurrent encodes. */
typedef struct mmcm_tokenizer_s* mmcm_tokenizer;
MMCM_API int mmcm_tokenizer_create(const char* vocab_json_path, const char* merges_txt_path, mmcm_tokenizer* out);
MMCM_API int mmcm_tokenizer_destroy(mmcm_tokenizer t);
MMCM_API int mmcm_tokenizer_info(mmcm_tokenizer t, int32_t* vocab_size, int32_t* bos_id, int32_t* eos_id, int32_t* pad_id);
MMCM_API int mmcm_tokenizer_encode(mmcm_tokenizer t, const char* const* texts, const int64_t* lengths, int32_t n,
                                   int32_t max_len, int64_t* input_ids_out, int64_t* attention_mask_out,
                                   int32_t n_threads);
/* Dev tool: when device_buffer != NULL, mmcm_gemm_bf16 (impl 0) writes 16 clock64 stamps per CTA into it
 * (>= 148 * 16 int64); NULL switches tracing off. */
MMCM_API int mmcm_debug_set_gemm_trace(void* device_buffer);
/* fp32 -> bf16 with scale (device pointers), used by tests to prepare operands. */
MMCM_API int mmcm_cast_bf16(const float* src, void* dst, int64_t n, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMCM_H_ */
