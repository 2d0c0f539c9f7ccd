/* b2c.h — C-ABI of libb2c.so, the B200 (sm_100a) implementation of the one data-parallel hot path of
 * aiXander/CLIP_assisted_data_labeling: 4-crop CLIP-ViT image embedding and the all-to-all cosine
 * duplicate search.  All citations are file:line in the reference tree.
 *
 * Conventions:
 *   - plain C types only; every function returns int (0 = ok, <0 = error code below); the message of
 *     the last failure on the calling thread is b2c_last_error();  no C++ exception crosses this ABI;
 *   - the CALLER owns every tensor (device pointers, e.g. torch allocations, including workspaces whose
 *     size is queried first); the library owns only opaque handles (b2c_vit) and their converted weights;
 *   - every launch takes a stream (a cudaStream_t passed as void*; NULL = legacy default stream), is
 *     asynchronous and never synchronises the device;
 *   - there is no CPU fallback: without an sm_100 device the compute entry points fail with B2C_ERR_CUDA.
 */
#ifndef B2C_H_
#define B2C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2C_OK 0
#define B2C_ERR_ARG (-1)         /* bad argument / unsupported shape */
#define B2C_ERR_CUDA (-2)        /* a CUDA runtime or driver call failed (incl. no usable device) */
#define B2C_ERR_STATE (-3)       /* handle not ready (e.g. missing weights) */
#define B2C_ERR_WORKSPACE (-4)   /* workspace too small */
#define B2C_ERR_UNSUPPORTED (-5) /* valid input this path does not cover (b2c_jpeg_*: keep the file on the host decoder) */

typedef void* b2c_stream; /* cudaStream_t */

/* dtype codes used wherever a tensor's element type is passed */
#define B2C_F32 0
#define B2C_F16 1
#define B2C_BF16 2
#define B2C_U8 3

const char* b2c_last_error(void);
int b2c_version(void);
/* Number of kernels this library has launched since load (all threads); bench.py reports the delta
 * over its timed region as "gpu_launches". */
unsigned long long b2c_launch_count(void);

/* Stage timer (measurement aid, off by default).  While enabled, every stage of the hot path the library
 * launches is bracketed by a pair of CUDA events recorded on the launching stream; b2c_prof_read()
 * synchronises those events and returns, per stage kind, the summed device time in milliseconds and the number
 * of bracketed stages since b2c_prof_enable(1).  bench.py uses it to report the live duration of the dominant
 * kernel ("roofline.achieved") and each stage's share of a step next to the ncu launch list. */
#define B2C_PROF_PREPROCESS 0
#define B2C_PROF_PATCH_EMBED 1
#define B2C_PROF_LAYERNORM 2
#define B2C_PROF_IN_PROJ 3
#define B2C_PROF_ATTENTION 4
#define B2C_PROF_OUT_PROJ 5
#define B2C_PROF_C_FC 6
#define B2C_PROF_C_PROJ 7
#define B2C_PROF_HEAD 8
#define B2C_PROF_DEDUP 9
#define B2C_PROF_OTHER 10
#define B2C_PROF_KINDS 11
int b2c_prof_enable(int on);  /* on != 0: drop earlier records and start recording; 0: stop */
int b2c_prof_read(double* ms /*[B2C_PROF_KINDS]*/, unsigned long long* stages /*[B2C_PROF_KINDS]*/);
const char* b2c_prof_kind_name(int kind);

/* ---------------------------------------------------------------------------------------------
 * K0 — 4-crop geometry + PIL-exact bicubic resize + normalise.
 * Replaces CustomImageDataset.extract_crops (utils/embedder.py:184-251) and the open_clip val
 * transform applied per crop (utils/embedder.py:90-92,173: Resize(R,BICUBIC) -> CenterCrop(R) ->
 * ToTensor -> Normalize(mean,std)).
 * ------------------------------------------------------------------------------------------- */

/* One crop as the reference builds it: a (cw x ch) canvas whose pixel (x,y) is image pixel
 * (x+dx, y+dy) when that lies inside the W x H image and black otherwise (square_padded_crop,
 * utils/embedder.py:204-212); the canvas is resized so that its shorter side is R (out_w x out_h,
 * torchvision Resize semantics) and the centre R x R window starting at (off_x, off_y) is kept. */
typedef struct {
  int32_t cw, ch;       /* canvas size before resize; 0,0 = crop dropped as empty (embedder.py:243-247) */
  int32_t dx, dy;       /* canvas -> image offset */
  int32_t out_w, out_h; /* size after Resize(R) */
  int32_t off_x, off_y; /* CenterCrop(R) offset inside the resized canvas */
} b2c_crop;

/* Host-only, pure integer/double arithmetic: the 4 crops [centre_crop, square_padded_crop, subcrop1,
 * subcrop2] of a W x H image at model resolution R (utils/embedder.py:196-236 + torchvision
 * Resize/CenterCrop rounding). */
int b2c_crop_geometry(int W, int H, int R, b2c_crop out[4]);

/* Bytes of device workspace b2c_preprocess_4crop needs for B images whose sides are <= max_side. */
int b2c_preprocess_workspace_bytes(int B, int max_side, int R, size_t* bytes);

/* out_layout */
#define B2C_OUT_NCHW_F32 0   /* f32[B,4,3,R,R]  — bit-identical to torch.stack([preprocess(c)...]) (embedder.py:173) */
#define B2C_OUT_PATCH_BF16 1 /* bf16[B*4, (R/patch)^2, Kp], Kp = round_up(3*patch*patch, 64), k = c*p*p + py*p + px:
                                the A operand of the patch-embed GEMM (conv1 with stride = kernel) */

/* img_ptrs: HOST array of B DEVICE pointers to uint8 RGB images, HWC, row pitch pitch[i] bytes.
 * H, W, pitch: HOST int arrays.  mean/std: HOST float[3].  Crops that the reference would drop
 * (zero area) are written as zeros. */
int b2c_preprocess_4crop(const uint8_t* const* img_ptrs, const int* H, const int* W, const int* pitch, int B,
                         int R, int patch, const float* mean, const float* std, int out_layout, void* out,
                         void* ws, size_t ws_bytes, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K1..K8 — the open_clip VisionTransformer forward + L2 normalise.
 * Replaces CLIP_Encoder.encode_image (utils/embedder.py:94-100), i.e. open_clip's
 * model.encode_image (third-party, un-vendored; architecture restated in SURVEY.md App. A).
 * ------------------------------------------------------------------------------------------- */
typedef struct b2c_vit b2c_vit;

#define B2C_ACT_QUICK_GELU 0 /* x*sigmoid(1.702x): open_clip tags ending in /openai */
#define B2C_ACT_GELU 1       /* exact erf GELU: LAION tags (ViT-H-14/laion2b_s32b_b79k) */

typedef struct {
  int32_t image;  /* R, e.g. 224 / 336 */
  int32_t patch;  /* 14 / 32 */
  int32_t width;  /* d */
  int32_t layers; /* L */
  int32_t heads;  /* width / heads must be 64 or 80 */
  int32_t mlp;    /* hidden width of the MLP (4d) */
  int32_t embed;  /* E */
  int32_t act;    /* B2C_ACT_* */
} b2c_vit_cfg;

int b2c_vit_create(const b2c_vit_cfg* cfg, b2c_vit** out);
int b2c_vit_destroy(b2c_vit* vit);
/* key: open_clip visual state-dict name without the 'visual.' prefix, e.g.
 * "conv1.weight", "transformer.resblocks.3.attn.in_proj_weight", "proj".  dev_ptr: DEVICE pointer,
 * contiguous, dtype B2C_F32/F16/BF16.  The handle keeps its own converted copy; the caller may free
 * the source after the stream work completes (the copy runs on the legacy default stream and is
 * synchronised before return). */
int b2c_vit_set_weight(b2c_vit* vit, const char* key, const void* dev_ptr, int dtype, const int64_t* shape,
                       int ndim);
/* 0 when every tensor of the architecture has been set, else B2C_ERR_STATE (message names one missing key). */
int b2c_vit_ready(const b2c_vit* vit);
/* Lanes (1..4, default 2; env B2C_VIT_LANES): a pass over up to `chunk` crops is split into `lanes` independent
 * sub-batches, each on its own internal stream forked from / joined to the caller's stream with events, so the
 * HBM-bound stages of one lane run under the tensor-bound GEMMs of another.  Results do not depend on it (crops are
 * independent units).  Set it BEFORE querying the workspace size.  The stage timer forces 1 lane while it is on. */
int b2c_vit_set_lanes(b2c_vit* vit, int lanes);
/* LayerNorm fusion (default on; env B2C_VIT_FUSED_LN=0): ln_1 / ln_2 are folded into in_proj / c_fc
 * (gamma into the weight, mean / rstd applied to the accumulator per row) and the out_proj / c_proj epilogues own
 * the residual update, leaving a bf16 copy of x and its row statistics for the next GEMM.  0 selects the stand-alone
 * LayerNorm kernels.  Both compute LN(x)·Wᵀ + b of utils/embedder.py:98's tower within the same tolerance. */
int b2c_vit_set_fused_ln(b2c_vit* vit, int on);
/* Opt-in, off by default (env B2C_VIT_GRAPH=1): a forward call whose buffers (input, output, workspace), crop count and
 * switches repeat is captured into a CUDA graph on its second occurrence and replayed from then on — one
 * cudaGraphLaunch instead of ~250 kernel launches and the lanes' fork / join events.  Same kernels, same results.  A
 * weight that is re-set keeps its device buffer, so captured graphs stay valid; 0 drops them.  Up to 16 graphs are
 * kept per handle.  The stage timer bypasses it. */
int b2c_vit_set_graph(b2c_vit* vit, int on);
/* Opt-in, off by default: in the LAST block evaluate only the class-token row — the only row ln_post / proj read
 * (utils/embedder.py:98 returns the pooled class token).  K and V of every token are still computed; the block's
 * attention, out_proj, ln_2, c_fc and c_proj run on one row per crop (3.3 % fewer FLOPs for ViT-L/14).  The embedding is
 * the same quantity, rounded at slightly different points (within the tower's tolerance).  Needs head dim 64 and the
 * fused LayerNorm path; otherwise the block runs in full.  bench.py's headline keeps this off. */
int b2c_vit_set_cls_only_last_block(b2c_vit* vit, int on);
int b2c_vit_workspace_bytes(const b2c_vit* vit, int n_crops, size_t* bytes);
/* pixels: [n,3,R,R] (B2C_F32 / B2C_F16 / B2C_BF16), already normalised — the tensor the reference
 * feeds encode_image (utils/embedder.py:95-98).  out: f32[n,E], unit-norm rows (embedder.py:99). */
int b2c_vit_forward_pixels(b2c_vit* vit, const void* pixels, int dtype, int n_crops, float* out, void* ws,
                           size_t ws_bytes, b2c_stream stream);
/* patches: bf16[n, g*g, Kp] as written by b2c_preprocess_4crop(B2C_OUT_PATCH_BF16). */
int b2c_vit_forward_patches(b2c_vit* vit, const void* patches, int n_crops, float* out, void* ws,
                            size_t ws_bytes, b2c_stream stream);

/* Operator-level entry points (unit parity tests; each is one kernel of the tower). */
#define B2C_EPI_BIAS_BF16 0      /* out bf16[M,N]  = A·Wᵀ + bias                    (K3 in_proj)          */
#define B2C_EPI_BIAS_QGELU_BF16 1 /* out bf16[M,N] = quick_gelu(A·Wᵀ + bias)        (K6 c_fc, openai)     */
#define B2C_EPI_BIAS_GELU_BF16 2 /* out bf16[M,N]  = gelu_erf(A·Wᵀ + bias)          (K6 c_fc, laion)      */
#define B2C_EPI_BIAS_RESID_F32 3 /* out f32[M,N]  += A·Wᵀ + bias                    (K5 out_proj, K7 c_proj) */
/* A: bf16[M,K] row-major, W: bf16[N,K] row-major (torch Linear layout), bias: f32[N] or NULL.
 * K % 64 == 0, N % 256 == 0. */
int b2c_gemm_bf16(const void* A, const void* W, const float* bias, void* out, int64_t M, int N, int K, int epilogue,
                  b2c_stream stream);
/* y bf16[M,d] = LayerNorm(x f32[M,d]; gamma,beta f32[d], eps) — ln_1 / ln_2 (K2). d % 128 == 0, d <= 2048. */
int b2c_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d, float eps,
                       b2c_stream stream);
/* qkv: bf16[n*T, 3d] (q | k | v packed like nn.MultiheadAttention.in_proj), out: bf16[n*T, d]. hd in {64,80}. */
int b2c_attention_bf16(const void* qkv, void* out, int n, int T, int heads, int hd, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K9 — duplicate search.  Replaces the body of find_near_duplicates (_2_remove_duplicates.py:63-80):
 * row normalise (67), S = E·Eᵀ (69), where(triu(S,1) > thr) (74), S[i,j] per pair (80) — without
 * ever materialising S.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t i, j; /* i < j */
  float sim;    /* fp32 accumulator of the fp16 x fp16 dot product */
} b2c_pair;

/* out f16[n, E_pad] = in[n,E] / ||in||_2 per row (fp32 math; _2_remove_duplicates.py:67), zero-padded to
 * E_pad = round_up(E, 64) columns.  in_dtype: B2C_F32 or B2C_F16. */
int b2c_normalize_rows_f16(const void* in, int in_dtype, int64_t n, int E, void* out_f16, b2c_stream stream);

#define B2C_CMP_FP32 0 /* pair kept iff fp32 accumulator > threshold */
#define B2C_CMP_REF_FP16 1 /* pair kept iff fp16(acc) > fp16(threshold): the reference's comparison on its
                              fp16 similarity matrix (_2_remove_duplicates.py:38,69,74) */
#define B2C_CMP_EUCLID 2 /* sim_type='euclidean' (_2_remove_duplicates.py:70-71, unreachable from its CLI): pair kept iff
                            the distance of the unit rows, sqrt(max(2 - 2 acc, 0)), exceeds threshold; that distance is
                            what b2c_pair.sim then holds */
/* emb: f16[n_total, E_pad] unit-norm rows (all ranks' shards gathered), E_pad % 64 == 0.
 * Emits every pair (i,j), row_begin <= i < row_end, i < j < n_total, whose similarity passes the
 * comparison.  out/count are DEVICE memory; *count is incremented atomically once per pair and may
 * exceed capacity (overflow: pairs beyond capacity are dropped, re-run with a larger buffer).
 * Pair order is unspecified (the host sorts by (i,j) to reproduce torch.where's row-major order). */
int b2c_dedup_pairs(const void* emb_f16, int64_t n_total, int E_pad, int64_t row_begin, int64_t row_end,
                    float threshold, int compare_mode, b2c_pair* out, unsigned long long capacity,
                    unsigned long long* count, b2c_stream stream);
/* Same, restricted to columns col_begin <= j < col_end (<= n_total): one block of the upper triangle.  The multi-GPU
 * search uses it to work on the block of its OWN shard (rows and columns local) while the all-gather of the peers'
 * shards is in flight, and to leave that block out of the bands it owns afterwards. */
int b2c_dedup_pairs_block(const void* emb_f16, int64_t n_total, int E_pad, int64_t row_begin, int64_t row_end,
                          int64_t col_begin, int64_t col_end, float threshold, int compare_mode, b2c_pair* out,
                          unsigned long long capacity, unsigned long long* count, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K10 — SimpleFC regressor forward (utils/nn_model.py:23-41; called at _5_predict_labels.py:135):
 * Linear -> LeakyReLU(slope) [-> Dropout: identity in eval] per hidden layer, Linear -> Sigmoid.
 * ------------------------------------------------------------------------------------------- */
#define B2C_MLP_MAX_LAYERS 8
typedef struct {
  int32_t n_layers;                      /* number of Linear layers (hidden + 1) */
  int32_t dims[B2C_MLP_MAX_LAYERS + 1];  /* dims[0] = input, dims[n_layers] = output (1) */
  const float* weight[B2C_MLP_MAX_LAYERS]; /* DEVICE f32[dims[l+1], dims[l]] (torch Linear layout) */
  const float* bias[B2C_MLP_MAX_LAYERS];   /* DEVICE f32[dims[l+1]] */
  float leaky_slope;                     /* 0.01 (nn.LeakyReLU default, utils/nn_model.py:28) */
} b2c_mlp_weights;

/* feats: DEVICE f32[B, dims[0]]; out: DEVICE f32[B, dims[n_layers]]. Hidden widths <= 1024. */
int b2c_mlp_score(const float* feats, int64_t B, const b2c_mlp_weights* w, float* out, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K11 — similarity-search variants over stored embeddings.
 * Replaces the per-sample loop of tools/find_similar_imgs.py:96-137 (compute_distance :88-94 + the topN
 * bookkeeping :67-85) and the greedy loop of diversity_ordered_image_files (_3_label_images.py:128-177).
 * HBM-bound streaming kernels (one pass over the N x E embeddings per context vector), not GEMMs.
 * ------------------------------------------------------------------------------------------- */
#define B2C_MEASURE_COSINE_DIST 0 /* (1 - cos(c,x)) / 2            tools/find_similar_imgs.py:90 */
#define B2C_MEASURE_L2 1          /* || c - x + 1e-6 ||_2           tools/find_similar_imgs.py:92 (F.pairwise_distance) */
#define B2C_MEASURE_COSINE_SIM 2  /* cos(c,x)                       _3_label_images.py:124-127 */
#define B2C_COMBINE_STORE 0       /* out[i]  = score_i */
#define B2C_COMBINE_MAX 1         /* out[i]  = max(out[i], score_i) */

/* emb: DEVICE [n, E] rows of dtype B2C_F32 / B2C_F16, row i at emb + i*row_stride elements (row_stride >= E, so one
 * crop of a packed [N, C, E] store is addressed in place).  Context vector: ctx = DEVICE f32[E], or ctx == NULL and
 * ctx_row = DEVICE int32 holding the index of the row of emb to use (chosen by an earlier kernel, no host round
 * trip).  skip: optional DEVICE u8[n]; rows with skip != 0 get +INF.  out: DEVICE f32[n]. */
int b2c_context_scores(const void* emb, int dtype, int64_t n, int E, int64_t row_stride, const float* ctx,
                       const int32_t* ctx_row, int measure, int combine, const uint8_t* skip, float* out,
                       b2c_stream stream);

/* The k smallest entries of scores f32[n] (DEVICE), k <= min(n, 4096): out_idx int32[k] / out_val f32[k] (DEVICE),
 * ascending by (value, index) — exact, ties go to the smaller index (the reference's strict `<` keeps the earlier
 * sample, tools/find_similar_imgs.py:83).  ws: DEVICE scratch, 16-byte aligned, size from b2c_topk_workspace_bytes. */
int b2c_topk_workspace_bytes(int k, size_t* bytes);
int b2c_topk_smallest(const float* scores, int64_t n, int k, int32_t* out_idx, float* out_val, void* ws,
                      size_t ws_bytes, b2c_stream stream);

/* Greedy diversity ordering (_3_label_images.py:128-177): order[0] = first_row; for step s < steps:
 * order[s+1] = the row r among samples[s*S .. s*S+S) (DEVICE int32, the host draws them with the reference's
 * random.sample sequence) whose maximum cosine similarity to rows order[0..s] is smallest (first such position).
 * maxsim: DEVICE f32[n] scratch (on return: max similarity of every row to the selected set); order: DEVICE
 * int32[steps+1]. */
int b2c_diversity_order(const void* emb, int dtype, int64_t n, int E, int64_t row_stride, int32_t first_row,
                        const int32_t* samples, int steps, int S, float* maxsim, int32_t* order, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K13 — the 22 img_stat_* scalars.  Replaces ImageFeaturizer.process
 * (utils/image_features.py:52-94, called at utils/embedder.py:170): cv2.resize(INTER_AREA) to ~768^2 pixels,
 * COLOR_BGR2GRAY / COLOR_BGR2HSV, channel means / standard deviations, colourfulness, grey-histogram entropy and
 * Laplacian variance — on the device-resident uint8 image the 4-crop preprocess reads, bit-exact in the integer
 * stages (OpenCV 4.13 arithmetic) and to float64 round-off in the final formulas.
 * ------------------------------------------------------------------------------------------- */
#define B2C_IMG_STATS 22 /* order of utils/image_features.py:63-86: width, height, aspect_ratio, mean_color, std_color,
                            mean_red/green/blue, std_red/green/blue, mean_gray, std_gray, mean_hue/sat/val,
                            std_hue/sat/val, colorfulness, image_entropy, laplacian_variance */
/* Size the reference resizes a W x H image to before measuring it (utils/image_features.py:57-60, including its
 * width/height swap).  Host-only. */
int b2c_image_stats_target_size(int W, int H, int* new_w, int* new_h);
/* Bytes of DEVICE workspace b2c_image_stats needs for these B images (H, W: HOST int arrays). */
int b2c_image_stats_workspace_bytes(const int* H, const int* W, int B, size_t* bytes);
/* img_ptrs / H / W / pitch as in b2c_preprocess_4crop.  out: DEVICE f64[B, 22].  ws: 256-byte aligned. */
int b2c_image_stats(const uint8_t* const* img_ptrs, const int* H, const int* W, const int* pitch, int B, double* out,
                    void* ws, size_t ws_bytes, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K12 — SimpleFC training step on the device.  Replaces the inner loop of
 * _4_train_model.py:196-204 (zero_grad / forward / MSELoss / backward / Adam.step) for utils/nn_model.SimpleFC.
 * The host keeps the reference's data order, split, initialisation and learning-rate schedule (trainer.py).
 * ------------------------------------------------------------------------------------------- */
typedef struct b2c_trainer b2c_trainer;

typedef struct {
  int32_t n_layers;                     /* number of Linear layers (hidden + 1), <= B2C_MLP_MAX_LAYERS */
  int32_t dims[B2C_MLP_MAX_LAYERS + 1]; /* dims[0] = input width, dims[n_layers] = outputs (1) */
  int32_t max_batch;                    /* largest batch a step will see, <= 64 (_4_train_model.py:250 default 16) */
  float leaky_slope;                    /* nn.LeakyReLU default 0.01 (utils/nn_model.py:28) */
  float dropout_p;                      /* nn.Dropout p after every hidden activation (utils/nn_model.py:29), train mode */
  uint64_t seed;                        /* dropout stream: Philox4x32-10, key = seed, counter = (unit/4, layer, step) */
} b2c_trainer_cfg;

typedef struct {
  float lr, beta1, beta2, eps, weight_decay; /* torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay) */
} b2c_adam;

int b2c_trainer_create(const b2c_trainer_cfg* cfg, b2c_trainer** out); /* weights and Adam moments start at zero */
int b2c_trainer_destroy(b2c_trainer* t);
/* W: DEVICE f32[dims[layer+1], dims[layer]] (torch Linear layout), bias: DEVICE f32[dims[layer+1]]; copied on `stream`. */
int b2c_trainer_set_layer(b2c_trainer* t, int layer, const float* W, const float* bias, b2c_stream stream);
int b2c_trainer_get_layer(b2c_trainer* t, int layer, float* W, float* bias, b2c_stream stream);
/* Fills `out` with pointers to the trainer's live weights, for b2c_mlp_score (test-set loss, _4_train_model.py:129-166). */
int b2c_trainer_weights(b2c_trainer* t, b2c_mlp_weights* out);
unsigned long long b2c_trainer_steps(const b2c_trainer* t); /* optimiser steps taken so far */
/* One pass over `n` samples in the given order, `batch` per step (the last step takes the remainder, like
 * DataLoader(drop_last=False)).  feats: DEVICE f32 rows of width dims[0], row i at feats + i*feat_stride; labels: DEVICE
 * f32 indexed like feats; order: DEVICE int32[n] sample indices (the epoch's shuffled permutation).  loss_sum: DEVICE
 * f32, += the mean squared error of every step (train_loss of _4_train_model.py:203) or NULL.  Asynchronous. */
int b2c_trainer_epoch(b2c_trainer* t, const float* feats, int64_t feat_stride, const float* labels, const int32_t* order,
                      int64_t n, int batch, const b2c_adam* hyper, float* loss_sum, b2c_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K14 — JPEG decode ahead of K0.  Replaces `Image.open(path).convert('RGB')` of
 * CustomImageDataset.__getitem__ (utils/embedder.py:167) for Huffman-coded 8-bit JPEG files — baseline, extended
 * sequential (single- or multi-scan) and progressive — bit-exactly with Pillow /
 * libjpeg-turbo at its defaults (islow inverse DCT, "fancy" chroma upsampling, 16-bit fixed-point YCbCr -> RGB).
 * Split along the one serial stage: Huffman decoding runs on the host (b2c_jpeg_decode_coefs, thread-safe, no CUDA
 * calls — DataLoader workers use it), everything after it on the device (b2c_jpeg_reconstruct, batched).
 * Streams this path does not cover — arithmetic-coded, lossless, 12-bit, CMYK / Adobe RGB, sampling other than
 * 4:4:4 / 4:2:2 / 4:2:0 / grey — return B2C_ERR_UNSUPPORTED and stay on the caller's Pillow path.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t width, height, ncomp;     /* ncomp 1 (grey) or 3 (YCbCr) */
  int32_t hs[3], vs[3];             /* sampling factors per component */
  int32_t mcus_x, mcus_y;
  int32_t blocks_w[3], blocks_h[3]; /* 8x8 blocks per component, whole MCUs */
  int32_t comp_w[3], comp_h[3];     /* real samples per component (libjpeg's downsampled_width / _height) */
  int32_t restart_interval;
  int32_t adobe_transform0;
  int32_t progressive;              /* SOF2: several scans refine the same coefficient buffer */
  int32_t reserved_;
  int64_t coef_offset[3];           /* component c's blocks start at coefs + coef_offset[c]; block (by,bx) at +(by*blocks_w+bx)*64 */
  int64_t coef_count;               /* int16 elements the (dense) coefficient buffer needs */
  /* packed form (b2c_jpeg_decode_packed): value offsets u32[nblocks] | counts u8[nblocks] | values i16[...] at these byte
   * offsets of one buffer; block b keeps its first counts[b] coefficients in scan (zigzag) order at element offs[b] of
   * the values (blocks of a single-scan baseline file are appended in decode order, hence explicit offsets).  Blocks
   * are numbered component by component, row-major. */
  int32_t nblocks, reserved2_;
  int64_t offs_off, counts_off, vals_off;
  int64_t packed_capacity;          /* bytes a buffer must have for b2c_jpeg_decode_packed (worst case) */
  int64_t packed_bytes;             /* bytes actually used (multiple of 16), set by b2c_jpeg_decode_packed */
  uint16_t qt[3][64];               /* quantisation table of each component, natural (row-major) order */
} b2c_jpeg_info;

/* Host only: marker parse up to the first scan; fills `info` (sizes for the coefficient buffer; the quantisation
 * tables are final only after b2c_jpeg_decode_coefs). */
int b2c_jpeg_parse(const uint8_t* data, size_t len, b2c_jpeg_info* info);
/* Host only: parse + Huffman decode of every scan into `coefs` (HOST memory, ideally pinned; capacity in int16 elements):
 * de-zigzagged, not dequantised. */
int b2c_jpeg_decode_coefs(const uint8_t* data, size_t len, b2c_jpeg_info* info, int16_t* coefs, size_t capacity);
/* Host only: like b2c_jpeg_decode_coefs, but the result is the packed form — typically 4-10x smaller than the dense
 * buffer, which is what travels from the worker to the main process and over PCIe.  `scratch`: coef_count int16 of host
 * scratch (the dense decode target; NULL = allocate internally per call). */
int b2c_jpeg_decode_packed(const uint8_t* data, size_t len, b2c_jpeg_info* info, uint8_t* packed, size_t capacity,
                           int16_t* scratch);
int b2c_jpeg_workspace_bytes(const b2c_jpeg_info* infos, int n, size_t* bytes);
/* Device: coefs[i] = DEVICE copy of image i's coefficient buffer, outs[i] = DEVICE uint8 [height, width, 3] with row
 * pitch out_pitch[i] bytes.  infos / pointer arrays are host arrays.  Two launches for the whole batch. */
int b2c_jpeg_reconstruct(const b2c_jpeg_info* infos, const int16_t* const* coefs, uint8_t* const* outs,
                         const int* out_pitch, int n, void* ws, size_t ws_bytes, b2c_stream stream);
/* Same from the packed form: packed[i] = DEVICE copy (16-byte aligned) of image i's packed buffer. */
int b2c_jpeg_reconstruct_packed(const b2c_jpeg_info* infos, const uint8_t* const* packed, uint8_t* const* outs,
                                const int* out_pitch, int n, void* ws, size_t ws_bytes, b2c_stream stream);

/* ---- K14b: the Huffman stage on the device (baseline / extended sequential files whose single scan interleaves all
 * components — what cameras and encoders write by default).  The host only parses the markers (b2c_jpeg_huff_prepare);
 * the file's bytes go to the device as they are, and one CTA per image (i) removes the 0xFF00 byte stuffing and finds
 * the end of the entropy-coded segment, (ii) decodes 128-byte sub-sequences of it in parallel from speculative states
 * and lets each thread run on into its successors until its state (bit position, block of the MCU, zigzag index)
 * coincides with theirs — Huffman streams self-synchronise after a few code words —, (iii) re-decodes every
 * sub-sequence from its now known entry state, writing coefficients (natural order, dense: the layout
 * b2c_jpeg_reconstruct takes) and CHECKING that it reproduces the recorded exit state, so a stream that did not
 * synchronise, is damaged or ends early is reported in status[i] instead of decoded wrongly, and (iv) turns the DC
 * differences into DC values with a prefix sum per component.  Progressive files, multi-scan files and streams the
 * device reports keep the host stage (b2c_jpeg_decode_packed). */
typedef struct {
  uint16_t look[512];  /* 9-bit lookahead: (code length << 8) | symbol, 0 = the code is longer than 9 bits */
  int32_t maxcode[18]; /* largest code of each length (-1: none), [17] = sentinel */
  int32_t valoffset[18];
  uint8_t vals[256];
} b2c_jpeg_hufftab;
typedef struct {
  int64_t scan_begin;       /* offset of the first entropy-coded byte in the file */
  int64_t scan_bytes;       /* bytes from there to the end of the file (the device finds the terminating marker) */
  int32_t restart_interval; /* MCUs between RSTn markers, 0 = none */
  int32_t reserved_;
  b2c_jpeg_hufftab tab[4];  /* DC luma, AC luma, DC chroma, AC chroma */
} b2c_jpeg_huff;

/* Host only: marker parse; fills `info` (final, including the quantisation tables) and `huff`.  B2C_ERR_UNSUPPORTED
 * for anything but one interleaved sequential Huffman scan (keep those on b2c_jpeg_decode_packed). */
int b2c_jpeg_huff_prepare(const uint8_t* data, size_t len, b2c_jpeg_info* info, b2c_jpeg_huff* huff);
int b2c_jpeg_huff_workspace_bytes(const b2c_jpeg_huff* huffs, int n, size_t* bytes);
/* Device: files[i] = DEVICE copy of image i's file bytes (at least scan_begin + scan_bytes of them); coefs[i] = DEVICE
 * int16[coef_count] (128-byte aligned), filled in the dense natural-order form; status: DEVICE int32[n], 0 = decoded,
 * non-zero = not decodable here (B2C_JPEG_HUFF_*).  One launch for the batch; asynchronous. */
#define B2C_JPEG_HUFF_CORRUPT 1   /* bad code, run past the block, block count mismatch */
#define B2C_JPEG_HUFF_TRUNCATED 2 /* the entropy-coded segment ends before the last MCU */
#define B2C_JPEG_HUFF_NOSYNC 4    /* a sub-sequence's verification failed (no synchronisation) */
int b2c_jpeg_huff_decode(const b2c_jpeg_info* infos, const b2c_jpeg_huff* huffs, const uint8_t* const* files,
                         int16_t* const* coefs, int32_t* status, int n, void* ws, size_t ws_bytes, b2c_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H_ */
