// b2c_vit.cu — the open_clip VisionTransformer forward as a sequence of sm_100a kernels behind an
// opaque handle.  Replaces model.encode_image + the in-place L2 normalise of
// CLIP_Encoder.encode_image (utils/embedder.py:94-100).  Architecture per SURVEY.md App. A:
//   conv1 (no bias) -> [cls ; patches] + pos -> ln_pre -> L x { x += out_proj(MHA(ln_1 x)) ;
//   x += c_proj(act(c_fc(ln_2 x))) } -> ln_post(cls) @ proj -> / ||.||
// The residual stream and all LayerNorm statistics are fp32; GEMM/attention operands are bf16 with
// fp32 accumulation.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "b2c_launch.h"
#include "b2c_umma_pipeline.cuh"

using namespace b2c;

struct b2c_vit_layer {
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  __nv_bfloat16 *w_qkv = nullptr, *w_out = nullptr, *w_fc = nullptr, *w_proj = nullptr;
  float *b_qkv = nullptr, *b_out = nullptr, *b_fc = nullptr, *b_proj = nullptr;
  CUtensorMap tm_qkv_h, tm_out_h, tm_fc_h, tm_proj_h;  // box 128 rows: each CTA of a pair loads half of the tile's N
  // LayerNorm folded into in_proj / c_fc (ln_fold_kernel): gamma-scaled weights, their column sums, beta·Wᵀ + b
  __nv_bfloat16 *wf_qkv = nullptr, *wf_fc = nullptr;
  float *cs_qkv = nullptr, *bf_qkv = nullptr, *cs_fc = nullptr, *bf_fc = nullptr;
  float *src_qkv = nullptr, *src_fc = nullptr;  // f32 staging of the original weights until the first fold
  CUtensorMap tm_qkvf_h, tm_fcf_h;
};

struct b2c_vit {
  b2c_vit_cfg cfg;
  int T, g, G2, Kp, hd;
  int chunk;  // crops processed per pass over the layers
  // A chunk is split into `lanes` independent sub-chunks, each with its own workspace slice and its own stream, so
  // that the HBM-bound stages of one lane (LayerNorm, the residual reduce-add tail, patchify) and the kernel heads /
  // tails of every stage run under the tensor-bound GEMMs of the other lane.
  int lanes = 2;
  // LayerNorm fused into the GEMMs on either side of it (default; B2C_VIT_FUSED_LN=0 or b2c_vit_set_fused_ln(0)
  // selects the stand-alone LayerNorm kernels + TMA reduce-add residual epilogue)
  bool fused_ln = true;
  bool fold_dirty = true;
  // Opt-in (b2c_vit_set_cls_only_last_block): the last block evaluates only what ln_post / proj read — the class-token
  // row.  K and V of all tokens are still computed (in_proj runs in full); attention, out_proj, ln_2, c_fc and c_proj run
  // on one row per crop.  Off by default: bench.py's headline executes every block in full.
  bool cls_only_last = false;
  // Opt-in (b2c_vit_set_graph / B2C_VIT_GRAPH=1): a pass whose buffers, size and switches repeat is captured once into a
  // CUDA graph (second call with the same key; the first runs eagerly so that every one-time attribute is set) and
  // replayed afterwards — ~250 launches and their fork/join events become one cudaGraphLaunch.
  bool use_graph = false;
  struct GraphEntry {
    const void* in;
    const float* out;
    const void* ws;
    int n, dtype, lanes;
    bool pixels, fused, cls;
    cudaGraphExec_t exec;  // nullptr: key seen once, not captured yet
    unsigned long long launches;
  };
  std::vector<GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;  // captures run on an internal stream (the caller's may be the legacy default stream)
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_join[3] = {nullptr, nullptr, nullptr};
  __nv_bfloat16* conv1 = nullptr;  // [d, Kp]
  float *cls = nullptr, *pos = nullptr, *proj = nullptr;
  float *ln_pre_w = nullptr, *ln_pre_b = nullptr, *ln_post_w = nullptr, *ln_post_b = nullptr;
  CUtensorMap tm_conv1;
  std::vector<b2c_vit_layer> layers;
  std::map<std::string, bool> have;
  std::vector<void*> allocs;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  size_t x, h, big, patches, xb, stats, stats2, shift, head, total;
};

WsLayout ws_layout(const b2c_vit* v, int nc, bool need_patches) {
  const size_t M = static_cast<size_t>(nc) * v->T;
  const size_t d = v->cfg.width;
  const size_t wide = std::max<size_t>(3 * d, v->cfg.mlp);
  WsLayout w;
  size_t off = 0;
  w.x = off;
  off = align_up(off + M * d * sizeof(float), 1024);
  w.h = off;
  off = align_up(off + M * d * 2, 1024);
  w.big = off;
  off = align_up(off + M * wide * 2, 1024);
  w.patches = off;
  if (need_patches) off = align_up(off + static_cast<size_t>(nc) * v->G2 * v->Kp * 2, 1024);
  w.xb = off;  // bf16 copy of the residual stream + per-row statistics partials (LayerNorm-fused layer loop)
  off = align_up(off + M * d * 2, 1024);
  w.stats = off;  // two statistics buffers: a residual update reads the previous update's and writes its own
  off = align_up(off + M * (d / 256) * sizeof(float2), 1024);
  w.stats2 = off;
  off = align_up(off + M * (d / 256) * sizeof(float2), 1024);
  w.shift = off;
  off = align_up(off + M * sizeof(float), 1024);
  w.head = off;  // partial squared norms of the head's column blocks
  off = align_up(off + head_part_floats(nc, v->cfg.embed) * sizeof(float), 1024);
  w.total = off;
  return w;
}

constexpr int kMinLaneCrops = 64;  // below this a lane's GEMMs no longer fill the machine

int lanes_for(const b2c_vit* v, int nc) {
  int l = v->lanes;
  while (l > 1 && nc / l < kMinLaneCrops) --l;
  return l;
}

// workspace of one pass over `nc` crops: `lanes` slices, and never less than the single-lane layout (the stage timer
// runs single-lane inside the same buffer)
size_t ws_bytes_for(const b2c_vit* v, int nc, bool need_patches, size_t* lane_bytes) {
  const int l = lanes_for(v, nc);
  const int per = (nc + l - 1) / l;
  const size_t lb = ws_layout(v, per, need_patches).total;
  if (lane_bytes) *lane_bytes = lb;
  return std::max(lb * l, ws_layout(v, nc, need_patches).total);
}

int ensure_lanes(b2c_vit* v) {
  if (v->ev_fork) return 0;
  B2C_CHECK_CUDA(cudaEventCreateWithFlags(&v->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 3; ++i) {
    B2C_CHECK_CUDA(cudaStreamCreateWithFlags(&v->side[i], cudaStreamNonBlocking));
    B2C_CHECK_CUDA(cudaEventCreateWithFlags(&v->ev_join[i], cudaEventDisableTiming));
  }
  return 0;
}

std::vector<std::string> expected_keys(const b2c_vit* v) {
  std::vector<std::string> k = {"class_embedding", "positional_embedding", "proj",        "conv1.weight",
                                "ln_pre.weight",   "ln_pre.bias",          "ln_post.weight", "ln_post.bias"};
  for (int i = 0; i < v->cfg.layers; ++i) {
    const std::string p = "transformer.resblocks." + std::to_string(i) + ".";
    for (const char* s : {"ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight",
                          "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight",
                          "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"})
      k.push_back(p + s);
  }
  return k;
}

int alloc_dev(b2c_vit* v, void** p, size_t bytes) {
  if (*p) return 0;
  B2C_CHECK_CUDA(cudaMalloc(p, bytes));
  v->allocs.push_back(*p);
  return 0;
}

int store_f32(b2c_vit* v, float** dst, const void* src, int dtype, int64_t count, int64_t expect, const char* key) {
  B2C_REQUIRE(count == expect, "set_weight(%s): %lld elements, expected %lld", key, (long long)count, (long long)expect);
  B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(dst), static_cast<size_t>(count) * sizeof(float)));
  return convert_launch(src, dtype, *dst, B2C_F32, count, nullptr);
}

int store_bf16(b2c_vit* v, __nv_bfloat16** dst, const void* src, int dtype, int64_t rows, int cols, int cols_padded,
               int64_t count, const char* key) {
  B2C_REQUIRE(count == rows * cols, "set_weight(%s): %lld elements, expected %lld", key, (long long)count,
              (long long)(rows * cols));
  B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(dst), static_cast<size_t>(rows) * cols_padded * 2));
  return pad_rows_bf16_launch(src, dtype, *dst, rows, cols, cols_padded, nullptr);
}

int stage_f32(float** dst, const void* src, int dtype, int64_t count) {
  if (!*dst) B2C_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), static_cast<size_t>(count) * sizeof(float)));
  return convert_launch(src, dtype, *dst, B2C_F32, count, nullptr);
}

// (re)compute the gamma-folded weights of every block; the f32 staging copies are released after their first use
int ensure_folded(b2c_vit* v, cudaStream_t stream) {
  if (!v->fold_dirty) return 0;
  const b2c_vit_cfg& c = v->cfg;
  const int d = c.width;
  bool staged = false;
  for (b2c_vit_layer& L : v->layers) {
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.wf_qkv), static_cast<size_t>(3) * d * d * 2));
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.wf_fc), static_cast<size_t>(c.mlp) * d * 2));
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.cs_qkv), static_cast<size_t>(3) * d * 4));
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.bf_qkv), static_cast<size_t>(3) * d * 4));
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.cs_fc), static_cast<size_t>(c.mlp) * 4));
    B2C_TRY(alloc_dev(v, reinterpret_cast<void**>(&L.bf_fc), static_cast<size_t>(c.mlp) * 4));
    B2C_TRY(ln_fold_launch(L.src_qkv ? static_cast<const void*>(L.src_qkv) : L.w_qkv, L.src_qkv ? B2C_F32 : B2C_BF16, L.ln1_w,
                           L.ln1_b, L.b_qkv, L.wf_qkv, L.cs_qkv, L.bf_qkv, 3 * d, d, stream));
    B2C_TRY(ln_fold_launch(L.src_fc ? static_cast<const void*>(L.src_fc) : L.w_fc, L.src_fc ? B2C_F32 : B2C_BF16, L.ln2_w,
                           L.ln2_b, L.b_fc, L.wf_fc, L.cs_fc, L.bf_fc, c.mlp, d, stream));
    B2C_TRY(make_tmap_2d(&L.tm_qkvf_h, L.wf_qkv, 3 * d, d, static_cast<uint64_t>(d) * 2, kBM, 1));
    B2C_TRY(make_tmap_2d(&L.tm_fcf_h, L.wf_fc, c.mlp, d, static_cast<uint64_t>(d) * 2, kBM, 1));
    staged = staged || L.src_qkv || L.src_fc;
  }
  if (staged) {
    B2C_CHECK_CUDA(cudaStreamSynchronize(stream));
    for (b2c_vit_layer& L : v->layers) {
      if (L.src_qkv) cudaFree(L.src_qkv);
      if (L.src_fc) cudaFree(L.src_fc);
      L.src_qkv = L.src_fc = nullptr;
    }
  }
  v->fold_dirty = false;
  return 0;
}

// The layer loop with LayerNorm fused into the GEMMs on either side of it: no stand-alone LayerNorm launches, and the
// residual stream is read once per residual update (by the GEMM epilogue that rewrites it) instead of twice.
int forward_chunk_fused(b2c_vit* v, const void* patches, int nc, float* out, uint8_t* ws, const WsLayout& w,
                        cudaStream_t stream) {
  const b2c_vit_cfg& c = v->cfg;
  const int d = c.width;
  const int nblk = d / 256;
  const int64_t M = static_cast<int64_t>(nc) * v->T;
  float* x = reinterpret_cast<float*>(ws + w.x);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(ws + w.h);
  __nv_bfloat16* big = reinterpret_cast<__nv_bfloat16*>(ws + w.big);
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws + w.xb);
  float2* stats_a = reinterpret_cast<float2*>(ws + w.stats);   // statistics at the start of a block (and after c_proj)
  float2* stats_b = reinterpret_cast<float2*>(ws + w.stats2);  // statistics after out_proj
  float* shift = reinterpret_cast<float*>(ws + w.shift);
  const float eps = 1e-5f;

  CUtensorMap tm_patches, tm_h, tm_xb, tm_mlp, st_qkv, st_mlp, st_x, st_xb;
  B2C_TRY(make_tmap_2d(&tm_patches, patches, static_cast<uint64_t>(nc) * v->G2, v->Kp, static_cast<uint64_t>(v->Kp) * 2,
                       kBM, 1));
  B2C_TRY(make_tmap_2d(&tm_h, h, M, d, static_cast<uint64_t>(d) * 2, kBM, 1));
  B2C_TRY(make_tmap_2d(&tm_xb, xb, M, d, static_cast<uint64_t>(d) * 2, kBM, 1));
  B2C_TRY(make_tmap_2d(&tm_mlp, big, M, c.mlp, static_cast<uint64_t>(c.mlp) * 2, kBM, 1));
  B2C_TRY(make_out_tmap(&st_qkv, big, M, 3 * d, 3 * d, kGemmBiasBf16));
  B2C_TRY(make_out_tmap(&st_mlp, big, M, c.mlp, c.mlp, kGemmBiasBf16));
  B2C_TRY(make_out_tmap(&st_x, x, M, d, d, kGemmResidLnF32));
  B2C_TRY(make_out_tmap(&st_xb, xb, M, d, d, kGemmResidLnBf16Copy));

  {
    ProfScope ps(B2C_PROF_PATCH_EMBED, stream);
    GemmLaunch gl{};
    gl.tmap_a = tm_patches; gl.tmap_b = v->tm_conv1; gl.tmap_b_half = v->tm_conv1; gl.tmap_out = tm_patches;
    gl.M = static_cast<int64_t>(nc) * v->G2; gl.N = d; gl.K = v->Kp; gl.mode = kGemmPatchEmbedF32;
    gl.out = x; gl.ldo = d; gl.pos = v->pos; gl.T = v->T; gl.G2 = v->G2;
    B2C_TRY(gemm_launch(gl, stream));
  }
  {
    ProfScope ps(B2C_PROF_LAYERNORM, stream);
    B2C_TRY(cls_pos_launch(x, v->cls, v->pos, nc, v->T, d, stream));
    B2C_TRY(layernorm_pre_launch(x, v->ln_pre_w, v->ln_pre_b, xb, stats_a, shift, M, d, eps, stream));
  }
  auto ln_gemm = [&](int kind, const CUtensorMap& tm_w, const CUtensorMap& st, int N, int mode, const float* bias_f,
                     const float* colsum, int64_t ldo, float2* stats) -> int {
    ProfScope ps(kind, stream);
    GemmLaunch gl{};
    gl.tmap_a = tm_xb; gl.tmap_b = tm_w; gl.tmap_b_half = tm_w; gl.tmap_out = st; gl.M = M; gl.N = N; gl.K = d;
    gl.mode = mode; gl.bias = bias_f; gl.colsum = colsum; gl.stats = stats; gl.shift = shift; gl.nblk = nblk; gl.eps = eps;
    gl.out = big; gl.ldo = ldo;
    return gemm_launch(gl, stream);
  };
  auto resid_gemm = [&](int kind, const CUtensorMap& tm_a, const CUtensorMap& tm_w_full, const CUtensorMap& tm_w, int K,
                        const float* bias, bool last, const float2* stats_in, float2* stats) -> int {
    ProfScope ps(kind, stream);
    GemmLaunch gl{};
    gl.tmap_a = tm_a; gl.tmap_b = tm_w_full; gl.tmap_b_half = tm_w; gl.tmap_out = st_x; gl.tmap_out2 = st_xb;
    gl.M = M; gl.N = d; gl.K = K; gl.bias = bias; gl.out = x; gl.ldo = d;
    gl.stats = stats; gl.stats_in = stats_in; gl.shift = shift; gl.nblk = nblk; gl.eps = eps;
    // nothing consumes the bf16 copy / statistics of the last residual update: plain reduce-add there
    gl.mode = last ? kGemmBiasResidF32 : kGemmResidLnF32;
    return gemm_launch(gl, stream);
  };
  const int fc_mode = c.act == B2C_ACT_GELU ? kGemmLnBiasGeluBf16 : kGemmLnBiasQGeluBf16;
  for (int li = 0; li < c.layers; ++li) {
    const b2c_vit_layer& L = v->layers[li];
    B2C_TRY(ln_gemm(B2C_PROF_IN_PROJ, L.tm_qkvf_h, st_qkv, 3 * d, kGemmLnBiasBf16, L.bf_qkv, L.cs_qkv, 3 * d, stats_a));
    if (li + 1 == c.layers && v->cls_only_last && v->hd == 64) {
      // class-token rows only: the GEMMs address row crop*T of h / x in place through strided tensor maps (M = nc)
      const int64_t ldr = static_cast<int64_t>(v->T) * d;
      CUtensorMap tm_hc, tm_xbc, tm_bigc, st_xc, st_bigc;
      B2C_TRY(make_tmap_2d(&tm_hc, h, nc, d, static_cast<uint64_t>(ldr) * 2, kBM, 1));
      B2C_TRY(make_tmap_2d(&tm_xbc, xb, nc, d, static_cast<uint64_t>(d) * 2, kBM, 1));
      B2C_TRY(make_tmap_2d(&tm_bigc, big, nc, c.mlp, static_cast<uint64_t>(c.mlp) * 2, kBM, 1));
      B2C_TRY(make_out_tmap(&st_xc, x, nc, d, ldr, kGemmBiasResidF32));
      B2C_TRY(make_out_tmap(&st_bigc, big, nc, c.mlp, c.mlp, kGemmBiasBf16));
      auto row_gemm = [&](int kind, const CUtensorMap& ta, const CUtensorMap& tw_full, const CUtensorMap& tw, const CUtensorMap& st,
                          int N, int K, int mode, const float* bias, void* o, int64_t ldo) -> int {
        ProfScope ps(kind, stream);
        GemmLaunch gl{};
        gl.tmap_a = ta; gl.tmap_b = tw_full; gl.tmap_b_half = tw; gl.tmap_out = st; gl.M = nc; gl.N = N; gl.K = K;
        gl.mode = mode; gl.bias = bias; gl.out = o; gl.ldo = ldo;
        return gemm_launch(gl, stream);
      };
      {
        ProfScope ps(B2C_PROF_ATTENTION, stream);
        B2C_TRY(attention_cls_launch(big, h, nc, v->T, c.heads, v->hd, stream));
      }
      B2C_TRY(row_gemm(B2C_PROF_OUT_PROJ, tm_hc, L.tm_out_h, L.tm_out_h, st_xc, d, d, kGemmBiasResidF32, L.b_out, x, ldr));
      {
        ProfScope ps(B2C_PROF_LAYERNORM, stream);
        B2C_TRY(layernorm_bf16_strided_launch(x, ldr, L.ln2_w, L.ln2_b, xb, nc, d, eps, stream));
      }
      B2C_TRY(row_gemm(B2C_PROF_C_FC, tm_xbc, L.tm_fc_h, L.tm_fc_h, st_bigc, c.mlp, d,
                       c.act == B2C_ACT_GELU ? kGemmBiasGeluBf16 : kGemmBiasQGeluBf16, L.b_fc, big, c.mlp));
      B2C_TRY(row_gemm(B2C_PROF_C_PROJ, tm_bigc, L.tm_proj_h, L.tm_proj_h, st_xc, d, c.mlp, kGemmBiasResidF32, L.b_proj, x, ldr));
      break;
    }
    {
      ProfScope ps(B2C_PROF_ATTENTION, stream);
      B2C_TRY(attention_launch(big, h, nc, v->T, c.heads, v->hd, stream));
    }
    B2C_TRY(resid_gemm(B2C_PROF_OUT_PROJ, tm_h, L.tm_out_h, L.tm_out_h, d, L.b_out, false, stats_a, stats_b));
    B2C_TRY(ln_gemm(B2C_PROF_C_FC, L.tm_fcf_h, st_mlp, c.mlp, fc_mode, L.bf_fc, L.cs_fc, c.mlp, stats_b));
    B2C_TRY(resid_gemm(B2C_PROF_C_PROJ, tm_mlp, L.tm_proj_h, L.tm_proj_h, c.mlp, L.b_proj, li + 1 == c.layers, stats_b, stats_a));
  }
  ProfScope ps(B2C_PROF_HEAD, stream);
  return head_launch(x, v->ln_post_w, v->ln_post_b, v->proj, out, reinterpret_cast<float*>(ws + w.head), nc, v->T, d, c.embed,
                     eps, stream);
}

int forward_chunk(b2c_vit* v, const void* patches, int nc, float* out, uint8_t* ws, const WsLayout& w,
                  cudaStream_t stream) {
  const b2c_vit_cfg& c = v->cfg;
  const int d = c.width;
  const int64_t M = static_cast<int64_t>(nc) * v->T;
  float* x = reinterpret_cast<float*>(ws + w.x);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(ws + w.h);
  __nv_bfloat16* big = reinterpret_cast<__nv_bfloat16*>(ws + w.big);
  const float eps = 1e-5f;

  CUtensorMap tm_patches, tm_h, tm_mlp;
  B2C_TRY(make_tmap_2d(&tm_patches, patches, static_cast<uint64_t>(nc) * v->G2, v->Kp, static_cast<uint64_t>(v->Kp) * 2,
                       kBM, 1));
  B2C_TRY(make_tmap_2d(&tm_h, h, M, d, static_cast<uint64_t>(d) * 2, kBM, 1));
  B2C_TRY(make_tmap_2d(&tm_mlp, big, M, c.mlp, static_cast<uint64_t>(c.mlp) * 2, kBM, 1));
  // epilogue store maps: qkv / mlp-hidden (bf16 tiled stores), residual stream (f32 reduce-add at L2)
  CUtensorMap st_qkv, st_mlp, st_x;
  B2C_TRY(make_out_tmap(&st_qkv, big, M, 3 * d, 3 * d, kGemmBiasBf16));
  B2C_TRY(make_out_tmap(&st_mlp, big, M, c.mlp, c.mlp, kGemmBiasBf16));
  B2C_TRY(make_out_tmap(&st_x, x, M, d, d, kGemmBiasResidF32));

  // K1: conv1 as a GEMM over patch rows, + positional embedding, scattered past the class token
  {
    ProfScope ps(B2C_PROF_PATCH_EMBED, stream);
    GemmLaunch gl{};
    gl.tmap_a = tm_patches;
    gl.tmap_b = v->tm_conv1;
    gl.tmap_b_half = v->tm_conv1;  // unused: patch-embed runs on the single-CTA kernel
    gl.tmap_out = tm_patches;  // unused by the patch-embed epilogue (direct scattered stores)
    gl.M = static_cast<int64_t>(nc) * v->G2;
    gl.N = d;
    gl.K = v->Kp;
    gl.mode = kGemmPatchEmbedF32;
    gl.out = x;
    gl.ldo = d;
    gl.pos = v->pos;
    gl.T = v->T;
    gl.G2 = v->G2;
    B2C_TRY(gemm_launch(gl, stream));
  }
  {
    ProfScope ps(B2C_PROF_LAYERNORM, stream);
    B2C_TRY(cls_pos_launch(x, v->cls, v->pos, nc, v->T, d, stream));
    B2C_TRY(layernorm_f32_inplace_launch(x, v->ln_pre_w, v->ln_pre_b, M, d, eps, stream));
  }

  for (int li = 0; li < c.layers; ++li) {
    const b2c_vit_layer& L = v->layers[li];
    // attention half
    {
      ProfScope ps(B2C_PROF_LAYERNORM, stream);
      B2C_TRY(layernorm_bf16_launch(x, L.ln1_w, L.ln1_b, h, M, d, eps, stream));
    }
    {
      ProfScope ps(B2C_PROF_IN_PROJ, stream);
      GemmLaunch gl{};
      gl.tmap_a = tm_h; gl.tmap_b = L.tm_qkv_h; gl.tmap_b_half = L.tm_qkv_h; gl.tmap_out = st_qkv; gl.M = M; gl.N = 3 * d; gl.K = d;
      gl.mode = kGemmBiasBf16; gl.bias = L.b_qkv; gl.out = big; gl.ldo = 3 * d;
      B2C_TRY(gemm_launch(gl, stream));
    }
    {
      ProfScope ps(B2C_PROF_ATTENTION, stream);
      B2C_TRY(attention_launch(big, h, nc, v->T, c.heads, v->hd, stream));
    }
    {
      ProfScope ps(B2C_PROF_OUT_PROJ, stream);
      GemmLaunch gl{};
      gl.tmap_a = tm_h; gl.tmap_b = L.tm_out_h; gl.tmap_b_half = L.tm_out_h; gl.tmap_out = st_x; gl.M = M; gl.N = d; gl.K = d;
      gl.mode = kGemmBiasResidF32; gl.bias = L.b_out; gl.out = x; gl.ldo = d;
      B2C_TRY(gemm_launch(gl, stream));
    }
    // MLP half
    {
      ProfScope ps(B2C_PROF_LAYERNORM, stream);
      B2C_TRY(layernorm_bf16_launch(x, L.ln2_w, L.ln2_b, h, M, d, eps, stream));
    }
    {
      ProfScope ps(B2C_PROF_C_FC, stream);
      GemmLaunch gl{};
      gl.tmap_a = tm_h; gl.tmap_b = L.tm_fc_h; gl.tmap_b_half = L.tm_fc_h; gl.tmap_out = st_mlp; gl.M = M; gl.N = c.mlp; gl.K = d;
      gl.mode = c.act == B2C_ACT_GELU ? kGemmBiasGeluBf16 : kGemmBiasQGeluBf16;
      gl.bias = L.b_fc; gl.out = big; gl.ldo = c.mlp;
      B2C_TRY(gemm_launch(gl, stream));
    }
    {
      ProfScope ps(B2C_PROF_C_PROJ, stream);
      GemmLaunch gl{};
      gl.tmap_a = tm_mlp; gl.tmap_b = L.tm_proj_h; gl.tmap_b_half = L.tm_proj_h; gl.tmap_out = st_x; gl.M = M; gl.N = d; gl.K = c.mlp;
      gl.mode = kGemmBiasResidF32; gl.bias = L.b_proj; gl.out = x; gl.ldo = d;
      B2C_TRY(gemm_launch(gl, stream));
    }
  }
  ProfScope ps(B2C_PROF_HEAD, stream);
  return head_launch(x, v->ln_post_w, v->ln_post_b, v->proj, out, reinterpret_cast<float*>(ws + w.head), nc, v->T, d, c.embed,
                     eps, stream);
}

int forward_eager(b2c_vit* v, const void* pixels, int dtype, const void* patches, int n, float* out, void* ws,
                  size_t ws_bytes, cudaStream_t stream) {
  const bool fused = v->fused_ln;
  const int nc_max = n < v->chunk ? n : v->chunk;
  size_t lane_bytes = 0;
  const size_t need = ws_bytes_for(v, nc_max, pixels != nullptr, &lane_bytes);
  if (ws_bytes < need)
    return set_error(B2C_ERR_WORKSPACE, "vit_forward: workspace %zu B < required %zu B", ws_bytes, need);
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "vit_forward: workspace must be 1024-byte aligned");
  uint8_t* wsb = static_cast<uint8_t*>(ws);
  const int R = v->cfg.image;
  const size_t px_elt = dtype == B2C_F32 ? 4 : 2;
  // one sub-chunk [c0, c0 + nc) on `st` inside the workspace slice `wsl`
  auto run = [&](int c0, int nc, uint8_t* wsl, const WsLayout& w, cudaStream_t st) -> int {
    const void* pch;
    if (pixels) {
      const uint8_t* px = static_cast<const uint8_t*>(pixels) + static_cast<size_t>(c0) * 3 * R * R * px_elt;
      ProfScope ps(B2C_PROF_OTHER, st);
      B2C_TRY(patchify_launch(px, dtype, wsl + w.patches, nc, R, v->cfg.patch, v->Kp, st));
      pch = wsl + w.patches;
    } else {
      pch = static_cast<const uint8_t*>(patches) + static_cast<size_t>(c0) * v->G2 * v->Kp * 2;
    }
    float* o = out + static_cast<size_t>(c0) * v->cfg.embed;
    return fused ? forward_chunk_fused(v, pch, nc, o, wsl, w, st) : forward_chunk(v, pch, nc, o, wsl, w, st);
  };
  for (int c0 = 0; c0 < n; c0 += nc_max) {
    const int nc = (n - c0) < nc_max ? (n - c0) : nc_max;
    // the stage timer brackets stages with events on the launching stream: meaningful only without overlap
    const int lanes = g_prof_on.load(std::memory_order_relaxed) ? 1 : lanes_for(v, nc);
    if (lanes == 1) {
      B2C_TRY(run(c0, nc, wsb, ws_layout(v, nc_max, pixels != nullptr), stream));
      continue;
    }
    B2C_TRY(ensure_lanes(v));
    const int per = (nc + lanes - 1) / lanes;
    const WsLayout w = ws_layout(v, (nc_max + lanes - 1) / lanes, pixels != nullptr);
    B2C_REQUIRE(w.total <= lane_bytes && lane_bytes * lanes <= ws_bytes, "vit_forward: lane layout exceeds the workspace");
    B2C_CHECK_CUDA(cudaEventRecord(v->ev_fork, stream));
    int rc = 0;
    int used = 0;
    for (int l = 0; l < lanes && rc == 0; ++l) {
      const int b = l * per;
      const int cnt = (nc - b) < per ? (nc - b) : per;
      if (cnt <= 0) break;
      cudaStream_t st = l == 0 ? stream : v->side[l - 1];
      if (l > 0) B2C_CHECK_CUDA(cudaStreamWaitEvent(st, v->ev_fork, 0));
      rc = run(c0 + b, cnt, wsb + static_cast<size_t>(l) * lane_bytes, w, st);
      used = l + 1;
    }
    // always join what was forked, also on error, so the caller's stream order covers every lane
    for (int l = 1; l < used; ++l) {
      cudaEventRecord(v->ev_join[l - 1], v->side[l - 1]);
      cudaStreamWaitEvent(stream, v->ev_join[l - 1], 0);
    }
    if (rc != 0) return rc;
  }
  return 0;
}

int forward_impl(b2c_vit* v, const void* pixels, int dtype, const void* patches, int n, float* out, void* ws,
                 size_t ws_bytes, cudaStream_t stream) {
  B2C_REQUIRE(v && out && ws, "vit_forward: null pointer");
  B2C_REQUIRE(n > 0, "vit_forward: n_crops must be positive");
  B2C_TRY(b2c_vit_ready(v));
  if (v->fused_ln) B2C_TRY(ensure_folded(v, stream));  // (may synchronise: never inside a capture)
  if (!v->use_graph || g_prof_on.load(std::memory_order_relaxed))
    return forward_eager(v, pixels, dtype, patches, n, out, ws, ws_bytes, stream);
  const void* in = pixels ? pixels : patches;
  b2c_vit::GraphEntry* e = nullptr;
  for (b2c_vit::GraphEntry& g : v->graphs)
    if (g.in == in && g.out == out && g.ws == ws && g.n == n && g.dtype == dtype && g.lanes == v->lanes &&
        g.pixels == (pixels != nullptr) && g.fused == v->fused_ln && g.cls == v->cls_only_last)
      e = &g;
  if (e && e->exec) {
    B2C_CHECK_CUDA(cudaGraphLaunch(e->exec, stream));
    g_launches.fetch_add(e->launches, std::memory_order_relaxed);  // the graph's kernel nodes
    return 0;
  }
  if (!e) {  // first sight of this key: eager pass (sets function attributes, creates the lanes' streams)
    if (v->graphs.size() >= 16) {
      if (v->graphs.front().exec) cudaGraphExecDestroy(v->graphs.front().exec);
      v->graphs.erase(v->graphs.begin());
    }
    v->graphs.push_back({in, out, ws, n, dtype, v->lanes, pixels != nullptr, v->fused_ln, v->cls_only_last, nullptr, 0});
    return forward_eager(v, pixels, dtype, patches, n, out, ws, ws_bytes, stream);
  }
  B2C_TRY(ensure_lanes(v));
  if (!v->cap_stream) B2C_CHECK_CUDA(cudaStreamCreateWithFlags(&v->cap_stream, cudaStreamNonBlocking));
  const unsigned long long before = g_launches.load(std::memory_order_relaxed);
  B2C_CHECK_CUDA(cudaStreamBeginCapture(v->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = forward_eager(v, pixels, dtype, patches, n, out, ws, ws_bytes, v->cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(v->cap_stream, &graph);
  const unsigned long long nodes = g_launches.load(std::memory_order_relaxed) - before;
  g_launches.fetch_sub(nodes, std::memory_order_relaxed);  // nothing has run yet
  if (rc != 0 || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    if (rc != 0) return rc;
    return set_error(B2C_ERR_CUDA, "vit_forward: stream capture failed: %s", cudaGetErrorString(ce));
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return set_error(B2C_ERR_CUDA, "vit_forward: cudaGraphInstantiate: %s", cudaGetErrorString(ie));
  e->exec = exec;
  e->launches = nodes;
  B2C_CHECK_CUDA(cudaGraphLaunch(exec, stream));
  g_launches.fetch_add(nodes, std::memory_order_relaxed);
  return 0;
}

}  // namespace

extern "C" int b2c_vit_create(const b2c_vit_cfg* cfg, b2c_vit** out) {
  B2C_REQUIRE(cfg && out, "b2c_vit_create: null pointer");
  B2C_REQUIRE(cfg->patch > 0 && cfg->image > 0 && cfg->image % cfg->patch == 0, "b2c_vit_create: image %d / patch %d",
              cfg->image, cfg->patch);
  B2C_REQUIRE(cfg->width > 0 && cfg->width % 256 == 0, "b2c_vit_create: width %d must be a multiple of 256", cfg->width);
  B2C_REQUIRE(cfg->mlp > 0 && cfg->mlp % 256 == 0, "b2c_vit_create: mlp %d must be a multiple of 256", cfg->mlp);
  B2C_REQUIRE(cfg->heads > 0 && cfg->width % cfg->heads == 0, "b2c_vit_create: width %d / heads %d", cfg->width,
              cfg->heads);
  const int hd = cfg->width / cfg->heads;
  B2C_REQUIRE(hd == 64 || hd == 80, "b2c_vit_create: head dim %d unsupported (64 or 80)", hd);
  B2C_REQUIRE(cfg->layers > 0 && cfg->layers <= 64, "b2c_vit_create: layers %d", cfg->layers);
  B2C_REQUIRE(cfg->embed > 0 && cfg->embed <= 4096, "b2c_vit_create: embed %d out of range", cfg->embed);
  B2C_REQUIRE(cfg->act == B2C_ACT_QUICK_GELU || cfg->act == B2C_ACT_GELU, "b2c_vit_create: act %d", cfg->act);
  b2c_vit* v = new b2c_vit();
  v->cfg = *cfg;
  v->g = cfg->image / cfg->patch;
  v->G2 = v->g * v->g;
  v->T = v->G2 + 1;
  v->Kp = (3 * cfg->patch * cfg->patch + kBK - 1) / kBK * kBK;
  v->hd = hd;
  v->chunk = 1024;
  if (const char* e = getenv("B2C_VIT_CHUNK")) {
    const int cv = atoi(e);
    if (cv > 0) v->chunk = cv;
  }
  if (const char* e = getenv("B2C_VIT_FUSED_LN")) v->fused_ln = atoi(e) != 0;
  if (const char* e = getenv("B2C_VIT_GRAPH")) v->use_graph = atoi(e) != 0;
  if (const char* e = getenv("B2C_VIT_LANES")) {
    const int lv = atoi(e);
    if (lv >= 1 && lv <= 4) v->lanes = lv;
  }
  v->layers.resize(cfg->layers);
  *out = v;
  return 0;
}

extern "C" int b2c_vit_destroy(b2c_vit* v) {
  if (!v) return 0;
  for (b2c_vit::GraphEntry& g : v->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (v->cap_stream) cudaStreamDestroy(v->cap_stream);
  for (void* p : v->allocs) cudaFree(p);
  for (b2c_vit_layer& L : v->layers) {
    if (L.src_qkv) cudaFree(L.src_qkv);
    if (L.src_fc) cudaFree(L.src_fc);
  }
  for (int i = 0; i < 3; ++i) {
    if (v->side[i]) {
      cudaStreamSynchronize(v->side[i]);
      cudaStreamDestroy(v->side[i]);
    }
    if (v->ev_join[i]) cudaEventDestroy(v->ev_join[i]);
  }
  if (v->ev_fork) cudaEventDestroy(v->ev_fork);
  delete v;
  return 0;
}

extern "C" int b2c_vit_set_weight(b2c_vit* v, const char* key, const void* dev_ptr, int dtype, const int64_t* shape,
                                  int ndim) {
  B2C_REQUIRE(v && key && dev_ptr && shape, "b2c_vit_set_weight: null pointer");
  B2C_REQUIRE(dtype == B2C_F32 || dtype == B2C_F16 || dtype == B2C_BF16, "b2c_vit_set_weight(%s): dtype %d", key, dtype);
  int64_t count = 1;
  for (int i = 0; i < ndim; ++i) count *= shape[i];
  const b2c_vit_cfg& c = v->cfg;
  const int64_t d = c.width;
  const std::string k(key);
  int rc = B2C_ERR_ARG;
  const std::string pre = "transformer.resblocks.";
  if (k == "class_embedding") rc = store_f32(v, &v->cls, dev_ptr, dtype, count, d, key);
  else if (k == "positional_embedding") rc = store_f32(v, &v->pos, dev_ptr, dtype, count, v->T * d, key);
  else if (k == "proj") rc = store_f32(v, &v->proj, dev_ptr, dtype, count, d * c.embed, key);
  else if (k == "ln_pre.weight") rc = store_f32(v, &v->ln_pre_w, dev_ptr, dtype, count, d, key);
  else if (k == "ln_pre.bias") rc = store_f32(v, &v->ln_pre_b, dev_ptr, dtype, count, d, key);
  else if (k == "ln_post.weight") rc = store_f32(v, &v->ln_post_w, dev_ptr, dtype, count, d, key);
  else if (k == "ln_post.bias") rc = store_f32(v, &v->ln_post_b, dev_ptr, dtype, count, d, key);
  else if (k == "conv1.weight") {
    rc = store_bf16(v, &v->conv1, dev_ptr, dtype, d, 3 * c.patch * c.patch, v->Kp, count, key);
    if (rc == 0) rc = make_tmap_2d(&v->tm_conv1, v->conv1, d, v->Kp, static_cast<uint64_t>(v->Kp) * 2, kBN, 1);
  } else if (k.compare(0, pre.size(), pre) == 0) {
    const size_t dot = k.find('.', pre.size());
    B2C_REQUIRE(dot != std::string::npos, "b2c_vit_set_weight: malformed key %s", key);
    const int li = atoi(k.substr(pre.size(), dot - pre.size()).c_str());
    B2C_REQUIRE(li >= 0 && li < c.layers, "b2c_vit_set_weight: layer index out of range in %s", key);
    b2c_vit_layer& L = v->layers[li];
    const std::string s = k.substr(dot + 1);
    if (s.compare(0, 3, "ln_") == 0 || s == "attn.in_proj_weight" || s == "attn.in_proj_bias" || s == "mlp.c_fc.weight" ||
        s == "mlp.c_fc.bias")
      v->fold_dirty = true;
    if (s == "ln_1.weight") rc = store_f32(v, &L.ln1_w, dev_ptr, dtype, count, d, key);
    else if (s == "ln_1.bias") rc = store_f32(v, &L.ln1_b, dev_ptr, dtype, count, d, key);
    else if (s == "ln_2.weight") rc = store_f32(v, &L.ln2_w, dev_ptr, dtype, count, d, key);
    else if (s == "ln_2.bias") rc = store_f32(v, &L.ln2_b, dev_ptr, dtype, count, d, key);
    else if (s == "attn.in_proj_bias") rc = store_f32(v, &L.b_qkv, dev_ptr, dtype, count, 3 * d, key);
    else if (s == "attn.out_proj.bias") rc = store_f32(v, &L.b_out, dev_ptr, dtype, count, d, key);
    else if (s == "mlp.c_fc.bias") rc = store_f32(v, &L.b_fc, dev_ptr, dtype, count, c.mlp, key);
    else if (s == "mlp.c_proj.bias") rc = store_f32(v, &L.b_proj, dev_ptr, dtype, count, d, key);
    else if (s == "attn.in_proj_weight") {
      rc = store_bf16(v, &L.w_qkv, dev_ptr, dtype, 3 * d, d, d, count, key);
      if (rc == 0) rc = stage_f32(&L.src_qkv, dev_ptr, dtype, count);
      if (rc == 0) rc = make_tmap_2d(&L.tm_qkv_h, L.w_qkv, 3 * d, d, d * 2, kBM, 1);
    } else if (s == "attn.out_proj.weight") {
      rc = store_bf16(v, &L.w_out, dev_ptr, dtype, d, d, d, count, key);
      if (rc == 0) rc = make_tmap_2d(&L.tm_out_h, L.w_out, d, d, d * 2, kBM, 1);
    } else if (s == "mlp.c_fc.weight") {
      rc = store_bf16(v, &L.w_fc, dev_ptr, dtype, c.mlp, d, d, count, key);
      if (rc == 0) rc = stage_f32(&L.src_fc, dev_ptr, dtype, count);
      if (rc == 0) rc = make_tmap_2d(&L.tm_fc_h, L.w_fc, c.mlp, d, d * 2, kBM, 1);
    } else if (s == "mlp.c_proj.weight") {
      rc = store_bf16(v, &L.w_proj, dev_ptr, dtype, d, c.mlp, c.mlp, count, key);
      if (rc == 0) rc = make_tmap_2d(&L.tm_proj_h, L.w_proj, d, c.mlp, static_cast<uint64_t>(c.mlp) * 2, kBM, 1);
    } else {
      return set_error(B2C_ERR_ARG, "b2c_vit_set_weight: unknown key %s", key);
    }
  } else {
    return set_error(B2C_ERR_ARG, "b2c_vit_set_weight: unknown key %s", key);
  }
  if (rc != 0) return rc;
  B2C_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  v->have[k] = true;
  return 0;
}

extern "C" int b2c_vit_ready(const b2c_vit* v) {
  B2C_REQUIRE(v, "b2c_vit_ready: null handle");
  for (const std::string& k : expected_keys(v))
    if (!v->have.count(k)) return set_error(B2C_ERR_STATE, "vit: weight '%s' has not been set", k.c_str());
  return 0;
}

extern "C" int b2c_vit_set_lanes(b2c_vit* v, int lanes) {
  B2C_REQUIRE(v, "b2c_vit_set_lanes: null handle");
  B2C_REQUIRE(lanes >= 1 && lanes <= 4, "b2c_vit_set_lanes: lanes %d out of range (1..4)", lanes);
  v->lanes = lanes;
  return 0;
}

extern "C" int b2c_vit_set_cls_only_last_block(b2c_vit* v, int on) {
  B2C_REQUIRE(v, "b2c_vit_set_cls_only_last_block: null handle");
  v->cls_only_last = on != 0;
  return 0;
}

extern "C" int b2c_vit_set_graph(b2c_vit* v, int on) {
  B2C_REQUIRE(v, "b2c_vit_set_graph: null handle");
  v->use_graph = on != 0;
  if (!v->use_graph) {
    for (b2c_vit::GraphEntry& g : v->graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    v->graphs.clear();
  }
  return 0;
}

extern "C" int b2c_vit_set_fused_ln(b2c_vit* v, int on) {
  B2C_REQUIRE(v, "b2c_vit_set_fused_ln: null handle");
  v->fused_ln = on != 0;
  return 0;
}

extern "C" int b2c_vit_workspace_bytes(const b2c_vit* v, int n_crops, size_t* bytes) {
  B2C_REQUIRE(v && bytes, "b2c_vit_workspace_bytes: null pointer");
  B2C_REQUIRE(n_crops > 0, "b2c_vit_workspace_bytes: n_crops must be positive");
  *bytes = ws_bytes_for(v, n_crops < v->chunk ? n_crops : v->chunk, true, nullptr);
  return 0;
}

extern "C" int b2c_vit_forward_pixels(b2c_vit* v, const void* pixels, int dtype, int n_crops, float* out, void* ws,
                                      size_t ws_bytes, b2c_stream stream) {
  B2C_REQUIRE(pixels, "b2c_vit_forward_pixels: null pixels");
  B2C_REQUIRE(dtype == B2C_F32 || dtype == B2C_F16 || dtype == B2C_BF16, "b2c_vit_forward_pixels: dtype %d", dtype);
  return forward_impl(v, pixels, dtype, nullptr, n_crops, out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int b2c_vit_forward_patches(b2c_vit* v, const void* patches, int n_crops, float* out, void* ws,
                                       size_t ws_bytes, b2c_stream stream) {
  B2C_REQUIRE(patches, "b2c_vit_forward_patches: null patches");
  return forward_impl(v, nullptr, 0, patches, n_crops, out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
