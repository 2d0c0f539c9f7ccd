// b2c_launch.h — internal (non-ABI) launch functions shared between the translation units.
#pragma once
#include <atomic>

#include "b2c_host.h"

namespace b2c {

extern std::atomic<unsigned long long> g_launches;

#define B2C_POST_LAUNCH(name)                                                                     \
  do {                                                                                            \
    ::b2c::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return ::b2c::set_error(B2C_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// ---- stage timer (b2c_host.cu): records a CUDA-event pair around a stage when b2c_prof_enable(1) is active
extern std::atomic<int> g_prof_on;
void prof_begin(int kind, cudaStream_t stream);
void prof_end(cudaStream_t stream);
struct ProfScope {
  cudaStream_t s;
  bool on;
  ProfScope(int kind, cudaStream_t stream) : s(stream), on(g_prof_on.load(std::memory_order_relaxed) != 0) {
    if (on) prof_begin(kind, s);
  }
  ~ProfScope() {
    if (on) prof_end(s);
  }
};

// ---- tensor-core GEMM family (b2c_gemm.cu) -----------------------------------------------------
enum GemmMode {
  kGemmBiasBf16 = B2C_EPI_BIAS_BF16,
  kGemmBiasQGeluBf16 = B2C_EPI_BIAS_QGELU_BF16,
  kGemmBiasGeluBf16 = B2C_EPI_BIAS_GELU_BF16,
  kGemmBiasResidF32 = B2C_EPI_BIAS_RESID_F32,
  kGemmPatchEmbedF32 = 4,  // x[(crop*T + 1 + p), :] = A·Wᵀ + pos[1+p, :]   (conv1 + positional embedding)
  // LayerNorm fused into the GEMMs on either side of it (b2c_umma_pipeline2.cuh).  A = bf16 copy of the residual
  // stream, W = gamma-folded weight, bias = beta·Wᵀ + b, colsum = row sums of the folded weight:
  // A = bf16(x − shift) per row (see GemmLaunch::shift):
  kGemmLnBiasBf16 = 5,       // out bf16 = rstd·(A·W'ᵀ − (mean − shift)·colsum) + bias        (ln_1 + in_proj)
  kGemmLnBiasQGeluBf16 = 6,  // same + QuickGELU                                               (ln_2 + c_fc, openai)
  kGemmLnBiasGeluBf16 = 7,   // same + erf GELU                                                (ln_2 + c_fc, laion)
  // x (f32, read-modify-write by the epilogue) += A·Wᵀ + bias; shift = row mean before the update; out2 = bf16(x − shift);
  // stats[row][n_block] = (mean, M2) of the new values
  kGemmResidLnF32 = 8,
  kGemmResidLnBf16Copy = 9,  // make_out_tmap only: the store map of out2
  kGemmResidLnDeepF32 = 10,  // internal: kGemmResidLnF32 with six mainloop stages and a one-slab x ring (large K)
  kGemmResidLnWideF32 = 11,  // internal: kGemmResidLnF32 with eight epilogue warps (two per row: 128 columns each), x ring of two
};

struct GemmLaunch {
  CUtensorMap tmap_a;  // [M, K] bf16, box 128 x 64
  CUtensorMap tmap_b;  // [N, K] bf16, box 256 x 64 (patch-embed only: the single-CTA kernel loads the whole tile's N)
  CUtensorMap tmap_b_half;  // same tensor, box 128 x 64: each CTA of a pair loads half of the tile's N
  CUtensorMap tmap_out;  // epilogue store map over `out`: bf16 modes box 32 rows x 64 cols, f32 residual mode
                         // box 32 rows x 32 cols (make_out_tmap); unused by the patch-embed mode
  int64_t M;
  int N, K;
  int mode;
  const float* bias;  // [N] or nullptr
  void* out;          // bf16 or f32, row stride ldo elements
  int64_t ldo;
  const float* pos;   // patch-embed only: positional embedding [T, N]
  int T, G2;          // patch-embed only: tokens per crop, patches per crop
  // LayerNorm-fused modes only
  CUtensorMap tmap_out2;  // kGemmResidLnF32: store map of the bf16 copy (make_out_tmap(kGemmResidLnBf16Copy))
  float2* stats;          // [M, nblk] (mean, M2) per 256-column block of the residual stream (kGemmLn*: read;
                          // kGemmResidLnF32: written for the NEW rows)
  const float2* stats_in; // kGemmResidLnF32: the statistics of the rows BEFORE this update (a different buffer)
  float* shift;           // [M] what was subtracted from a row before its bf16 copy was rounded: the row mean before the
                          // update that wrote the copy (kGemmResidLnF32 writes it, kGemmLn* read it)
  int nblk;               // width / 256
  float eps;
  const float* colsum;    // kGemmLn*: [N]
};
int gemm_launch(const GemmLaunch& g, cudaStream_t stream);
// store map for a GEMM output [M, N] (row stride ldo elements) matching `mode`
int make_out_tmap(CUtensorMap* out, void* base, int64_t M, int N, int64_t ldo, int mode);

// ---- row-wise kernels (b2c_rowops.cu) ----------------------------------------------------------
int layernorm_bf16_launch(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d, float eps,
                          cudaStream_t stream);
// same, input rows ldx elements apart (e.g. only the class-token rows of the residual stream), output rows dense
int layernorm_bf16_strided_launch(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t M, int d,
                                  float eps, cudaStream_t stream);
// x[crop*T + 0, :] = cls + pos[0, :]; then ln_pre over all rows in place (f32 -> f32)
int cls_pos_launch(float* x, const float* cls, const float* pos, int n, int T, int d, cudaStream_t stream);
int layernorm_f32_inplace_launch(float* x, const float* gamma, const float* beta, int64_t M, int d, float eps,
                                 cudaStream_t stream);
// LayerNorm-fused layer loop: ln_pre in place + per-256-column (mean, M2) partials of the new rows + shift[row] = row
// mean + bf16 copy of (row - shift)
int layernorm_pre_launch(float* x, const float* gamma, const float* beta, void* xb, float2* stats, float* shift, int64_t M,
                         int d, float eps, cudaStream_t stream);
// wf = bf16(gamma ⊙ w) [N,K], colsum = row sums of wf, bias_f = bias + w·beta;  w: f32 or bf16 [N,K]
int ln_fold_launch(const void* w, int w_dtype, const float* gamma, const float* beta, const float* bias, void* wf,
                   float* colsum, float* bias_f, int N, int K, cudaStream_t stream);
// pixels [n,3,R,R] (f32/f16/bf16) -> patches bf16[n, g*g, Kp]
int patchify_launch(const void* pixels, int dtype, void* patches, int n, int R, int patch, int Kp, cudaStream_t stream);
// out f32[n,E] = l2norm( LN(x[crop*T + 0, :]) @ proj[d,E] ); part: head_part_floats(n, E) floats of scratch
int head_launch(const float* x, const float* gamma, const float* beta, const float* proj, float* out, float* part, int n,
                int T, int d, int E, float eps, cudaStream_t stream);
size_t head_part_floats(int n, int E);
// generic dtype conversion used by set_weight: dst bf16/f32 <- src (f32/f16/bf16)
int convert_launch(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t count, cudaStream_t stream);
// conv1.weight [d,3,p,p] -> bf16 [d, Kp] zero padded
int pad_rows_bf16_launch(const void* src, int src_dtype, void* dst, int64_t rows, int cols, int cols_padded,
                         cudaStream_t stream);

// ---- attention (b2c_attention.cu) --------------------------------------------------------------
int attention_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream);
// only the class-token query row of every (crop, head) -> out[crop*T, head*64 ..] (head dim 64)
int attention_cls_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream);

}  // namespace b2c
