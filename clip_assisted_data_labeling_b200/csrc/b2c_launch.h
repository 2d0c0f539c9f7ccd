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
};

struct GemmLaunch {
  CUtensorMap tmap_a;  // [M, K] bf16, box 128 x 64
  CUtensorMap tmap_b;  // [N, K] bf16, box 256 x 64
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
};
int gemm_launch(const GemmLaunch& g, cudaStream_t stream);
// store map for a GEMM output [M, N] (row stride ldo elements) matching `mode`
int make_out_tmap(CUtensorMap* out, void* base, int64_t M, int N, int64_t ldo, int mode);

// ---- row-wise kernels (b2c_rowops.cu) ----------------------------------------------------------
int layernorm_bf16_launch(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d, float eps,
                          cudaStream_t stream);
// x[crop*T + 0, :] = cls + pos[0, :]; then ln_pre over all rows in place (f32 -> f32)
int cls_pos_launch(float* x, const float* cls, const float* pos, int n, int T, int d, cudaStream_t stream);
int layernorm_f32_inplace_launch(float* x, const float* gamma, const float* beta, int64_t M, int d, float eps,
                                 cudaStream_t stream);
// pixels [n,3,R,R] (f32/f16/bf16) -> patches bf16[n, g*g, Kp]
int patchify_launch(const void* pixels, int dtype, void* patches, int n, int R, int patch, int Kp, cudaStream_t stream);
// out f32[n,E] = l2norm( LN(x[crop*T + 0, :]) @ proj[d,E] )
int head_launch(const float* x, const float* gamma, const float* beta, const float* proj, float* out, int n, int T,
                int d, int E, float eps, cudaStream_t stream);
// generic dtype conversion used by set_weight: dst bf16/f32 <- src (f32/f16/bf16)
int convert_launch(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t count, cudaStream_t stream);
// conv1.weight [d,3,p,p] -> bf16 [d, Kp] zero padded
int pad_rows_bf16_launch(const void* src, int src_dtype, void* dst, int64_t rows, int cols, int cols_padded,
                         cudaStream_t stream);

// ---- attention (b2c_attention.cu) --------------------------------------------------------------
int attention_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream);

}  // namespace b2c
