// b2c_gemm.cu — the ViT tower's dense contractions on tcgen05 tensor cores with fused epilogues.
//   K1 patch-embed (conv1, stride = kernel, no bias) + positional embedding      -> f32 residual stream
//   K3 attn.in_proj  (+bias)                                                     -> bf16 qkv
//   K5 attn.out_proj (+bias, += residual)                                        -> f32 residual stream
//   K6 mlp.c_fc      (+bias, QuickGELU | erf-GELU)                               -> bf16 hidden
//   K7 mlp.c_proj    (+bias, += residual)                                        -> f32 residual stream
// Arithmetic restated from open_clip's VisionTransformer (third-party), which the
// reference calls at utils/embedder.py:98.  bf16 operands, fp32 accumulation in TMEM.
#include <cuda_bf16.h>

#include "b2c_launch.h"
#include <stdlib.h>

#include "b2c_umma_pipeline.cuh"
#include "b2c_umma_pipeline2.cuh"

namespace b2c {

struct GemmParams {
  int num_tiles;   // 128 x 256 tiles (single-CTA kernel)
  int num_tiles2;  // 256 x 256 tiles (CTA-pair kernel)
  int k_blocks;
  int n_blocks;
  int M, N;
  const float* bias;
  void* out;
  long long ldo;
  const float* pos;
  int T, G2;
  // LayerNorm fusion (b2c_umma_pipeline2.cuh): per-row (mean, M2) partials of the residual stream, one per 256-column
  // block; written by the kGemmResidLnF32 epilogue, merged by the kGemmLn* epilogues
  float2* stats;
  const float2* stats_in;  // kStoreRmwLn: statistics of the rows before this update
  float* shift;            // per-row offset of the bf16 copy (see GemmLaunch)
  int nblk;
  float eps;
  const float* colsum;  // kGemmLn*: column sums of the folded weight, s_n = sum_k (gamma_k W_nk)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// QuickGELU x*sigmoid(1.702x) with sigmoid(z) = 0.5 + 0.5*tanh(z/2): one MUFU op (tanh.approx, rel. error ~2^-11,
// far below the bf16 rounding of the output) instead of exp + reciprocal — the c_fc epilogue is MUFU-bound otherwise.
__device__ __forceinline__ float quick_gelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int MODE>
struct GemmPolicy {
  using Params = GemmParams;

  // n-block fastest: the CTAs running concurrently share A tiles (L2 hits) and the weight matrix
  // (<= 10 MB) stays L2-resident.
  __device__ static __forceinline__ bool tile(const Params& p, int t, int& a_row, int& b_row) {
    const int mb = t / p.n_blocks;
    const int nb = t - mb * p.n_blocks;
    a_row = mb * kBM;
    b_row = nb * kBN;
    return true;
  }

  // CTA-pair kernel: 256 x 256 tiles, n-block fastest
  __device__ static __forceinline__ bool tile2(const Params& p, int t, int& m_row, int& n_row) {
    const int mb2 = t / p.n_blocks;
    const int nb = t - mb2 * p.n_blocks;
    m_row = mb2 * 2 * kBM;
    n_row = nb * kBN;
    return true;
  }

  static constexpr int kStore = (MODE == kGemmPatchEmbedF32) ? kStoreDirect
                                : (MODE == kGemmBiasResidF32) ? kStoreTmaAddF32
                                : (MODE == kGemmResidLnF32 || MODE == kGemmResidLnDeepF32 || MODE == kGemmResidLnWideF32) ? kStoreRmwLn
                                                                                                                        : kStoreTmaBf16;
  static constexpr int kRmwRing = MODE == kGemmResidLnDeepF32 ? 1 : MODE == kGemmResidLnWideF32 ? 2 : 5;  // see b2c_umma_pipeline2.cuh
  // CTA-pair kernel: two epilogue warps per TMEM lane quarter, except for the read-modify-write epilogue (its slab
  // ring needs the shared memory a second set of warps would stage through)
  static constexpr int kEpiWarps = (MODE == kGemmResidLnF32 || MODE == kGemmResidLnDeepF32) ? 4 : 8;
  static constexpr bool kLnFold = MODE == kGemmLnBiasBf16 || MODE == kGemmLnBiasQGeluBf16 || MODE == kGemmLnBiasGeluBf16;
  static constexpr bool kQGelu = MODE == kGemmBiasQGeluBf16 || MODE == kGemmLnBiasQGeluBf16;
  static constexpr bool kGelu = MODE == kGemmBiasGeluBf16 || MODE == kGemmLnBiasGeluBf16;

  // bias (+ activation) on one 32-column chunk of an accumulator row; the kernel then stages and TMA-stores it
  __device__ static __forceinline__ void transform(const Params& p, int col, float (&f)[32]) {
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
      }
    }
    activate(f);
  }

  __device__ static __forceinline__ void activate(float (&f)[32]) {
    if constexpr (kQGelu) {
      // x·sigmoid(1.702x) = hx + hx·tanh(0.851x), two elements per FMUL2 / FFMA2
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float z0, z1, h0, h1, t0, t1;
        fmul2(z0, z1, f[e], f[e + 1], 0.851f, 0.851f);
        fmul2(h0, h1, f[e], f[e + 1], 0.5f, 0.5f);
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
        ffma2(f[e], f[e + 1], h0, h1, t0, t1, h0, h1);
      }
    } else if constexpr (kGelu) {
#pragma unroll
      for (int e = 0; e < 32; ++e) f[e] = gelu_erf(f[e]);
    }
  }

  // CTA-pair kernel: lane l keeps column col0 + 32 c + l of the tile's bias (and, kLnFold, column-sum) vector
  template <int NC>
  __device__ static __forceinline__ void load_cols(const Params& p, int col, float (&b)[NC], float (&s)[NC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      b[c] = p.bias ? __ldg(p.bias + col + 32 * c) : 0.f;
      if constexpr (kLnFold) s[c] = __ldg(p.colsum + col + 32 * c);
    }
  }

  // kLnFold: the row's per-block (mean, M2) partials and the shift its bf16 copy was rounded with (requested a tile
  // ahead), merged in block order -> a = rstd, b = -(mean - shift) * rstd
  __device__ static __forceinline__ void load_row_stats(const Params& p, int row, float2 (&st)[8], float& sh) {
    const float2* src = p.stats + static_cast<size_t>(row) * p.nblk;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < p.nblk && row < p.M) st[j] = src[j];
    if (row < p.M) sh = p.shift[row];
  }
  // kStoreRmwLn: the row's mean before this update (from the previous update's partials) = the shift of its new bf16 copy
  __device__ static __forceinline__ float load_old_mean(const Params& p, int row) {
    if (row >= p.M) return 0.f;
    const float2* src = p.stats_in + static_cast<size_t>(row) * p.nblk;
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < p.nblk) m += src[j].x;
    return m / static_cast<float>(p.nblk);
  }
  __device__ static __forceinline__ void merge_row_stats(const Params& p, int row, const float2 (&st)[8], float sh, float& a,
                                                         float& b) {
    a = 0.f;
    b = 0.f;
    if (row >= p.M) return;
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < p.nblk) mean += st[j].x;
    mean /= static_cast<float>(p.nblk);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < p.nblk) {
        const float dlt = st[j].x - mean;
        m2 += fmaf(256.0f * dlt, dlt, st[j].y);
      }
    }
    const float rstd = rsqrtf(m2 / (256.0f * static_cast<float>(p.nblk)) + p.eps);
    a = rstd;
    b = -(mean - sh) * rstd;
  }

  // kStoreRmwLn: (mean, M2) of the row's 256 new values in this tile's column block
  __device__ static __forceinline__ void store_row_stats(const Params& p, int row, int b_row, float mean, float m2, float sh) {
    if (row < p.M) {
      p.stats[static_cast<size_t>(row) * p.nblk + (b_row >> 8)] = make_float2(mean, m2);
      if (b_row == 0) p.shift[row] = sh;  // every column block of the row used the same value
    }
  }

  // patch-embed only: rows are scattered past each crop's class token, so the stores stay per-thread
  __device__ static __forceinline__ void epilogue(const Params& p, int a_row, int b_row, int row_in_tile, int col0,
                                                  const uint32_t (&acc)[32]) {
    const int row = a_row + row_in_tile;
    if (row >= p.M) return;
    const int col = b_row + col0;
    const int crop = row / p.G2;
    const int pidx = row - crop * p.G2;
    const size_t orow = static_cast<size_t>(crop) * p.T + 1 + pidx;
    const float* addend = p.pos + static_cast<size_t>(1 + pidx) * p.N + col;
    float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(addend + j));
      float4 v;
      v.x = __uint_as_float(acc[j + 0]) + b.x;
      v.y = __uint_as_float(acc[j + 1]) + b.y;
      v.z = __uint_as_float(acc[j + 2]) + b.z;
      v.w = __uint_as_float(acc[j + 3]) + b.w;
      *reinterpret_cast<float4*>(o + j) = v;
    }
  }
};

template <int MODE>
static int gemm_launch_pair(const GemmLaunch& g, const GemmParams& p, cudaStream_t stream) {
  auto kern = umma2_tile_kernel<GemmPolicy<MODE>>;
  static PerDeviceFlag attr_once;
  if (attr_once.first_use()) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kUmma2SmemBytes));
  }
  const int sms = num_sms();
  B2C_REQUIRE(sms >= 2, "no CUDA device");
  int grid = 2 * p.num_tiles2 < sms ? 2 * p.num_tiles2 : (sms & ~1);
  kern<<<grid, 64 + 32 * GemmPolicy<MODE>::kEpiWarps, kUmma2SmemBytes, stream>>>(g.tmap_a, g.tmap_b_half, g.tmap_out, g.tmap_out2, p,
                                                        make_idesc_f16(2 * kBM, kBN, 1));
  B2C_POST_LAUNCH("umma2_tile_kernel<gemm>");
  return 0;
}

// patch-embed (scattered per-thread stores past each crop's class token) runs on the single-CTA kernel; every other mode
// on CTA pairs
static int gemm_launch_patch_embed(const GemmLaunch& g, const GemmParams& p, cudaStream_t stream) {
  auto kern = umma_tile_kernel<GemmPolicy<kGemmPatchEmbedF32>>;
  static PerDeviceFlag attr_once;
  if (attr_once.first_use()) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kUmmaSmemBytes));
  }
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  kern<<<grid, kUmmaThreads, kUmmaSmemBytes, stream>>>(g.tmap_a, g.tmap_b, g.tmap_out, p, make_idesc_f16(kBM, kBN, 1));
  B2C_POST_LAUNCH("umma_tile_kernel<patch-embed>");
  return 0;
}

template <int MODE>
static int gemm_launch_mode(const GemmLaunch& g, const GemmParams& p, cudaStream_t stream) {
  if constexpr (MODE >= kGemmLnBiasBf16) {
    B2C_REQUIRE(p.stats && p.shift && p.nblk > 0 && p.nblk <= 8, "gemm: LayerNorm-fused mode %d needs the row statistics / shift buffers", MODE);
    if constexpr (GemmPolicy<MODE>::kStore == kStoreRmwLn)
      B2C_REQUIRE(p.stats_in && p.stats_in != p.stats, "gemm: mode %d needs the previous update's statistics in a separate buffer", MODE);
    if constexpr (GemmPolicy<MODE>::kLnFold) B2C_REQUIRE(p.colsum && p.bias, "gemm: mode %d needs colsum and the folded bias", MODE);
  }
  return gemm_launch_pair<MODE>(g, p, stream);
}

int make_out_tmap(CUtensorMap* out, void* base, int64_t M, int N, int64_t ldo, int mode) {
  if (mode == kGemmResidLnBf16Copy)  // the bf16 copy of the residual stream: 32 x 32 half slabs, 64-B swizzle
    return make_tmap_2d_sw(out, base, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldo) * 2, 32, 32,
                           B2C_BF16, 64);
  if (mode == kGemmBiasResidF32 || mode == kGemmResidLnF32)
    return make_tmap_2d_ex(out, base, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldo) * 4, 32, 32,
                           B2C_F32);
  // bf16 tiles are stored per 32-column chunk: 32 x 32 half slabs, 64-B swizzle
  return make_tmap_2d_sw(out, base, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldo) * 2, 32, 32,
                         B2C_BF16, 64);
}

int gemm_launch(const GemmLaunch& g, cudaStream_t stream) {
  B2C_REQUIRE(g.M > 0 && g.M < (1ll << 31) - kBM, "gemm: M=%lld out of range", (long long)g.M);
  B2C_REQUIRE(g.N > 0 && g.N % kBN == 0, "gemm: N=%d must be a positive multiple of %d", g.N, kBN);
  B2C_REQUIRE(g.K > 0 && g.K % kBK == 0, "gemm: K=%d must be a positive multiple of %d", g.K, kBK);
  GemmParams p;
  p.k_blocks = g.K / kBK;
  p.n_blocks = g.N / kBN;
  const long long m_blocks = (g.M + kBM - 1) / kBM;
  B2C_REQUIRE(m_blocks * p.n_blocks < (1ll << 31), "gemm: too many tiles");
  p.num_tiles = static_cast<int>(m_blocks * p.n_blocks);
  p.num_tiles2 = static_cast<int>(((m_blocks + 1) / 2) * p.n_blocks);
  p.M = static_cast<int>(g.M);
  p.N = g.N;
  p.bias = g.bias;
  p.out = g.out;
  p.ldo = g.ldo;
  p.pos = g.pos;
  p.T = g.T;
  p.G2 = g.G2;
  p.stats = g.stats;
  p.stats_in = g.stats_in;
  p.shift = g.shift;
  p.nblk = g.nblk;
  p.eps = g.eps;
  p.colsum = g.colsum;
  switch (g.mode) {
    case kGemmBiasBf16: return gemm_launch_mode<kGemmBiasBf16>(g, p, stream);
    case kGemmBiasQGeluBf16: return gemm_launch_mode<kGemmBiasQGeluBf16>(g, p, stream);
    case kGemmBiasGeluBf16: return gemm_launch_mode<kGemmBiasGeluBf16>(g, p, stream);
    case kGemmBiasResidF32: return gemm_launch_mode<kGemmBiasResidF32>(g, p, stream);
    case kGemmPatchEmbedF32:
      B2C_REQUIRE(g.pos && g.T > 0 && g.G2 > 0, "gemm: patch-embed mode needs pos/T/G2");
      return gemm_launch_patch_embed(g, p, stream);
    case kGemmLnBiasBf16: return gemm_launch_mode<kGemmLnBiasBf16>(g, p, stream);
    case kGemmLnBiasQGeluBf16: return gemm_launch_mode<kGemmLnBiasQGeluBf16>(g, p, stream);
    case kGemmLnBiasGeluBf16: return gemm_launch_mode<kGemmLnBiasGeluBf16>(g, p, stream);
    case kGemmResidLnF32:
      B2C_REQUIRE(g.N == g.nblk * kBN, "gemm: the residual+statistics mode needs N == 256 * nblk (whole rows)");
      // K > 2048 (c_proj): keep all six mainloop stages, the epilogue has the slack to wait for each x slab
      if (g.K > 2048) return gemm_launch_mode<kGemmResidLnDeepF32>(g, p, stream);
      {
        // K <= 2048 (out_proj): the epilogue is the long pole — two warps per row (B2C_RESID_WIDE=0: one, the round-1 form)
        static const bool wide = [] { const char* e = getenv("B2C_RESID_WIDE"); return !e || atoi(e) != 0; }();
        return wide ? gemm_launch_mode<kGemmResidLnWideF32>(g, p, stream) : gemm_launch_mode<kGemmResidLnF32>(g, p, stream);
      }
    default: return set_error(B2C_ERR_ARG, "gemm: unknown epilogue mode %d", g.mode);
  }
}

}  // namespace b2c

extern "C" int b2c_gemm_bf16(const void* A, const void* W, const float* bias, void* out, int64_t M, int N, int K,
                             int epilogue, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(A && W && out, "b2c_gemm_bf16: null pointer");
  B2C_REQUIRE(epilogue >= B2C_EPI_BIAS_BF16 && epilogue <= B2C_EPI_BIAS_RESID_F32, "b2c_gemm_bf16: bad epilogue %d",
              epilogue);
  B2C_REQUIRE(M > 0 && N > 0 && K > 0, "b2c_gemm_bf16: empty problem");
  GemmLaunch g{};
  B2C_TRY(make_tmap_2d(&g.tmap_a, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K),
                       static_cast<uint64_t>(K) * 2, kBM, 1));
  B2C_TRY(make_tmap_2d(&g.tmap_b, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                       static_cast<uint64_t>(K) * 2, kBN, 1));
  B2C_TRY(make_tmap_2d(&g.tmap_b_half, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                       static_cast<uint64_t>(K) * 2, kBM, 1));
  g.M = M;
  g.N = N;
  g.K = K;
  g.mode = epilogue;
  g.bias = bias;
  g.out = out;
  g.ldo = N;
  B2C_TRY(make_out_tmap(&g.tmap_out, out, M, N, N, epilogue));
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}
