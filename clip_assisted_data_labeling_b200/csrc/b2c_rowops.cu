// b2c_rowops.cu — the HBM-bound row kernels of the tower (warp-shuffle reductions, 128-bit accesses):
//   K2  LayerNorm (ln_pre in place f32; ln_1/ln_2 f32 -> bf16), eps 1e-5, fp32 statistics
//   K1' class token + positional embedding row, pixel -> patch-major re-index (conv1 im2col, stride = kernel)
//   K8  CLS pool -> ln_post -> @ proj -> L2 normalise            (utils/embedder.py:98-99)
//   K10 SimpleFC forward                                          (utils/nn_model.py:23-41)
// plus dtype conversion used when weights are loaded.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "b2c_launch.h"

namespace b2c {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, size_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the whole row lives in registers (NV float4 per lane, d = 128*NV).
// Two-pass statistics (mean, then centred sum of squares) in fp32 like torch's LayerNorm kernel.
// OUT_BF16: y is a separate bf16 tensor; otherwise y aliases x (in-place f32).
// ------------------------------------------------------------------------------------------------
template <int NV, bool OUT_BF16>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, void* y,
                                                        long long M, int d, float eps, int nv_rt, long long ldx) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nv = NV > 0 ? NV : nv_rt;
  constexpr int CAP = NV > 0 ? NV : 16;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);  // input rows ldx elements apart, output rows dense
  float4 v[CAP];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      v[i] = xr[i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (i < nv) {
      const float4 g = __ldg(g4 + i * 32 + lane);
      const float4 b = __ldg(b4 + i * 32 + lane);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (OUT_BF16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&lo);
        w.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * d)[i * 32 + lane] = w;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * d)[i * 32 + lane] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ln_pre for the LayerNorm-fused layer loop: x = LN(x) in place (f32) and, for the first block's fused ln_1,
// the bf16 copy of the new row plus its (mean, M2) per 256-column block — what the kGemmResidLnF32 epilogue
// (b2c_umma_pipeline2.cuh) leaves behind after every later residual update.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_pre_kernel(float* x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ xb,
                                                            float2* __restrict__ stats, float* __restrict__ shift, long long M,
                                                            int d, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float4* xr = reinterpret_cast<float4*>(x + row * d);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + i * 32 + lane);
    const float4 b = __ldg(b4 + i * 32 + lane);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    v[i] = o;
    xr[i * 32 + lane] = o;
  }
  // float4 group i covers columns [128 i, 128 i + 128): block j of 256 columns = groups 2j, 2j + 1
  float msum = 0.f;
#pragma unroll
  for (int j = 0; j < NV / 2; ++j) {
    const float4 a = v[2 * j], b = v[2 * j + 1];
    const float bm = warp_sum(((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * (1.0f / 256.0f);
    float bq = 0.f;
    bq += (a.x - bm) * (a.x - bm) + (a.y - bm) * (a.y - bm) + (a.z - bm) * (a.z - bm) + (a.w - bm) * (a.w - bm);
    bq += (b.x - bm) * (b.x - bm) + (b.y - bm) * (b.y - bm) + (b.z - bm) * (b.z - bm) + (b.w - bm) * (b.w - bm);
    bq = warp_sum(bq);
    msum += bm;
    if (lane == 0) stats[row * (NV / 2) + j] = make_float2(bm, bq);
  }
  // bf16 copy of the row, centred on its mean (the consumer GEMM's epilogue knows the shift): rounding x - m instead of
  // x keeps the bf16 error relative to the row's spread, not to its offset
  const float sh = msum / static_cast<float>(NV / 2);
  if (lane == 0) shift[row] = sh;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x - sh, v[i].y - sh);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[i].z - sh, v[i].w - sh);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&lo);
    w.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(xb + row * d)[i * 32 + lane] = w;
  }
}

int layernorm_pre_launch(float* x, const float* gamma, const float* beta, void* xb, float2* stats, float* shift, int64_t M,
                         int d, float eps, cudaStream_t stream) {
  B2C_REQUIRE(d % 256 == 0 && d >= 256 && d <= 2048, "layernorm_pre: d=%d must be a multiple of 256 in [256,2048]", d);
  B2C_REQUIRE(M > 0, "layernorm_pre: M must be positive");
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(xb);
  switch (d / 128) {
    case 2: layernorm_pre_kernel<2><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 4: layernorm_pre_kernel<4><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 6: layernorm_pre_kernel<6><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 8: layernorm_pre_kernel<8><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 10: layernorm_pre_kernel<10><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 12: layernorm_pre_kernel<12><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    case 14: layernorm_pre_kernel<14><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
    default: layernorm_pre_kernel<16><<<grid, 256, 0, stream>>>(x, gamma, beta, o, stats, shift, M, d, eps); break;
  }
  B2C_POST_LAUNCH("layernorm_pre_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Weight folding for the fused LayerNorm:  LN(x)·Wᵀ + b = rstd·(x·W'ᵀ − mean·s) + b'   with
//   W'[n,k] = bf16(gamma_k · W[n,k]),  s_n = sum_k float(W'[n,k]) (of the ROUNDED weights the tensor core multiplies),
//   b'_n = b_n + sum_k beta_k · W[n,k].
// One block per output row n.  SRC = float (the staged original) or __nv_bfloat16 (re-fold from the stored copy).
// ------------------------------------------------------------------------------------------------
template <typename SRC>
__global__ void __launch_bounds__(256) ln_fold_kernel(const SRC* __restrict__ w, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                                      __nv_bfloat16* __restrict__ wf, float* __restrict__ colsum,
                                                      float* __restrict__ bias_f, int K) {
  const int n = blockIdx.x;
  const SRC* wr = w + static_cast<size_t>(n) * K;
  __nv_bfloat16* wo = wf + static_cast<size_t>(n) * K;
  float s = 0.f, bs = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = load_as_float<SRC>(wr, k);
    const __nv_bfloat16 r = __float2bfloat16_rn(gamma[k] * wv);
    wo[k] = r;
    s += __bfloat162float(r);
    bs = fmaf(beta[k], wv, bs);
  }
  __shared__ float red[2][8];
  s = warp_sum(s);
  bs = warp_sum(bs);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = bs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    colsum[n] = a;
    bias_f[n] = (bias ? bias[n] : 0.f) + b;
  }
}

int ln_fold_launch(const void* w, int w_dtype, const float* gamma, const float* beta, const float* bias, void* wf,
                   float* colsum, float* bias_f, int N, int K, cudaStream_t stream) {
  B2C_REQUIRE(w && gamma && beta && wf && colsum && bias_f && N > 0 && K > 0, "ln_fold: bad arguments");
  if (w_dtype == B2C_F32)
    ln_fold_kernel<float><<<N, 256, 0, stream>>>(static_cast<const float*>(w), gamma, beta, bias,
                                                 static_cast<__nv_bfloat16*>(wf), colsum, bias_f, K);
  else if (w_dtype == B2C_BF16)
    ln_fold_kernel<__nv_bfloat16><<<N, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(w), gamma, beta, bias,
                                                         static_cast<__nv_bfloat16*>(wf), colsum, bias_f, K);
  else
    return set_error(B2C_ERR_ARG, "ln_fold: source dtype %d", w_dtype);
  B2C_POST_LAUNCH("ln_fold_kernel");
  return 0;
}

template <bool OUT_BF16>
static int layernorm_dispatch(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d,
                              float eps, cudaStream_t stream, int64_t ldx = 0) {
  if (ldx == 0) ldx = d;
  B2C_REQUIRE(ldx % 4 == 0 && (OUT_BF16 || ldx == d), "layernorm: bad input row stride");
  B2C_REQUIRE(d % 128 == 0 && d >= 128 && d <= 2048, "layernorm: d=%d must be a multiple of 128 in [128,2048]", d);
  B2C_REQUIRE(M > 0, "layernorm: M must be positive");
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  const int nv = d / 128;
  switch (nv) {
    case 6: layernorm_kernel<6, OUT_BF16><<<grid, 256, 0, stream>>>(x, gamma, beta, y, M, d, eps, nv, ldx); break;
    case 8: layernorm_kernel<8, OUT_BF16><<<grid, 256, 0, stream>>>(x, gamma, beta, y, M, d, eps, nv, ldx); break;
    case 10: layernorm_kernel<10, OUT_BF16><<<grid, 256, 0, stream>>>(x, gamma, beta, y, M, d, eps, nv, ldx); break;
    default: layernorm_kernel<0, OUT_BF16><<<grid, 256, 0, stream>>>(x, gamma, beta, y, M, d, eps, nv, ldx); break;
  }
  B2C_POST_LAUNCH("layernorm_kernel");
  return 0;
}

int layernorm_bf16_launch(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d, float eps,
                          cudaStream_t stream) {
  return layernorm_dispatch<true>(x, gamma, beta, y, M, d, eps, stream);
}
int layernorm_bf16_strided_launch(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t M, int d,
                                  float eps, cudaStream_t stream) {
  return layernorm_dispatch<true>(x, gamma, beta, y, M, d, eps, stream, ldx);
}
int layernorm_f32_inplace_launch(float* x, const float* gamma, const float* beta, int64_t M, int d, float eps,
                                 cudaStream_t stream) {
  return layernorm_dispatch<false>(x, gamma, beta, x, M, d, eps, stream);
}

// ------------------------------------------------------------------------------------------------
// x[crop*T + 0, :] = class_embedding + positional_embedding[0, :]
// ------------------------------------------------------------------------------------------------
__global__ void cls_pos_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                               int T, int d) {
  float* row = x + static_cast<size_t>(blockIdx.x) * T * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) row[i] = cls[i] + pos[i];
}
int cls_pos_launch(float* x, const float* cls, const float* pos, int n, int T, int d, cudaStream_t stream) {
  cls_pos_kernel<<<n, 256, 0, stream>>>(x, cls, pos, T, d);
  B2C_POST_LAUNCH("cls_pos_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pixels [n,3,R,R] -> patches bf16 [n*g*g, Kp], k = c*p*p + py*p + px (the flattening of conv1.weight
// [d,3,p,p]); columns k >= 3*p*p are zero.  Only the encode_image(Tensor) surface uses this; the fused
// u8 path writes patch-major directly from the resize kernel.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ px, __nv_bfloat16* __restrict__ out, int R, int p, int g, int Kp) {
  const int prow = blockIdx.x;  // crop*g*g + gy*g + gx
  const int crop = prow / (g * g);
  const int pi = prow - crop * g * g;
  const int gy = pi / g, gx = pi - gy * g;
  const int pp = p * p;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float v = 0.f;
    if (k < 3 * pp) {
      const int c = k / pp;
      const int r = k - c * pp;
      const int py = r / p, pxx = r - py * p;
      v = load_as_float<T>(px, ((static_cast<size_t>(crop) * 3 + c) * R + (gy * p + py)) * R + gx * p + pxx);
    }
    out[static_cast<size_t>(prow) * Kp + k] = __float2bfloat16_rn(v);
  }
}
int patchify_launch(const void* pixels, int dtype, void* patches, int n, int R, int patch, int Kp,
                    cudaStream_t stream) {
  const int g = R / patch;
  const unsigned grid = static_cast<unsigned>(n) * g * g;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(patches);
  switch (dtype) {
    case B2C_F32: patchify_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(pixels), o, R, patch, g, Kp); break;
    case B2C_F16: patchify_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(pixels), o, R, patch, g, Kp); break;
    case B2C_BF16: patchify_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pixels), o, R, patch, g, Kp); break;
    default: return set_error(B2C_ERR_ARG, "patchify: unsupported pixel dtype %d", dtype);
  }
  B2C_POST_LAUNCH("patchify_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Head: for CPB crops per block: y = LN(x[crop*T]) ; e = y @ proj[d,E] ; out = e / ||e||_2.
// proj is read once per block (coalesced over E) and reused for the CPB crops.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadCPB = 8;       // crops per block
constexpr int kHeadThreads = 256;  // = embedding columns per block

// grid (crop groups of 8, column blocks of 256).  Warp w normalises the CLS row of crop w into shared memory, stored
// k-major ([k][8 crops]) so that one k step is two 16-byte broadcast loads; thread t then owns column cb*256 + t for all
// eight crops (proj is read once per crop group, coalesced).  The squared norm of each (crop, column block) goes to
// `part` and head_norm_kernel divides by the full-row norm, summed in a fixed order (deterministic).
__global__ void __launch_bounds__(kHeadThreads) head_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ proj, float* __restrict__ out,
                                                            float* __restrict__ part, int n, int T, int d, int E, float eps) {
  extern __shared__ __align__(16) float sm[];  // [d][kHeadCPB] normalised CLS rows, then [8 warps][kHeadCPB] partial norms
  float* ys = sm;
  float* red = sm + kHeadCPB * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crop0 = blockIdx.x * kHeadCPB;
  {
    const int crop = crop0 + warp;
    if (crop < n) {
      const float* xr = x + static_cast<size_t>(crop) * T * d;
      float s = 0.f;
      for (int i = lane; i < d; i += 32) s += xr[i];
      const float mean = warp_sum(s) / d;
      float q = 0.f;
      for (int i = lane; i < d; i += 32) { const float c = xr[i] - mean; q += c * c; }
      const float rstd = rsqrtf(warp_sum(q) / d + eps);
      for (int i = lane; i < d; i += 32) ys[i * kHeadCPB + warp] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
    } else {
      for (int i = lane; i < d; i += 32) ys[i * kHeadCPB + warp] = 0.f;
    }
  }
  __syncthreads();
  const int e = blockIdx.y * kHeadThreads + threadIdx.x;
  float acc[kHeadCPB];
#pragma unroll
  for (int c = 0; c < kHeadCPB; ++c) acc[c] = 0.f;
  if (e < E) {
    const float* pw = proj + e;
#pragma unroll 4
    for (int k = 0; k < d; ++k) {
      const float w = __ldg(pw + static_cast<size_t>(k) * E);
      const float4 ya = *reinterpret_cast<const float4*>(ys + k * kHeadCPB);
      const float4 yb = *reinterpret_cast<const float4*>(ys + k * kHeadCPB + 4);
      acc[0] = fmaf(ya.x, w, acc[0]); acc[1] = fmaf(ya.y, w, acc[1]); acc[2] = fmaf(ya.z, w, acc[2]); acc[3] = fmaf(ya.w, w, acc[3]);
      acc[4] = fmaf(yb.x, w, acc[4]); acc[5] = fmaf(yb.y, w, acc[5]); acc[6] = fmaf(yb.z, w, acc[6]); acc[7] = fmaf(yb.w, w, acc[7]);
    }
  }
#pragma unroll
  for (int c = 0; c < kHeadCPB; ++c) {
    const float s = warp_sum(acc[c] * acc[c]);
    if (lane == 0) red[warp * kHeadCPB + c] = s;
    if (e < E && crop0 + c < n) out[static_cast<size_t>(crop0 + c) * E + e] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < kHeadCPB && crop0 + threadIdx.x < n) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * kHeadCPB + threadIdx.x];
    part[static_cast<size_t>(crop0 + threadIdx.x) * gridDim.y + blockIdx.y] = s;
  }
}

// out[crop, :] /= sqrt(sum of the crop's column-block partials)        (utils/embedder.py:99, no eps)
__global__ void __launch_bounds__(256) head_norm_kernel(float* __restrict__ out, const float* __restrict__ part, int n,
                                                        int E, int nparts) {
  const int crop = blockIdx.x;
  float s = 0.f;
  for (int j = 0; j < nparts; ++j) s += part[static_cast<size_t>(crop) * nparts + j];
  const float inv = 1.0f / sqrtf(s);
  for (int e = threadIdx.x; e < E; e += blockDim.x) out[static_cast<size_t>(crop) * E + e] *= inv;
}

int head_launch(const float* x, const float* gamma, const float* beta, const float* proj, float* out, float* part, int n,
                int T, int d, int E, float eps, cudaStream_t stream) {
  B2C_REQUIRE(part, "head: null partial-norm buffer");
  const size_t smem = (static_cast<size_t>(kHeadCPB) * d + 8 * kHeadCPB) * sizeof(float);
  B2C_REQUIRE(smem <= 48 * 1024, "head: d=%d too wide", d);
  const dim3 grid((n + kHeadCPB - 1) / kHeadCPB, (E + kHeadThreads - 1) / kHeadThreads);
  head_kernel<<<grid, kHeadThreads, smem, stream>>>(x, gamma, beta, proj, out, part, n, T, d, E, eps);
  B2C_POST_LAUNCH("head_kernel");
  head_norm_kernel<<<n, 256, 0, stream>>>(out, part, n, E, static_cast<int>(grid.y));
  B2C_POST_LAUNCH("head_norm_kernel");
  return 0;
}
// floats of scratch head_launch needs for n crops and E columns
size_t head_part_floats(int n, int E) { return static_cast<size_t>(n) * ((E + kHeadThreads - 1) / kHeadThreads); }

// ------------------------------------------------------------------------------------------------
// dtype conversion / row padding for weights
// ------------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, long long rows, int cols, int cols_padded) {
  const long long total = rows * cols_padded;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols_padded;
    const int c = static_cast<int>(i - r * cols_padded);
    const float v = c < cols ? load_as_float<S>(src, static_cast<size_t>(r) * cols + c) : 0.f;
    if constexpr (sizeof(D) == 2) dst[i] = __float2bfloat16_rn(v);
    else dst[i] = v;
  }
}
template <typename D>
static int convert_rows(const void* src, int src_dtype, D* dst, int64_t rows, int cols, int cols_padded,
                        cudaStream_t stream) {
  const long long total = rows * static_cast<long long>(cols_padded);
  if (total == 0) return 0;
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 > 4096 ? 4096 : total / 256 + 1);
  switch (src_dtype) {
    case B2C_F32: convert_kernel<float, D><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), dst, rows, cols, cols_padded); break;
    case B2C_F16: convert_kernel<__half, D><<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), dst, rows, cols, cols_padded); break;
    case B2C_BF16: convert_kernel<__nv_bfloat16, D><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), dst, rows, cols, cols_padded); break;
    default: return set_error(B2C_ERR_ARG, "convert: unsupported source dtype %d", src_dtype);
  }
  B2C_POST_LAUNCH("convert_kernel");
  return 0;
}
int convert_launch(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t count, cudaStream_t stream) {
  if (dst_dtype == B2C_BF16) return convert_rows<__nv_bfloat16>(src, src_dtype, static_cast<__nv_bfloat16*>(dst), 1, (int)count, (int)count, stream);
  if (dst_dtype == B2C_F32) return convert_rows<float>(src, src_dtype, static_cast<float*>(dst), 1, (int)count, (int)count, stream);
  return set_error(B2C_ERR_ARG, "convert: unsupported destination dtype %d", dst_dtype);
}
int pad_rows_bf16_launch(const void* src, int src_dtype, void* dst, int64_t rows, int cols, int cols_padded,
                         cudaStream_t stream) {
  return convert_rows<__nv_bfloat16>(src, src_dtype, static_cast<__nv_bfloat16*>(dst), rows, cols, cols_padded, stream);
}

// ------------------------------------------------------------------------------------------------
// K10: SimpleFC forward.  IPB images per block share every weight read; activations ping-pong in smem.
// ------------------------------------------------------------------------------------------------
constexpr int kMlpIPB = 4;
constexpr int kMlpThreads = 256;

__global__ void __launch_bounds__(kMlpThreads) mlp_kernel(const float* __restrict__ feats, long long B,
                                                          const b2c_mlp_weights w, float* __restrict__ out, int max_dim) {
  extern __shared__ float sm[];  // 2 x [kMlpIPB][max_dim]
  float* cur = sm;
  float* nxt = sm + kMlpIPB * max_dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long img0 = static_cast<long long>(blockIdx.x) * kMlpIPB;
  const int d0 = w.dims[0];
  for (int c = 0; c < kMlpIPB; ++c) {
    const long long img = img0 + c;
    for (int i = threadIdx.x; i < d0; i += kMlpThreads) cur[c * max_dim + i] = img < B ? feats[img * d0 + i] : 0.f;
  }
  __syncthreads();
  for (int l = 0; l < w.n_layers; ++l) {
    const int din = w.dims[l], dout = w.dims[l + 1];
    const float* W = w.weight[l];
    const float* bvec = w.bias[l];
    const bool last = (l == w.n_layers - 1);
    for (int j = warp; j < dout; j += kMlpThreads / 32) {
      float acc[kMlpIPB];
#pragma unroll
      for (int c = 0; c < kMlpIPB; ++c) acc[c] = 0.f;
      const float* wr = W + static_cast<size_t>(j) * din;
      for (int k = lane; k < din; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int c = 0; c < kMlpIPB; ++c) acc[c] = fmaf(wv, cur[c * max_dim + k], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < kMlpIPB; ++c) acc[c] = warp_sum(acc[c]);
      if (lane == 0) {
        const float b = bvec ? bvec[j] : 0.f;
#pragma unroll
        for (int c = 0; c < kMlpIPB; ++c) {
          float v = acc[c] + b;
          if (last) {
            v = 1.0f / (1.0f + expf(-v));
            const long long img = img0 + c;
            if (img < B) out[img * dout + j] = v;
          } else {
            v = v > 0.f ? v : v * w.leaky_slope;
            nxt[c * max_dim + j] = v;
          }
        }
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace b2c

extern "C" int b2c_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int64_t M, int d,
                                  float eps, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(x && gamma && beta && y, "b2c_layernorm_bf16: null pointer");
  return layernorm_bf16_launch(x, gamma, beta, y, M, d, eps, static_cast<cudaStream_t>(stream));
}

extern "C" int b2c_mlp_score(const float* feats, int64_t B, const b2c_mlp_weights* w, float* out, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(feats && w && out, "b2c_mlp_score: null pointer");
  B2C_REQUIRE(w->n_layers >= 1 && w->n_layers <= B2C_MLP_MAX_LAYERS, "b2c_mlp_score: n_layers=%d out of range", w->n_layers);
  if (B <= 0) return 0;
  int max_dim = 0;
  for (int l = 0; l <= w->n_layers; ++l) {
    B2C_REQUIRE(w->dims[l] > 0, "b2c_mlp_score: dims[%d] must be positive", l);
    if (w->dims[l] > max_dim) max_dim = w->dims[l];
  }
  for (int l = 0; l < w->n_layers; ++l) B2C_REQUIRE(w->weight[l], "b2c_mlp_score: weight[%d] is null", l);
  const size_t smem = 2ull * kMlpIPB * max_dim * sizeof(float);
  B2C_REQUIRE(smem <= 200 * 1024, "b2c_mlp_score: layer width %d too large", max_dim);
  static PerDeviceFlag attr_once;
  if (attr_once.first_use()) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const unsigned grid = static_cast<unsigned>((B + kMlpIPB - 1) / kMlpIPB);
  mlp_kernel<<<grid, kMlpThreads, smem, static_cast<cudaStream_t>(stream)>>>(feats, B, *w, out, max_dim);
  B2C_POST_LAUNCH("mlp_kernel");
  return 0;
}
