// b2c_attention.cu — K4: multi-head self-attention of one ViT block, softmax(q·kᵀ / sqrt(hd)) · v,
// no mask, no dropout (nn.MultiheadAttention inside open_clip's ResidualAttentionBlock; the reference
// reaches it through utils/embedder.py:98).
//
// One CTA per (crop, head).  K and V of that head (T <= 592 tokens) are staged once in shared memory
// (cp.async, 16-byte chunks, rows padded by 16 B so ldmatrix is bank-conflict-free); each warp owns
// 16-query tiles and streams the keys in chunks of 64 with an online (running max / running sum)
// softmax in fp32 registers.  Tensor work is bf16 mma.sync m16n8k16 with fp32 accumulation; this is
// 4 % of the tower's FLOPs.
#include <cuda_bf16.h>

#include "b2c_launch.h"

namespace b2c {

constexpr int kAttnWarps = 6;
constexpr int kAttnThreads = kAttnWarps * 32;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ out, int T, int Tp,
                                                                 int heads, float scale_log2) {
  constexpr int LDS = HD + 8;        // padded row (elements)
  constexpr int KS = HD / 16;        // k-steps of q·kᵀ
  constexpr int NO = HD / 8;         // n8 blocks of the output
  constexpr int CPR = HD / 8;        // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Vs = Ks + static_cast<size_t>(Tp) * LDS;

  const int crop = blockIdx.x / heads;
  const int head = blockIdx.x - crop * heads;
  const int d = heads * HD;
  const size_t row_stride = static_cast<size_t>(3) * d;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(crop) * T * row_stride + head * HD;

  // ---- stage K and V
  for (int idx = threadIdx.x; idx < Tp * CPR; idx += kAttnThreads) {
    const int t = idx / CPR;
    const int c = idx - t * CPR;
    __nv_bfloat16* kd = Ks + t * LDS + c * 8;
    __nv_bfloat16* vd = Vs + t * LDS + c * 8;
    if (t < T) {
      const __nv_bfloat16* src = base + t * row_stride + c * 8;
      cp_async16(kd, src + d);
      cp_async16(vd, src + 2 * d);
    } else {
      *reinterpret_cast<uint4*>(kd) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vd) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t ks_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
  const uint32_t vs_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Vs));
  // ldmatrix lane -> (matrix, row) decomposition
  const int lm = lane >> 3, lr = lane & 7;

  const int n_mtiles = (T + 15) / 16;
  for (int mt = warp; mt < n_mtiles; mt += kAttnWarps) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // ---- Q fragments straight from global (each element is read exactly once)
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c0 = ks * 16 + 2 * tq;
      qf[ks][0] = r0 < T ? *reinterpret_cast<const uint32_t*>(base + r0 * row_stride + c0) : 0u;
      qf[ks][1] = r1 < T ? *reinterpret_cast<const uint32_t*>(base + r1 * row_stride + c0) : 0u;
      qf[ks][2] = r0 < T ? *reinterpret_cast<const uint32_t*>(base + r0 * row_stride + c0 + 8) : 0u;
      qf[ks][3] = r1 < T ? *reinterpret_cast<const uint32_t*>(base + r1 * row_stride + c0 + 8) : 0u;
    }
    float o[NO][4];
#pragma unroll
    for (int f = 0; f < NO; ++f) o[f][0] = o[f][1] = o[f][2] = o[f][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kc = 0; kc < Tp; kc += 64) {
      const int nb = (Tp - kc) >= 64 ? 8 : (Tp - kc) / 8;  // n8 key blocks in this chunk (warp-uniform, even)
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        if (j < nb) {
          const int key0 = kc + j * 8;
          // matrices: (keys key0..+8) x (feats 0-7 | 8-15 | 16-23 | 24-31) for k-steps 2i, 2i+1
#pragma unroll
          for (int kp = 0; kp + 1 < KS; kp += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(ks_addr + ((key0 + lr) * LDS + kp * 16 + lm * 8) * 2, b0, b1, b2, b3);
            mma_bf16_16816(s[j], qf[kp], b0, b1);
            mma_bf16_16816(s[j], qf[kp + 1], b2, b3);
          }
          if constexpr (KS & 1) {
            uint32_t b0, b1;
            ldsm_x2(ks_addr + ((key0 + lr) * LDS + (KS - 1) * 16 + (lm & 1) * 8) * 2, b0, b1);
            mma_bf16_16816(s[j], qf[KS - 1], b0, b1);
          }
        }
      }
      // ---- mask padded keys, running max
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          const int key = kc + j * 8 + 2 * tq;
          if (key >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (key + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
          cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1]));
          cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3]));
        }
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);  // finite: every chunk holds >= 1 real key
      const float a0 = exp2f((m0 - mn0) * scale_log2), a1 = exp2f((m1 - mn1) * scale_log2);
      m0 = mn0; m1 = mn1;
      const float ms0 = mn0 * scale_log2, ms1 = mn1 * scale_log2;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          s[j][0] = exp2f(fmaf(s[j][0], scale_log2, -ms0));
          s[j][1] = exp2f(fmaf(s[j][1], scale_log2, -ms0));
          s[j][2] = exp2f(fmaf(s[j][2], scale_log2, -ms1));
          s[j][3] = exp2f(fmaf(s[j][3], scale_log2, -ms1));
          ps0 += s[j][0] + s[j][1];
          ps1 += s[j][2] + s[j][3];
        }
      }
      l0 = l0 * a0 + ps0;
      l1 = l1 * a1 + ps1;
#pragma unroll
      for (int f = 0; f < NO; ++f) { o[f][0] *= a0; o[f][1] *= a0; o[f][2] *= a1; o[f][3] *= a1; }
      // ---- O += P · V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (2 * kk < nb) {
          uint32_t pa[4];
          pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          const int key0 = kc + kk * 16;
          // matrices (transposed on load): m0 = keys 0-7 x feats f, m1 = keys 8-15 x feats f,
          //                                m2 = keys 0-7 x feats f+8, m3 = keys 8-15 x feats f+8
#pragma unroll
          for (int f = 0; f < NO; f += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_trans(vs_addr + ((key0 + (lm & 1) * 8 + lr) * LDS + (f + (lm >> 1)) * 8) * 2, b0, b1, b2, b3);
            mma_bf16_16816(o[f], pa, b0, b1);
            mma_bf16_16816(o[f + 1], pa, b2, b3);
          }
        }
      }
    }
    // ---- finalise: divide by the row sums (quad-reduced) and store bf16
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __nv_bfloat16* ob = out + static_cast<size_t>(crop) * T * d + head * HD;
#pragma unroll
    for (int f = 0; f < NO; ++f) {
      const int c = f * 8 + 2 * tq;
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * d + c) = pack2(o[f][0] * i0, o[f][1] * i0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * d + c) = pack2(o[f][2] * i1, o[f][3] * i1);
    }
  }
}

template <int HD>
static int attention_launch_hd(const void* qkv, void* out, int n, int T, int heads, cudaStream_t stream) {
  const int Tp = (T + 15) / 16 * 16;
  const size_t smem = 2ull * Tp * (HD + 8) * sizeof(__nv_bfloat16);
  B2C_REQUIRE(smem <= 227 * 1024, "attention: T=%d does not fit shared memory", T);
  auto kern = attention_kernel<HD>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  kern<<<static_cast<unsigned>(n) * heads, kAttnThreads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), T, Tp, heads, scale_log2);
  B2C_POST_LAUNCH("attention_kernel");
  return 0;
}

int attention_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream) {
  B2C_REQUIRE(n > 0 && T > 0 && heads > 0, "attention: empty problem");
  if (hd == 64) return attention_launch_hd<64>(qkv, out, n, T, heads, stream);
  if (hd == 80) return attention_launch_hd<80>(qkv, out, n, T, heads, stream);
  return set_error(B2C_ERR_ARG, "attention: head dim %d unsupported (64 or 80)", hd);
}

}  // namespace b2c

extern "C" int b2c_attention_bf16(const void* qkv, void* out, int n, int T, int heads, int hd, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(qkv && out, "b2c_attention_bf16: null pointer");
  return attention_launch(qkv, out, n, T, heads, hd, static_cast<cudaStream_t>(stream));
}
