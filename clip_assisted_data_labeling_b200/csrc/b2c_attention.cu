// b2c_attention.cu — K4: multi-head self-attention of one ViT block, softmax(q·kᵀ / sqrt(hd)) · v,
// no mask, no dropout (nn.MultiheadAttention inside open_clip's ResidualAttentionBlock; the reference
// reaches it through utils/embedder.py:98).
//
// Kernels, by shape (attention_launch picks):
//   T = 257, head dim 64  (ViT-L/14-224)  attention_umma5_kernel  tcgen05, S / P / O in tensor memory, 16 softmax warps,
//                                          class token on 4 dedicated warps (section 1e)
//   T = 257, head dim 80  (ViT-H/14)      attention_umma2_kernel<80>  tcgen05, 8 softmax warps, head dim as a 64- plus a
//                                          16-column slab (section 1b)
//   T = 577, head dim 64  (ViT-L/14-336)  attention_umma6_kernel  tcgen05, K / V of the head resident in shared memory,
//                                          key blocks of 96 visited twice, 16 softmax + 4 class-token warps (section 1c)
//   anything else (ViT-B/32: T = 50)      attention_kernel<HD>  bf16 mma.sync, online softmax (section 2)
//   class-token query row only            attention_cls_kernel  (b2c_vit_set_cls_only_last_block)
// In every tcgen05 kernel the class-token KEY is a rank-1 term (s0 = q·k0 by FMA, p0·v0 added in the epilogue) so that
// the tensor-core problem is exactly the patch keys with no padding, P is written back as packed bf16 over the consumed
// S columns, O = P·V runs with P read from TMEM and V as an MN-major operand straight from the TMA tile, and the row
// sums of the ROUNDED probabilities come from the tensor core (L = P·1).
//
// (2): One CTA per (crop, head).  K and V of that head (T <= 592 tokens) are staged once in shared memory
// (cp.async, 16-byte chunks, rows padded by 16 B so ldmatrix is bank-conflict-free); each warp owns
// 16-query tiles and streams the keys in chunks of 64 with an online (running max / running sum)
// softmax in fp32 registers.  Tensor work is bf16 mma.sync m16n8k16 with fp32 accumulation.
#include <type_traits>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "b2c_launch.h"
#include "b2c_umma_pipeline.cuh"

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


constexpr int kAuKeys = 256;  // patch tokens of the 224-pixel towers = keys handled by one N=256 tensor-core tile

// The class-token query row of every (crop, head): one warp each, SIMT.  1/257 of the attention work.
__global__ void __launch_bounds__(128) attention_cls_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int n_ch, int T, int heads,
                                                            float scale_log2) {
  const int ch = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (ch >= n_ch) return;
  const int lane = threadIdx.x & 31;
  const int crop = ch / heads, head = ch - crop * heads;
  const int d = heads * 64;
  const size_t row_stride = static_cast<size_t>(3) * d;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(crop) * T * row_stride + head * 64;
  // q0 in registers (every lane holds all 64 values)
  float q[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 a = __ldg(qp + j);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(a2[e]);
        q[8 * j + 2 * e] = f.x;
        q[8 * j + 2 * e + 1] = f.y;
      }
    }
  }
  // scores: lane owns keys lane, lane+32, ...
  constexpr int SLOTS = 19;  // ceil(592 / 32)
  float s[SLOTS];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    const int key = i * 32 + lane;
    s[i] = -INFINITY;
    if (key < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + key * row_stride + d);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 a = __ldg(kp + j);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(a2[e]);
          acc = fmaf(q[8 * j + 2 * e], f.x, acc);
          acc = fmaf(q[8 * j + 2 * e + 1], f.y, acc);
        }
      }
      s[i] = acc;
      m = fmaxf(m, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    s[i] = exp2f((s[i] - m) * scale_log2);  // -inf slots become 0
    l += s[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  // output: lane owns columns 2*lane, 2*lane+1; p broadcast from its owner lane
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    if (i * 32 >= T) break;
    for (int src = 0; src < 32; ++src) {
      const int key = i * 32 + src;
      const float p = __shfl_sync(0xffffffffu, s[i], src);
      if (key < T) {
        const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(base + key * row_stride + 2 * d + 2 * lane);
        const float2 f = __bfloat1622float2(vv);
        o0 = fmaf(p, f.x, o0);
        o1 = fmaf(p, f.y, o1);
      }
    }
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(crop) * T * d + head * 64 + 2 * lane) = pack2(o0 * inv, o1 * inv);
}

// ================================================================================================
// (1b) tcgen05 attention v2 for T = 257, head dim 64 (ViT-L/14) or 80 (ViT-H/14): persistent, pipelined, P kept in
//      tensor memory.
//   CTA (one per SM, 320 threads) loops over (crop, head) pairs:
//     warp 0      TMA producer: Q (both 128-row tiles), K, V of the NEXT head stream into the other smem stage
//     warp 1      MMA issuer:   S_w = Q_w·Kᵀ (SS) into TMEM region w;  O_w = P_w·V (A = P from TMEM, V MN-major);
//                               L_w = P_w·1 (row sums of the ROUNDED probabilities, fp32, from the tensor core)
//     warps 2-5   softmax group 0 (query tile 0), warps 6-9 softmax group 1 (query tile 1); one TMEM lane =
//                 one query row per thread: max pass, exp2 pass, P (bf16x2) written back over S with tcgen05.st,
//                 O and L read back, class-key rank-1 term added, normalised, stored.
//   The class-token QUERY row of the head is computed by one of the two groups (alternating) with plain FMAs on
//   the K/V tiles already in shared memory, inside the time it would otherwise wait for the tensor core.
//   TMEM region w (256 columns): S [0,256) -> P [0,128) | O [128,128+HD) | L [128+HD, +16).
//   Head dim 80 = one 64-column slab (128-byte rows, 128-B swizzle) + one 16-column slab (32-byte rows, 32-B swizzle)
//   per operand: the QKᵀ contraction is 4 + 1 K16 steps, P·V is an N=64 plus an N=16 MMA per key step.  Q/K stay
//   double-buffered; V (needed only after the softmax) is single-buffered for HD = 80 to fit 227 KB.
// ================================================================================================
constexpr int kA2Threads = 320;

template <int HD>
struct A2L {
  static constexpr int kX = HD - 64;                       // columns in the second slab (0 or 16)
  static constexpr int kSlab1 = 256 * 128;                 // [256 rows x 128 B]
  static constexpr int kSlabX = kX ? 256 * 32 : 0;         // [256 rows x 32 B]
  static constexpr int kOp = kSlab1 + kSlabX;              // one operand tile (Q, K or V)
  static constexpr int kQKStage = 2 * kOp;                 // Q | K
  static constexpr int kVStages = kX ? 1 : 2;
  static constexpr int kOffV = 2 * kQKStage;
  static constexpr int kOffOnes = kOffV + kVStages * kOp;  // 8 KB of bf16 1.0: the B operand of the row-sum MMA
  static constexpr int kOffMisc = kOffOnes + 8 * 1024;
  static constexpr int kColO = 128, kColL = 128 + HD;
};

template <int HD>
struct A2Misc {
  uint64_t full_qk[2], full_v[2], empty_qk[2], empty_v[2];
  uint64_t s_full[2], p_full[2], o_full[2], tmem_free[2];
  uint32_t tmem_slot;
  uint32_t pad[3];
  float red[2][8];          // per group: cross-warp max / sum scratch
  float vec[2][3][HD];      // per group: q0, k0, v0 of the current head as fp32
  float part[2][16][HD];    // per group: 16 key-slices of the class-row output
  float p_cls[2][264];      // per group: class-row probabilities (256 patch keys + class key)
};

template <int HD>
constexpr int a2_smem_bytes() {
  return A2L<HD>::kOffMisc + static_cast<int>((sizeof(A2Misc<HD>) + 1023) / 1024 * 1024) + 1024;
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// dot of 8 bf16 values (one uint4) with 8 fp32 values in shared memory, accumulated into (acc0, acc1)
__device__ __forceinline__ void dot8(const uint4& a, const float* __restrict__ q, float& acc0, float& acc1) {
  const float4 qa = *reinterpret_cast<const float4*>(q);
  const float4 qb = *reinterpret_cast<const float4*>(q + 4);
  ffma2(acc0, acc1, bf16_lo(a.x), bf16_hi(a.x), qa.x, qa.y, acc0, acc1);
  ffma2(acc0, acc1, bf16_lo(a.y), bf16_hi(a.y), qa.z, qa.w, acc0, acc1);
  ffma2(acc0, acc1, bf16_lo(a.z), bf16_hi(a.z), qb.x, qb.y, acc0, acc1);
  ffma2(acc0, acc1, bf16_lo(a.w), bf16_hi(a.w), qb.z, qb.w, acc0, acc1);
}

// dot over the head dim of row `row` of an operand tile in shared memory (slab 1: 128-B swizzle, slab 2: 32-B swizzle)
// with an fp32 vector in shared memory.  `tile` points at the operand's slab 1; slab 2 follows at +kSlab1.
template <int HD>
__device__ __forceinline__ float dot_row(const uint8_t* tile, int row, const float* __restrict__ q) {
  float acc0 = 0.f, acc1 = 0.f;
  const uint8_t* rowp = tile + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) dot8(*reinterpret_cast<const uint4*>(rowp + ((j ^ (row & 7)) << 4)), q + 8 * j, acc0, acc1);
  if constexpr (HD > 64) {
    const uint8_t* rowx = tile + A2L<HD>::kSlab1 + row * 32;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      dot8(*reinterpret_cast<const uint4*>(rowx + ((j ^ ((row >> 2) & 1)) << 4)), q + 64 + 8 * j, acc0, acc1);
  }
  return acc0 + acc1;
}

__device__ __forceinline__ uint32_t tmem_ld_1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Shared-memory descriptor for a tile of 32-byte rows with the 32-byte swizzle (CU_TENSOR_MAP_SWIZZLE_32B): 8-row groups
// are 256 B apart (SBO).  Used K-major (Q, K: 16 contraction elements per row) and MN-major (V: 16 head-dim columns
// per key row); layout code 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t make_sw32_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}

template <int HD>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_umma2_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmx,
                       const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n_ch, int T, int heads,
                       float scale_log2) {
  using L = A2L<HD>;
  using Misc = A2Misc<HD>;
  extern __shared__ uint8_t smem_a2_raw[];
  uint8_t* smem = smem_a2_raw + ((1024u - (smem_u32(smem_a2_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST
  Misc* mb = reinterpret_cast<Misc*>(smem + L::kOffMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const size_t row_stride = static_cast<size_t>(3) * d;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    if constexpr (HD > 64) tma_prefetch_desc(&tmx);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&mb->full_qk[s], 1);
      mbar_init(&mb->full_v[s], 1);
      // MMA commit + the class-row owner group (one arrival after its group barrier: own Q rows and K read) + the four
      // warps of the other group (one arrival each once the warp has read its Q rows)
      mbar_init(&mb->empty_qk[s], 6);
      mbar_init(&mb->empty_v[s], 2);   // MMA commit + the group that computed the class row from V
      mbar_init(&mb->s_full[s], 1);
      mbar_init(&mb->p_full[s], 4);
      mbar_init(&mb->o_full[s], 1);
      mbar_init(&mb->tmem_free[s], 4);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 8 * 1024 / 4; i += kA2Threads)
    reinterpret_cast<uint32_t*>(smem + L::kOffOnes)[i] = 0x3F803F80u;  // bf16 1.0 x2
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(&mb->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mb->tmem_slot;

  // stage / phase of head k:  Q,K: stage k&1, use (k>>1);   V: stage k % kVStages, use k / kVStages
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int k = 0;
      for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
        const int s = k & 1;
        const uint32_t u = (k >> 1) & 1;
        const int vs = k % L::kVStages;
        const uint32_t vu = (k / L::kVStages) & 1;
        const int crop = ch / heads, head = ch - crop * heads;
        const int row0 = crop * T + 1;
        uint8_t* st = smem + s * L::kQKStage;
        uint8_t* sv = smem + L::kOffV + vs * L::kOp;
        mbar_wait(&mb->empty_qk[s], u ^ 1);
        mbar_arrive_expect_tx(&mb->full_qk[s], 2 * L::kOp);
        tma_load_2d(st, &tm, &mb->full_qk[s], head * HD, row0);
        tma_load_2d(st + L::kOp, &tm, &mb->full_qk[s], d + head * HD, row0);
        if constexpr (HD > 64) {
          tma_load_2d(st + L::kSlab1, &tmx, &mb->full_qk[s], head * HD + 64, row0);
          tma_load_2d(st + L::kOp + L::kSlab1, &tmx, &mb->full_qk[s], d + head * HD + 64, row0);
        }
        mbar_wait(&mb->empty_v[vs], vu ^ 1);
        mbar_arrive_expect_tx(&mb->full_v[vs], L::kOp);
        tma_load_2d(sv, &tm, &mb->full_v[vs], 2 * d + head * HD, row0);
        if constexpr (HD > 64) tma_load_2d(sv + L::kSlab1, &tmx, &mb->full_v[vs], 2 * d + head * HD + 64, row0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform control flow (all lanes wait and build the uniform descriptors, one elected lane issues): from inside
    // a `lane == 0` branch ptxas wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY waterfall, and the MMAs of a
    // head then cost more issue time than they take to execute.
    const uint32_t idesc_s = make_idesc_f16(128, 256, 1);
    const uint32_t idesc_o = make_idesc_f16(128, 64, 1) | (1u << 16);  // B (= V) is MN-major
    const uint32_t idesc_ox = make_idesc_f16(128, 16, 1) | (1u << 16);
    const uint32_t idesc_l = make_idesc_f16(128, 16, 1);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t ones_desc = make_sw128_kmajor_desc(smem_base + L::kOffOnes);
    int k = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t u = (k >> 1) & 1, kp = k & 1;
      const int vs = k % L::kVStages;
      const uint32_t vu = (k / L::kVStages) & 1;
      const uint32_t sbase = smem_base + s * L::kQKStage;
      const uint32_t vbase = smem_base + L::kOffV + vs * L::kOp;
      mbar_wait(&mb->full_qk[s], u);
      const uint64_t k_desc = make_sw128_kmajor_desc(sbase + L::kOp);
      const uint64_t kx_desc = make_sw32_desc(sbase + L::kOp + L::kSlab1);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(&mb->tmem_free[w], kp ^ 1);  // group w has read the previous O out of its region
        tc_fence_after();
        const uint64_t q_desc = make_sw128_kmajor_desc(sbase + w * 16384);
        const uint64_t qx_desc = make_sw32_desc(sbase + L::kSlab1 + w * 4096);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tmem + w * 256, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
          if constexpr (HD > 64) umma_f16(tmem + w * 256, qx_desc, kx_desc, idesc_s, 1);
          umma_commit(&mb->s_full[w]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&mb->empty_qk[s]);
      __syncwarp();
      mbar_wait(&mb->full_v[vs], vu);
      const uint64_t v_desc0 = make_sw128_kmajor_desc(vbase);
      const uint64_t vx_desc0 = make_sw32_desc(vbase + L::kSlab1);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(&mb->p_full[w], kp);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            // 16 keys further down the V tile: 16 rows of 128 B (+128 in the addr>>4 field) / of 32 B (+32)
            umma_f16_ts(tmem + w * 256 + L::kColO, tmem + w * 256 + kk * 8, v_desc0 + kk * 128, idesc_o, kk != 0);
            if constexpr (HD > 64)
              umma_f16_ts(tmem + w * 256 + L::kColO + 64, tmem + w * 256 + kk * 8, vx_desc0 + kk * 32, idesc_ox, kk != 0);
          }
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)  // every element of the ones tile is 1.0, so any 16 x 16 slice will do
            umma_f16_ts(tmem + w * 256 + L::kColL, tmem + w * 256 + kk * 8, ones_desc + 2 * (kk & 3), idesc_l, kk != 0);
          umma_commit(&mb->o_full[w]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&mb->empty_v[vs]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int w = (warp - 2) >> 2;          // group = query tile
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;      // query row in the tile = TMEM lane
    const int gt = (warp - 2 - 4 * w) * 32 + lane;  // thread index inside the group, 0..127
    const uint32_t taddr = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    float* red = mb->red[w];
    float* pcls = mb->p_cls[w];
    float* q0f = mb->vec[w][0];
    float* k0f = mb->vec[w][1];
    float* v0f = mb->vec[w][2];
    // this thread's elements of the head's class-token q/k/v are fetched one iteration ahead as RAW bf16 bits (converting
    // at load time would stall on the load right away); its own query row is read from the Q tile TMA already staged
    uint32_t cq = 0, ck = 0, cv = 0;
    auto prefetch = [&](int ch) {
      if (gt < HD) {
        const int crop = ch / heads, head = ch - crop * heads;
        const unsigned short* cls_row =
            reinterpret_cast<const unsigned short*>(qkv + static_cast<size_t>(crop) * T * row_stride + head * HD);
        cq = __ldg(cls_row + gt);
        ck = __ldg(cls_row + d + gt);
        cv = __ldg(cls_row + 2 * d + gt);
      }
    };
    if (static_cast<int>(blockIdx.x) < n_ch) prefetch(blockIdx.x);
    int k = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t u = (k >> 1) & 1, kp = k & 1;
      const int vs = k % L::kVStages;
      const uint32_t vu = (k / L::kVStages) & 1;
      const int crop = ch / heads, head = ch - crop * heads;
      const int tok0 = crop * T;
      const int token = tok0 + 1 + w * 128 + r;
      const uint8_t* st = smem + s * L::kQKStage;
      const uint8_t* sv = smem + L::kOffV + vs * L::kOp;
      const bool cls_owner = ((k & 1) == w);

      if (gt < HD) {
        q0f[gt] = __uint_as_float(cq << 16);
        k0f[gt] = __uint_as_float(ck << 16);
        v0f[gt] = __uint_as_float(cv << 16);
      }
      if (ch + static_cast<int>(gridDim.x) < n_ch) prefetch(ch + gridDim.x);
      named_bar_sync(1 + w, 128);
      // class-token KEY for this thread's query row: s0 = q_r·k0, q_r from the (swizzled) Q tile in shared memory
      mbar_wait(&mb->full_qk[s], u);
      const float s0 = dot_row<HD>(st, w * 128 + r, k0f);
      if (!cls_owner) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&mb->empty_qk[s]);
      }

      // ---- class-token QUERY row, part 1 (scores + softmax statistics) from the K tile in smem
      float cls_l = 0.f;
      if (cls_owner) {
        float sc = -INFINITY;
        const float sa = dot_row<HD>(st + L::kOp, gt, q0f);
        const float sb = dot_row<HD>(st + L::kOp, gt + 128, q0f);
        if (gt == 0) {
          float acc = 0.f;
          for (int c = 0; c < HD; ++c) acc = fmaf(q0f[c], k0f[c], acc);
          sc = acc;
        }
        float m = fmaxf(fmaxf(sa, sb), sc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[gt >> 5] = m;
        named_bar_sync(1 + w, 128);  // also: every Q and K read of the group is done
        if (gt == 0) mbar_arrive(&mb->empty_qk[s]);
        m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * scale_log2;
        const float pa = ex2_ftz(fmaf(sa, scale_log2, -m)), pb = ex2_ftz(fmaf(sb, scale_log2, -m));
        float psum = pa + pb;
        pcls[gt] = pa;
        pcls[gt + 128] = pb;
        if (gt == 0) {
          const float pc = ex2_ftz(fmaf(sc, scale_log2, -m));
          pcls[256] = pc;
          psum += pc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        if (lane == 0) red[4 + (gt >> 5)] = psum;
        named_bar_sync(1 + w, 128);
        cls_l = (red[4] + red[5]) + (red[6] + red[7]);
      }

      // ---- own row: S in TMEM -> max -> P (bf16x2 packed) back into TMEM over S
      mbar_wait(&mb->s_full[w], kp);
      tc_fence_after();
      // Both passes keep one TMEM load in flight while the previous 32-column chunk is processed (two register buffers).
      float m = s0;
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + (c + 1) * 32, vb);
#pragma unroll
        for (int e = 0; e < 32; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(va[e]), __uint_as_float(va[e + 1])));
        tmem_ld_wait();
        tmem_ld_32x32(taddr + ((c + 2) & 7) * 32, va);  // last iteration: chunk 0 again, for the second pass
#pragma unroll
        for (int e = 0; e < 32; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(vb[e]), __uint_as_float(vb[e + 1])));
      }
      const float ms = m * scale_log2;
      const float p0 = ex2_ftz(fmaf(s0, scale_log2, -ms));
      auto emit_p = [&](const uint32_t (&v)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          pk[e] = pack2(ex2_ftz(fmaf(__uint_as_float(v[2 * e]), scale_log2, -ms)),
                        ex2_ftz(fmaf(__uint_as_float(v[2 * e + 1]), scale_log2, -ms)));
        tmem_st_32x16(taddr + c * 16, pk);  // columns [16c, 16c+16) <= columns already consumed
      };
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + (c + 1) * 32, vb);
        emit_p(va, c);
        tmem_ld_wait();
        if (c + 2 < 8) tmem_ld_32x32(taddr + (c + 2) * 32, va);
        emit_p(vb, c + 1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mb->p_full[w]);

      // ---- class-token QUERY row, part 2: O_cls = P_cls·V from the V tile in smem.
      //      thread = (16-key slice ks, 8-column chunk cc): one 16-byte read per key (two for cc < 2 when HD = 80).
      if (cls_owner) {
        mbar_wait(&mb->full_v[vs], vu);
        const int cc = gt & 7, ks = gt >> 3;
        const float* pp = pcls + ks * 16;
        auto accumulate = [&](const uint8_t* base, int pitch, int chunk_of_key0, int phase_shift, float* dst) {
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
          for (int key = 0; key < 16; ++key) {
            // ks*16 + key has the same low 4 bits as key, so the swizzle phase depends on `key` only
            const int phase = phase_shift == 0 ? (key & 7) : ((key >> 2) & 1);
            const uint4 a = *reinterpret_cast<const uint4*>(base + key * pitch + ((chunk_of_key0 ^ phase) << 4));
            const float pkey = pp[key];
            acc[0] = fmaf(pkey, bf16_lo(a.x), acc[0]);
            acc[1] = fmaf(pkey, bf16_hi(a.x), acc[1]);
            acc[2] = fmaf(pkey, bf16_lo(a.y), acc[2]);
            acc[3] = fmaf(pkey, bf16_hi(a.y), acc[3]);
            acc[4] = fmaf(pkey, bf16_lo(a.z), acc[4]);
            acc[5] = fmaf(pkey, bf16_hi(a.z), acc[5]);
            acc[6] = fmaf(pkey, bf16_lo(a.w), acc[6]);
            acc[7] = fmaf(pkey, bf16_hi(a.w), acc[7]);
          }
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        };
        accumulate(sv + (ks * 16) * 128, 128, cc, 0, &mb->part[w][ks][cc * 8]);
        if constexpr (HD > 64) {
          if (cc < 2) accumulate(sv + L::kSlab1 + (ks * 16) * 32, 32, cc, 1, &mb->part[w][ks][64 + cc * 8]);
        }
        named_bar_sync(1 + w, 128);  // also: every V read of the group is done
        if (gt == 0) mbar_arrive(&mb->empty_v[vs]);
        if (gt < HD) {
          float o = pcls[256] * v0f[gt];
#pragma unroll
          for (int i = 0; i < 16; ++i) o += mb->part[w][i][gt];
          out[static_cast<size_t>(tok0) * d + head * HD + gt] = __float2bfloat16_rn(o / cls_l);
        }
      }

      // ---- own row: (O + p0·v0) / (L + p0) -> bf16
      mbar_wait(&mb->o_full[w], kp);
      tc_fence_after();
      const float lsum = __uint_as_float(tmem_ld_1(taddr + L::kColL));
      __nv_bfloat16* orow = out + static_cast<size_t>(token) * d + head * HD;
      auto store8 = [&](const uint32_t* v, const float* v0, __nv_bfloat16* dst, float inv, float p0i) {
        const float4 va4 = *reinterpret_cast<const float4*>(v0);
        const float4 vb4 = *reinterpret_cast<const float4*>(v0 + 4);
        uint4 o4;
        o4.x = pack2(fmaf(__uint_as_float(v[0]), inv, p0i * va4.x), fmaf(__uint_as_float(v[1]), inv, p0i * va4.y));
        o4.y = pack2(fmaf(__uint_as_float(v[2]), inv, p0i * va4.z), fmaf(__uint_as_float(v[3]), inv, p0i * va4.w));
        o4.z = pack2(fmaf(__uint_as_float(v[4]), inv, p0i * vb4.x), fmaf(__uint_as_float(v[5]), inv, p0i * vb4.y));
        o4.w = pack2(fmaf(__uint_as_float(v[6]), inv, p0i * vb4.z), fmaf(__uint_as_float(v[7]), inv, p0i * vb4.w));
        *reinterpret_cast<uint4*>(dst) = o4;
      };
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + L::kColO + c * 32, v);
        tmem_ld_wait();
        const float inv = 1.0f / (lsum + p0);
        const float p0i = p0 * inv;
#pragma unroll
        for (int j = 0; j < 4; ++j) store8(v + 8 * j, v0f + c * 32 + 8 * j, orow + c * 32 + 8 * j, inv, p0i);
      }
      if constexpr (HD > 64) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + L::kColO + 64, v);
        tmem_ld_wait();
        const float inv = 1.0f / (lsum + p0);
        const float p0i = p0 * inv;
#pragma unroll
        for (int j = 0; j < 2; ++j) store8(v + 8 * j, v0f + 64 + 8 * j, orow + 64 + 8 * j, inv, p0i);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mb->tmem_free[w]);
      // the group's scratch (q0f/k0f/v0f, p_cls, part) is rewritten next iteration: everyone must be done with it
      named_bar_sync(1 + w, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int HD>
static int attention_umma2_launch(const void* qkv, void* out, int n, int T, int heads, cudaStream_t stream) {
  const int d = heads * HD;
  CUtensorMap tm, tmx;
  B2C_TRY(make_tmap_2d(&tm, qkv, static_cast<uint64_t>(n) * T, 3ull * d, 3ull * d * 2, 256, 1));
  tmx = tm;
  if (HD > 64)
    B2C_TRY(make_tmap_2d_sw(&tmx, qkv, static_cast<uint64_t>(n) * T, 3ull * d, 3ull * d * 2, 256, HD - 64, B2C_BF16, 32));
  constexpr int smem_bytes = a2_smem_bytes<HD>();
  static_assert(smem_bytes <= 227 * 1024, "attention v2 shared memory exceeds 227 KB");
  auto kern = attention_umma2_kernel<HD>;
  static PerDeviceFlag attr_once;
  if (attr_once.first_use()) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  }
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  const int n_ch = n * heads;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));  // log2(e) / sqrt(hd)
  kern<<<n_ch < sms ? n_ch : sms, kA2Threads, smem_bytes, stream>>>(
      tm, tmx, static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n_ch, T, heads, scale_log2);
  B2C_POST_LAUNCH("attention_umma2_kernel");
  return 0;
}

constexpr int kA3KB = 96;                       // keys per block (T = 577: six blocks)
constexpr int kA3BlockBytes = kA3KB * 128;      // one K or V block: 96 rows x 128 B
constexpr int kA3QStage = 256 * 128;
constexpr int kA3MaxNB = 6;

// first traced iteration of the development phase trace (b2c_debug_attn_trace_start)
__device__ int g_attn_trace_k0 = 4;

// exp2 on the FMA pipe for part of a row's elements (the XU pipe, 16 ex2 per clock per SM, is what the exp2 pass
// saturates): x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5·2^23, 2^f by a degree-3 minimax
// polynomial on [-0.5, 0.5] (relative error 7.6e-5, a fiftieth of the bf16 rounding P gets next), 2^n added into the
// exponent field.  x is clamped to >= -126 so the exponent cannot wrap; x <= 0 always (the row maximum was subtracted).
__device__ __forceinline__ void exp2_poly2(float& y0, float& y1, float x0, float x1) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, kMagic, kMagic);
  fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
  fadd2(f0, f1, x0, x1, -n0, -n1);
  ffma2(p0, p1, f0, f1, 0.05520550534129143f, 0.05520550534129143f, 0.24261397123336792f, 0.24261397123336792f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.6932547688484192f, 0.6932547688484192f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.9999276995658875f, 0.9999276995658875f);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}


// ================================================================================================
// (1c) tcgen05 attention for longer sequences (ViT-L/14-336: T = 577 = class token + 576 patches), head dim 64,
//      patch count a multiple of 96.  Two-pass exact softmax over key blocks of 96:
//   CTA (one per SM, 22 warps) loops over (crop, head); K and V of the head stay resident in shared memory (2 x 72 KB
//   for 576 patches), query tiles stream through a 2-stage ring, 256 rows (two 128-row tiles, one per softmax group) per
//   stage.  Per query tile the key blocks are visited twice: pass A (S_b = Q·K_bᵀ, running row max only) and pass B (S_b
//   again, p = exp2(s·c − m·c), P_b in bf16 written over S_b in TMEM, O += P_b·V_b on the tensor core, row sums in
//   registers).  Recomputing S costs tensor time that is idle anyway and keeps the softmax exact with no rescaling of
//   O.  S is double-buffered per group — TMEM region g (256 columns): S/P buffer 0 [0,96) | buffer 1 [96,192) | O
//   [192,256) — and S tiles are issued two tiles ahead; PV_b and the next S into the same buffer rely on tcgen05.mma
//   executing in issue order.
//     warp 0       TMA producer            warp 1   MMA issuer
//     warps 2-17   softmax: group (query tile) x column half (48 of a block's 96 keys) x TMEM lane quarter, so every
//                  SM sub-partition hosts four softmax warps; the halves exchange their row maximum once per query tile
//                  (after pass A) and their row sums once (after pass B) through 64-thread barriers.  Half h of a buffer
//                  owns S columns [48h, 48h+48) and writes its packed P over them at [48h, 48h+24).
//     warps 18-21  class-token warps: s0[r] = q_r·k0 for every query-tile pair (the class-token KEY, a rank-1 term as in
//                  the other tcgen05 kernels), and the class-token QUERY row of the head from the resident K and V.
//   Measured against its 8-softmax-warp predecessor (one thread per full row, class row as a second kernel re-reading
//   K and V from global memory): 1.18 -> 0.93 ms per launch at 256 crops; ViT-L/14-336 step 604 -> 629 images/s.
// ================================================================================================
constexpr int kA6Threads = 64 + 512 + 128;
constexpr int kA6ClsWarp0 = 18;

struct A6Misc {
  uint64_t k_full, k_empty, v_full, v_empty, q_full[2], q_empty[2];
  uint64_t s_full[2][2], s_free[2][2], p_full[2][2], o_full[2], o_free[2];
  uint64_t cls_done[2];
  uint32_t tmem_slot;
  uint32_t pad[3];
  float vec[3][64];        // q0, k0 as packed bf16 (32 words each), v0 as fp32
  float s0[2][256];        // per Q stage: q_r·k0 of the stage's 256 query rows
  float mx[2][2][128];     // per group, column half: partial row maxima (pass A)
  float ls[2][2][128];     // per group, column half: partial row sums (pass B)
  float red[8];
  float p_cls[592];        // class-row probabilities: patch keys, then the class key at [G2]
  float part[16][64];
};

constexpr int a6_smem_bytes(int NB) {
  return 2 * NB * kA3BlockBytes + 2 * kA3QStage + static_cast<int>((sizeof(A6Misc) + 1023) / 1024 * 1024) + 1024;
}

__global__ void __launch_bounds__(kA6Threads, 1)
attention_umma6_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                       const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n_ch, int T, int heads,
                       int NB, float scale_log2) {
  extern __shared__ uint8_t smem_a6_raw[];
  uint8_t* smem = smem_a6_raw + ((1024u - (smem_u32(smem_a6_raw) & 1023u)) & 1023u);  // keeps the address space: LDS/STS
  const int off_v = NB * kA3BlockBytes;
  const int off_q = 2 * NB * kA3BlockBytes;
  A6Misc* mb = reinterpret_cast<A6Misc*>(smem + off_q + 2 * kA3QStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * 64;
  const size_t row_stride = static_cast<size_t>(3) * d;
  const int G2 = NB * kA3KB;
  const int NQ = (G2 + 255) / 256;  // query-tile pairs per head

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_q);
    mbar_init(&mb->k_full, 1);
    mbar_init(&mb->k_empty, 1 + 4);  // MMA commit + class warps (class-row scores read K)
    mbar_init(&mb->v_full, 1);
    mbar_init(&mb->v_empty, 1 + 4);  // MMA commit + class warps (P_cls·V reads V)
    for (int s = 0; s < 2; ++s) {
      mbar_init(&mb->q_full[s], 1);
      mbar_init(&mb->q_empty[s], 1 + 4 + 16);  // MMA commit + class warps (Q rows read) + softmax warps (s0 of the stage read)
      mbar_init(&mb->o_full[s], 1);
      mbar_init(&mb->o_free[s], 8);
      mbar_init(&mb->cls_done[s], 4);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&mb->s_full[s][b], 1);
        mbar_init(&mb->s_free[s][b], 8);
        mbar_init(&mb->p_full[s][b], 8);
      }
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&mb->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mb->tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int hk = 0, qk = 0;
      for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++hk) {
        const int crop = ch / heads, head = ch - crop * heads;
        const int row0 = crop * T + 1;
        auto load_q = [&](int pair) {
          const int s = qk & 1;
          mbar_wait(&mb->q_empty[s], ((qk >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&mb->q_full[s], kA3QStage);
          tma_load_2d(smem + off_q + s * kA3QStage, &tm_q, &mb->q_full[s], head * 64, row0 + pair * 256);
          ++qk;
        };
        mbar_wait(&mb->k_empty, (hk & 1) ^ 1);
        mbar_arrive_expect_tx(&mb->k_full, NB * kA3BlockBytes);
        for (int b = 0; b < NB; ++b) tma_load_2d(smem + b * kA3BlockBytes, &tm_kv, &mb->k_full, d + head * 64, row0 + b * kA3KB);
        load_q(0);
        mbar_wait(&mb->v_empty, (hk & 1) ^ 1);
        mbar_arrive_expect_tx(&mb->v_full, NB * kA3BlockBytes);
        for (int b = 0; b < NB; ++b)
          tma_load_2d(smem + off_v + b * kA3BlockBytes, &tm_kv, &mb->v_full, 2 * d + head * 64, row0 + b * kA3KB);
        for (int pair = 1; pair < NQ; ++pair) load_q(pair);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues)
    const uint32_t idesc_s = make_idesc_f16(128, kA3KB, 1);
    const uint32_t idesc_o = make_idesc_f16(128, 64, 1) | (1u << 16);  // B (= V) is MN-major
    const uint32_t smem_base = smem_u32(smem);
    int hk = 0, qk = 0;
    uint32_t n_sfree[2][2] = {{0, 0}, {0, 0}}, n_pfull[2][2] = {{0, 0}, {0, 0}}, n_o[2] = {0, 0};
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++hk) {
      mbar_wait(&mb->k_full, hk & 1);
      bool v_ready = false;
      for (int pair = 0; pair < NQ; ++pair, ++qk) {
        const int s = qk & 1;
        mbar_wait(&mb->q_full[s], (qk >> 1) & 1);
        tc_fence_after();
        const int nact = (pair * 256 + 128 < G2) ? 2 : 1;
        const uint32_t qbase = smem_base + off_q + s * kA3QStage;
        // P_b·V_b for tile i (a pass-B tile) of group w, accumulated into O
        auto issue_pv = [&](int i, int w) {
          const int b = i - NB, buf = i & 1;
          mbar_wait(&mb->p_full[w][buf], n_pfull[w][buf] & 1);
          ++n_pfull[w][buf];
          if (b == 0) {
            mbar_wait(&mb->o_free[w], (n_o[w] & 1) ^ 1);  // the group has read the previous O of its region
            ++n_o[w];
            if (!v_ready) {
              mbar_wait(&mb->v_full, hk & 1);
              v_ready = true;
            }
          }
          tc_fence_after();
          const uint64_t v_desc0 = make_sw128_kmajor_desc(smem_base + off_v + b * kA3BlockBytes);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kA3KB / 16; ++kk) {
              // half 0's packed P at buffer columns [0,24), half 1's at [48,72): three 16-key steps each
              const uint32_t pcol = buf * kA3KB + (kk < 3 ? 0 : 48) + (kk % 3) * 8;
              umma_f16_ts(tmem + w * 256 + 192, tmem + w * 256 + pcol, v_desc0 + kk * 128, idesc_o, (b | kk) != 0);
            }
          }
          __syncwarp();
        };
        for (int i = 0; i < 2 * NB; ++i) {
          const int b = i < NB ? i : i - NB, buf = i & 1;
          const uint64_t k_desc = make_sw128_kmajor_desc(smem_base + b * kA3BlockBytes);
          for (int w = 0; w < nact; ++w) {
            if (i >= 2) {
              if (i - 2 < NB) {  // the buffer held a pass-A tile: wait until the group has scanned it
                mbar_wait(&mb->s_free[w][buf], n_sfree[w][buf] & 1);
                ++n_sfree[w][buf];
                tc_fence_after();
              } else {           // it held a pass-B tile: its P·V goes first (same thread, executes in order)
                issue_pv(i - 2, w);
              }
            }
            const uint64_t q_desc = make_sw128_kmajor_desc(qbase + w * 16384);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem + w * 256 + buf * kA3KB, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
              umma_commit(&mb->s_full[w][buf]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          umma_commit(&mb->q_empty[s]);                       // every S of this pair has been issued
          if (pair == NQ - 1) umma_commit(&mb->k_empty);      // ... and of this head
        }
        __syncwarp();
        for (int i = 2 * NB - 2; i < 2 * NB; ++i)
          for (int w = 0; w < nact; ++w) {
            issue_pv(i, w);
            if (i == 2 * NB - 1) {
              if (elect_one()) umma_commit(&mb->o_full[w]);
              __syncwarp();
            }
          }
      }
      if (elect_one()) umma_commit(&mb->v_empty);
      __syncwarp();
    }
  } else if (warp < kA6ClsWarp0) {
    // ------------------------------------------------------------------ softmax: 2 groups x 2 column halves x 4 warps
    const int sw = warp - 2;
    const int w = sw >> 3;                  // group = query tile of the pair
    const int h = (sw >> 2) & 1;            // column half: keys [48h, 48h + 48) of every block
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;      // query row in the tile = TMEM lane
    const uint32_t taddr = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const int pair_bar = 3 + w * 4 + quarter;
    const float* v0f = mb->vec[2];
    int qk = 0;
    uint32_t n_sfull[2] = {0, 0}, n_ofull = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x) {
      const int crop = ch / heads, head = ch - crop * heads;
      const int tok0 = crop * T;
      for (int pair = 0; pair < NQ; ++pair, ++qk) {
        const int s = qk & 1;
        const int qrow = pair * 256 + w * 128 + r;
        const bool active = pair * 256 + w * 128 < G2;  // group-uniform
        if (!active) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&mb->q_empty[s]);
          continue;
        }

        // rows of the tile past the head's last patch (the last query tile of 576 patches holds 64 rows) carry no work:
        // the warp keeps the barrier protocol going and skips the arithmetic (warp-uniform: a warp is 32 consecutive rows)
        const bool rows_live = pair * 256 + w * 128 + quarter * 32 < G2;

        // ---- pass A: running row max over this half's 48 keys of every block
        float m = -INFINITY;
        for (int b = 0; b < NB; ++b) {
          const int buf = b & 1;
          mbar_wait(&mb->s_full[w][buf], n_sfull[buf] & 1);
          ++n_sfull[buf];
          tc_fence_after();
          if (rows_live) {
            uint32_t va[32], vb[16];
            const uint32_t ta = taddr + buf * kA3KB + 48 * h;
            tmem_ld_32x32(ta, va);
            tmem_ld_32x16(ta + 32, vb);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(va[e]), __uint_as_float(va[e + 1])));
#pragma unroll
            for (int e = 0; e < 16; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(vb[e]), __uint_as_float(vb[e + 1])));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mb->s_free[w][buf]);
        }
        // the class-token KEY joins the maximum here: s0 comes from the class warps
        mbar_wait(&mb->cls_done[s], (qk >> 1) & 1);
        const float s0 = mb->s0[s][w * 128 + r];
        mb->mx[w][h][r] = m;
        __syncwarp();
        if (lane == 0) mbar_arrive(&mb->q_empty[s]);  // s0 of the stage has been read by this warp
        named_bar_sync(pair_bar, 64);
        m = fmaxf(fmaxf(m, mb->mx[w][h ^ 1][r]), s0);
        const float ms = m * scale_log2;
        const float nms = -ms;
        const float p0 = ex2_ftz(fmaf(s0, scale_log2, nms));
        float lsum = 0.f;

        // ---- pass B: probabilities (bf16x2) over this half's S columns, partial row sum in registers
        for (int b = 0; b < NB; ++b) {
          const int buf = (NB + b) & 1;
          mbar_wait(&mb->s_full[w][buf], n_sfull[buf] & 1);
          ++n_sfull[buf];
          tc_fence_after();
          if (!rows_live) {  // nothing reads these rows of O
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mb->p_full[w][buf]);
            continue;
          }
          uint32_t va[32], vb[16];
          const uint32_t ta = taddr + buf * kA3KB + 48 * h;
          tmem_ld_32x32(ta, va);
          tmem_ld_32x16(ta + 32, vb);
          tmem_ld_wait();
          uint32_t pk[16], pk2[8];
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float t0, t1;
            ffma2(t0, t1, __uint_as_float(va[2 * e]), __uint_as_float(va[2 * e + 1]), scale_log2, scale_log2, nms, nms);
            const float pa = ex2_ftz(t0), pb = ex2_ftz(t1);
            acc0 += pa;
            acc1 += pb;
            pk[e] = pack2(pa, pb);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float t0, t1;
            ffma2(t0, t1, __uint_as_float(vb[2 * e]), __uint_as_float(vb[2 * e + 1]), scale_log2, scale_log2, nms, nms);
            const float pa = ex2_ftz(t0), pb = ex2_ftz(t1);
            acc0 += pa;
            acc1 += pb;
            pk2[e] = pack2(pa, pb);
          }
          lsum += acc0 + acc1;
          tmem_st_32x16(ta, pk);        // packed columns [48h, 48h + 16): over S columns this thread has consumed
          tmem_st_32x8(ta + 16, pk2);   // ... [48h + 16, 48h + 24)
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mb->p_full[w][buf]);
        }
        mb->ls[w][h][r] = lsum;
        named_bar_sync(pair_bar, 64);
        lsum += mb->ls[w][h ^ 1][r] + p0;

        // ---- own row, own 32 output columns: (O + p0·v0) / (L + p0) -> bf16
        mbar_wait(&mb->o_full[w], n_ofull & 1);
        ++n_ofull;
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32(taddr + 192 + h * 32, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mb->o_free[w]);
          const float inv = 1.0f / lsum;
          const float p0i = p0 * inv;
          if (qrow < G2) {
            __nv_bfloat16* orow = out + (static_cast<size_t>(tok0) + 1 + qrow) * d + head * 64 + h * 32;
            const float* v0 = v0f + h * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 va4 = *reinterpret_cast<const float4*>(v0 + 8 * j);
              const float4 vb4 = *reinterpret_cast<const float4*>(v0 + 8 * j + 4);
              uint4 o4;
              o4.x = pack2(fmaf(__uint_as_float(v[8 * j + 0]), inv, p0i * va4.x), fmaf(__uint_as_float(v[8 * j + 1]), inv, p0i * va4.y));
              o4.y = pack2(fmaf(__uint_as_float(v[8 * j + 2]), inv, p0i * va4.z), fmaf(__uint_as_float(v[8 * j + 3]), inv, p0i * va4.w));
              o4.z = pack2(fmaf(__uint_as_float(v[8 * j + 4]), inv, p0i * vb4.x), fmaf(__uint_as_float(v[8 * j + 5]), inv, p0i * vb4.y));
              o4.w = pack2(fmaf(__uint_as_float(v[8 * j + 6]), inv, p0i * vb4.z), fmaf(__uint_as_float(v[8 * j + 7]), inv, p0i * vb4.w));
              *reinterpret_cast<uint4*>(orow + 8 * j) = o4;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ class-token warps (128 threads, barrier 11)
    const int t = static_cast<int>(threadIdx.x) - kA6ClsWarp0 * 32;
    const int cw = warp - kA6ClsWarp0;
    float* red = mb->red;
    float* pcls = mb->p_cls;
    uint32_t* q0b = reinterpret_cast<uint32_t*>(mb->vec[0]);
    uint32_t* k0b = reinterpret_cast<uint32_t*>(mb->vec[1]);
    float* v0f = mb->vec[2];
    uint32_t cq = 0, ck = 0, cv = 0;  // two bf16 each (threads 0..31)
    auto prefetch = [&](int ch) {
      if (t < 32) {
        const int crop = ch / heads, head = ch - crop * heads;
        const uint32_t* cls_row = reinterpret_cast<const uint32_t*>(qkv + static_cast<size_t>(crop) * T * row_stride + head * 64);
        cq = __ldg(cls_row + t);
        ck = __ldg(cls_row + d / 2 + t);
        cv = __ldg(cls_row + d + t);
      }
    };
    // dots of rows `ra` and `rb` of a swizzled 128-byte-row tile with a packed bf16 vector (32 words in shared memory)
    auto dot2_packed = [](const uint8_t* tile, int ra, int rb, const uint32_t* vec, float& da, float& db) {
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
      const uint8_t* pa = tile + ra * 128;
      const uint8_t* pb = tile + rb * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 v[4], xa[4], xb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = reinterpret_cast<const uint4*>(vec)[4 * half + j];
          xa[j] = *reinterpret_cast<const uint4*>(pa + (((4 * half + j) ^ (ra & 7)) << 4));
          xb[j] = *reinterpret_cast<const uint4*>(pb + (((4 * half + j) ^ (rb & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t vw[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          const uint32_t aw[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w};
          const uint32_t bw[4] = {xb[j].x, xb[j].y, xb[j].z, xb[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float vl = bf16_lo(vw[e]), vh = bf16_hi(vw[e]);
            ffma2(a0, a1, bf16_lo(aw[e]), bf16_hi(aw[e]), vl, vh, a0, a1);
            ffma2(b0, b1, bf16_lo(bw[e]), bf16_hi(bw[e]), vl, vh, b0, b1);
          }
        }
      }
      da = a0 + a1;
      db = b0 + b1;
    };
    if (static_cast<int>(blockIdx.x) < n_ch) prefetch(blockIdx.x);
    int hk = 0, qk = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++hk) {
      const int crop = ch / heads, head = ch - crop * heads;
      const int tok0 = crop * T;
      // K of the head has landed => K/V/vec of the previous head have been released by everyone (k_empty / v_empty)
      mbar_wait(&mb->k_full, hk & 1);
      named_bar_sync(11, 128);  // the previous head's class row (vec, p_cls, part, red) is finished in every class warp
      if (t < 32) {
        q0b[t] = cq;
        k0b[t] = ck;
        *reinterpret_cast<float2*>(v0f + 2 * t) = make_float2(bf16_lo(cv), bf16_hi(cv));
      }
      if (ch + static_cast<int>(gridDim.x) < n_ch) prefetch(ch + gridDim.x);
      named_bar_sync(11, 128);
      // class-token KEY for one query-tile pair: s0[r] = q_r·k0 -> softmax warps
      auto key_scores = [&]() {
        const int s = qk & 1;
        mbar_wait(&mb->q_full[s], (qk >> 1) & 1);
        float a, b;
        dot2_packed(smem + off_q + s * kA3QStage, t, t + 128, k0b, a, b);
        mb->s0[s][t] = a;
        mb->s0[s][t + 128] = b;
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&mb->cls_done[s]);
          mbar_arrive(&mb->q_empty[s]);
        }
        ++qk;
      };
      int pair = 0;
      for (; pair < NQ && pair < 2; ++pair) key_scores();  // both Q stages are prefetched at the start of a head
      // class-token QUERY row: scores over the resident K (rows t, t+128, ... < G2), class key, softmax
      float sa[5];
      {
        const uint8_t* ktile = smem;
        float x, y;
        dot2_packed(ktile, t, t + 128, q0b, sa[0], sa[1]);
        if (t + 384 < G2) dot2_packed(ktile, t + 256, t + 384, q0b, sa[2], sa[3]);
        else if (t + 256 < G2) { dot2_packed(ktile, t + 256, t + 256, q0b, sa[2], y); sa[3] = -INFINITY; }
        else { sa[2] = sa[3] = -INFINITY; }
        if (t + 512 < G2) { dot2_packed(ktile, t + 512, t + 512, q0b, sa[4], x); } else sa[4] = -INFINITY;
        if (t + 128 >= G2) sa[1] = -INFINITY;
        if (t >= G2) sa[0] = -INFINITY;
      }
      float sc;
      {
        const uint32_t qa = q0b[lane], ka = k0b[lane];
        sc = fmaf(bf16_lo(qa), bf16_lo(ka), bf16_hi(qa) * bf16_hi(ka));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      float cm;
      {
        const float mine = fmaxf(fmaxf(fmaxf(sa[0], sa[1]), fmaxf(sa[2], sa[3])), fmaxf(sa[4], sc));
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(cm) : "f"(mine));
      }
      if (lane == 0) red[cw] = cm;
      named_bar_sync(11, 128);
      cm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * scale_log2;
      const float pc = ex2_ftz(fmaf(sc, scale_log2, -cm));
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float pv = ex2_ftz(fmaf(sa[i], scale_log2, -cm));  // -inf -> 0
        if (t + 128 * i < G2) pcls[t + 128 * i] = pv;
        psum += pv;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      if (lane == 0) red[4 + cw] = psum;
      named_bar_sync(11, 128);
      // every K read of the class warps is done
      if (lane == 0) mbar_arrive(&mb->k_empty);
      // O_cls = P_cls·V over the resident V: thread = (key slice ks of G2/16 keys, 8-column chunk cc)
      mbar_wait(&mb->v_full, hk & 1);
      {
        const int cc = t & 7, ks = t >> 3;
        const int per = G2 >> 4;  // keys per slice (36 for 576)
        const uint8_t* vt = smem + off_v;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int k0 = 0; k0 < per; k0 += 4) {
          uint4 row[4];
          float pk4[4];
#pragma unroll
          for (int key = 0; key < 4; ++key) {
            const int j = ks * per + k0 + key;
            const bool ok = k0 + key < per;
            row[key] = ok ? *reinterpret_cast<const uint4*>(vt + j * 128 + ((cc ^ (j & 7)) << 4)) : make_uint4(0, 0, 0, 0);
            pk4[key] = ok ? pcls[j] : 0.f;
          }
#pragma unroll
          for (int key = 0; key < 4; ++key) {
            const uint4 a = row[key];
            const float pkey = pk4[key];
            ffma2(acc[0], acc[1], bf16_lo(a.x), bf16_hi(a.x), pkey, pkey, acc[0], acc[1]);
            ffma2(acc[2], acc[3], bf16_lo(a.y), bf16_hi(a.y), pkey, pkey, acc[2], acc[3]);
            ffma2(acc[4], acc[5], bf16_lo(a.z), bf16_hi(a.z), pkey, pkey, acc[4], acc[5]);
            ffma2(acc[6], acc[7], bf16_lo(a.w), bf16_hi(a.w), pkey, pkey, acc[6], acc[7]);
          }
        }
        *reinterpret_cast<float4*>(&mb->part[ks][cc * 8]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(&mb->part[ks][cc * 8 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      named_bar_sync(11, 128);
      if (lane == 0) mbar_arrive(&mb->v_empty);  // every V read of the class warps is done
      if (t < 64) {
        const float cls_l = ((red[4] + red[5]) + (red[6] + red[7])) + pc;
        float o = pc * v0f[t];
#pragma unroll
        for (int i = 0; i < 16; ++i) o += mb->part[i][t];
        out[static_cast<size_t>(tok0) * d + head * 64 + t] = __float2bfloat16_rn(o / cls_l);
      }
      for (; pair < NQ; ++pair) key_scores();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int attention_umma6_launch(const void* qkv, void* out, int n, int T, int heads, cudaStream_t stream) {
  const int d = heads * 64;
  const int NB = (T - 1) / kA3KB;
  CUtensorMap tm_kv, tm_q;
  const uint64_t rows = static_cast<uint64_t>(n) * T;
  B2C_TRY(make_tmap_2d(&tm_kv, qkv, rows, 3ull * d, 3ull * d * 2, kA3KB, 1));
  B2C_TRY(make_tmap_2d(&tm_q, qkv, rows, 3ull * d, 3ull * d * 2, 256, 1));
  const int smem_bytes = a6_smem_bytes(NB);
  B2C_REQUIRE(smem_bytes <= 227 * 1024, "attention v6: T=%d does not fit shared memory", T);
  B2C_REQUIRE((NB * kA3KB) % 16 == 0 && NB * kA3KB <= 576, "attention v6: %d patch keys unsupported", NB * kA3KB);
  auto kern = attention_umma6_kernel;
  static PerDeviceMax smem_set;
  if (smem_set.raise(static_cast<long long>(smem_bytes)))
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  const int n_ch = n * heads;
  const float scale_log2 = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64)
  kern<<<n_ch < sms ? n_ch : sms, kA6Threads, smem_bytes, stream>>>(
      tm_kv, tm_q, static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n_ch, T, heads, NB, scale_log2);
  B2C_POST_LAUNCH("attention_umma6_kernel");
  return 0;
}

// ================================================================================================
// (1e) tcgen05 attention for T = 257, head dim 64 (ViT-L/14-224): persistent, one CTA per SM, 22 warps.
//     warp 0       TMA producer: Q (two 128-row tiles), K, V of the NEXT head into the other of two 96 KB stages
//     warp 1       MMA issuer: S_g = Q_g·Kᵀ (M128 N256 K64, SS), O_g = P_g·V (P from TMEM, V MN-major from the TMA tile),
//                  L_g = P_g·1.  Issue order S0(k) PV1(k-1) S1(k) PV0(k): the two query tiles run half a period apart, so
//                  one tile's exp2 pass (XU pipe) overlaps the other's tensor-core / epilogue / max phases instead of its
//                  exp2 pass.  The tile index is a compile-time constant in every operand (a run-time index sends each
//                  operand through R2UR).
//     warps 2-17   softmax: group g = (warp-2)/8 (query tile), column half h = ((warp-2)/4)&1 (keys [128h, 128h+128)),
//                  TMEM lane quarter = warp & 3.  max pass -> 64-thread exchange with the other column half -> exp2 pass
//                  -> P (packed bf16 over the consumed S columns) -> epilogue.  No group barrier and no class-token
//                  arithmetic: p0 = exp2((s0 - m)·c) with s0 read from shared memory.
//     warps 18-21  class-token warps, one per SM sub-partition, running one head AHEAD of the softmax warps (they need
//                  only the operand tiles, which the producer prefetches a head ahead): s0[r] = q_r·k0 for the 256
//                  query rows, the class QUERY row's scores, softmax and P_cls·V, and its output row.
//   TMEM region g (256 columns): S [0,256) -> P_0 [0,64) | O [64,128) | P_1 [128,192) | L [192,208).
//   Operand stage s is released (empty_qk) by the MMA commit, the 4 class warps and the 16 softmax warps (end of their
//   iteration: they read s0 / v0 of the stage until the epilogue).
//   Why this shape (phase traces with clock64 at phase boundaries, profiles/r2_attn_trace_v4_v5.txt): in the predecessor
//   (16 softmax warps that also did the class token, both S issued back to back) the group that owned the class-token
//   QUERY row spent 3-5 k cycles on it between its P and its epilogue — its LDS / SHFL traffic queues behind the other
//   group's MUFU.EX2 in the same MIO queue — which delayed tmem_free and the next S of that tile; every phase boundary was
//   a 256-thread barrier; and both groups ran their exp2 passes at the same time (4 k cycles XU-bound) with the
//   tensor-core / epilogue phases exposed.  12.7 k -> 8.1 k cycles per (crop, head); energy per launch -14 %.
// ================================================================================================
constexpr int kA5Threads = 64 + 512 + 128;
constexpr int kA5ClsWarp0 = 18;

struct A5Misc {
  uint64_t full_qk[2], full_v[2], empty_qk[2], empty_v[2];
  uint64_t s_full[2], p_full[2], o_full[2], tmem_free[2];
  uint64_t cls_done[2];
  uint32_t tmem_slot;
  uint32_t pad[3];
  float vec[2][3][64];      // per stage: q0, k0, v0 of the head as fp32
  float s0[2][256];         // per stage: q_r·k0 of every patch query row
  float mx[2][2][2][128];   // per iteration parity, group, column half: partial row maxima
  float red[8];             // class warps: max / sum scratch
  float p_cls[264];         // class-row probabilities (256 patch keys + class key)
  float part[16][64];       // 16 key-slices of the class-row output
};

constexpr int a5_smem_bytes() {
  return A2L<64>::kOffMisc + static_cast<int>((sizeof(A5Misc) + 1023) / 1024 * 1024) + 1024;
}

__device__ long long g_attn5_trace[22][8][16];
__device__ long long g_attn5_cta[160][4];  // per CTA: globaltimer at start / end, clock64 at start / end (trace builds)
#define A5_TRACE(ev)                                                                                    \
  do {                                                                                                  \
    if constexpr (kTrace) {                                                                             \
      if (blockIdx.x == 0 && lane == 0 && k >= g_attn_trace_k0 && k < g_attn_trace_k0 + 8) g_attn5_trace[warp][k - g_attn_trace_k0][ev] = clock64(); \
    }                                                                                                   \
  } while (0)

// kVar: bit 0 staggered MMA issue order | bits 2-3 exp2 on the FMA pipe for every 4th / 3rd / 2nd pair | bit 4 trace
template <int kVar>
// 704 threads: the register file is allocated as for 768, i.e. 80 registers per thread (88 does not launch)
__global__ void __launch_bounds__(kA5Threads, 1)
attention_umma5_kernel(const __grid_constant__ CUtensorMap tm, const __nv_bfloat16* __restrict__ qkv,
                       __nv_bfloat16* __restrict__ out, int n_ch, int T, int heads, float scale_log2) {
  constexpr int HD = 64;
  constexpr bool kStagger = (kVar & 1) != 0;
  constexpr int kPolyMod = ((kVar >> 2) & 3) == 0 ? 0 : 5 - ((kVar >> 2) & 3);  // 0 | 4 | 3 | 2
  constexpr bool kTrace = (kVar & 16) != 0;
  using L = A2L<HD>;
  constexpr int kColO4 = 64, kColL4 = 192;
  extern __shared__ uint8_t smem_a5_raw[];
  uint8_t* smem = smem_a5_raw + ((1024u - (smem_u32(smem_a5_raw) & 1023u)) & 1023u);  // keeps the address space: LDS/STS
  A5Misc* mb = reinterpret_cast<A5Misc*>(smem + L::kOffMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * HD;
  const size_t row_stride = static_cast<size_t>(3) * d;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&mb->full_qk[s], 1);
      mbar_init(&mb->full_v[s], 1);
      mbar_init(&mb->empty_qk[s], 1 + 4 + 16);  // MMA commit + class warps + softmax warps
      mbar_init(&mb->empty_v[s], 1 + 4);        // MMA commit + class warps
      mbar_init(&mb->s_full[s], 1);
      mbar_init(&mb->p_full[s], 8);
      mbar_init(&mb->o_full[s], 1);
      mbar_init(&mb->tmem_free[s], 8);
      mbar_init(&mb->cls_done[s], 4);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 8 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem + L::kOffOnes)[i] = 0x3F803F80u;  // bf16 1.0 x2
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(&mb->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mb->tmem_slot;
  if constexpr (kTrace) {
    if (threadIdx.x == 0 && blockIdx.x < 160) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      g_attn5_cta[blockIdx.x][0] = gt;
      g_attn5_cta[blockIdx.x][2] = clock64();
    }
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int k = 0;
      for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
        const int s = k & 1;
        const uint32_t u = (k >> 1) & 1;
        const int crop = ch / heads, head = ch - crop * heads;
        const int row0 = crop * T + 1;
        uint8_t* st = smem + s * L::kQKStage;
        uint8_t* sv = smem + L::kOffV + s * L::kOp;
        mbar_wait(&mb->empty_qk[s], u ^ 1);
        mbar_arrive_expect_tx(&mb->full_qk[s], 2 * L::kOp);
        tma_load_2d(st, &tm, &mb->full_qk[s], head * HD, row0);
        tma_load_2d(st + L::kOp, &tm, &mb->full_qk[s], d + head * HD, row0);
        mbar_wait(&mb->empty_v[s], u ^ 1);
        mbar_arrive_expect_tx(&mb->full_v[s], L::kOp);
        tma_load_2d(sv, &tm, &mb->full_v[s], 2 * d + head * HD, row0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform control flow; the query
    // tile index is a compile-time constant in every tcgen05.mma operand, see v4)
    const uint32_t idesc_s = make_idesc_f16(128, 256, 1);
    const uint32_t idesc_o = make_idesc_f16(128, 64, 1) | (1u << 16);  // B (= V) is MN-major
    const uint32_t idesc_l = make_idesc_f16(128, 16, 1);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t ones_desc = make_sw128_kmajor_desc(smem_base + L::kOffOnes);
    auto issue_s = [&](auto wc, uint32_t sbase, uint64_t k_desc, uint32_t kp) {
      constexpr int w = decltype(wc)::value;
      const uint32_t treg = tmem + w * 256;
      const uint64_t q_desc = make_sw128_kmajor_desc(sbase + w * 16384);
      mbar_wait(&mb->tmem_free[w], kp ^ 1);  // both halves of group w have read the previous O out of the region
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(treg, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk != 0);
        umma_commit(&mb->s_full[w]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](auto wc, uint64_t v_desc0, uint32_t kp) {
      constexpr int w = decltype(wc)::value;
      const uint32_t treg = tmem + w * 256;
      mbar_wait(&mb->p_full[w], kp);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint32_t pcol = (kk < 8 ? 0u : 128u) + (kk & 7) * 8;  // P_0 at [0,64), P_1 at [128,192)
          umma_f16_ts(treg + kColO4, treg + pcol, v_desc0 + kk * 128, idesc_o, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const uint32_t pcol = (kk < 8 ? 0u : 128u) + (kk & 7) * 8;
          umma_f16_ts(treg + kColL4, treg + pcol, ones_desc + 2 * (kk & 3), idesc_l, kk != 0);
        }
        umma_commit(&mb->o_full[w]);
      }
      __syncwarp();
    };
    using W0 = std::integral_constant<int, 0>;
    using W1 = std::integral_constant<int, 1>;
    int k = 0;
    uint64_t v_desc_prev = 0;
    int s_prev = 0;
    uint32_t kp_prev = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t u = (k >> 1) & 1, kp = k & 1;
      const uint32_t sbase = smem_base + s * L::kQKStage;
      const uint32_t vbase = smem_base + L::kOffV + s * L::kOp;
      A5_TRACE(0);
      mbar_wait(&mb->full_qk[s], u);
      const uint64_t k_desc = make_sw128_kmajor_desc(sbase + L::kOp);
      A5_TRACE(1);
      issue_s(W0{}, sbase, k_desc, kp);
      A5_TRACE(2);
      if constexpr (kStagger) {
        if (k > 0) {  // tile 1 of the previous head: its softmax ran while tile 0's P·V, epilogue and this S were in flight
          issue_pv(W1{}, v_desc_prev, kp_prev);
          if (elect_one()) umma_commit(&mb->empty_v[s_prev]);
          __syncwarp();
        }
      }
      A5_TRACE(3);
      issue_s(W1{}, sbase, k_desc, kp);
      A5_TRACE(4);
      if (elect_one()) umma_commit(&mb->empty_qk[s]);
      __syncwarp();
      mbar_wait(&mb->full_v[s], u);
      const uint64_t v_desc0 = make_sw128_kmajor_desc(vbase);
      A5_TRACE(5);
      issue_pv(W0{}, v_desc0, kp);
      A5_TRACE(6);
      if constexpr (kStagger) {
        v_desc_prev = v_desc0;
        s_prev = s;
        kp_prev = kp;
      } else {
        issue_pv(W1{}, v_desc0, kp);
        if (elect_one()) umma_commit(&mb->empty_v[s]);
        __syncwarp();
      }
      A5_TRACE(7);
    }
    if constexpr (kStagger) {
      if (k > 0) {
        issue_pv(W1{}, v_desc_prev, kp_prev);
        if (elect_one()) umma_commit(&mb->empty_v[s_prev]);
        __syncwarp();
      }
    }
  } else if (warp < kA5ClsWarp0) {
    // ------------------------------------------------------------------ softmax: 2 groups x 2 column halves x 4 warps
    const int sw = warp - 2;
    const int w = sw >> 3;                  // group = query tile
    const int h = (sw >> 2) & 1;            // column half: keys [128h, 128h + 128)
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;      // query row in the tile = TMEM lane
    const uint32_t taddr = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sbase_col = 128u * h;    // this half's S columns; its P goes over them at [128h, 128h + 64)
    int k = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t u = (k >> 1) & 1, kp = k & 1;
      const int crop = ch / heads, head = ch - crop * heads;
      const int token = crop * T + 1 + w * 128 + r;

      // ---- own row, own 128 keys: partial max -> exchange with the other half -> P (bf16x2 packed) back into TMEM over S
      A5_TRACE(0);
      mbar_wait(&mb->s_full[w], kp);
      tc_fence_after();
      A5_TRACE(1);
      float m = -INFINITY;
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr + sbase_col, va);
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + sbase_col + (c + 1) * 32, vb);
#pragma unroll
        for (int e = 0; e < 32; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(va[e]), __uint_as_float(va[e + 1])));
        tmem_ld_wait();
        tmem_ld_32x32(taddr + sbase_col + ((c + 2) & 3) * 32, va);  // last iteration: chunk 0 again, for the second pass
#pragma unroll
        for (int e = 0; e < 32; e += 2) m = fmaxf(m, fmaxf(__uint_as_float(vb[e]), __uint_as_float(vb[e + 1])));
      }
      mb->mx[kp][w][h][r] = m;
      A5_TRACE(2);
      named_bar_sync(3 + w * 4 + quarter, 64);  // the two warps that share these 32 rows
      m = fmaxf(m, mb->mx[kp][w][h ^ 1][r]);
      A5_TRACE(3);
      const float ms = m * scale_log2;
      const float nms = -ms;
      auto emit_p = [&](const uint32_t (&v)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t0, t1;  // one FFMA2 per pair of scores
          ffma2(t0, t1, __uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]), scale_log2, scale_log2, nms, nms);
          if (kPolyMod != 0 && (e % (kPolyMod ? kPolyMod : 1)) == kPolyMod - 1) {
            float y0, y1;
            exp2_poly2(y0, y1, t0, t1);
            pk[e] = pack2(y0, y1);
          } else {
            pk[e] = pack2(ex2_ftz(t0), ex2_ftz(t1));
          }
        }
        tmem_st_32x16(taddr + sbase_col + c * 16, pk);  // columns already consumed by this thread
      };
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + sbase_col + (c + 1) * 32, vb);
        emit_p(va, c);
        tmem_ld_wait();
        if (c + 2 < 4) tmem_ld_32x32(taddr + sbase_col + (c + 2) * 32, va);
        emit_p(vb, c + 1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mb->p_full[w]);
      A5_TRACE(4);

      // ---- class-token KEY: p0 = exp2((q_r·k0 − m)·c), the dot product comes from the class warps
      mbar_wait(&mb->cls_done[s], u);
      A5_TRACE(5);
      const float p0 = ex2_ftz(fminf(fmaf(mb->s0[s][w * 128 + r], scale_log2, nms), 126.f));
      const float* v0 = mb->vec[s][2] + h * 32;

      // ---- own row, own 32 output columns: (O + p0·v0) / (L + p0) -> bf16
      A5_TRACE(8);
      mbar_wait(&mb->o_full[w], kp);
      tc_fence_after();
      A5_TRACE(9);
      const float lsum = __uint_as_float(tmem_ld_1(taddr + kColL4));
      {
        uint32_t v[32];
        tmem_ld_32x32(taddr + kColO4 + h * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&mb->tmem_free[w]);  // O and L are in registers: the region can take the next S
        const float inv = 1.0f / (lsum + p0);
        const float p0i = p0 * inv;
        __nv_bfloat16* orow = out + static_cast<size_t>(token) * d + head * HD + h * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 va4 = *reinterpret_cast<const float4*>(v0 + 8 * j);
          const float4 vb4 = *reinterpret_cast<const float4*>(v0 + 8 * j + 4);
          uint4 o4;
          float t[8], o[8];
          fmul2(t[0], t[1], va4.x, va4.y, p0i, p0i);
          fmul2(t[2], t[3], va4.z, va4.w, p0i, p0i);
          fmul2(t[4], t[5], vb4.x, vb4.y, p0i, p0i);
          fmul2(t[6], t[7], vb4.z, vb4.w, p0i, p0i);
#pragma unroll
          for (int e = 0; e < 8; e += 2)
            ffma2(o[e], o[e + 1], __uint_as_float(v[8 * j + e]), __uint_as_float(v[8 * j + e + 1]), inv, inv, t[e], t[e + 1]);
          o4.x = pack2(o[0], o[1]);
          o4.y = pack2(o[2], o[3]);
          o4.z = pack2(o[4], o[5]);
          o4.w = pack2(o[6], o[7]);
          *reinterpret_cast<uint4*>(orow + 8 * j) = o4;
        }
      }
      A5_TRACE(10);
      // s0 / v0 of stage s are no longer needed by this warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&mb->empty_qk[s]);
      A5_TRACE(11);
    }
  } else {
    // ------------------------------------------------------------------ class-token warps (128 threads, barrier 11)
    // Everything here goes through the MIO queue (LDS, SHFL) that the softmax warps keep full of MUFU.EX2, so a round
    // trip costs hundreds of cycles: loads are issued in few, wide, independent rounds (the vectors stay packed bf16 in
    // registers, a row's eight 16-byte chunks are requested together) and the max reduction uses CREDUX.
    const int t = static_cast<int>(threadIdx.x) - kA5ClsWarp0 * 32;
    const int cw = warp - kA5ClsWarp0;
    float* red = mb->red;
    float* pcls = mb->p_cls;
    uint32_t cq = 0, ck = 0, cv = 0;  // two bf16 each (threads 0..31)
    auto prefetch = [&](int ch) {
      if (t < HD / 2) {
        const int crop = ch / heads, head = ch - crop * heads;
        const uint32_t* cls_row = reinterpret_cast<const uint32_t*>(qkv + static_cast<size_t>(crop) * T * row_stride + head * HD);
        cq = __ldg(cls_row + t);
        ck = __ldg(cls_row + d / 2 + t);
        cv = __ldg(cls_row + d + t);
      }
    };
    // dots of rows `ra` and `rb` of a swizzled 128-byte-row tile with a packed bf16 vector (32 words in shared memory):
    // two rounds of 12 independent 16-byte loads (half the vector + half of either row), 48 registers in flight
    auto dot2_packed = [](const uint8_t* tile, int ra, int rb, const uint32_t* vec, float& da, float& db) {
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
      const uint8_t* pa = tile + ra * 128;
      const uint8_t* pb = tile + rb * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 v[4], xa[4], xb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = reinterpret_cast<const uint4*>(vec)[4 * half + j];
          xa[j] = *reinterpret_cast<const uint4*>(pa + (((4 * half + j) ^ (ra & 7)) << 4));
          xb[j] = *reinterpret_cast<const uint4*>(pb + (((4 * half + j) ^ (rb & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t vw[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          const uint32_t aw[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w};
          const uint32_t bw[4] = {xb[j].x, xb[j].y, xb[j].z, xb[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float vl = bf16_lo(vw[e]), vh = bf16_hi(vw[e]);
            ffma2(a0, a1, bf16_lo(aw[e]), bf16_hi(aw[e]), vl, vh, a0, a1);
            ffma2(b0, b1, bf16_lo(bw[e]), bf16_hi(bw[e]), vl, vh, b0, b1);
          }
        }
      }
      da = a0 + a1;
      db = b0 + b1;
    };
    if (static_cast<int>(blockIdx.x) < n_ch) prefetch(blockIdx.x);
    int k = 0;
    for (int ch = blockIdx.x; ch < n_ch; ch += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t u = (k >> 1) & 1;
      const int crop = ch / heads, head = ch - crop * heads;
      const int tok0 = crop * T;
      const uint8_t* st = smem + s * L::kQKStage;
      const uint8_t* sv = smem + L::kOffV + s * L::kOp;
      uint32_t* q0b = reinterpret_cast<uint32_t*>(mb->vec[s][0]);  // q0 / k0 packed bf16 (32 words each); v0 as fp32
      uint32_t* k0b = reinterpret_cast<uint32_t*>(mb->vec[s][1]);
      float* v0f = mb->vec[s][2];
      A5_TRACE(0);
      mbar_wait(&mb->full_qk[s], u);  // the stage (tiles, vec[s], s0[s]) was released by everyone who used it two heads ago
      A5_TRACE(1);
      if (t < HD / 2) {
        q0b[t] = cq;
        k0b[t] = ck;
        *reinterpret_cast<float2*>(v0f + 2 * t) = make_float2(bf16_lo(cv), bf16_hi(cv));
      }
      if (ch + static_cast<int>(gridDim.x) < n_ch) prefetch(ch + gridDim.x);
      named_bar_sync(11, 128);
      // class-token KEY: q_r·k0 for the 256 patch query rows -> the softmax warps
      float sq0, sq1;
      dot2_packed(st, t, t + 128, k0b, sq0, sq1);
      mb->s0[s][t] = sq0;
      mb->s0[s][t + 128] = sq1;
      // class key x class query (each warp redundantly): lane l takes the packed pair l
      float sc;
      {
        const uint32_t qa = q0b[lane], ka = k0b[lane];
        sc = fmaf(bf16_lo(qa), bf16_lo(ka), bf16_hi(qa) * bf16_hi(ka));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mb->cls_done[s]);  // release: s0[s] and vec[s] (written before barrier 11) are visible
      A5_TRACE(2);
      // class-token QUERY row: scores of the 256 patch keys (two per thread) and of the class key
      float sa0, sa1;
      dot2_packed(st + L::kOp, t, t + 128, q0b, sa0, sa1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      float cm;
      {
        const float mine = fmaxf(fmaxf(sa0, sa1), sc);
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(cm) : "f"(mine));
      }
      if (lane == 0) red[cw] = cm;
      named_bar_sync(11, 128);
      cm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * scale_log2;
      const float pa0 = ex2_ftz(fmaf(sa0, scale_log2, -cm));
      const float pa1 = ex2_ftz(fmaf(sa1, scale_log2, -cm));
      const float pc = ex2_ftz(fmaf(sc, scale_log2, -cm));
      pcls[t] = pa0;
      pcls[t + 128] = pa1;
      float psum = pa0 + pa1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      if (lane == 0) red[4 + cw] = psum;
      named_bar_sync(11, 128);
      A5_TRACE(3);
      // O_cls = P_cls·V: thread = (16-key slice ks, 8-column chunk cc), one 16-byte read per key, 8 keys per round
      mbar_wait(&mb->full_v[s], u);
      const int cc = t & 7, ks = t >> 3;
      const float* pp = pcls + ks * 16;
      const uint8_t* base = sv + (ks * 16) * 128;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float4 pA = *reinterpret_cast<const float4*>(pp + 8 * half);
        const float4 pB = *reinterpret_cast<const float4*>(pp + 8 * half + 4);
        const float pk8[8] = {pA.x, pA.y, pA.z, pA.w, pB.x, pB.y, pB.z, pB.w};
        // ks*16 + key has the same low 3 bits as key, so the swizzle phase depends on `key` only
        uint4 row[8];
#pragma unroll
        for (int key = 0; key < 8; ++key) row[key] = *reinterpret_cast<const uint4*>(base + (8 * half + key) * 128 + ((cc ^ key) << 4));
#pragma unroll
        for (int key = 0; key < 8; ++key) {
          const uint4 a = row[key];
          const float pkey = pk8[key];
          ffma2(acc[0], acc[1], bf16_lo(a.x), bf16_hi(a.x), pkey, pkey, acc[0], acc[1]);
          ffma2(acc[2], acc[3], bf16_lo(a.y), bf16_hi(a.y), pkey, pkey, acc[2], acc[3]);
          ffma2(acc[4], acc[5], bf16_lo(a.z), bf16_hi(a.z), pkey, pkey, acc[4], acc[5]);
          ffma2(acc[6], acc[7], bf16_lo(a.w), bf16_hi(a.w), pkey, pkey, acc[6], acc[7]);
        }
      }
      *reinterpret_cast<float4*>(&mb->part[ks][cc * 8]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(&mb->part[ks][cc * 8 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      named_bar_sync(11, 128);
      // every Q / K / V read of the class warps is done
      if (lane == 0) {
        mbar_arrive(&mb->empty_qk[s]);
        mbar_arrive(&mb->empty_v[s]);
      }
      if (t < HD) {
        const float cls_l = ((red[4] + red[5]) + (red[6] + red[7])) + pc;
        float o = pc * v0f[t];
#pragma unroll
        for (int i = 0; i < 16; ++i) o += mb->part[i][t];
        out[static_cast<size_t>(tok0) * d + head * HD + t] = __float2bfloat16_rn(o / cls_l);
      }
      A5_TRACE(4);
      named_bar_sync(11, 128);  // red / p_cls / part are rewritten next iteration
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kTrace) {
    if (threadIdx.x == 0 && blockIdx.x < 160) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      g_attn5_cta[blockIdx.x][1] = gt;
      g_attn5_cta[blockIdx.x][3] = clock64();
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

constexpr int kA5DefaultVar = 1;  // staggered MMA issue order, all exp2 on the XU pipe

static int attention_umma5_launch(const void* qkv, void* out, int n, int T, int heads, int var, cudaStream_t stream) {
  constexpr int HD = 64;
  const int d = heads * HD;
  CUtensorMap tm;
  B2C_TRY(make_tmap_2d(&tm, qkv, static_cast<uint64_t>(n) * T, 3ull * d, 3ull * d * 2, 256, 1));
  constexpr int smem_bytes = a5_smem_bytes();
  static_assert(smem_bytes <= 227 * 1024, "attention v5 shared memory exceeds 227 KB");
  using Kern = void (*)(const CUtensorMap, const __nv_bfloat16*, __nv_bfloat16*, int, int, int, float);
  Kern kern = nullptr;
  switch (var) {
    case 0: kern = attention_umma5_kernel<0>; break;    // both tiles' S issued back to back (v4's order)
    case 1: kern = attention_umma5_kernel<1>; break;    // default
    case 5: kern = attention_umma5_kernel<5>; break;    // + every 4th pair's exp2 on the FMA pipe
    case 17: kern = attention_umma5_kernel<17>; break;  // default + phase trace
    default: return set_error(B2C_ERR_ARG, "attention v5: variant %d is not built", var);
  }
  B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  const int n_ch = n * heads;
  const float scale_log2 = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64)
  kern<<<n_ch < sms ? n_ch : sms, kA5Threads, smem_bytes, stream>>>(
      tm, static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n_ch, T, heads, scale_log2);
  B2C_POST_LAUNCH("attention_umma5_kernel");
  return 0;
}

static int g_attn5_var_override = -1;  // development (b2c_debug_set_attn5): < 0 = follow B2C_ATTN5_VAR (default kA5DefaultVar), >= 0 = that variant

template <int HD>
static int attention_launch_hd(const void* qkv, void* out, int n, int T, int heads, cudaStream_t stream) {
  const int Tp = (T + 15) / 16 * 16;
  const size_t smem = 2ull * Tp * (HD + 8) * sizeof(__nv_bfloat16);
  B2C_REQUIRE(smem <= 227 * 1024, "attention: T=%d does not fit shared memory", T);
  auto kern = attention_kernel<HD>;
  static PerDeviceMax smem_set;
  if (smem_set.raise(static_cast<long long>(smem))) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  kern<<<static_cast<unsigned>(n) * heads, kAttnThreads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), T, Tp, heads, scale_log2);
  B2C_POST_LAUNCH("attention_kernel");
  return 0;
}

int attention_cls_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream) {
  B2C_REQUIRE(n > 0 && T > 0 && heads > 0 && hd == 64, "attention_cls: head dim %d unsupported (64)", hd);
  const float scale_log2 = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64)
  attention_cls_kernel<<<(n * heads + 3) / 4, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                                                n * heads, T, heads, scale_log2);
  B2C_POST_LAUNCH("attention_cls_kernel");
  return 0;
}

int attention_launch(const void* qkv, void* out, int n, int T, int heads, int hd, cudaStream_t stream) {
  B2C_REQUIRE(n > 0 && T > 0 && heads > 0, "attention: empty problem");
  // B2C_ATTN=legacy forces the mma.sync kernel (A/B reference); B2C_ATTN5_VAR / b2c_debug_set_attn5 pick a v5 variant
  static const bool legacy = [] { const char* e = getenv("B2C_ATTN"); return e && e[0] == 'l'; }();
  static const int v5env = [] { const char* e = getenv("B2C_ATTN5_VAR"); return e ? atoi(e) : kA5DefaultVar; }();
  const int v5var = g_attn5_var_override >= 0 ? g_attn5_var_override : v5env;
  if (!legacy) {
    if (hd == 64 && T == kAuKeys + 1) return attention_umma5_launch(qkv, out, n, T, heads, v5var, stream);
    if (hd == 80 && T == kAuKeys + 1) return attention_umma2_launch<80>(qkv, out, n, T, heads, stream);
    if (hd == 64 && T > kAuKeys + 1 && (T - 1) % kA3KB == 0 && (T - 1) / kA3KB <= kA3MaxNB)
      return attention_umma6_launch(qkv, out, n, T, heads, stream);
  }
  if (hd == 64) return attention_launch_hd<64>(qkv, out, n, T, heads, stream);
  if (hd == 80) return attention_launch_hd<80>(qkv, out, n, T, heads, stream);
  return set_error(B2C_ERR_ARG, "attention: head dim %d unsupported (64 or 80)", hd);
}

}  // namespace b2c

// development only (not part of include/b2c.h): phase trace of attention_umma5_kernel<kVar | 16> (tools/attn_trace.py)
extern "C" int b2c_debug_attn_trace_start(int k0) {
  using namespace b2c;
  B2C_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace_k0, &k0, sizeof(int)));
  return 0;
}
extern "C" int b2c_debug_set_attn5(int var) {
  b2c::g_attn5_var_override = var;
  return 0;
}
extern "C" int b2c_debug_attn5_cta(void* out, size_t bytes) {
  using namespace b2c;
  B2C_REQUIRE(out && bytes == sizeof(g_attn5_cta), "b2c_debug_attn5_cta: need %zu bytes", sizeof(g_attn5_cta));
  B2C_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_attn5_cta, bytes));
  return 0;
}
extern "C" int b2c_debug_attn5_trace(void* out, size_t bytes) {
  using namespace b2c;
  B2C_REQUIRE(out && bytes == sizeof(g_attn5_trace), "b2c_debug_attn5_trace: need %zu bytes", sizeof(g_attn5_trace));
  B2C_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_attn5_trace, bytes));
  return 0;
}

extern "C" int b2c_attention_bf16(const void* qkv, void* out, int n, int T, int heads, int hd, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(qkv && out, "b2c_attention_bf16: null pointer");
  return attention_launch(qkv, out, n, T, heads, hd, static_cast<cudaStream_t>(stream));
}
