// b2c_similar.cu — K11: similarity-search variants on stored embeddings.
//
//   * context scores: distance of every stored embedding to ONE context vector — the per-sample
//     compute_distance of tools/find_similar_imgs.py:88-94 ((1 - cos)/2 or ||c - x + 1e-6||_2) for all N rows at once.
//     This is a matrix-vector product: N*E*4 bytes are read once and 2*N*E FLOPs are done on them (0.5 FLOP/byte), so
//     the kernel is HBM-bound by three orders of magnitude and stays on the SIMT path — one warp per row, 16-byte
//     coalesced loads, the context vector in shared memory.  It is NOT reshaped into a GEMM.
//   * top-k smallest: the topN bookkeeping of tools/find_similar_imgs.py:67-85 as an exact radix select over the
//     composite key (orderable float bits << index bits | index): unique keys, ties go to the smaller index, result
//     sorted ascending.  All passes read the 4 B/row score vector (L2-resident), not the embeddings.
//   * greedy diversity ordering (_3_label_images.py:128-177): the reference recomputes cos(selected set, sample) every
//     step (O(i*S*E)); here max-similarity-to-the-selected-set is kept for ALL rows and updated with one context-score
//     launch per step (combine = max), so a step is one streaming pass plus an argmin over the S sampled rows.
#include <cuda_fp16.h>
#include <math.h>

#include "b2c_launch.h"

namespace b2c {

constexpr int kSimThreads = 256;
constexpr int kSimMaxE = 8192;

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int kN = 4;
  using Raw = float4;
  __device__ static void unpack(const Raw& r, float* f) {
    f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
  }
};
template <>
struct Vec<__half> {
  static constexpr int kN = 8;
  using Raw = uint4;
  __device__ static void unpack(const Raw& r, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(h[i]);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row, grid-stride.  VEC: 16-byte loads (requires aligned base, E and row_stride multiples of Vec::kN).
template <typename T, bool VEC>
__global__ void __launch_bounds__(kSimThreads)
context_scores_kernel(const T* __restrict__ emb, long long n, int E, long long row_stride, const float* __restrict__ ctx,
                      const int* __restrict__ ctx_row, int measure, int combine, const unsigned char* __restrict__ skip,
                      float* __restrict__ out) {
  extern __shared__ float s_ctx[];
  __shared__ float s_red[kSimThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage the context vector (given directly, or as a row of emb chosen on the device by an earlier kernel)
  float part = 0.f;
  if (ctx != nullptr) {
    for (int k = threadIdx.x; k < E; k += kSimThreads) {
      const float v = ctx[k];
      s_ctx[k] = v;
      part = fmaf(v, v, part);
    }
  } else {
    const T* r = emb + static_cast<long long>(*ctx_row) * row_stride;
    for (int k = threadIdx.x; k < E; k += kSimThreads) {
      const float v = static_cast<float>(r[k]);
      s_ctx[k] = v;
      part = fmaf(v, v, part);
    }
  }
  part = warp_sum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  float cc = 0.f;
#pragma unroll
  for (int w = 0; w < kSimThreads / 32; ++w) cc += s_red[w];
  const float c_norm = fmaxf(sqrtf(cc), 1e-8f);  // torch.cosine_similarity clamps each norm at eps = 1e-8

  const long long warps_total = static_cast<long long>(gridDim.x) * (kSimThreads / 32);
  for (long long row = static_cast<long long>(blockIdx.x) * (kSimThreads / 32) + warp; row < n; row += warps_total) {
    if (skip != nullptr && skip[row]) {
      if (lane == 0) out[row] = INFINITY;
      continue;
    }
    const T* r = emb + row * row_stride;
    float dot = 0.f, xx = 0.f, dd = 0.f;
    if constexpr (VEC) {
      constexpr int V = Vec<T>::kN;
      using Raw = typename Vec<T>::Raw;
      const Raw* rv = reinterpret_cast<const Raw*>(r);
      const int nv = E / V;
      // batches of 4 independent 16-byte loads per lane keep ~64 B x 2048 threads in flight per SM (Little's law for
      // ~6.5 TB/s needs > 35 KB per SM); the FMAs of a batch start only after its loads have been issued
      for (int k0 = lane; k0 < nv; k0 += 128) {
        Raw raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k0 + 32 * u < nv) raw[u] = __ldcs(rv + k0 + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k0 + 32 * u < nv) {
            float f[V];
            Vec<T>::unpack(raw[u], f);
            const float* cs = s_ctx + (k0 + 32 * u) * V;
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const float c = cs[e];
              dot = fmaf(c, f[e], dot);
              xx = fmaf(f[e], f[e], xx);
              const float d = (c - f[e]) + 1e-6f;  // F.pairwise_distance adds eps to the difference
              dd = fmaf(d, d, dd);
            }
          }
        }
      }
    } else {
      for (int k = lane; k < E; k += 32) {
        const float x = static_cast<float>(r[k]);
        const float c = s_ctx[k];
        dot = fmaf(c, x, dot);
        xx = fmaf(x, x, xx);
        const float d = (c - x) + 1e-6f;
        dd = fmaf(d, d, dd);
      }
    }
    dot = warp_sum(dot);
    xx = warp_sum(xx);
    dd = warp_sum(dd);
    if (lane == 0) {
      float v;
      if (measure == B2C_MEASURE_L2) {
        v = sqrtf(dd);
      } else {
        const float cs = dot / (c_norm * fmaxf(sqrtf(xx), 1e-8f));
        v = measure == B2C_MEASURE_COSINE_DIST ? (1.0f - cs) * 0.5f : cs;
      }
      out[row] = combine == B2C_COMBINE_MAX ? fmaxf(out[row], v) : v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ top-k (radix select)
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;
constexpr int kTopkMax = 4096;

struct TopkState {
  unsigned long long prefix;  // value of the key bits fixed so far
  unsigned long long mask;    // which bits are fixed
  unsigned int k_rem;         // rank (1-based) of the wanted key among the keys matching the prefix
  unsigned int taken;         // slots handed out by the gather pass
};

__device__ __forceinline__ unsigned long long topk_key(float v, long long idx, int idx_bits) {
  v += 0.0f;  // -0 -> +0
  unsigned int u = __float_as_uint(v);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // monotone float -> unsigned
  return (static_cast<unsigned long long>(u) << idx_bits) | static_cast<unsigned long long>(idx);
}

__global__ void topk_init_kernel(TopkState* st, unsigned int* hist, int k) {
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) {
    st->prefix = 0;
    st->mask = 0;
    st->k_rem = static_cast<unsigned int>(k);
    st->taken = 0;
  }
}

__global__ void __launch_bounds__(256)
topk_hist_kernel(const float* __restrict__ scores, long long n, int idx_bits, int shift, int width,
                 const TopkState* __restrict__ st, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned long long prefix = st->prefix, mask = st->mask;
  const unsigned int dmask = (1u << width) - 1u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long key = topk_key(scores[i], i, idx_bits);
    if ((key & mask) == prefix) atomicAdd(&sh[static_cast<unsigned int>(key >> shift) & dmask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one block of 1024 threads: find the digit that holds the k_rem-th key, fix it in the prefix, clear the histogram
__global__ void __launch_bounds__(1024) topk_scan_kernel(TopkState* st, unsigned int* hist, int shift, int width) {
  __shared__ unsigned int cum[kBins];
  const int t = threadIdx.x;
  const unsigned int a = hist[2 * t], b = hist[2 * t + 1];
  cum[2 * t] = a;
  cum[2 * t + 1] = a + b;
  __syncthreads();
  // inclusive scan over the pair sums (Hillis-Steele on 1024 values held at the odd slots)
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned int add = 0;
    if (t >= off) add = cum[2 * (t - off) + 1];
    __syncthreads();
    if (t >= off) {
      cum[2 * t + 1] += add;
    }
    __syncthreads();
  }
  const unsigned int before_pair = t ? cum[2 * t - 1] : 0u;
  cum[2 * t] = before_pair + a;
  __syncthreads();
  const unsigned int k = st->k_rem;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int d = 2 * t + e;
    const unsigned int incl = cum[d], excl = d ? cum[d - 1] : 0u;
    if (excl < k && k <= incl) {  // exactly one digit satisfies this
      st->prefix |= static_cast<unsigned long long>(d) << shift;
      st->mask |= static_cast<unsigned long long>((1u << width) - 1u) << shift;
      st->k_rem = k - excl;
    }
  }
  hist[2 * t] = 0;
  hist[2 * t + 1] = 0;
}

__global__ void __launch_bounds__(256)
topk_gather_kernel(const float* __restrict__ scores, long long n, int idx_bits, TopkState* st,
                   unsigned long long* __restrict__ keys, int k) {
  const unsigned long long kth = st->prefix;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long key = topk_key(scores[i], i, idx_bits);
    if (key <= kth) {
      const unsigned int slot = atomicAdd(&st->taken, 1u);
      if (slot < static_cast<unsigned int>(k)) keys[slot] = key;
    }
  }
}

// single block: bitonic sort of the k gathered keys (padded to a power of two), then split key -> (value, index)
__global__ void __launch_bounds__(1024)
topk_sort_kernel(const unsigned long long* __restrict__ keys, int k, int idx_bits, int* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  extern __shared__ unsigned long long sk[];
  int P = 1;
  while (P < k) P <<= 1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sk[i] = i < k ? keys[i] : ~0ull;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < P / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long x = sk[lo], y = sk[hi];
        if ((x > y) == up) {
          sk[lo] = y;
          sk[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = sk[i];
    unsigned int u = static_cast<unsigned int>(key >> idx_bits);
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    out_idx[i] = static_cast<int>(key & ((1ull << idx_bits) - 1ull));
    out_val[i] = __uint_as_float(u);
  }
}

// ------------------------------------------------------------------------------------------------ diversity ordering
// argmin over the S sampled rows of maxsim (first position wins ties, like torch.argmin on CPU); one block.
__global__ void __launch_bounds__(256)
diversity_pick_kernel(const float* __restrict__ maxsim, const int* __restrict__ sample, int S, int* __restrict__ picked) {
  __shared__ float sv[256];
  __shared__ int sp[256];
  float best = INFINITY;
  int pos = 0x7fffffff;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = maxsim[sample[i]];
    if (v < best || (v == best && i < pos) || pos == 0x7fffffff) {
      best = v;
      pos = i;
    }
  }
  sv[threadIdx.x] = best;
  sp[threadIdx.x] = pos;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const float v = sv[threadIdx.x + off];
      const int p = sp[threadIdx.x + off];
      if (p != 0x7fffffff && (sp[threadIdx.x] == 0x7fffffff || v < sv[threadIdx.x] ||
                              (v == sv[threadIdx.x] && p < sp[threadIdx.x]))) {
        sv[threadIdx.x] = v;
        sp[threadIdx.x] = p;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *picked = sample[sp[0]];
}

__global__ void set_int_kernel(int* dst, int v) { *dst = v; }

static int scores_launch(const void* emb, int dtype, int64_t n, int E, int64_t row_stride, const float* ctx,
                         const int32_t* ctx_row, int measure, int combine, const uint8_t* skip, float* out,
                         cudaStream_t st) {
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  const long long blocks_needed = (n + kSimThreads / 32 - 1) / (kSimThreads / 32);
  const unsigned grid = static_cast<unsigned>(blocks_needed < 8ll * sms ? blocks_needed : 8ll * sms);
  const size_t smem = static_cast<size_t>(E) * sizeof(float);
  const uintptr_t base = reinterpret_cast<uintptr_t>(emb);
  if (dtype == B2C_F32) {
    const bool vec = (E % 4 == 0) && (row_stride % 4 == 0) && (base % 16 == 0);
    auto k = vec ? context_scores_kernel<float, true> : context_scores_kernel<float, false>;
    k<<<grid, kSimThreads, smem, st>>>(static_cast<const float*>(emb), n, E, row_stride, ctx, ctx_row, measure, combine,
                                       skip, out);
  } else {
    const bool vec = (E % 8 == 0) && (row_stride % 8 == 0) && (base % 16 == 0);
    auto k = vec ? context_scores_kernel<__half, true> : context_scores_kernel<__half, false>;
    k<<<grid, kSimThreads, smem, st>>>(static_cast<const __half*>(emb), n, E, row_stride, ctx, ctx_row, measure,
                                       combine, skip, out);
  }
  B2C_POST_LAUNCH("context_scores_kernel");
  return 0;
}

static int check_emb_args(const char* who, const void* emb, int dtype, int64_t n, int E, int64_t row_stride) {
  B2C_REQUIRE(emb != nullptr, "%s: null embeddings", who);
  B2C_REQUIRE(dtype == B2C_F32 || dtype == B2C_F16, "%s: dtype %d (want f32 or f16)", who, dtype);
  B2C_REQUIRE(E > 0 && E <= kSimMaxE, "%s: E=%d out of range (1..%d)", who, E, kSimMaxE);
  B2C_REQUIRE(row_stride >= E, "%s: row_stride %lld < E", who, (long long)row_stride);
  B2C_REQUIRE(n >= 0 && n < (1ll << 31), "%s: n=%lld out of range", who, (long long)n);
  return 0;
}

}  // namespace b2c

extern "C" int b2c_context_scores(const void* emb, int dtype, int64_t n, int E, int64_t row_stride, const float* ctx,
                                  const int32_t* ctx_row, int measure, int combine, const uint8_t* skip, float* out,
                                  b2c_stream stream) {
  using namespace b2c;
  B2C_TRY(check_emb_args("b2c_context_scores", emb, dtype, n, E, row_stride));
  B2C_REQUIRE((ctx != nullptr) != (ctx_row != nullptr), "b2c_context_scores: give exactly one of ctx / ctx_row");
  B2C_REQUIRE(measure >= B2C_MEASURE_COSINE_DIST && measure <= B2C_MEASURE_COSINE_SIM, "b2c_context_scores: measure %d", measure);
  B2C_REQUIRE(combine == B2C_COMBINE_STORE || combine == B2C_COMBINE_MAX, "b2c_context_scores: combine %d", combine);
  B2C_REQUIRE(out != nullptr, "b2c_context_scores: null output");
  if (n == 0) return 0;
  return scores_launch(emb, dtype, n, E, row_stride, ctx, ctx_row, measure, combine, skip, out,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int b2c_topk_workspace_bytes(int k, size_t* bytes) {
  using namespace b2c;
  B2C_REQUIRE(bytes != nullptr, "b2c_topk_workspace_bytes: null pointer");
  B2C_REQUIRE(k > 0 && k <= kTopkMax, "b2c_topk: k=%d out of range (1..%d)", k, kTopkMax);
  *bytes = 256 + kBins * sizeof(unsigned int) + static_cast<size_t>(kTopkMax) * sizeof(unsigned long long);
  return 0;
}

extern "C" int b2c_topk_smallest(const float* scores, int64_t n, int k, int32_t* out_idx, float* out_val, void* ws,
                                 size_t ws_bytes, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(scores && out_idx && out_val && ws, "b2c_topk_smallest: null pointer");
  B2C_REQUIRE(k > 0 && k <= kTopkMax, "b2c_topk_smallest: k=%d out of range (1..%d)", k, kTopkMax);
  B2C_REQUIRE(n >= k && n < (1ll << 31), "b2c_topk_smallest: need k <= n < 2^31 (n=%lld, k=%d)", (long long)n, k);
  size_t need = 0;
  B2C_TRY(b2c_topk_workspace_bytes(k, &need));
  if (ws_bytes < need) return set_error(B2C_ERR_WORKSPACE, "b2c_topk_smallest: workspace %zu < %zu", ws_bytes, need);
  B2C_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 16 == 0, "b2c_topk_smallest: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");
  auto* state = static_cast<TopkState*>(ws);
  auto* hist = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 256);
  auto* keys = reinterpret_cast<unsigned long long*>(static_cast<char*>(ws) + 256 + kBins * sizeof(unsigned int));
  int idx_bits = 1;
  while ((1ll << idx_bits) < n) ++idx_bits;
  const int total_bits = 32 + idx_bits;
  const long long blocks_needed = (n + 255) / 256;
  const unsigned grid = static_cast<unsigned>(blocks_needed < 4ll * sms ? blocks_needed : 4ll * sms);
  topk_init_kernel<<<1, 256, 0, st>>>(state, hist, k);
  B2C_POST_LAUNCH("topk_init_kernel");
  for (int hi = total_bits; hi > 0;) {
    const int shift = hi > kDigitBits ? hi - kDigitBits : 0;
    const int width = hi - shift;
    topk_hist_kernel<<<grid, 256, 0, st>>>(scores, n, idx_bits, shift, width, state, hist);
    B2C_POST_LAUNCH("topk_hist_kernel");
    topk_scan_kernel<<<1, 1024, 0, st>>>(state, hist, shift, width);
    B2C_POST_LAUNCH("topk_scan_kernel");
    hi = shift;
  }
  topk_gather_kernel<<<grid, 256, 0, st>>>(scores, n, idx_bits, state, keys, k);
  B2C_POST_LAUNCH("topk_gather_kernel");
  int P = 1;
  while (P < k) P <<= 1;
  topk_sort_kernel<<<1, 1024, static_cast<size_t>(P) * sizeof(unsigned long long), st>>>(keys, k, idx_bits, out_idx, out_val);
  B2C_POST_LAUNCH("topk_sort_kernel");
  return 0;
}

extern "C" int b2c_diversity_order(const void* emb, int dtype, int64_t n, int E, int64_t row_stride, int32_t first_row,
                                   const int32_t* samples, int steps, int S, float* maxsim, int32_t* order,
                                   b2c_stream stream) {
  using namespace b2c;
  B2C_TRY(check_emb_args("b2c_diversity_order", emb, dtype, n, E, row_stride));
  B2C_REQUIRE(samples && maxsim && order, "b2c_diversity_order: null pointer");
  B2C_REQUIRE(n > 0 && first_row >= 0 && first_row < n, "b2c_diversity_order: first_row %d outside [0,%lld)", first_row, (long long)n);
  B2C_REQUIRE(steps >= 0 && S > 0, "b2c_diversity_order: steps=%d S=%d", steps, S);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // order[0] = first_row; maxsim = cos(e_first, .)
  set_int_kernel<<<1, 1, 0, st>>>(order, first_row);
  B2C_POST_LAUNCH("set_int_kernel");
  B2C_TRY(scores_launch(emb, dtype, n, E, row_stride, nullptr, order, B2C_MEASURE_COSINE_SIM, B2C_COMBINE_STORE, nullptr,
                        maxsim, st));
  for (int s = 0; s < steps; ++s) {
    diversity_pick_kernel<<<1, 256, 0, st>>>(maxsim, samples + static_cast<size_t>(s) * S, S, order + s + 1);
    B2C_POST_LAUNCH("diversity_pick_kernel");
    B2C_TRY(scores_launch(emb, dtype, n, E, row_stride, nullptr, order + s + 1, B2C_MEASURE_COSINE_SIM, B2C_COMBINE_MAX,
                          nullptr, maxsim, st));
  }
  return 0;
}
