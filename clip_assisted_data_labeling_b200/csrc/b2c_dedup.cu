// b2c_dedup.cu — K9: all-to-all cosine-similarity duplicate search without materialising S.
// Replaces the body of find_near_duplicates (_2_remove_duplicates.py:63-80):
//   E / ||E|| (67)  ->  S = E·Eᵀ (69)  ->  where(triu(S, 1) > thr) (74)  ->  S[i, j] per pair (80)
// The similarity tiles (128 rows x 256 columns, K = E) run on the same tcgen05/TMA pipeline as the
// ViT GEMMs with fp16 operands like the reference (_2:38) and fp32 accumulation; the threshold test
// and the (i, j, sim) emission are the epilogue, so only upper-triangle tiles are computed and the
// N x N matrix never exists.  Work is issued in bands of 16 row blocks so the band's A tiles stay
// L2-resident while the B tiles stream through once per band.
#include <cuda_fp16.h>

#include "b2c_launch.h"
#include <stdlib.h>

#include "b2c_umma_pipeline.cuh"
#include "b2c_umma_pipeline2.cuh"

namespace b2c {

constexpr int kDedupBandBlocks = 16;  // 128-row blocks per band (2048 rows, 3 MB of fp16 x 768)

struct DedupParams {
  int num_tiles;
  int num_tiles2;  // CTA-pair kernel: 256 x 256 tiles in this band
  int k_blocks;
  int GI;   // row blocks in this band
  int bi0;  // first row block of the band (units of kBM rows)
  int bj0;  // first column block that can hold j > i for this band (units of kBN rows)
  long long n_total, row_begin, row_end;
  long long col_begin, col_end;  // only pairs with col_begin <= j < col_end (b2c_dedup_pairs_block)
  float thr_quick;  // necessary condition for a hit, checked on every accumulator element: sgn * acc > thr_quick
  float sgn;        // +1 (similarity above a threshold) | -1 (B2C_CMP_EUCLID: distance above a threshold)
  float thr;
  int mode;
  b2c_pair* out;
  unsigned long long capacity;
  unsigned long long* count;
};

struct DedupPolicy {
  using Params = DedupParams;
  static constexpr int kStore = kStoreDirect;  // the epilogue emits pairs itself; nothing is stored as a tile
  static constexpr bool kLnFold = false;
  static constexpr int kEpiWarps = 4;
  static constexpr int kRmwRing = 1;
  __device__ static __forceinline__ void transform(const Params&, int, float (&)[32]) {}

  // consecutive tiles walk down the band's row blocks for one column block: the CTAs running at the
  // same time share B tiles and the band's A tiles.
  __device__ static __forceinline__ bool tile(const Params& p, int t, int& a_row, int& b_row) {
    const int cj = t / p.GI;
    const int ci = t - cj * p.GI;
    a_row = (p.bi0 + ci) * kBM;
    b_row = (p.bj0 + cj) * kBN;
    return b_row + (kBN - 1) > a_row;  // tile holds at least one j > i
  }

  // CTA-pair kernel: the band is GI/2 blocks of 256 rows; tile (ci, cj) -> rows (bi0/2 + ci)*256, cols (bj0 + cj)*256
  __device__ static __forceinline__ bool tile2(const Params& p, int t, int& m_row, int& n_row) {
    const int g2 = (p.GI + 1) >> 1;
    const int cj = t / g2;
    const int ci = t - cj * g2;
    m_row = p.bi0 * kBM + ci * 2 * kBM;
    n_row = (p.bj0 + cj) * kBN;
    return n_row + (kBN - 1) > m_row;
  }

  __device__ static __forceinline__ bool passes(const Params& p, float v) {
    if (p.mode == B2C_CMP_REF_FP16)
      // the reference thresholds its fp16 similarity matrix: fp16(S) > fp16(thr)
      return __half2float(__float2half_rn(v)) > p.thr;  // p.thr already holds float(fp16(threshold))
    if (p.mode == B2C_CMP_EUCLID) return euclid(v) > p.thr;
    return v > p.thr;
  }
  // distance of two unit vectors from their cosine (the reference's sim_type='euclidean': cdist of the normalised rows)
  __device__ static __forceinline__ float euclid(float cosine) { return sqrtf(fmaxf(2.0f - 2.0f * cosine, 0.0f)); }

  __device__ static __forceinline__ void epilogue(const Params& p, int a_row, int b_row, int row_in_tile, int col0,
                                                  const uint32_t (&acc)[32]) {
    // thr_quick is a necessary condition on the raw accumulator; for the distance mode the sense is reversed
    // (dist > thr  <=>  cos < 1 - thr^2 / 2) and sgn = -1 turns it into the same comparison
    bool any = false;
    if (p.sgn > 0.f) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) any |= __uint_as_float(acc[jj]) > p.thr_quick;
    } else {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) any |= __uint_as_float(acc[jj]) < -p.thr_quick;
    }
    if (!any) return;
    const long long i = static_cast<long long>(a_row) + row_in_tile;
    if (i < p.row_begin || i >= p.row_end) return;
    const long long j0 = static_cast<long long>(b_row) + col0;
#pragma unroll  // static indices keep the accumulator chunk in registers; the body is a rarely-taken branch
    for (int jj = 0; jj < 32; ++jj) {
      const float v = __uint_as_float(acc[jj]);
      const long long j = j0 + jj;
      if (p.sgn * v > p.thr_quick && j > i && j >= p.col_begin && j < p.col_end && passes(p, v)) {
        const unsigned long long slot = atomicAdd(p.count, 1ull);
        if (slot < p.capacity) {
          b2c_pair pr;
          pr.i = static_cast<int32_t>(i);
          pr.j = static_cast<int32_t>(j);
          pr.sim = p.mode == B2C_CMP_EUCLID ? euclid(v) : v;
          p.out[slot] = pr;
        }
      }
    }
  }
};

// one warp per row: out = in / ||in||_2 (fp32 math), zero-padded to E_pad columns
template <typename T>
__global__ void normalize_rows_kernel(const T* __restrict__ in, __half* __restrict__ out, long long n, int E, int E_pad) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const T* r = in + row * E;
  float s = 0.f;
  for (int k = lane; k < E; k += 32) {
    const float v = static_cast<float>(r[k]);
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / sqrtf(s);  // zero rows become NaN exactly like the reference's 0/0
  __half* o = out + row * E_pad;
  for (int k = lane; k < E_pad; k += 32) o[k] = k < E ? __float2half_rn(static_cast<float>(r[k]) * inv) : __float2half_rn(0.f);
}

}  // namespace b2c

extern "C" int b2c_normalize_rows_f16(const void* in, int in_dtype, int64_t n, int E, void* out_f16, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(in && out_f16, "b2c_normalize_rows_f16: null pointer");
  B2C_REQUIRE(E > 0, "b2c_normalize_rows_f16: E must be positive");
  if (n <= 0) return 0;
  const int E_pad = (E + 63) / 64 * 64;
  const unsigned grid = static_cast<unsigned>((n + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == B2C_F32)
    normalize_rows_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), static_cast<__half*>(out_f16), n, E, E_pad);
  else if (in_dtype == B2C_F16)
    normalize_rows_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(in), static_cast<__half*>(out_f16), n, E, E_pad);
  else
    return set_error(B2C_ERR_ARG, "b2c_normalize_rows_f16: unsupported dtype %d", in_dtype);
  B2C_POST_LAUNCH("normalize_rows_kernel");
  return 0;
}

extern "C" int b2c_dedup_pairs(const void* emb_f16, int64_t n_total, int E_pad, int64_t row_begin, int64_t row_end,
                               float threshold, int compare_mode, b2c_pair* out, unsigned long long capacity,
                               unsigned long long* count, b2c_stream stream) {
  return b2c_dedup_pairs_block(emb_f16, n_total, E_pad, row_begin, row_end, 0, n_total, threshold, compare_mode, out, capacity,
                               count, stream);
}

extern "C" int b2c_dedup_pairs_block(const void* emb_f16, int64_t n_total, int E_pad, int64_t row_begin, int64_t row_end,
                                     int64_t col_begin, int64_t col_end, float threshold, int compare_mode, b2c_pair* out,
                                     unsigned long long capacity, unsigned long long* count, b2c_stream stream_) {
  using namespace b2c;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B2C_REQUIRE(emb_f16 && count, "b2c_dedup_pairs: null pointer");
  B2C_REQUIRE(out || capacity == 0, "b2c_dedup_pairs: null output with non-zero capacity");
  B2C_REQUIRE(E_pad > 0 && E_pad % kBK == 0, "b2c_dedup_pairs: E_pad=%d must be a positive multiple of %d", E_pad, kBK);
  B2C_REQUIRE(n_total >= 0 && n_total < (1ll << 31) - kBN, "b2c_dedup_pairs: n_total=%lld out of range", (long long)n_total);
  B2C_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n_total, "b2c_dedup_pairs: bad row range [%lld,%lld)",
              (long long)row_begin, (long long)row_end);
  B2C_REQUIRE(compare_mode == B2C_CMP_FP32 || compare_mode == B2C_CMP_REF_FP16 || compare_mode == B2C_CMP_EUCLID,
              "b2c_dedup_pairs: compare_mode %d", compare_mode);
  B2C_REQUIRE(col_begin >= 0 && col_begin <= col_end && col_end <= n_total, "b2c_dedup_pairs: bad column range [%lld,%lld)",
              (long long)col_begin, (long long)col_end);
  if (row_end == row_begin || n_total < 2 || col_end <= col_begin || col_end <= row_begin + 1) return 0;

  CUtensorMap tm_a;  // box 128 rows x 64 columns: both operands of the CTA-pair kernel
  B2C_TRY(make_tmap_2d(&tm_a, emb_f16, n_total, E_pad, static_cast<uint64_t>(E_pad) * 2, kBM, 0));

  const int sms = num_sms();
  B2C_REQUIRE(sms > 0, "no CUDA device");

  DedupParams p;
  p.k_blocks = E_pad / kBK;
  p.n_total = n_total;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.col_begin = col_begin;
  p.col_end = col_end;
  p.mode = compare_mode;
  p.sgn = 1.0f;
  if (compare_mode == B2C_CMP_REF_FP16) {
    p.thr = __half2float(__float2half_rn(threshold));
    p.thr_quick = p.thr;  // fp16(v) > h  implies  v > h  (round-to-nearest is monotone)
  } else if (compare_mode == B2C_CMP_EUCLID) {
    // sqrt(max(2 - 2c, 0)) > thr  =>  c < 1 - thr^2/2 (+ a rounding margin; `passes` decides); thr < 0 matches everything
    p.thr = threshold;
    p.sgn = -1.0f;
    p.thr_quick = threshold <= 0.f ? -4.0f : -(1.0f - 0.5f * threshold * threshold) - 1e-5f;
  } else {
    p.thr = threshold;
    p.thr_quick = threshold;
  }
  p.out = out;
  p.capacity = capacity;
  p.count = count;

  auto kern2 = umma2_tile_kernel<DedupPolicy>;
  static PerDeviceFlag attr_once;
  if (attr_once.first_use())
    B2C_CHECK_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, kUmma2SmemBytes));
  // bands start on 256-row boundaries so that both kernels see the same band grid
  const long long bi_begin = (row_begin / (2 * kBM)) * 2;
  const long long bi_end = (row_end + kBM - 1) / kBM;
  const long long NJ = (col_end + kBN - 1) / kBN;
  ProfScope ps(B2C_PROF_DEDUP, stream);
  for (long long b = bi_begin; b < bi_end; b += kDedupBandBlocks) {
    const int gi = static_cast<int>(bi_end - b < kDedupBandBlocks ? bi_end - b : kDedupBandBlocks);
    long long bj0 = (b * kBM) / kBN;  // first column block that can hold j > i ...
    if (col_begin / kBN > bj0) bj0 = col_begin / kBN;  // ... inside the requested column range
    const long long tiles = gi * (NJ - bj0);
    if (tiles <= 0) continue;
    B2C_REQUIRE(tiles < (1ll << 31), "b2c_dedup_pairs: too many tiles in one band");
    p.GI = gi;
    p.bi0 = static_cast<int>(b);
    p.bj0 = static_cast<int>(bj0);
    p.num_tiles = static_cast<int>(tiles);
    p.num_tiles2 = static_cast<int>(((gi + 1) / 2) * (NJ - bj0));
    const int grid = 2 * p.num_tiles2 < sms ? 2 * p.num_tiles2 : (sms & ~1);
    kern2<<<grid, kUmmaThreads, kUmma2SmemBytes, stream>>>(tm_a, tm_a, tm_a, tm_a, p, make_idesc_f16(2 * kBM, kBN, 0));
    B2C_POST_LAUNCH("umma2_tile_kernel<dedup>");
  }
  return 0;
}
