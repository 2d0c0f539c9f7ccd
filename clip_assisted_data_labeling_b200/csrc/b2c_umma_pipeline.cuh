// b2c_umma_pipeline.cuh — the one tensor-core mainloop every dense contraction of the path uses
// (patch-embed, QKV, out-proj, MLP GEMMs of the ViT tower; the embedding x embedding^T similarity
// tiles of the duplicate search).
//
//   D[128 x 256] (fp32, TMEM)  =  A[128 x K] (smem, K-major)  x  B[256 x K]^T (smem, K-major)
//
// Persistent, warp-specialised CTA (one per SM, 192 threads):
//   warp 0        TMA producer   : cp.async.bulk.tensor 128B-swizzled tiles into a 4-deep smem ring
//   warp 1        MMA issuer     : one elected lane issues tcgen05.mma (M128 N256 K16), commits to mbarriers
//   warps 2..5    epilogue       : tcgen05.ld the accumulator (one TMEM lane quarter per warp) and hand
//                                  32-column chunks to Policy::epilogue()
// Two accumulator stages (2 x 256 TMEM columns = all 512) let the epilogue of tile i overlap the
// mainloop of tile i+1.
//
// A Policy supplies:  struct Params { int num_tiles; int k_blocks; ... };
//   __device__ static bool tile(const Params&, int t, int& a_row, int& b_row)   // false = skip tile
//   __device__ static void epilogue(const Params&, int a_row, int b_row, int row_in_tile, int col0,
//                                   const uint32_t (&acc)[32])
#pragma once
#include "b2c_ptx.cuh"

namespace b2c {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;
constexpr int kBBytes = kBN * kBK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBN;  // 512
constexpr int kUmmaThreads = 192;
// ring + 1 KB slack for manual 1024-B alignment + barriers
constexpr int kUmmaSmemBytes = kStages * kStageBytes + 1024 + 256;

template <class Policy>
__global__ void __launch_bounds__(kUmmaThreads, 1)
umma_tile_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const typename Policy::Params p, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                           // [kStages]
  uint64_t* empty_bar = bars + kStages;                // [kStages]
  uint64_t* acc_full_bar = bars + 2 * kStages;         // [kAccStages]
  uint64_t* acc_empty_bar = bars + 2 * kStages + kAccStages;  // [kAccStages]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int a_row, b_row;
        if (!Policy::tile(p, t, a_row, b_row)) continue;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBK, a_row);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBK, b_row);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int a_row, b_row;
        if (!Policy::tile(p, t, a_row, b_row)) continue;
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t a_desc = make_sw128_kmajor_desc(sa);
          const uint64_t b_desc = make_sw128_kmajor_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in the addr>>4 field
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int a_row, b_row;
      if (!Policy::tile(p, t, a_row, b_row)) continue;
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kBN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        Policy::epilogue(p, a_row, b_row, row_in_tile, c * 32, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace b2c
