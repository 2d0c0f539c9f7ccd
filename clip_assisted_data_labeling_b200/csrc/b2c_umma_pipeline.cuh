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
//   static constexpr int kStore = kStoreDirect | kStoreTmaBf16 | kStoreTmaAddF32
//   kStoreDirect   : __device__ static void epilogue(const Params&, int a_row, int b_row, int row_in_tile,
//                                                    int col0, const uint32_t (&acc)[32])   // policy writes itself
//   kStoreTma*     : __device__ static void transform(const Params&, int col, float (&v)[32])   // bias / activation
//                    the kernel stages the 32-row x 128-byte slab of each epilogue warp in 128B-swizzled shared
//                    memory and hands it to the TMA unit: a plain tiled store (bf16) or a reduce-add performed
//                    at L2 (f32 residual stream: x += tile without the SM ever loading x).
#pragma once
#include <cuda_bf16.h>

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
constexpr int kStoreDirect = 0, kStoreTmaBf16 = 1, kStoreTmaAddF32 = 2;
constexpr int kSlabBytes = 32 * 128;          // one epilogue warp's staging slab: 32 rows x 128 B
constexpr int kSlabsPerWarp = 2;              // double buffered against the TMA unit's reads
constexpr int kStagingBytes = 4 * kSlabsPerWarp * kSlabBytes;
// ring + epilogue staging + 1 KB slack for manual 1024-B alignment + barriers
constexpr int kUmmaSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;

template <class Policy>
__global__ void __launch_bounds__(kUmmaThreads, 1)
umma_tile_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const typename Policy::Params p, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST
  uint8_t* staging = smem + kStages * kStageBytes;  // [4 warps][kSlabsPerWarp][32 rows][128 B], 1024-B aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
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
    if (Policy::kStore != kStoreDirect) tma_prefetch_desc(&tmap_out);
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
    // warp-uniform control flow, one elected lane issues (see b2c_umma_pipeline2.cuh for why)
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t smem_base = smem_u32(smem);
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int a_row, b_row;
      if (!Policy::tile(p, t, a_row, b_row)) continue;
      mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint64_t a_desc = make_sw128_kmajor_desc(sa);
        const uint64_t b_desc = make_sw128_kmajor_desc(sa + kABytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle row: +2 in the addr>>4 field
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full_bar[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
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
      if constexpr (Policy::kStore == kStoreDirect) {
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          Policy::epilogue(p, a_row, b_row, row_in_tile, c * 32, v);
        }
      } else {
        // lane = row of the slab; 16-byte chunk j of the row lands at chunk (j ^ (lane & 7)): the layout
        // CU_TENSOR_MAP_SWIZZLE_128B expects, and conflict-free for the 8 lanes of a store phase.
        uint8_t* my_slabs = staging + (warp - 2) * (kSlabsPerWarp * kSlabBytes);
        const int out_row = a_row + quarter * 32;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
          Policy::transform(p, b_row + c * 32, f);
          if constexpr (Policy::kStore == kStoreTmaAddF32) {
            uint8_t* slab = my_slabs + (c & 1) * kSlabBytes;
            if (lane == 0) tma_store_wait_read<kSlabsPerWarp - 1>();  // the slab's previous store has been read out
            __syncwarp();
            uint8_t* rowp = slab + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) =
                  make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmap_out, slab, b_row + c * 32, out_row);
              tma_store_commit();
            }
          } else {
            // bf16: two 32-column chunks fill one 128-byte row (64 columns) before the store is issued
            uint8_t* slab = my_slabs + ((c >> 1) & 1) * kSlabBytes;
            if ((c & 1) == 0) {
              if (lane == 0) tma_store_wait_read<kSlabsPerWarp - 1>();
              __syncwarp();
            }
            uint8_t* rowp = slab + lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
              w.x = *reinterpret_cast<uint32_t*>(&t0);
              w.y = *reinterpret_cast<uint32_t*>(&t1);
              w.z = *reinterpret_cast<uint32_t*>(&t2);
              w.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(rowp + ((((c & 1) * 4 + j) ^ (lane & 7)) << 4)) = w;
            }
            if (c & 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmap_out, slab, b_row + (c - 1) * 32, out_row);
                tma_store_commit();
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (Policy::kStore != kStoreDirect && lane == 0) tma_store_wait<0>();  // all bulk stores retired before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace b2c
