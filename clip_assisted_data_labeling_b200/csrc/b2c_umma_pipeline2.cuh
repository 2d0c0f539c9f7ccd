// b2c_umma_pipeline2.cuh — the CTA-pair (cta_group::2) variant of the GEMM mainloop.
//
//   D[256 x 256] (fp32, TMEM of two SMs)  =  A[256 x K]  x  B[256 x K]^T
//
// A cluster of two CTAs (one SM each, same TPC) owns one 256 x 256 tile: CTA r loads rows [128r, 128r+128) of A and
// rows [128r, 128r+128) of B (half of the tile's N) per k-block — 32 KB per stage instead of the 48 KB the
// single-CTA 128 x 256 tile needs for half the FLOPs, i.e. 1.5x the arithmetic intensity against L2 and half the
// shared-memory operand traffic per MMA.  The leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) and
// multicasts its commits to the barriers of both CTAs; each CTA's epilogue warps drain their own 128 TMEM lanes
// through the same swizzled-slab TMA store / reduce-add path as the single-CTA kernel.
//
// Policy: as in b2c_umma_pipeline.cuh, plus
//   __device__ static bool tile2(const Params&, int t, int& m_row, int& n_row)   // 256 x 256 tile origin; false = skip
// Params must provide num_tiles2 and k_blocks.
#pragma once
#include "b2c_umma_pipeline.cuh"

namespace b2c {

constexpr int kStages2 = 6;
constexpr int kStage2Bytes = kABytes + kBM * kBK * 2;  // A 128 x 64 + B half 128 x 64 = 32 KB
constexpr int kUmma2SmemBytes = kStages2 * kStage2Bytes + kStagingBytes + 1024 + 256;

template <class Policy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kUmmaThreads, 1)
umma2_tile_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const typename Policy::Params p, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = smem_raw2 + ((1024u - (smem_u32(smem_raw2) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST
  uint8_t* staging = smem + kStages2 * kStage2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;                                     // [kStages2]   (only the leader's are waited on)
  uint64_t* empty_bar = bars + kStages2;                         // [kStages2]   (each CTA waits on its own)
  uint64_t* acc_full_bar = bars + 2 * kStages2;                  // [kAccStages] (each CTA waits on its own)
  uint64_t* acc_empty_bar = bars + 2 * kStages2 + kAccStages;    // [kAccStages] (leader's: 8 arrivals, 4 per CTA)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages2 + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (Policy::kStore != kStoreDirect) tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], 8);
    }
    mbar_fence_init();
  }
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  if (warp == 1) tmem_alloc_2sm(tmem_base_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int cid = blockIdx.x >> 1;
  const int ncl = gridDim.x >> 1;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    // The whole warp runs the loop so that every address / coordinate stays in the uniform datapath; one elected
    // lane issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t leader_full0 = mapa_shared(smem_u32(&full_bar[0]), 0);
    for (int t = cid; t < p.num_tiles2; t += ncl) {
      int m_row, n_row;
      if (!Policy::tile2(p, t, m_row, n_row)) continue;
      const int a_row = m_row + static_cast<int>(rank) * kBM;
      const int b_row = n_row + static_cast<int>(rank) * (kBN / 2);
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * kStage2Bytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t leader_full = leader_full0 + stage * 8;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStage2Bytes);  // bytes of BOTH CTAs
          tma_load_2d_2sm(sa, &tmap_a, leader_full, kb * kBK, a_row);
          tma_load_2d_2sm(sb, &tmap_b, leader_full, kb * kBK, b_row);
        }
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA only)
    // Warp-uniform control flow: all 32 lanes wait on the barriers and compute the (uniform) descriptors, one elected
    // lane issues.  Issuing from inside a `lane == 0` branch makes ptxas treat every operand as divergent and wrap each
    // tcgen05.mma in an ELECT / R2UR / BRA.U.ANY waterfall (~25 instructions per MMA), which made the issue loop
    // itself as long as the 4 MMAs of a k-block take to execute.
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int t = cid; t < p.num_tiles2; t += ncl) {
        int m_row, n_row;
        if (!Policy::tile2(p, t, m_row, n_row)) continue;
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStage2Bytes;
          const uint64_t a_desc = make_sw128_kmajor_desc(sa);
          const uint64_t b_desc = make_sw128_kmajor_desc(sa + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(&empty_bar[stage], 3);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(&acc_full_bar[acc], 3);
        __syncwarp();
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* my_slabs = staging + (warp - 2) * (kSlabsPerWarp * kSlabBytes);
    for (int t = cid; t < p.num_tiles2; t += ncl) {
      int m_row, b_row;
      if (!Policy::tile2(p, t, m_row, b_row)) continue;
      const int a_row = m_row + static_cast<int>(rank) * kBM;
      const int out_row = a_row + quarter * 32;
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kBN + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (Policy::kStore == kStoreDirect) {
        // four 32-column TMEM loads in flight per wait: the load latency (contended by the running MMAs) is paid
        // twice per tile instead of eight times
#pragma unroll 1
        for (int c4 = 0; c4 < kBN / 128; ++c4) {
          uint32_t v0[32], v1[32], v2[32], v3[32];
          tmem_ld_32x32(taddr + c4 * 128, v0);
          tmem_ld_32x32(taddr + c4 * 128 + 32, v1);
          tmem_ld_32x32(taddr + c4 * 128 + 64, v2);
          tmem_ld_32x32(taddr + c4 * 128 + 96, v3);
          tmem_ld_wait();
          Policy::epilogue(p, a_row, b_row, quarter * 32 + lane, c4 * 128, v0);
          Policy::epilogue(p, a_row, b_row, quarter * 32 + lane, c4 * 128 + 32, v1);
          Policy::epilogue(p, a_row, b_row, quarter * 32 + lane, c4 * 128 + 64, v2);
          Policy::epilogue(p, a_row, b_row, quarter * 32 + lane, c4 * 128 + 96, v3);
        }
      }
#pragma unroll 1
      for (int c = 0; c < (Policy::kStore == kStoreDirect ? 0 : kBN / 32); ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        Policy::transform(p, b_row + c * 32, f);
        if constexpr (Policy::kStore == kStoreTmaAddF32) {
          uint8_t* slab = my_slabs + (c & 1) * kSlabBytes;
          if (lane == 0) tma_store_wait_read<kSlabsPerWarp - 1>();
          __syncwarp();
          uint8_t* rowp = slab + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmap_out, slab, b_row + c * 32, out_row);
            tma_store_commit();
          }
        } else {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty_bar[acc]), 0));  // the leader's barrier
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (Policy::kStore != kStoreDirect && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM or exits while the peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

}  // namespace b2c
