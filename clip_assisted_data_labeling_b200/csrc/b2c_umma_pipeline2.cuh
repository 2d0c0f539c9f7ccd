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
//
// LayerNorm fusion (GemmPolicy modes kGemmLn* / kGemmResidLnF32, b2c_gemm.cu):
//   * kStoreRmwLn (out_proj, c_proj): the epilogue owns the residual update.  Each epilogue warp streams its 32-row x
//     32-column fp32 slabs of x through a 3-deep shared-memory ring: TMA load of the old values (issued two slabs
//     ahead, across tile boundaries) -> += accumulator + bias in place -> TMA store of the new fp32 values, plus a
//     bf16 copy of them (the A operand of the next GEMM) through two 64-B-swizzled half slabs.  Every thread owns one
//     row and keeps (mean, M2) of its 256 columns (exact two-pass per 32-column slab, Chan's merge across slabs);
//     the partials go to stats[row][n_block] — merged in a fixed order by the consumer, so results are deterministic.
//     The bf16 copy is x̃ = bf16(x − m̂) with m̂ = the row's mean BEFORE this update (from the previous update's partials,
//     which live in the other of two statistics buffers), stored to shift[row]: rounding the centred value keeps the
//     copy's error relative to the row's spread even when a trained model's residual rows sit far from zero.
//   * Policy::kLnFold (in_proj, c_fc): LN(x)·Wᵀ = rstd·(x̃·(γ⊙W)ᵀ − (μ − m̂)·colsum(γ⊙W)) + (β·Wᵀ + b) — the epilogue
//     thread merges its row's partials into (μ, rstd) once per tile and applies them with two FMAs per element.
//   Together they remove the stand-alone LayerNorm kernels (and one fp32 read of x per LayerNorm) from the layer loop.
#pragma once
#include "b2c_umma_pipeline.cuh"

namespace b2c {

constexpr int kStoreRmwLn = 3;
constexpr int kStages2 = 6;
constexpr int kStage2Bytes = kABytes + kBM * kBK * 2;  // A 128 x 64 + B half 128 x 64 = 32 KB
constexpr int kUmma2BarBytes = 320;
constexpr int kUmma2XchgBytes = 4 * 32 * 8;  // kStoreRmwLn with 8 epilogue warps: (mean, M2) of the upper column half, per row
constexpr int kUmma2SmemBytes = kStages2 * kStage2Bytes + kStagingBytes + 1024 + kUmma2BarBytes + kUmma2XchgBytes;
// kStoreRmwLn, same total: Policy::kRmwRing fp32 slabs (4 KB) + 2 bf16 half slabs (2 KB) per epilogue warp, and as many
// mainloop stages as still fit: ring 5 -> 4 stages (out_proj, K = d: the epilogue is the long pole, an x slab must be
// requested ~4 slabs ahead), ring 3 -> 5 stages, ring 1 -> 6 stages (c_proj, K = 4d: the mainloop needs all six stages
// to cover the L2 latency and leaves the epilogue four times the slack, so it can afford to wait for each slab).
constexpr int kRmwMaxRing = 5;
constexpr int kHalfSlabBytes = 32 * 64;
constexpr int rmw_warp_bytes(int ring) { return ring * kSlabBytes + 2 * kHalfSlabBytes; }
// mainloop stages that fit beside `warps` epilogue warps with an x ring of `ring` slabs each (same total as the other modes)
constexpr int rmw_stages(int ring, int warps) {
  return (kStages2 * kStage2Bytes + kStagingBytes - warps * rmw_warp_bytes(ring)) / kStage2Bytes;
}

__device__ __forceinline__ uint32_t pack_bf16_pair(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// (mean, M2) of n_a + n_b values from the two parts' (mean, M2)   [Chan et al.]
__device__ __forceinline__ void chan_merge(float& mean_a, float& m2_a, float n_a, float mean_b, float m2_b, float n_b) {
  const float n = n_a + n_b;
  const float delta = mean_b - mean_a;
  const float w = n_b / n;
  mean_a = fmaf(delta, w, mean_a);
  m2_a = m2_a + m2_b + delta * delta * n_a * w;
}

template <class Policy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * Policy::kEpiWarps, 1)
umma2_tile_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
                  const typename Policy::Params p, const uint32_t idesc) {
  constexpr bool kRmw = Policy::kStore == kStoreRmwLn;
  constexpr int kRmwRing = kRmw ? Policy::kRmwRing : 1;
  static_assert(kRmwRing >= 1 && kRmwRing <= kRmwMaxRing && Policy::kEpiWarps * kRmwRing <= 4 * kRmwMaxRing, "x slab ring");
  constexpr int kRmwWarpBytes = rmw_warp_bytes(kRmwRing);
  constexpr int kStages2 = kRmw ? rmw_stages(kRmwRing, Policy::kEpiWarps) : b2c::kStages2;  // shadows the namespace constant inside the kernel
  constexpr int kStagingBytes = kRmw ? Policy::kEpiWarps * kRmwWarpBytes : b2c::kStagingBytes;
  static_assert(kStages2 >= 3, "mainloop stages");
  static_assert(kStages2 * kStage2Bytes + kStagingBytes <= b2c::kStages2 * kStage2Bytes + b2c::kStagingBytes, "smem budget");
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = smem_raw2 + ((1024u - (smem_u32(smem_raw2) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST
  uint8_t* staging = smem + kStages2 * kStage2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full_bar = bars;                                     // [kStages2]   (only the leader's are waited on)
  uint64_t* empty_bar = bars + kStages2;                         // [kStages2]   (each CTA waits on its own)
  uint64_t* acc_full_bar = bars + 2 * kStages2;                  // [kAccStages] (each CTA waits on its own)
  uint64_t* acc_empty_bar = bars + 2 * kStages2 + kAccStages;    // [kAccStages] (leader's: 8 arrivals, 4 per CTA)
  uint64_t* x_bar = bars + 2 * kStages2 + 2 * kAccStages;        // [4 warps][kRmwMaxRing] (kStoreRmwLn: x slab has landed)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(x_bar + 4 * kRmwMaxRing);
  [[maybe_unused]] float2* xchg = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + kUmma2BarBytes);  // [4 quarters][32 rows]
  static_assert((2 * kStages2 + 2 * kAccStages + 4 * kRmwMaxRing) * 8 + 4 <= kUmma2BarBytes, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (Policy::kStore != kStoreDirect) tma_prefetch_desc(&tmap_out);
    if (kRmw) {
      tma_prefetch_desc(&tmap_out2);
      for (int s = 0; s < 4 * kRmwMaxRing; ++s) mbar_init(&x_bar[s], 1);
    }
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], 2 * Policy::kEpiWarps);  // one arrival per epilogue warp of both CTAs
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
    // Nothing on the per-chunk critical path goes to global memory: the tile's bias / column-sum vectors live in
    // registers (lane l holds columns l, l+32, ... and hands them out by shuffle; the next tile's are requested a tile
    // ahead), the row statistics of the next tile are requested a tile ahead, the TMEM load of chunk c+1 is in flight
    // while chunk c is processed, and (kStoreRmwLn) the old x slabs arrive through a ring filled kRmwRing-1 slabs ahead.
    // Policy::kEpiWarps = 4: one warp per TMEM lane quarter, all 8 column chunks each.  8: two warps per quarter (one
    // per SM sub-partition pair), 4 chunks each — the per-element epilogue work (bias / LayerNorm terms, activation,
    // rounding, staging) is a single dependent instruction stream per warp, and with one warp per sub-partition it
    // takes about as long as the tile's mainloop at K = 1024.
    constexpr int kEpiWarps = Policy::kEpiWarps;
    static_assert(kEpiWarps == 4 || kEpiWarps == 8, "4 or 8 epilogue warps");
    constexpr int kSplit = kEpiWarps / 4;
    constexpr int kWarpStaging = kRmw ? kRmwWarpBytes : kStagingBytes / kEpiWarps;
    const int ew = warp - 2;
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint8_t* my_slabs = staging + ew * kWarpStaging;
    const int row_off = static_cast<int>(rank) * kBM + quarter * 32;  // this warp's first row inside a 256-row tile
    constexpr int kChunks = kBN / 32 / kSplit;                         // column chunks of a tile this warp handles
    const int col_off = (ew >> 2) * kChunks * 32;                      // its first column inside the tile

    // ---- kStoreRmwLn: slab ring state.  g = running index of the slab being processed (slot g % ring, barrier parity
    // (g / ring) & 1); lane 0 also tracks the next slab to request: (ld_t, ld_c) with running index gi = g + max(ring - 1, 1).
    [[maybe_unused]] uint32_t g = 0, gi = 0;
    [[maybe_unused]] int ld_t = cid, ld_c = 0;
    [[maybe_unused]] uint64_t* xb = x_bar + ew * kRmwRing;
    auto issue_next_x = [&]() {
      if (ld_t < p.num_tiles2) {
        int m0 = 0, n0 = 0;
        Policy::tile2(p, ld_t, m0, n0);
        const uint32_t slot = gi % kRmwRing;
        mbar_arrive_expect_tx(&xb[slot], kSlabBytes);
        tma_load_2d(my_slabs + slot * kSlabBytes, &tmap_out, &xb[slot], n0 + col_off + ld_c * 32, m0 + row_off);
      }
      ++gi;
      if (++ld_c == kChunks) { ld_c = 0; ld_t += ncl; }
    };
    if constexpr (kRmw) {
      if (lane == 0)
        for (int i = 0; i < (kRmwRing > 1 ? kRmwRing - 1 : 1); ++i) issue_next_x();
    }

    // ---- per-tile column vectors and row statistics, requested one tile ahead
    [[maybe_unused]] float col_b[kChunks], col_s[kChunks], ncol_b[kChunks], ncol_s[kChunks];
    [[maybe_unused]] float ln_a = 1.f, ln_b = 0.f;
    [[maybe_unused]] float2 nst[8];
    [[maybe_unused]] float nsh = 0.f, sh = 0.f;  // shift of the row's bf16 copy: read (kLnFold) / produced (kRmw); next tile's, this tile's
    if constexpr (Policy::kStore != kStoreDirect) {
      if (cid < p.num_tiles2) {
        int m0 = 0, n0 = 0;
        Policy::tile2(p, cid, m0, n0);
        Policy::load_cols(p, n0 + col_off + lane, ncol_b, ncol_s);
        if constexpr (Policy::kLnFold) Policy::load_row_stats(p, m0 + row_off + lane, nst, nsh);
        if constexpr (kRmw) nsh = Policy::load_old_mean(p, m0 + row_off + lane);
      }
    }

    for (int t = cid; t < p.num_tiles2; t += ncl) {
      int m_row, b_row;
      if (!Policy::tile2(p, t, m_row, b_row)) continue;
      const int a_row = m_row + static_cast<int>(rank) * kBM;
      const int out_row = a_row + quarter * 32;
      if constexpr (Policy::kStore != kStoreDirect) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) { col_b[c] = ncol_b[c]; col_s[c] = ncol_s[c]; }
        if constexpr (Policy::kLnFold) Policy::merge_row_stats(p, out_row + lane, nst, nsh, ln_a, ln_b);
        if constexpr (kRmw) sh = nsh;
        if (t + ncl < p.num_tiles2) {
          int nm = 0, nn = 0;
          Policy::tile2(p, t + ncl, nm, nn);
          Policy::load_cols(p, nn + col_off + lane, ncol_b, ncol_s);
          if constexpr (Policy::kLnFold) Policy::load_row_stats(p, nm + row_off + lane, nst, nsh);
          if constexpr (kRmw) nsh = Policy::load_old_mean(p, nm + row_off + lane);
        }
      }
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kBN + col_off + (static_cast<uint32_t>(quarter * 32) << 16);
      const int col0 = b_row + col_off;
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
      } else {
        uint32_t vv[2][32];
        tmem_ld_32x32(taddr, vv[0]);
        [[maybe_unused]] float mean = 0.f, m2 = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          tmem_ld_wait();
          if (c + 1 < kChunks) tmem_ld_32x32(taddr + (c + 1) * 32, vv[(c + 1) & 1]);
          float f[32];
          // lane e of the warp holds this chunk's bias (and column sum) for column e
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float bs0 = __shfl_sync(0xffffffffu, col_b[c], e);
            const float bs1 = __shfl_sync(0xffffffffu, col_b[c], e + 1);
            if constexpr (Policy::kLnFold) {
              const float cs0 = __shfl_sync(0xffffffffu, col_s[c], e);
              const float cs1 = __shfl_sync(0xffffffffu, col_s[c], e + 1);
              float t0, t1;  // two elements per FFMA2
              ffma2(t0, t1, cs0, cs1, ln_b, ln_b, bs0, bs1);
              ffma2(f[e], f[e + 1], __uint_as_float(vv[c & 1][e]), __uint_as_float(vv[c & 1][e + 1]), ln_a, ln_a, t0, t1);
            } else {
              f[e] = __uint_as_float(vv[c & 1][e]) + bs0;
              f[e + 1] = __uint_as_float(vv[c & 1][e + 1]) + bs1;
            }
          }
          Policy::activate(f);
          if constexpr (kRmw) {
            const uint32_t slot = g % kRmwRing;
            mbar_wait(&xb[slot], (g / kRmwRing) & 1);  // the old x values of this slab have landed
            uint8_t* slab = my_slabs + slot * kSlabBytes;
            uint8_t* rowp = slab + lane * 128;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4* q = reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4));
              const float4 xo = *q;
              f[4 * j] += xo.x; f[4 * j + 1] += xo.y; f[4 * j + 2] += xo.z; f[4 * j + 3] += xo.w;
              *q = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              s += (f[4 * j] + f[4 * j + 1]) + (f[4 * j + 2] + f[4 * j + 3]);
            }
            // statistics of the new row values: exact two-pass over the 32 columns in registers, merged into the tile's
            const float cm = s * (1.0f / 32.0f);
            float cq = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float dlt = f[e] - cm;
              cq = fmaf(dlt, dlt, cq);
            }
            if (c == 0) { mean = cm; m2 = cq; }
            else chan_merge(mean, m2, 32.0f * c, cm, cq, 32.0f);
            // bf16 copy of (x - shift): 32 columns = 64 B per row, 16-byte chunk j at (j ^ ((row >> 1) & 3)):
            // CU_TENSOR_MAP_SWIZZLE_64B.  shift = the row's mean before this update, so the rounding error of the copy
            // scales with the row's spread, not with its offset.
            uint8_t* bslab = my_slabs + kRmwRing * kSlabBytes + (c & 1) * kHalfSlabBytes;
            uint8_t* browp = bslab + lane * 64;
            const float nsh_c = -sh;
#pragma unroll
            for (int e = 0; e < 32; e += 2) fadd2(f[e], f[e + 1], f[e], f[e + 1], nsh_c, nsh_c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16_pair(f[8 * j + 0], f[8 * j + 1]);
              w.y = pack_bf16_pair(f[8 * j + 2], f[8 * j + 3]);
              w.z = pack_bf16_pair(f[8 * j + 4], f[8 * j + 5]);
              w.w = pack_bf16_pair(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(browp + ((j ^ ((lane >> 1) & 3)) << 4)) = w;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_out, slab, col0 + c * 32, out_row);
              tma_store_2d(&tmap_out2, bslab, col0 + c * 32, out_row);
              tma_store_commit();
              // ring > 1: the PREVIOUS slab's stores have been read out -> its ring slot and half slab are free;
              // ring 1: this slab's own stores, the next load goes into the same slot
              tma_store_wait_read<(kRmwRing > 1 ? 1 : 0)>();
              issue_next_x();
            }
            __syncwarp();
            ++g;
          } else if constexpr (Policy::kStore == kStoreTmaAddF32) {
            constexpr int kAddSlabs = kWarpStaging / kSlabBytes;  // 1 (8 warps) or 2 (4 warps)
            uint8_t* slab = my_slabs + (c % kAddSlabs) * kSlabBytes;
            if (lane == 0) tma_store_wait_read<kAddSlabs - 1>();
            __syncwarp();
            uint8_t* rowp = slab + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmap_out, slab, col0 + c * 32, out_row);
              tma_store_commit();
            }
          } else {
            // bf16: one 32-column chunk = 64 B per row = one half slab (CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk j of
            // row r at j ^ ((r >> 1) & 3)), stored per chunk and double-buffered against the TMA unit's reads
            constexpr int kHalfSlabs = kWarpStaging / kHalfSlabBytes;
            uint8_t* hslab = my_slabs + (c % kHalfSlabs) * kHalfSlabBytes;
            if (lane == 0) tma_store_wait_read<kHalfSlabs - 1>();
            __syncwarp();
            uint8_t* rowp = hslab + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16_pair(f[8 * j + 0], f[8 * j + 1]);
              w.y = pack_bf16_pair(f[8 * j + 2], f[8 * j + 3]);
              w.z = pack_bf16_pair(f[8 * j + 4], f[8 * j + 5]);
              w.w = pack_bf16_pair(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(rowp + ((j ^ ((lane >> 1) & 3)) << 4)) = w;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_out, hslab, col0 + c * 32, out_row);
              tma_store_commit();
            }
          }
        }
        if constexpr (kRmw) {
          if constexpr (kSplit == 2) {
            // two warps share a row (128 columns each): the upper half's partial goes through shared memory to the
            // lower half's warp, which merges in a fixed order and stores.  Named barrier 1 + quarter, both warps; the
            // second barrier keeps the next tile's partial from overwriting one that has not been read.
            float2* slot = xchg + quarter * 32 + lane;
            if (ew >= 4) *slot = make_float2(mean, m2);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            if (ew < 4) {
              const float2 up = *slot;
              chan_merge(mean, m2, 128.0f, up.x, up.y, 128.0f);
              Policy::store_row_stats(p, out_row + lane, b_row, mean, m2, sh);
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
          } else {
            Policy::store_row_stats(p, out_row + lane, b_row, mean, m2, sh);
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
