// b2c_jpeg_huff.cu — K14b: the Huffman stage of K14 on the device.
// Replaces, for single-scan sequential files, the host stage of b2c_jpeg.cu (jdhuff.c's decode_mcu restated there): the
// file's bytes are copied to the device as they are and one CTA per image produces the dense coefficient buffer
// b2c_jpeg_reconstruct consumes.  A Huffman stream has no random access, but it SELF-SYNCHRONISES: a decoder started at a
// wrong bit position falls into step with the true code-word boundaries after a few symbols.  The kernel uses that the
// way Weißenberger & Schmidt describe for JPEG ("Accelerating JPEG decompression on GPUs", 2021), restated for one CTA
// per image and with an explicit verification pass:
//   1. destuff     : FF 00 -> FF, stop at the first marker other than RSTn; RSTn markers are dropped and their positions
//                    (in destuffed bytes) become the segment table.  CTA-wide stream compaction.
//   2. speculate   : thread i decodes sub-sequence i (128 destuffed bytes, more for large files) from the state (first bit, block 0 of the MCU,
//                    DC coefficient next) — true only for the first sub-sequence of a segment — and records its exit
//                    state (bit position, block-in-MCU, zigzag index) and the number of blocks it completed.
//   3. synchronise : every thread continues through the following sub-sequences from its own exit state, overwriting
//                    their records, until its exit state equals the recorded one (from then on it would only repeat its
//                    successor's work).  Typically one or two steps.
//   4. write       : every thread decodes its sub-sequence once more from its predecessor's recorded exit state, now
//                    knowing the absolute block index (prefix sum of the block counts), stores the coefficients and
//                    VERIFIES that it arrives at its own recorded exit state and block count.  By induction from the
//                    true start of the segment a stream that passes is decoded exactly; anything else is reported.
//   5. DC          : differences -> values, a prefix sum per component over the blocks in decode order (restarted at
//                    every restart interval).
// Sub-sequences of all restart segments are numbered consecutively and handled in rounds of kHuffThreads; a round that
// cuts a segment hands the true exit state and block count of its last sub-sequence to the next one.
#include <string.h>

#include <vector>

#include "b2c_launch.h"

namespace b2c {
namespace {

#ifndef B2C_HUFF_THREADS
#define B2C_HUFF_THREADS 512
#endif
constexpr int kHuffThreads = B2C_HUFF_THREADS;
// A sub-sequence is 128 destuffed bytes, doubled (up to 1 KB) until the image's sub-sequences fit one round of the CTA:
// short sub-sequences keep every thread of a small file busy, long ones spare a large file the per-round barriers and
// the re-decoding of synchronisation steps (measured per 256 files, 34 KB / 136 KB each: 1024 bits 0.74 / 2.42 ms,
// 2048 bits 0.98 / 2.07 ms, 4096 bits 1.49 / 1.91 ms).
constexpr uint32_t kSubBitsMin = 1024, kSubBitsMax = 8192;

struct HuffJobDev {
  const uint8_t* src;   // first entropy-coded byte (device)
  uint32_t raw_len;     // bytes from there to the end of the file
  uint32_t max_segs;    // capacity of the segment table
  uint8_t* clean;       // destuffed stream, 16-byte aligned, raw_len + 64 bytes
  uint32_t* seg_begin;  // [max_segs + 1] first destuffed byte of every restart segment
  uint32_t* sub_base;   // [max_segs + 1] sub-sequences in the segments before this one
  int16_t* coefs;       // dense output, natural order
  const b2c_jpeg_hufftab* tabs;  // 4 tables (device)
  int32_t ncomp, hs0, vs0, mcus_x, mcus_y;
  int32_t restart_interval;
  int32_t blocks_w[3];
  int32_t block_off[3];  // coef_offset[c] / 64
  int64_t coef_count;
};

__constant__ uint8_t kZigzagNat[80] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33,
                                       40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36,
                                       29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
                                       47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Bits {
  const uint32_t* w;  // destuffed stream as words
  uint64_t acc;       // unread bits, left-aligned
  int n;              // how many
  uint32_t next;      // next word to load
  __device__ __forceinline__ void init(const uint32_t* words, uint32_t bitpos) {
    w = words;
    next = bitpos >> 5;
    const uint32_t v = __byte_perm(w[next], 0, 0x0123);
    ++next;
    const int sh = bitpos & 31;
    acc = (static_cast<uint64_t>(v) << 32) << sh;
    n = 32 - sh;
  }
  __device__ __forceinline__ void fill() {  // afterwards n > 32: a code (<= 16 bits) and its magnitude bits (<= 15) fit
    if (n <= 32) {
      const uint32_t v = __byte_perm(w[next], 0, 0x0123);
      ++next;
      acc |= static_cast<uint64_t>(v) << (32 - n);
      n += 32;
    }
  }
  __device__ __forceinline__ uint32_t peek(int k) const { return static_cast<uint32_t>(acc >> (64 - k)); }
  __device__ __forceinline__ void drop(int k) {
    acc <<= k;
    n -= k;
  }
  __device__ __forceinline__ uint32_t pos() const { return next * 32u - static_cast<uint32_t>(n); }
};

// symbol, or -1 when no code of up to 16 bits matches (nothing is consumed then)
__device__ __forceinline__ int huff_symbol(Bits& br, const b2c_jpeg_hufftab& t) {
  const uint32_t e = t.look[br.peek(9)];
  if (e) {
    br.drop(e >> 8);
    return e & 0xFF;
  }
  int l = 10;
  int code = static_cast<int>(br.peek(10));
  while (l <= 16 && code > t.maxcode[l]) {
    ++l;
    code = static_cast<int>(br.peek(l));
  }
  if (l > 16) return -1;
  br.drop(l);
  return t.vals[(code + t.valoffset[l]) & 0xFF];
}

__device__ __forceinline__ int receive_extend(Bits& br, int s) {  // s in 1..15
  const int v = static_cast<int>(br.peek(s));
  br.drop(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// decode-order block d of the image -> its 64 coefficients in the dense buffer (component-major, row-major blocks)
__device__ __forceinline__ int16_t* block_ptr(const HuffJobDev& J, int bpm, int ylum, int d) {
  const int mcu = d / bpm, r = d - mcu * bpm;
  const int uy = mcu / J.mcus_x, ux = mcu - uy * J.mcus_x;
  int c, by, bx;
  if (r < ylum) {
    const int v = r / J.hs0, h = r - v * J.hs0;
    c = 0; by = uy * J.vs0 + v; bx = ux * J.hs0 + h;
  } else {
    c = r - ylum + 1; by = uy; bx = ux;
  }
  return J.coefs + (static_cast<int64_t>(J.block_off[c]) + static_cast<int64_t>(by) * J.blocks_w[c] + bx) * 64;
}

struct HState {
  uint32_t p;  // bit position of the next code word
  int b, z;    // block within the MCU, zigzag index of the next coefficient (0: the DC difference comes next)
  __device__ __forceinline__ uint64_t pack() const {
    return (static_cast<uint64_t>(p) << 16) | (static_cast<uint64_t>(b) << 8) | static_cast<uint64_t>(z);
  }
  __device__ __forceinline__ static HState unpack(uint64_t v) {
    HState s;
    s.p = static_cast<uint32_t>(v >> 16);
    s.b = static_cast<int>((v >> 8) & 0xFF);
    s.z = static_cast<int>(v & 0xFF);
    return s;
  }
};

// Decodes code words that START before end_bit, from state `st` (updated).  nblk = blocks completed.  WRITE: blocks are
// numbered from blk_abs (decode order), at most blk_limit of them exist; coefficients go to the dense buffer and
// violations of the format set err.  Without WRITE the function is a pure state transition that tolerates garbage
// (speculative starts): an unmatched code skips one bit, a run past the block ends the block.
template <bool WRITE>
__device__ __forceinline__ void decode_sub(const HuffJobDev& J, const b2c_jpeg_hufftab* tabs, const uint32_t* words,
                                           uint32_t end_bit, int bpm, int ylum, HState& st, int& nblk, int blk_abs,
                                           int blk_limit, int& err) {
  Bits br;
  br.init(words, st.p);
  int b = st.b, z = st.z, n = 0;
  int16_t* dst = nullptr;
  if (WRITE && blk_abs < blk_limit) dst = block_ptr(J, bpm, ylum, blk_abs);
  while (br.pos() < end_bit && (!WRITE || blk_abs + n < blk_limit)) {
    br.fill();
    const int cls = b < ylum ? 0 : 2;
    if (z == 0) {
      int s = huff_symbol(br, tabs[cls]);
      if (s < 0 || s > 11) {
        if (WRITE) err |= B2C_JPEG_HUFF_CORRUPT;
        if (s < 0) br.drop(1);
        s = 0;
      }
      const int diff = s ? receive_extend(br, s) : 0;
      if (WRITE) dst[0] = static_cast<int16_t>(diff);
      z = 1;
    } else {
      const int rs = huff_symbol(br, tabs[cls + 1]);
      if (rs < 0) {
        if (WRITE) err |= B2C_JPEG_HUFF_CORRUPT;
        br.drop(1);
      } else {
        const int r = rs >> 4, s = rs & 15;
        if (s == 0) {
          z = r == 15 ? z + 16 : 64;  // ZRL / end of block
        } else {
          z += r;
          const int v = receive_extend(br, s);
          if (z > 63) {
            if (WRITE) err |= B2C_JPEG_HUFF_CORRUPT;
          } else if (WRITE) {
            dst[kZigzagNat[z]] = static_cast<int16_t>(v);
          }
          ++z;
        }
      }
    }
    if (z >= 64) {
      z = 0;
      b = b + 1 == bpm ? 0 : b + 1;
      ++n;
      if (WRITE && blk_abs + n < blk_limit) dst = block_ptr(J, bpm, ylum, blk_abs + n);
    }
  }
  st.p = br.pos();
  st.b = b;
  st.z = z;
  nblk = n;
}

// CTA-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = sum over the CTA
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kHuffThreads / 32; ++i) {
    const int t = s_warp[i];
    if (i < wid) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kHuffThreads) jpeg_huff_kernel(const HuffJobDev* __restrict__ jobs,
                                                                 int32_t* __restrict__ status) {
  __shared__ b2c_jpeg_hufftab s_tabs[4];
  __shared__ uint64_t s_state[kHuffThreads];
  __shared__ int s_nblk[kHuffThreads];
  __shared__ int s_warp[kHuffThreads / 32];
  __shared__ uint32_t s_end;
  __shared__ HuffJobDev s_job;
  const int tid = threadIdx.x;
  if (tid == 0) s_job = jobs[blockIdx.x];
  __syncthreads();
  const HuffJobDev& J = s_job;
  const int ylum = J.hs0 * J.vs0, bpm = ylum + (J.ncomp == 3 ? 2 : 0);
  const int total_mcus = J.mcus_x * J.mcus_y;

  // ---- 0. tables -> shared memory, output -> zero (only non-zero coefficients are stored later)
  {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(J.tabs);
    uint32_t* s = reinterpret_cast<uint32_t*>(s_tabs);
    for (int i = tid; i < static_cast<int>(4 * sizeof(b2c_jpeg_hufftab) / 4); i += kHuffThreads) s[i] = g[i];
    uint4* o = reinterpret_cast<uint4*>(J.coefs);
    const int64_t n16 = J.coef_count / 8;
    for (int64_t i = tid; i < n16; i += kHuffThreads) o[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s_end = 0xFFFFFFFFu;
  }
  __syncthreads();

  // ---- 1. destuff + segment table
  uint32_t L = 0;     // destuffed bytes so far
  uint32_t nsegs = 1;  // segments begun
  if (tid == 0) J.seg_begin[0] = 0;
  for (uint32_t chunk = 0; chunk < J.raw_len; chunk += kHuffThreads * 4) {
    const uint32_t o = chunk + tid * 4;
    uint32_t b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int64_t idx = static_cast<int64_t>(o) - 1 + k;
      b[k] = idx < 0 ? 0u : (idx < J.raw_len ? static_cast<uint32_t>(J.src[idx]) : 1u);  // past the end: "a marker follows"
    }
    uint32_t myend = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      const bool rst = b[k + 2] >= 0xD0 && b[k + 2] <= 0xD7;
      if (o + k < J.raw_len && b[k + 1] == 0xFF && b[k + 2] != 0x00 && !rst) myend = o + k;
    }
    if (myend != 0xFFFFFFFFu) atomicMin(&s_end, myend);
    __syncthreads();
    const uint32_t end = s_end;
    uint32_t keep = 0, mark = 0;  // per byte: kept in the destuffed stream / second byte of an RSTn marker
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t pos = o + k;
      if (pos >= J.raw_len || pos >= end) continue;
      const bool stuffed = b[k + 1] == 0x00 && b[k] == 0xFF;
      const bool rst2 = b[k] == 0xFF && b[k + 1] >= 0xD0 && b[k + 1] <= 0xD7;
      const bool rst1 = b[k + 1] == 0xFF && b[k + 2] >= 0xD0 && b[k + 2] <= 0xD7;
      if (rst2) mark |= 1u << k;
      if (!stuffed && !rst1 && !rst2) keep |= 1u << k;
    }
    int tot = 0;
    const int packed = __popc(keep) | (__popc(mark) << 16);
    const int ex = block_exclusive_scan(packed, s_warp, &tot);
    uint32_t off = L + (ex & 0xFFFF), sidx = nsegs + (ex >> 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (keep & (1u << k)) J.clean[off++] = static_cast<uint8_t>(b[k + 1]);
      if (mark & (1u << k)) {
        if (sidx <= J.max_segs) J.seg_begin[sidx] = off;  // the segment after this marker starts at the next kept byte
        ++sidx;
      }
    }
    L += tot & 0xFFFF;
    nsegs += tot >> 16;
    if (end != 0xFFFFFFFFu) break;
  }
  for (int i = tid; i < 32; i += kHuffThreads) J.clean[L + i] = 0;
  int err = 0;
  const int interval = J.restart_interval;
  const uint32_t want_segs = interval ? static_cast<uint32_t>((total_mcus + interval - 1) / interval) : 1u;
  if (nsegs != want_segs || nsegs > J.max_segs) err |= B2C_JPEG_HUFF_CORRUPT;
  if (tid == 0 && nsegs <= J.max_segs) J.seg_begin[nsegs] = L;
  __syncthreads();
  const uint32_t* words = reinterpret_cast<const uint32_t*>(J.clean);

  // ---- 2-4. sub-sequences of ALL restart segments, numbered consecutively, in rounds of kHuffThreads
  // sub_base[s] = number of sub-sequences in the segments before s.  Each segment has its own grid, anchored at its
  // first bit, and its first sub-sequence starts from the true state — restart markers only add parallelism.
  uint32_t total_subs = 0;
  uint32_t kSubBits = kSubBitsMin;  // (uniform over the CTA)
  while (kSubBits < kSubBitsMax && (L * 8u + kSubBits - 1) / kSubBits > static_cast<uint32_t>(kHuffThreads)) kSubBits <<= 1;
  if (!err) {
    for (uint32_t s0 = 0; s0 < nsegs; s0 += kHuffThreads) {
      const uint32_t sg = s0 + tid;
      int cnt = 0;
      if (sg < nsegs) {
        const uint32_t bytes = J.seg_begin[sg + 1] - J.seg_begin[sg];
        cnt = static_cast<int>((bytes * 8u + kSubBits - 1) / kSubBits);
        if (cnt == 0) err |= B2C_JPEG_HUFF_CORRUPT;  // two markers back to back: a segment holds at least one MCU
      }
      int tot = 0;
      const int ex = block_exclusive_scan(cnt, s_warp, &tot);
      if (sg < nsegs) J.sub_base[sg] = total_subs + ex;
      total_subs += tot;
    }
    if (tid == 0) J.sub_base[nsegs] = total_subs;
    if (__syncthreads_or(err)) err |= B2C_JPEG_HUFF_CORRUPT;
  }
  if (!err) {
    __shared__ int s_excl[kHuffThreads];
    __shared__ int s_carry;
    HState ent;          // true entry state of the round's first sub-sequence when it continues a segment
    ent.p = 0; ent.b = 0; ent.z = 0;
    int carry = 0;       // blocks that segment completed in earlier rounds
    for (uint32_t base = 0; base < total_subs; base += kHuffThreads) {
      const int nact = static_cast<int>(min(static_cast<uint32_t>(kHuffThreads), total_subs - base));
      const bool active = tid < nact;
      const uint32_t g = base + tid;
      // which segment, which sub-sequence of it
      uint32_t seg = 0, idx = 0, nsub = 1, sb = 0, se = 0;
      int blk_first = 0, blk_limit = 0;
      if (active) {
        uint32_t lo = 0, hi = nsegs - 1;
        while (lo < hi) {
          const uint32_t mid = (lo + hi + 1) >> 1;
          if (J.sub_base[mid] <= g) lo = mid;
          else hi = mid - 1;
        }
        seg = lo;
        idx = g - J.sub_base[seg];
        nsub = J.sub_base[seg + 1] - J.sub_base[seg];
        sb = J.seg_begin[seg] * 8u;
        se = J.seg_begin[seg + 1] * 8u;
        const int first_mcu = interval ? static_cast<int>(seg) * interval : 0;
        const int seg_mcus = interval ? min(interval, total_mcus - first_mcu) : total_mcus;
        blk_first = first_mcu * bpm;
        blk_limit = blk_first + seg_mcus * bpm;
      }
      const bool seg_first = idx == 0, seg_last = idx + 1 == nsub;
      // -- speculate
      HState st;
      st.p = 0; st.b = 0; st.z = 0;
      int nb = 0, dummy = 0;
      if (active) {
        if (seg_first) { st.p = sb; }
        else if (tid == 0) st = ent;
        else { st.p = sb + idx * kSubBits; }
        decode_sub<false>(J, s_tabs, words, min(sb + (idx + 1) * kSubBits, se), bpm, ylum, st, nb, 0, 0, dummy);
        s_state[tid] = st.pack();
        s_nblk[tid] = nb;
      }
      // -- synchronise (a chain ends with its segment)
      bool done = !active || tid == nact - 1 || seg_last;
      for (int k = 1; k < nact; ++k) {
        __syncthreads();
        const int j = tid + k;
        const bool work = !done && j < nact && idx + k < nsub;
        bool same = false;
        if (work) {
          decode_sub<false>(J, s_tabs, words, min(sb + (idx + k + 1) * kSubBits, se), bpm, ylum, st, nb, 0, 0, dummy);
          same = st.pack() == s_state[j];
        }
        __syncthreads();
        if (work) {
          s_nblk[j] = nb;  // block counts of the earlier-started chain are the authoritative ones
          if (same) done = true;
          else s_state[j] = st.pack();
        } else {
          done = true;
        }
        if (!__syncthreads_or(!done && tid + k + 1 < nact && idx + k + 1 < nsub)) break;
      }
      __syncthreads();
      // -- write + verify
      const int mine = active ? s_nblk[tid] : 0;
      int round_blocks = 0;
      const int ex = block_exclusive_scan(mine, s_warp, &round_blocks);
      s_excl[tid] = ex;
      __syncthreads();
      // blocks of my segment completed before my sub-sequence
      const bool from_earlier_round = idx > static_cast<uint32_t>(tid);
      const int before = from_earlier_round ? carry + ex : ex - s_excl[tid - static_cast<int>(idx)];
      if (active) {
        HState in;
        if (seg_first) { in.p = sb; in.b = 0; in.z = 0; }
        else if (tid == 0) in = ent;
        else in = HState::unpack(s_state[tid - 1]);
        int got = 0;
        decode_sub<true>(J, s_tabs, words, min(sb + (idx + 1) * kSubBits, se), bpm, ylum, in, got, blk_first + before,
                         blk_limit, err);
        if (seg_last) {
          if (blk_first + before + got != blk_limit) err |= in.p >= se ? B2C_JPEG_HUFF_TRUNCATED : B2C_JPEG_HUFF_CORRUPT;
          if (in.p > se) err |= B2C_JPEG_HUFF_TRUNCATED;  // consumed bits the file does not have
        } else if (in.pack() != s_state[tid] || got != mine) {
          err |= B2C_JPEG_HUFF_NOSYNC;
        }
        if (tid == nact - 1) s_carry = before + mine;  // only used if the next round continues this segment
      }
      ent = HState::unpack(s_state[nact - 1]);
      __syncthreads();
      carry = s_carry;
    }
  }
  // combine the threads' verdicts (bit by bit: __syncthreads_or returns a predicate)
  int any = 0;
#pragma unroll
  for (int bit = 1; bit <= 4; bit <<= 1)
    if (__syncthreads_or(err & bit)) any |= bit;

  // ---- 5. DC differences -> DC values: prefix sum per component in decode order, restarted at every restart interval
  if (!any) {
    const int per = (total_mcus + kHuffThreads - 1) / kHuffThreads;  // MCUs per thread, contiguous
    const int m0 = min(total_mcus, tid * per), m1 = min(total_mcus, m0 + per);
    // a thread's range may span restart boundaries: sums since the last boundary inside the range, plus whether one occurred
    int sum[3] = {0, 0, 0};
    bool reset = false;
    for (int m = m0; m < m1; ++m) {
      if (interval && m % interval == 0) { sum[0] = sum[1] = sum[2] = 0; reset = true; }
      for (int r = 0; r < bpm; ++r) sum[r < ylum ? 0 : r - ylum + 1] += block_ptr(J, bpm, ylum, m * bpm + r)[0];
    }
    // serial carry over the threads' summaries (512 entries, three values each): cheap next to the decode
    __shared__ int s_sum[3][kHuffThreads];
    __shared__ uint8_t s_reset[kHuffThreads];
    s_sum[0][tid] = sum[0]; s_sum[1][tid] = sum[1]; s_sum[2][tid] = sum[2];
    s_reset[tid] = reset ? 1 : 0;
    __syncthreads();
    int carry[3] = {0, 0, 0};
    for (int t = tid - 1; t >= 0; --t) {  // walk back to the nearest range that contains a restart boundary
      carry[0] += s_sum[0][t]; carry[1] += s_sum[1][t]; carry[2] += s_sum[2][t];
      if (s_reset[t]) break;
    }
    for (int m = m0; m < m1; ++m) {
      if (interval && m % interval == 0) carry[0] = carry[1] = carry[2] = 0;
      for (int r = 0; r < bpm; ++r) {
        int16_t* blk = block_ptr(J, bpm, ylum, m * bpm + r);
        int& c = carry[r < ylum ? 0 : r - ylum + 1];
        c += blk[0];
        blk[0] = static_cast<int16_t>(c);
      }
    }
  }
  if (tid == 0) status[blockIdx.x] = any;
}

size_t h256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct HuffLayout {
  size_t jobs, tabs, total;
};
uint32_t max_segments(const b2c_jpeg_huff& h) {
  // an RSTn marker takes two bytes: a scan cannot hold more than scan_bytes / 2 of them
  return h.restart_interval ? static_cast<uint32_t>(h.scan_bytes / 2 + 2) : 1u;
}

}  // namespace
}  // namespace b2c

extern "C" int b2c_jpeg_huff_workspace_bytes(const b2c_jpeg_huff* huffs, int n, size_t* bytes) {
  using namespace b2c;
  B2C_REQUIRE(huffs && bytes && n > 0, "b2c_jpeg_huff_workspace_bytes: bad arguments");
  size_t total = h256(static_cast<size_t>(n) * sizeof(HuffJobDev)) + h256(static_cast<size_t>(n) * 4 * sizeof(b2c_jpeg_hufftab));
  for (int i = 0; i < n; ++i) {
    B2C_REQUIRE(huffs[i].scan_bytes > 0 && huffs[i].scan_bytes < (1ll << 28), "b2c_jpeg_huff_workspace_bytes: bad scan size for image %d", i);
    total += h256(static_cast<size_t>(huffs[i].scan_bytes) + 64);
    total += 2 * h256((static_cast<size_t>(max_segments(huffs[i])) + 2) * sizeof(uint32_t));
  }
  *bytes = total;
  return 0;
}

extern "C" int b2c_jpeg_huff_decode(const b2c_jpeg_info* infos, const b2c_jpeg_huff* huffs, const uint8_t* const* files,
                                    int16_t* const* coefs, int32_t* status, int n, void* ws, size_t ws_bytes,
                                    b2c_stream stream_) {
  using namespace b2c;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B2C_REQUIRE(infos && huffs && files && coefs && status && ws && n > 0, "b2c_jpeg_huff_decode: bad arguments");
  size_t need = 0;
  B2C_TRY(b2c_jpeg_huff_workspace_bytes(huffs, n, &need));
  if (ws_bytes < need) return set_error(B2C_ERR_WORKSPACE, "b2c_jpeg_huff_decode: workspace %zu B < required %zu B", ws_bytes, need);
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "b2c_jpeg_huff_decode: workspace must be 256-byte aligned");
  uint8_t* wsb = static_cast<uint8_t*>(ws);
  std::vector<HuffJobDev> jobs(n);
  std::vector<b2c_jpeg_hufftab> tabs(static_cast<size_t>(n) * 4);
  size_t off = 0;
  HuffJobDev* jd = reinterpret_cast<HuffJobDev*>(wsb + off);
  off += h256(static_cast<size_t>(n) * sizeof(HuffJobDev));
  b2c_jpeg_hufftab* td = reinterpret_cast<b2c_jpeg_hufftab*>(wsb + off);
  off += h256(static_cast<size_t>(n) * 4 * sizeof(b2c_jpeg_hufftab));
  for (int i = 0; i < n; ++i) {
    const b2c_jpeg_info& I = infos[i];
    const b2c_jpeg_huff& H = huffs[i];
    B2C_REQUIRE(files[i] && coefs[i], "b2c_jpeg_huff_decode: null buffer for image %d", i);
    B2C_REQUIRE((reinterpret_cast<uintptr_t>(coefs[i]) & 15) == 0, "b2c_jpeg_huff_decode: coefficient buffer %d is not 16-byte aligned", i);
    B2C_REQUIRE((I.ncomp == 1 || I.ncomp == 3) && I.mcus_x > 0 && I.mcus_y > 0 && I.nblocks > 0 &&
                    static_cast<int64_t>(I.nblocks) * 64 == I.coef_count && I.hs[0] >= 1 && I.hs[0] <= 2 && I.vs[0] >= 1 &&
                    I.vs[0] <= 2 && H.restart_interval >= 0,
                "b2c_jpeg_huff_decode: bad info for image %d", i);
    HuffJobDev& J = jobs[i];
    memset(&J, 0, sizeof(J));
    J.src = files[i] + H.scan_begin;
    J.raw_len = static_cast<uint32_t>(H.scan_bytes);
    J.max_segs = max_segments(H);
    J.clean = wsb + off;
    off += h256(static_cast<size_t>(H.scan_bytes) + 64);
    J.seg_begin = reinterpret_cast<uint32_t*>(wsb + off);
    off += h256((static_cast<size_t>(J.max_segs) + 2) * sizeof(uint32_t));
    J.sub_base = reinterpret_cast<uint32_t*>(wsb + off);
    off += h256((static_cast<size_t>(J.max_segs) + 2) * sizeof(uint32_t));
    J.coefs = coefs[i];
    J.tabs = td + static_cast<size_t>(i) * 4;
    J.ncomp = I.ncomp;
    J.hs0 = I.hs[0];
    J.vs0 = I.vs[0];
    J.mcus_x = I.mcus_x;
    J.mcus_y = I.mcus_y;
    J.restart_interval = H.restart_interval;
    for (int c = 0; c < I.ncomp; ++c) {
      J.blocks_w[c] = I.blocks_w[c];
      J.block_off[c] = static_cast<int32_t>(I.coef_offset[c] / 64);
    }
    J.coef_count = I.coef_count;
    memcpy(&tabs[static_cast<size_t>(i) * 4], H.tab, 4 * sizeof(b2c_jpeg_hufftab));
  }
  B2C_TRY(upload_async(jd, jobs.data(), static_cast<size_t>(n) * sizeof(HuffJobDev), stream));
  B2C_TRY(upload_async(td, tabs.data(), tabs.size() * sizeof(b2c_jpeg_hufftab), stream));
  ProfScope ps(B2C_PROF_OTHER, stream);
  jpeg_huff_kernel<<<static_cast<unsigned>(n), kHuffThreads, 0, stream>>>(jd, status);
  B2C_POST_LAUNCH("jpeg_huff_kernel");
  return 0;
}
