// b2c_jpeg.cu — K14: JPEG decode for the embedding path (the step immediately before K0).
// Replaces `Image.open(path).convert('RGB')` of CustomImageDataset.__getitem__ (utils/embedder.py:167) for Huffman-coded
// 8-bit JPEG files (baseline, extended sequential, progressive), bit-exactly with what Pillow returns (libjpeg-turbo,
// default settings: JDCT_ISLOW, fancy upsampling, YCbCr -> RGB with 16-bit fixed-point tables):
//   host   b2c_jpeg_parse / b2c_jpeg_decode_coefs : marker parse + Huffman decode of every scan (sequential blocks;
//          progressive DC / AC first and refinement passes with end-of-band runs) -> int16 coefficient blocks (natural
//          order) per component.  The only inherently serial stage; runs on the DataLoader workers.
//   device b2c_jpeg_reconstruct : dequantise + 8x8 inverse DCT (the 13-bit "islow" integer transform) -> component
//          planes; triangle-filter chroma upsampling (h2v1 / h2v2) + colour conversion -> uint8 [H,W,3] in HBM, the
//          layout b2c_preprocess_4crop and b2c_image_stats read.  Batched: one launch pair for a list of images.
// Anything else (arithmetic coding, lossless, 12-bit, CMYK / Adobe RGB, exotic sampling) is refused with
// B2C_ERR_UNSUPPORTED so that the caller keeps the file on the Pillow path.  The arithmetic below restates the
// published libjpeg algorithms (jidctint.c, jdsample.c, jdcolor.c, jdhuff.c, jdphuff.c of libjpeg-turbo 3.1, the library Pillow 12.2
// links); it is checked against Pillow itself in tests/test_jpeg.py.
#include <string.h>

#include <vector>

#include "b2c_launch.h"

namespace b2c {

namespace {

// ------------------------------------------------------------------------------------------------ host: parse
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  // 9-bit lookahead: (length << 8) | symbol, 0 = longer than 9 bits
  uint16_t look[512];
  int32_t maxcode[18];  // maxcode[l] = largest code of length l (-1 if none); maxcode[17] = sentinel
  int32_t valoffset[17];
  uint8_t vals[256];
  // AC tables only: a 9-bit window that holds a whole (code, magnitude bits) pair decodes in one lookup:
  // (value << 8) | (run << 4) | total bits, 0 = not applicable
  int16_t fast_ac[512];
};

struct Scan {
  int ns = 0;
  int ci[3];          // component indices of the frame
  int td[3], ta[3];   // Huffman table ids
  int ss = 0, se = 63, ah = 0, al = 0;
  size_t begin = 0;   // first entropy-coded byte
};

struct Parsed {
  b2c_jpeg_info info;
  HuffTable dc[4], ac[4];
  uint16_t qt[4][64];  // natural order
  bool qt_present[4] = {false, false, false, false};
  int comp_id[3], comp_tq[3];
  bool latched[3] = {false, false, false};  // quantisation table captured at the component's first scan (jdinput.c)
  bool have_sof = false;
  bool progressive = false;
};

int fail_unsupported(const char* what) { return set_error(B2C_ERR_UNSUPPORTED, "jpeg: unsupported stream (%s)", what); }
int fail_corrupt(const char* what) { return set_error(B2C_ERR_ARG, "jpeg: corrupt stream (%s)", what); }

int build_huff(HuffTable& t, const uint8_t* counts, const uint8_t* vals, int nvals) {
  t.present = true;
  memcpy(t.vals, vals, nvals);
  memset(t.look, 0, sizeof(t.look));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t.valoffset[l] = k - code;
    if (counts[l - 1]) {
      for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code) {
        if (l <= 9) {
          const int base = code << (9 - l);
          for (int j = 0; j < (1 << (9 - l)); ++j) t.look[base + j] = static_cast<uint16_t>((l << 8) | vals[k]);
        }
      }
      t.maxcode[l] = code - 1;
      if (code > (1 << l)) return fail_corrupt("bad Huffman code lengths");
    } else {
      t.maxcode[l] = -1;
    }
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
  for (int i = 0; i < 512; ++i) {
    t.fast_ac[i] = 0;
    const uint16_t e = t.look[i];
    if (!e) continue;
    const int len = e >> 8, rs = e & 0xFF, run = rs >> 4, mag = rs & 15;
    if (mag == 0 || len + mag > 9) continue;
    int k = ((i << len) & 511) >> (9 - mag);  // the magnitude bits that follow the code inside the window
    if (k < (1 << (mag - 1))) k += (-1 << mag) + 1;
    if (k >= -128 && k <= 127) t.fast_ac[i] = static_cast<int16_t>((k * 256) + (run * 16) + (len + mag));
  }
  return 0;
}

int frame_geometry(Parsed& P) {
  b2c_jpeg_info& I = P.info;
  int hmax = 1, vmax = 1;
  for (int c = 0; c < I.ncomp; ++c) {
    hmax = I.hs[c] > hmax ? I.hs[c] : hmax;
    vmax = I.vs[c] > vmax ? I.vs[c] : vmax;
  }
  if (I.ncomp == 1) {  // a single-component scan is never interleaved: the MCU is one block, sampling factors are moot
    I.hs[0] = I.vs[0] = 1;
    hmax = vmax = 1;
  } else {
    // supported chroma layouts: 4:4:4 (1x1), 4:2:2 (2x1), 4:2:0 (2x2); both chroma components alike
    const bool ok = I.hs[1] == 1 && I.vs[1] == 1 && I.hs[2] == 1 && I.vs[2] == 1 &&
                    ((I.hs[0] == 1 && I.vs[0] == 1) || (I.hs[0] == 2 && I.vs[0] == 1) || (I.hs[0] == 2 && I.vs[0] == 2));
    if (!ok) return fail_unsupported("chroma sampling other than 4:4:4, 4:2:2, 4:2:0");
  }
  I.mcus_x = (I.width + 8 * hmax - 1) / (8 * hmax);
  I.mcus_y = (I.height + 8 * vmax - 1) / (8 * vmax);
  int64_t off = 0;
  for (int c = 0; c < I.ncomp; ++c) {
    I.blocks_w[c] = I.mcus_x * I.hs[c];
    I.blocks_h[c] = I.mcus_y * I.vs[c];
    I.comp_w[c] = (I.width * I.hs[c] + hmax - 1) / hmax;   // libjpeg's downsampled_width / _height: the real samples
    I.comp_h[c] = (I.height * I.vs[c] + vmax - 1) / vmax;
    I.coef_offset[c] = off;
    off += static_cast<int64_t>(I.blocks_w[c]) * I.blocks_h[c] * 64;
  }
  I.coef_count = off;
  // packed form: [value offsets u32 x nblocks | counts u8 x nblocks | pad to 16 | values i16 ...], a multiple of 16 bytes
  I.nblocks = static_cast<int32_t>(off / 64);
  I.offs_off = 0;
  I.counts_off = 4ll * I.nblocks;
  I.vals_off = (I.counts_off + I.nblocks + 15) & ~15ll;
  I.packed_capacity = I.vals_off + 2 * off;  // worst case: every block keeps all 64 coefficients
  I.packed_bytes = 0;
  if (I.ncomp == 3 && I.comp_w[1] < 2 && I.hs[0] == 2) return fail_unsupported("image narrower than two chroma samples");
  return 0;
}

// Walks the marker segments from `pos` (tables, frame header, ...) up to and including the next SOS header.
// Returns 0 with `scan` filled, 1 at EOI / end of data, < 0 on error.
int next_scan(const uint8_t* d, size_t n, size_t& pos, Parsed& P, Scan& scan) {
  b2c_jpeg_info& I = P.info;
  while (true) {
    // skip anything up to the next marker (padding, or the unread tail of an entropy-coded segment)
    while (pos + 1 < n && !(d[pos] == 0xFF && d[pos + 1] != 0x00 && d[pos + 1] != 0xFF)) ++pos;
    if (pos + 1 >= n) return P.have_sof ? 1 : fail_corrupt("truncated before SOS");
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;  // standalone markers
    if (m == 0xD9) return 1;
    if (pos + 2 > n) return fail_corrupt("truncated segment");
    const size_t len = (static_cast<size_t>(d[pos]) << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) return fail_corrupt("segment length");
    const uint8_t* s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // baseline / extended sequential / progressive, Huffman
      if (P.have_sof) return fail_corrupt("two frame headers");
      if (sl < 6) return fail_corrupt("SOF length");
      if (s[0] != 8) return fail_unsupported("sample precision other than 8 bits");
      I.height = (s[1] << 8) | s[2];
      I.width = (s[3] << 8) | s[4];
      I.ncomp = s[5];
      if (I.height <= 0 || I.width <= 0) return fail_unsupported("zero dimension (DNL)");
      if (I.ncomp != 1 && I.ncomp != 3) return fail_unsupported("component count other than 1 or 3");
      if (sl < 6 + 3 * static_cast<size_t>(I.ncomp)) return fail_corrupt("SOF length");
      for (int c = 0; c < I.ncomp; ++c) {
        P.comp_id[c] = s[6 + 3 * c];
        I.hs[c] = s[7 + 3 * c] >> 4;
        I.vs[c] = s[7 + 3 * c] & 15;
        P.comp_tq[c] = s[8 + 3 * c];
        if (I.hs[c] < 1 || I.hs[c] > 4 || I.vs[c] < 1 || I.vs[c] > 4 || P.comp_tq[c] > 3) return fail_corrupt("sampling factors");
      }
      if (I.adobe_transform0 && I.ncomp == 3) return fail_unsupported("Adobe RGB (untransformed) colour space");
      P.have_sof = true;
      P.progressive = m == 0xC2;
      I.progressive = P.progressive ? 1 : 0;
      B2C_TRY(frame_geometry(P));
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return fail_unsupported("lossless / hierarchical / arithmetic frame type");
    } else if (m == 0xCC) {
      return fail_unsupported("arithmetic coding conditioning");
    } else if (m == 0xC4) {  // DHT
      size_t q = 0;
      while (q < sl) {
        if (q + 17 > sl) return fail_corrupt("DHT length");
        const int tc = s[q] >> 4, th = s[q] & 15;
        if (tc > 1 || th > 3) return fail_corrupt("DHT table id");
        int nv = 0;
        for (int i = 0; i < 16; ++i) nv += s[q + 1 + i];
        if (nv > 256 || q + 17 + nv > sl) return fail_corrupt("DHT values");
        B2C_TRY(build_huff(tc ? P.ac[th] : P.dc[th], s + q + 1, s + q + 17, nv));
        q += 17 + nv;
      }
    } else if (m == 0xDB) {  // DQT
      size_t q = 0;
      while (q < sl) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        if (tq > 3 || pq > 1) return fail_corrupt("DQT table id");
        const size_t need = 1 + (pq ? 128 : 64);
        if (q + need > sl) return fail_corrupt("DQT length");
        for (int i = 0; i < 64; ++i)
          P.qt[tq][kZigzag[i]] = pq ? static_cast<uint16_t>((s[q + 1 + 2 * i] << 8) | s[q + 2 + 2 * i]) : s[q + 1 + i];
        P.qt_present[tq] = true;
        q += need;
      }
    } else if (m == 0xDD) {  // DRI
      if (sl < 2) return fail_corrupt("DRI length");
      I.restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xEE) {  // APP14 Adobe: transform 0 with three components means RGB data, not YCbCr
      if (sl >= 12 && memcmp(s, "Adobe", 5) == 0 && s[11] == 0) {
        if (P.have_sof && I.ncomp == 3) return fail_unsupported("Adobe RGB (untransformed) colour space");
        I.adobe_transform0 = 1;
      }
    } else if (m == 0xDA) {  // SOS
      if (!P.have_sof) return fail_corrupt("SOS before SOF");
      if (sl < 1) return fail_corrupt("SOS length");
      scan.ns = s[0];
      if (scan.ns < 1 || scan.ns > I.ncomp) return fail_corrupt("scan component count");
      // an interleaved scan of SOME of the components has its own MCU geometry (JPEG A.2.3): not covered here
      if (scan.ns > 1 && scan.ns != I.ncomp) return fail_unsupported("interleaved scan of a subset of the components");
      if (sl < 1 + 2 * static_cast<size_t>(scan.ns) + 3) return fail_corrupt("SOS length");
      for (int k = 0; k < scan.ns; ++k) {
        int ci = -1;
        for (int c = 0; c < I.ncomp; ++c)
          if (P.comp_id[c] == s[1 + 2 * k]) ci = c;
        if (ci < 0 || (k > 0 && ci <= scan.ci[k - 1])) return fail_corrupt("scan component selector");
        scan.ci[k] = ci;
        scan.td[k] = s[2 + 2 * k] >> 4;
        scan.ta[k] = s[2 + 2 * k] & 15;
        if (scan.td[k] > 3 || scan.ta[k] > 3) return fail_corrupt("scan Huffman table id");
        if (!P.latched[ci]) {
          if (!P.qt_present[P.comp_tq[ci]]) return fail_corrupt("component refers to an undefined quantisation table");
          for (int i = 0; i < 64; ++i) I.qt[ci][i] = P.qt[P.comp_tq[ci]][i];
          P.latched[ci] = true;
        }
      }
      const uint8_t* tail = s + 1 + 2 * scan.ns;
      scan.ss = tail[0];
      scan.se = tail[1];
      scan.ah = tail[2] >> 4;
      scan.al = tail[2] & 15;
      if (P.progressive) {
        const bool dc = scan.ss == 0;
        if (scan.ss > scan.se || scan.se > 63 || (dc && scan.se != 0) || (!dc && scan.ns != 1) || scan.al > 13 ||
            (scan.ah != 0 && scan.ah != scan.al + 1))
          return fail_corrupt("progressive scan parameters");
      } else if (scan.ss != 0 || scan.se != 63 || scan.ah != 0 || scan.al != 0) {
        return fail_corrupt("spectral selection in a sequential frame");
      }
      for (int k = 0; k < scan.ns; ++k) {
        const bool need_dc = scan.ss == 0 && scan.ah == 0, need_ac = scan.se > 0;
        if ((need_dc && !P.dc[scan.td[k]].present) || (need_ac && !P.ac[scan.ta[k]].present))
          return fail_corrupt("scan refers to an undefined Huffman table");
      }
      pos += len;
      scan.begin = pos;
      return 0;
    }
    pos += len;
  }
}

int parse(const uint8_t* d, size_t n, Parsed& P, Scan& first, size_t& pos) {
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail_corrupt("no SOI marker");
  memset(&P.info, 0, sizeof(P.info));
  pos = 2;
  const int rc = next_scan(d, n, pos, P, first);
  if (rc == 1) return fail_corrupt("no scan");
  return rc;
}

// ------------------------------------------------------------------------------------------------ host: Huffman
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;  // bits left-aligned at bit 63
  int nbits = 0;
  bool hit_marker = false;
  int64_t pad_bits = 0;  // zero bits appended past a marker / the end of the data since the last restart

  // Bits the decoder consumed that were never in the file: libjpeg feeds zeros there and warns ("premature end of data
  // segment"), and Pillow raises OSError for a truncated file — such a stream is not ours to decode silently.
  bool consumed_padding() const { return pad_bits > nbits; }

  void fill() {
    // fast path: eight stream bytes without an 0xFF among them are appended whole
    if (!hit_marker && end - p >= 8 && nbits <= 56) {
      uint64_t v;
      memcpy(&v, p, 8);
      v = __builtin_bswap64(v);
      const uint64_t x = ~v;  // a zero byte of x is an 0xFF byte of v
      if (!((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull)) {
        const int take = (64 - nbits) >> 3;
        const int have = nbits + 8 * take;
        acc |= v >> nbits;
        if (have < 64) acc &= ~0ull << (64 - have);
        p += take;
        nbits = have;
        return;
      }
    }
    while (nbits <= 56) {
      int b = 0;
      if (!hit_marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            hit_marker = true;  // leave p at the marker; feed zeros like libjpeg does past the end of a segment
            b = 0;
            pad_bits += 8;
          }
        } else {
          ++p;
        }
      } else {
        pad_bits += 8;
      }
      acc |= static_cast<uint64_t>(b) << (56 - nbits);
      nbits += 8;
    }
  }
  inline uint32_t peek(int n) { return static_cast<uint32_t>(acc >> (64 - n)); }
  inline void drop(int n) {
    acc <<= n;
    nbits -= n;
  }
  inline int bit() {
    if (nbits < 1) fill();
    const int b = static_cast<int>(acc >> 63);
    drop(1);
    return b;
  }
  inline int bits(int n) {  // n in 1..16
    if (nbits < 16) fill();
    const int v = static_cast<int>(peek(n));
    drop(n);
    return v;
  }
  // byte-align and consume the expected RSTn marker
  int restart(int expect) {
    if (consumed_padding()) return fail_unsupported("entropy-coded segment ends early (truncated or damaged file)");
    nbits = 0;
    acc = 0;
    pad_bits = 0;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF)) ++p;
    if (p + 1 >= end || p[1] != 0xD0 + expect) return fail_corrupt("restart marker missing");
    p += 2;
    hit_marker = false;
    return 0;
  }
};

inline int decode_symbol(BitReader& br, const HuffTable& t) {
  if (br.nbits < 32) br.fill();
  const uint16_t e = t.look[br.peek(9)];
  if (e) {
    br.drop(e >> 8);
    return e & 0xFF;
  }
  int l = 10;
  int32_t code = static_cast<int32_t>(br.peek(10));
  while (code > t.maxcode[l]) {
    ++l;
    if (l > 16) return -1;
    code = static_cast<int32_t>(br.peek(l));
  }
  br.drop(l);
  return t.vals[(code + t.valoffset[l]) & 0xFF];
}

inline int receive_extend(BitReader& br, int s) {
  if (br.nbits < 16) br.fill();
  const int v = static_cast<int>(br.peek(s));
  br.drop(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// one sequential block: DC difference + AC run/size pairs up to EOB
inline int block_sequential(BitReader& br, const HuffTable& dct, const HuffTable& act, int& pred, int16_t* blk) {
  int s = decode_symbol(br, dct);
  if (s < 0 || s > 11) return fail_corrupt("bad DC code");
  if (s) pred += receive_extend(br, s);
  blk[0] = static_cast<int16_t>(pred);
  for (int k = 1; k < 64;) {
    if (br.nbits < 32) br.fill();
    const int16_t fa = act.fast_ac[br.peek(9)];
    if (fa) {  // code and magnitude bits inside the 9-bit window
      k += (fa >> 4) & 15;
      if (k > 63) return fail_corrupt("AC run past the block");
      br.drop(fa & 15);
      blk[k++] = static_cast<int16_t>(fa >> 8);
      continue;
    }
    const int rs = decode_symbol(br, act);
    if (rs < 0) return fail_corrupt("bad AC code");
    const int r = rs >> 4;
    s = rs & 15;
    if (s == 0) {
      if (r != 15) break;  // EOB
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) return fail_corrupt("AC run past the block");
    blk[k] = static_cast<int16_t>(receive_extend(br, s));
    ++k;
  }
  return 0;
}

// progressive: first pass over the AC band [ss, se] of one block (jdphuff.c decode_mcu_AC_first)
inline int block_ac_first(BitReader& br, const HuffTable& act, const Scan& sc, int& eobrun, int16_t* blk) {
  if (eobrun > 0) {
    --eobrun;
    return 0;
  }
  for (int k = sc.ss; k <= sc.se; ++k) {
    const int rs = decode_symbol(br, act);
    if (rs < 0) return fail_corrupt("bad AC code");
    const int r = rs >> 4, s = rs & 15;
    if (s) {
      k += r;
      if (k > 63) return fail_corrupt("AC run past the block");
      blk[k] = static_cast<int16_t>(receive_extend(br, s) * (1 << sc.al));
    } else if (r == 15) {
      k += 15;
    } else {
      eobrun = 1 << r;
      if (r) eobrun += br.bits(r);
      --eobrun;
      break;
    }
  }
  return 0;
}

// progressive: refinement pass over the AC band of one block (jdphuff.c decode_mcu_AC_refine)
inline int block_ac_refine(BitReader& br, const HuffTable& act, const Scan& sc, int& eobrun, int16_t* blk) {
  const int p1 = 1 << sc.al, m1 = -(1 << sc.al);
  int k = sc.ss;
  if (eobrun == 0) {
    for (; k <= sc.se; ++k) {
      const int rs = decode_symbol(br, act);
      if (rs < 0) return fail_corrupt("bad AC code");
      int r = rs >> 4, s = rs & 15;
      if (s) {
        s = br.bit() ? p1 : m1;  // the size of a newly non-zero coefficient is always 1
      } else if (r != 15) {
        eobrun = 1 << r;
        if (r) eobrun += br.bits(r);
        break;  // the rest of the band is handled as an end-of-band run below
      }
      // skip r still-zero coefficients, appending a correction bit to every already non-zero one on the way
      do {
        int16_t* c = blk + k;
        if (*c != 0) {
          if (br.bit() && (*c & p1) == 0) *c = static_cast<int16_t>(*c + (*c >= 0 ? p1 : m1));
        } else if (--r < 0) {
          break;
        }
        ++k;
      } while (k <= sc.se);
      if (s) {
        if (k > 63) return fail_corrupt("AC run past the block");
        blk[k] = static_cast<int16_t>(s);
      }
    }
  }
  if (eobrun > 0) {
    for (; k <= sc.se; ++k) {
      int16_t* c = blk + k;
      if (*c != 0 && br.bit() && (*c & p1) == 0) *c = static_cast<int16_t>(*c + (*c >= 0 ? p1 : m1));
    }
    --eobrun;
  }
  return 0;
}

// One scan of any kind.  Interleaved scans (ns > 1) walk MCUs; a single-component scan walks that component's own
// ceil(comp/8) block grid (JPEG A.2.3), which is smaller than the MCU-padded grid the coefficient buffer uses.
int decode_scan(const uint8_t* d, size_t n, const Parsed& P, const Scan& sc, int16_t* coefs, size_t& end_pos) {
  const b2c_jpeg_info& I = P.info;
  BitReader br{d + sc.begin, d + n};
  int pred[3] = {0, 0, 0};
  int eobrun = 0;
  int restarts_left = I.restart_interval;
  int next_rst = 0;
  const bool interleaved = sc.ns > 1;
  const int c0 = sc.ci[0];
  const int units_x = interleaved ? I.mcus_x : (I.comp_w[c0] + 7) / 8;
  const int units_y = interleaved ? I.mcus_y : (I.comp_h[c0] + 7) / 8;
  for (int uy = 0; uy < units_y; ++uy) {
    for (int ux = 0; ux < units_x; ++ux) {
      if (I.restart_interval && restarts_left == 0) {
        B2C_TRY(br.restart(next_rst));
        next_rst = (next_rst + 1) & 7;
        pred[0] = pred[1] = pred[2] = 0;
        eobrun = 0;
        restarts_left = I.restart_interval;
      }
      for (int k = 0; k < sc.ns; ++k) {
        const int c = sc.ci[k];
        const int nv = interleaved ? I.vs[c] : 1, nh = interleaved ? I.hs[c] : 1;
        for (int v = 0; v < nv; ++v) {
          for (int h = 0; h < nh; ++h) {
            const int by = interleaved ? uy * I.vs[c] + v : uy, bx = interleaved ? ux * I.hs[c] + h : ux;
            int16_t* blk = coefs + I.coef_offset[c] + (static_cast<int64_t>(by) * I.blocks_w[c] + bx) * 64;
            if (!P.progressive) {
              B2C_TRY(block_sequential(br, P.dc[sc.td[k]], P.ac[sc.ta[k]], pred[c], blk));
            } else if (sc.ss == 0) {
              if (sc.ah == 0) {  // DC first
                const int s = decode_symbol(br, P.dc[sc.td[k]]);
                if (s < 0 || s > 11) return fail_corrupt("bad DC code");
                if (s) pred[c] += receive_extend(br, s);
                blk[0] = static_cast<int16_t>(pred[c] * (1 << sc.al));
              } else if (br.bit()) {  // DC refinement: one more bit of precision
                blk[0] = static_cast<int16_t>(blk[0] | (1 << sc.al));
              }
            } else if (sc.ah == 0) {
              B2C_TRY(block_ac_first(br, P.ac[sc.ta[k]], sc, eobrun, blk));
            } else {
              B2C_TRY(block_ac_refine(br, P.ac[sc.ta[k]], sc, eobrun, blk));
            }
          }
        }
      }
      if (I.restart_interval) --restarts_left;
    }
  }
  if (br.consumed_padding()) return fail_unsupported("entropy-coded segment ends early (truncated or damaged file)");
  end_pos = static_cast<size_t>(br.p - d);
  return 0;
}

// every scan of the file, in order; sequential frames stop once each component has been coded.  Blocks come out in
// ZIGZAG order (coefficient k of the scan order at blk[k]): the packed form keeps that order, the dense API permutes.
int decode_all(const uint8_t* d, size_t n, Parsed& P, Scan scan, size_t pos, int16_t* coefs) {
  memset(coefs, 0, static_cast<size_t>(P.info.coef_count) * sizeof(int16_t));
  bool seen[3] = {false, false, false};
  for (int nscans = 0; nscans < 4096; ++nscans) {
    size_t end_pos = 0;
    B2C_TRY(decode_scan(d, n, P, scan, coefs, end_pos));
    for (int k = 0; k < scan.ns; ++k) seen[scan.ci[k]] = true;
    if (!P.progressive) {
      bool all = true;
      for (int c = 0; c < P.info.ncomp; ++c) all = all && seen[c];
      if (all) return 0;
    }
    pos = end_pos;
    const int rc = next_scan(d, n, pos, P, scan);
    if (rc < 0) return rc;
    if (rc == 1) break;
  }
  for (int c = 0; c < P.info.ncomp; ++c)
    if (!seen[c]) return fail_corrupt("a component has no scan");
  return 0;
}

// zigzag-ordered dense blocks -> natural (row-major) order in place: what the dense API and the oracle consume
void dezigzag_in_place(int16_t* coefs, int64_t nblocks) {
  int16_t tmp[64];
  for (int64_t b = 0; b < nblocks; ++b) {
    int16_t* blk = coefs + b * 64;
    memcpy(tmp, blk, sizeof(tmp));
    for (int k = 0; k < 64; ++k) blk[kZigzag[k]] = tmp[k];
  }
}

// zigzag-ordered dense blocks -> packed form (see frame_geometry); returns the bytes used (a multiple of 16)
int64_t pack_blocks(const b2c_jpeg_info& I, const int16_t* dense, uint8_t* out) {
  uint32_t* offs = reinterpret_cast<uint32_t*>(out + I.offs_off);
  uint8_t* counts = out + I.counts_off;
  int16_t* vals = reinterpret_cast<int16_t*>(out + I.vals_off);
  uint32_t nv = 0;
  for (int32_t b = 0; b < I.nblocks; ++b) {
    const int16_t* blk = dense + static_cast<int64_t>(b) * 64;
    int n = 64;
    while (n > 0) {  // last non-zero coefficient in scan order, four at a time
      uint64_t w;
      memcpy(&w, blk + n - 4, 8);
      if (w) break;
      n -= 4;
    }
    while (n > 0 && blk[n - 1] == 0) --n;
    offs[b] = nv;
    counts[b] = static_cast<uint8_t>(n);
    memcpy(vals + nv, blk, static_cast<size_t>(n) * 2);
    nv += n;
  }
  memset(out + I.counts_off + I.nblocks, 0, static_cast<size_t>(I.vals_off - I.counts_off - I.nblocks));
  const int64_t used = I.vals_off + 2ll * nv;
  const int64_t padded = (used + 15) & ~15ll;
  memset(out + used, 0, static_cast<size_t>(padded - used));
  return padded;
}

// Single interleaved sequential scan (what almost every baseline file is): blocks are appended to the packed value
// stream in decode order as they are decoded — no dense scratch, no second pass.  vout has room for 64 values.
inline int block_sequential_append(BitReader& br, const HuffTable& dct, const HuffTable& act, int& pred, int16_t* vout, int& count) {
  int s = decode_symbol(br, dct);
  if (s < 0 || s > 11) return fail_corrupt("bad DC code");
  if (s) pred += receive_extend(br, s);
  vout[0] = static_cast<int16_t>(pred);
  int filled = 1;                 // vout[0, filled) is written
  int last = pred != 0 ? 0 : -1;  // last non-zero position
  for (int k = 1; k < 64;) {
    if (br.nbits < 32) br.fill();
    const int16_t fa = act.fast_ac[br.peek(9)];
    int v;
    if (fa) {
      k += (fa >> 4) & 15;
      if (k > 63) return fail_corrupt("AC run past the block");
      br.drop(fa & 15);
      v = fa >> 8;
    } else {
      const int rs = decode_symbol(br, act);
      if (rs < 0) return fail_corrupt("bad AC code");
      const int r = rs >> 4;
      s = rs & 15;
      if (s == 0) {
        if (r != 15) break;  // EOB
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return fail_corrupt("AC run past the block");
      v = receive_extend(br, s);
    }
    while (filled < k) vout[filled++] = 0;
    vout[k] = static_cast<int16_t>(v);
    filled = k + 1;
    last = k;
    ++k;
  }
  count = last + 1;
  return 0;
}

int decode_scan_append(const uint8_t* d, size_t n, const Parsed& P, const Scan& sc, uint8_t* out, int64_t& used) {
  const b2c_jpeg_info& I = P.info;
  uint32_t* offs = reinterpret_cast<uint32_t*>(out + I.offs_off);
  uint8_t* counts = out + I.counts_off;
  int16_t* vals = reinterpret_cast<int16_t*>(out + I.vals_off);
  BitReader br{d + sc.begin, d + n};
  int pred[3] = {0, 0, 0};
  int restarts_left = I.restart_interval;
  int next_rst = 0;
  uint32_t nv = 0;
  const bool interleaved = sc.ns > 1;
  const int units_x = interleaved ? I.mcus_x : (I.comp_w[0] + 7) / 8;
  const int units_y = interleaved ? I.mcus_y : (I.comp_h[0] + 7) / 8;
  if (!interleaved || units_x * I.hs[0] != I.blocks_w[0] || units_y * I.vs[0] != I.blocks_h[0])
    memset(out, 0, static_cast<size_t>(I.vals_off));  // blocks the scan does not visit keep count 0
  for (int uy = 0; uy < units_y; ++uy) {
    for (int ux = 0; ux < units_x; ++ux) {
      if (I.restart_interval && restarts_left == 0) {
        B2C_TRY(br.restart(next_rst));
        next_rst = (next_rst + 1) & 7;
        pred[0] = pred[1] = pred[2] = 0;
        restarts_left = I.restart_interval;
      }
      for (int k = 0; k < sc.ns; ++k) {
        const int c = sc.ci[k];
        const int nvb = interleaved ? I.vs[c] : 1, nhb = interleaved ? I.hs[c] : 1;
        for (int v = 0; v < nvb; ++v) {
          for (int h = 0; h < nhb; ++h) {
            const int by = interleaved ? uy * I.vs[c] + v : uy, bx = interleaved ? ux * I.hs[c] + h : ux;
            const int64_t b = I.coef_offset[c] / 64 + static_cast<int64_t>(by) * I.blocks_w[c] + bx;
            int count = 0;
            B2C_TRY(block_sequential_append(br, P.dc[sc.td[k]], P.ac[sc.ta[k]], pred[c], vals + nv, count));
            offs[b] = nv;
            counts[b] = static_cast<uint8_t>(count);
            nv += count;
          }
        }
      }
      if (I.restart_interval) --restarts_left;
    }
  }
  if (br.consumed_padding()) return fail_unsupported("entropy-coded segment ends early (truncated or damaged file)");
  memset(out + I.counts_off + I.nblocks, 0, static_cast<size_t>(I.vals_off - I.counts_off - I.nblocks));
  const int64_t end = I.vals_off + 2ll * nv;
  used = (end + 15) & ~15ll;
  memset(out + end, 0, static_cast<size_t>(used - end));
  return 0;
}

// ------------------------------------------------------------------------------------------------ device
struct JpegJobDev {
  const int16_t* coefs;        // dense form (natural order) or nullptr
  const uint8_t* packed;       // packed form or nullptr
  int64_t offs_off, counts_off, vals_off;
  int32_t nblocks;
  uint8_t* out;
  int32_t out_pitch;
  int32_t width, height, ncomp;
  int32_t hs0, vs0;            // luma sampling (chroma is 1x1)
  int32_t blocks_w[3], blocks_h[3];
  int32_t comp_w[3], comp_h[3];
  int64_t coef_offset[3];
  int64_t plane_offset[3];     // bytes into the plane workspace; plane c is [blocks_h*8][blocks_w*8] uint8
  uint16_t qt[3][64];
  int32_t block_begin;         // first global 8x8-block index of this image (all components, in order; multiple of 32)
  int32_t pixel_tile_begin;    // first global 32x8-pixel tile of this image
};

#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// post-IDCT range limit table of libjpeg (sample_range_limit + CENTERJSAMPLE indexed with x & 1023)
__device__ __forceinline__ uint8_t range_limit(int x) {
  const int i = x & 1023;
  if (i < 128) return static_cast<uint8_t>(i + 128);
  if (i < 512) return 255;
  if (i < 896) return 0;
  return static_cast<uint8_t>(i - 896);
}

// one 1-D pass of the LL&M inverse DCT on eight inputs already scaled for the pass; outputs NOT descaled
__device__ __forceinline__ void idct_1d(const int (&in)[8], int (&o)[8]) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * FIX_0_541196100;
  int tmp2 = z1 + z3 * (-FIX_1_847759065);
  int tmp3 = z1 + z2 * FIX_0_765366865;
  z2 = in[0];
  z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * FIX_1_175875602;
  tmp0 *= FIX_0_298631336;
  tmp1 *= FIX_2_053119869;
  tmp2 *= FIX_3_072711026;
  tmp3 *= FIX_1_501321110;
  z1 *= -FIX_0_899976223;
  z2 *= -FIX_2_562915447;
  z3 *= -FIX_1_961570560;
  z4 *= -FIX_0_390180644;
  z3 += z5;
  z4 += z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  o[0] = tmp10 + tmp3;
  o[7] = tmp10 - tmp3;
  o[1] = tmp11 + tmp2;
  o[6] = tmp11 - tmp2;
  o[2] = tmp12 + tmp1;
  o[5] = tmp12 - tmp1;
  o[3] = tmp13 + tmp0;
  o[4] = tmp13 - tmp0;
}

constexpr int kIdctThreads = 256;  // 32 DCT blocks per CTA, 8 threads per block (a column, then a row)

// job lookup: jobs are sorted by block_begin / pixel_tile_begin; n is small (a batch), binary search
__device__ __forceinline__ int find_job(const JpegJobDev* jobs, int n, int idx, bool tiles) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int b = tiles ? jobs[mid].pixel_tile_begin : jobs[mid].block_begin;
    if (b <= idx) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__constant__ uint8_t kZigzagInv[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                                       41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                                       46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// One CTA = 32 consecutive blocks of ONE image (every image's block range starts at a multiple of 32).
__global__ void __launch_bounds__(kIdctThreads) jpeg_idct_kernel(const JpegJobDev* __restrict__ jobs, int njobs,
                                                                 int total_blocks, uint8_t* __restrict__ planes) {
  __shared__ int ws[32][8][9];  // [block][row][col], padded against bank conflicts
  const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int gb = blockIdx.x * 32 + lb;
  const JpegJobDev* J = jobs + find_job(jobs, njobs, blockIdx.x * 32, false);
  const int b = gb - J->block_begin;  // block index inside the image
  const bool live = gb < total_blocks && b < J->nblocks;
  int c = 0, bx = 0, by = 0;
  if (live) {
    int rem = b;
    while (c + 1 < J->ncomp && rem >= J->blocks_w[c] * J->blocks_h[c]) {
      rem -= J->blocks_w[c] * J->blocks_h[c];
      ++c;
    }
    by = rem / J->blocks_w[c];
    bx = rem - by * J->blocks_w[c];
    // pass 1: column t of the block, dequantised
    const uint16_t* q = J->qt[c];
    int in[8], o[8];
    if (J->packed != nullptr) {
      const int16_t* vals = reinterpret_cast<const int16_t*>(J->packed + J->vals_off) +
                            reinterpret_cast<const uint32_t*>(J->packed + J->offs_off)[b];
      const int cnt = J->packed[J->counts_off + b];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int p = kZigzagInv[r * 8 + t];  // position of natural coefficient (r, t) in scan order
        in[r] = p < cnt ? static_cast<int>(vals[p]) * static_cast<int>(q[r * 8 + t]) : 0;
      }
    } else {
      const int16_t* blk = J->coefs + static_cast<int64_t>(b) * 64;  // blocks are stored in image block order
#pragma unroll
      for (int r = 0; r < 8; ++r) in[r] = static_cast<int>(blk[r * 8 + t]) * static_cast<int>(q[r * 8 + t]);
    }
    idct_1d(in, o);  // (an all-zero AC column gives dc << 2 through the same arithmetic as libjpeg's shortcut)
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[lb][r][t] = descale(o[r], 13 - 2);
  }
  __syncwarp();
  if (live) {
    // pass 2: row t
    int in[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = ws[lb][t][k];
    idct_1d(in, o);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= static_cast<uint32_t>(range_limit(descale(o[k], 13 + 2 + 3))) << (8 * k);
      hi |= static_cast<uint32_t>(range_limit(descale(o[4 + k], 13 + 2 + 3))) << (8 * k);
    }
    const int pw = J->blocks_w[c] * 8;
    uint8_t* dst = planes + J->plane_offset[c] + static_cast<int64_t>(by * 8 + t) * pw + bx * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}

// colour conversion tables of jdcolor.c (SCALEBITS 16): evaluated, not stored
__device__ __forceinline__ int cr_r(int cr) { return (91881 * (cr - 128) + 32768) >> 16; }
__device__ __forceinline__ int cb_b(int cb) { return (116130 * (cb - 128) + 32768) >> 16; }
__device__ __forceinline__ int cr_g(int cr) { return -46802 * (cr - 128); }
__device__ __forceinline__ int cb_g(int cb) { return -22554 * (cb - 128) + 32768; }
__device__ __forceinline__ uint8_t clamp8(int v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// chroma sample for output pixel (x, y): libjpeg's "fancy" triangle filters (jdsample.c)
__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ pl, int pw, int cw, int ch, int hs, int vs, int x,
                                         int y) {
  if (hs == 1) return pl[static_cast<int64_t>(y) * pw + x];
  const int i = x >> 1;
  if (vs == 1) {  // h2v1
    const uint8_t* row = pl + static_cast<int64_t>(y) * pw;
    const int v = row[i];
    if (x & 1) return i == cw - 1 ? v : (3 * v + row[i + 1] + 2) >> 2;
    return i == 0 ? v : (3 * v + row[i - 1] + 1) >> 2;
  }
  // h2v2: nearer row j, further row above (even y) or below (odd y), replicated at the image edges
  const int j = y >> 1;
  int jf = (y & 1) ? j + 1 : j - 1;
  jf = jf < 0 ? 0 : (jf > ch - 1 ? ch - 1 : jf);
  const uint8_t* r0 = pl + static_cast<int64_t>(j) * pw;
  const uint8_t* r1 = pl + static_cast<int64_t>(jf) * pw;
  const int cur = 3 * r0[i] + r1[i];
  if (x & 1) {
    if (i == cw - 1) return (cur * 4 + 7) >> 4;
    return (cur * 3 + (3 * r0[i + 1] + r1[i + 1]) + 7) >> 4;
  }
  if (i == 0) return (cur * 4 + 8) >> 4;
  return (cur * 3 + (3 * r0[i - 1] + r1[i - 1]) + 8) >> 4;
}

constexpr int kPixTileW = 32, kPixTileH = 8;  // one CTA of 256 threads = one 32 x 8 pixel tile

__global__ void __launch_bounds__(kPixTileW* kPixTileH) jpeg_color_kernel(const JpegJobDev* __restrict__ jobs, int njobs,
                                                                          const uint8_t* __restrict__ planes) {
  const JpegJobDev* J = jobs + find_job(jobs, njobs, blockIdx.x, true);
  const int tile = blockIdx.x - J->pixel_tile_begin;
  const int tiles_x = (J->width + kPixTileW - 1) / kPixTileW;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int x = tx * kPixTileW + (threadIdx.x & 31), y = ty * kPixTileH + (threadIdx.x >> 5);
  if (x >= J->width || y >= J->height) return;
  const int yv = planes[J->plane_offset[0] + static_cast<int64_t>(y) * (J->blocks_w[0] * 8) + x];
  uint8_t* o = J->out + static_cast<int64_t>(y) * J->out_pitch + 3 * x;
  if (J->ncomp == 1) {  // Pillow: mode L -> convert('RGB') replicates the sample
    o[0] = o[1] = o[2] = static_cast<uint8_t>(yv);
    return;
  }
  const int cb = chroma_at(planes + J->plane_offset[1], J->blocks_w[1] * 8, J->comp_w[1], J->comp_h[1], J->hs0, J->vs0, x, y);
  const int cr = chroma_at(planes + J->plane_offset[2], J->blocks_w[2] * 8, J->comp_w[2], J->comp_h[2], J->hs0, J->vs0, x, y);
  o[0] = clamp8(yv + cr_r(cr));
  o[1] = clamp8(yv + ((cb_g(cb) + cr_g(cr)) >> 16));
  o[2] = clamp8(yv + cb_b(cb));
}

size_t a256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace
}  // namespace b2c

extern "C" int b2c_jpeg_parse(const uint8_t* data, size_t len, b2c_jpeg_info* info) {
  using namespace b2c;
  B2C_REQUIRE(data && info, "b2c_jpeg_parse: null pointer");
  Parsed P;
  Scan first;
  size_t pos = 0;
  B2C_TRY(parse(data, len, P, first, pos));
  *info = P.info;
  return 0;
}

extern "C" int b2c_jpeg_huff_prepare(const uint8_t* data, size_t len, b2c_jpeg_info* info, b2c_jpeg_huff* huff) {
  using namespace b2c;
  B2C_REQUIRE(data && info && huff, "b2c_jpeg_huff_prepare: null pointer");
  Parsed P;
  Scan first;
  size_t pos = 0;
  B2C_TRY(parse(data, len, P, first, pos));
  if (P.progressive) return fail_unsupported("device Huffman stage: progressive frame");
  if (first.ns != P.info.ncomp) return fail_unsupported("device Huffman stage: the first scan does not hold every component");
  for (int k = 0; k < first.ns; ++k)
    if (first.ci[k] != k) return fail_corrupt("scan component order");
  if (P.info.ncomp == 3 && (first.td[1] != first.td[2] || first.ta[1] != first.ta[2]))
    return fail_unsupported("device Huffman stage: Cb and Cr use different Huffman tables");
  if (len - first.begin >= (1ull << 28)) return fail_unsupported("device Huffman stage: scan larger than 256 MB");
  memset(huff, 0, sizeof(*huff));
  huff->scan_begin = static_cast<int64_t>(first.begin);
  huff->scan_bytes = static_cast<int64_t>(len - first.begin);
  huff->restart_interval = P.info.restart_interval;
  const int cc = P.info.ncomp == 3 ? 1 : 0;  // grey: the chroma slots repeat the luma tables
  const HuffTable* src[4] = {&P.dc[first.td[0]], &P.ac[first.ta[0]], &P.dc[first.td[cc]], &P.ac[first.ta[cc]]};
  for (int t = 0; t < 4; ++t) {
    b2c_jpeg_hufftab& d = huff->tab[t];
    memcpy(d.look, src[t]->look, sizeof(d.look));
    memcpy(d.maxcode, src[t]->maxcode, sizeof(d.maxcode));
    d.valoffset[0] = 0;
    memcpy(d.valoffset + 1, src[t]->valoffset + 1, 16 * sizeof(int32_t));
    d.valoffset[17] = 0;
    memcpy(d.vals, src[t]->vals, sizeof(d.vals));
  }
  *info = P.info;
  return 0;
}

extern "C" int b2c_jpeg_decode_coefs(const uint8_t* data, size_t len, b2c_jpeg_info* info, int16_t* coefs,
                                     size_t capacity) {
  using namespace b2c;
  B2C_REQUIRE(data && info && coefs, "b2c_jpeg_decode_coefs: null pointer");
  Parsed P;
  Scan first;
  size_t pos = 0;
  B2C_TRY(parse(data, len, P, first, pos));
  B2C_REQUIRE(static_cast<size_t>(P.info.coef_count) <= capacity, "b2c_jpeg_decode_coefs: buffer holds %zu coefficients, %lld needed",
              capacity, (long long)P.info.coef_count);
  B2C_TRY(decode_all(data, len, P, first, pos, coefs));
  dezigzag_in_place(coefs, P.info.nblocks);
  *info = P.info;  // quantisation tables are latched scan by scan
  return 0;
}

extern "C" int b2c_jpeg_decode_packed(const uint8_t* data, size_t len, b2c_jpeg_info* info, uint8_t* packed, size_t capacity,
                                      int16_t* scratch) {
  using namespace b2c;
  B2C_REQUIRE(data && info && packed, "b2c_jpeg_decode_packed: null pointer");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "b2c_jpeg_decode_packed: buffer must be 16-byte aligned");
  Parsed P;
  Scan first;
  size_t pos = 0;
  B2C_TRY(parse(data, len, P, first, pos));
  B2C_REQUIRE(static_cast<size_t>(P.info.packed_capacity) <= capacity, "b2c_jpeg_decode_packed: buffer holds %zu bytes, %lld needed",
              capacity, (long long)P.info.packed_capacity);
  if (!P.progressive && first.ns == P.info.ncomp) {
    // one interleaved (or single-component) sequential scan covers the image: append blocks as they are decoded
    for (int k = 0; k < first.ns; ++k)
      if (first.ci[k] != k) return fail_corrupt("scan component order");
    int64_t used = 0;
    B2C_TRY(decode_scan_append(data, len, P, first, packed, used));
    P.info.packed_bytes = used;
    *info = P.info;
    return 0;
  }
  std::vector<int16_t> own;
  if (!scratch) {
    own.resize(static_cast<size_t>(P.info.coef_count));
    scratch = own.data();
  }
  B2C_TRY(decode_all(data, len, P, first, pos, scratch));
  P.info.packed_bytes = pack_blocks(P.info, scratch, packed);
  *info = P.info;
  return 0;
}

extern "C" int b2c_jpeg_workspace_bytes(const b2c_jpeg_info* infos, int n, size_t* bytes) {
  using namespace b2c;
  B2C_REQUIRE(infos && bytes && n > 0, "b2c_jpeg_workspace_bytes: bad arguments");
  size_t total = a256(static_cast<size_t>(n) * sizeof(JpegJobDev));
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < infos[i].ncomp; ++c) total += a256(static_cast<size_t>(infos[i].blocks_w[c]) * infos[i].blocks_h[c] * 64);
  *bytes = total;
  return 0;
}

namespace b2c {
namespace {
int reconstruct_impl(const b2c_jpeg_info* infos, const int16_t* const* coefs, const uint8_t* const* packed, uint8_t* const* outs,
                     const int* out_pitch, int n, void* ws, size_t ws_bytes, cudaStream_t stream) {
  B2C_REQUIRE(infos && (coefs || packed) && outs && out_pitch && ws && n > 0, "b2c_jpeg_reconstruct: bad arguments");
  size_t need = 0;
  B2C_TRY(b2c_jpeg_workspace_bytes(infos, n, &need));
  if (ws_bytes < need) return set_error(B2C_ERR_WORKSPACE, "b2c_jpeg_reconstruct: workspace %zu B < required %zu B", ws_bytes, need);
  std::vector<JpegJobDev> jobs(n);
  size_t off = a256(static_cast<size_t>(n) * sizeof(JpegJobDev));
  long long blocks = 0, tiles = 0;
  for (int i = 0; i < n; ++i) {
    const b2c_jpeg_info& I = infos[i];
    const void* src = packed ? static_cast<const void*>(packed[i]) : static_cast<const void*>(coefs[i]);
    B2C_REQUIRE(src && outs[i] && out_pitch[i] >= 3 * I.width, "b2c_jpeg_reconstruct: bad buffers for image %d", i);
    B2C_REQUIRE((I.ncomp == 1 || I.ncomp == 3) && I.width > 0 && I.height > 0 && I.nblocks > 0 &&
                    static_cast<int64_t>(I.nblocks) * 64 == I.coef_count,
                "b2c_jpeg_reconstruct: bad info for image %d", i);
    if (packed) B2C_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0, "b2c_jpeg_reconstruct_packed: buffer %d is not 16-byte aligned", i);
    JpegJobDev& J = jobs[i];
    memset(&J, 0, sizeof(J));
    J.coefs = packed ? nullptr : coefs[i];
    J.packed = packed ? packed[i] : nullptr;
    J.offs_off = I.offs_off;
    J.counts_off = I.counts_off;
    J.vals_off = I.vals_off;
    J.nblocks = I.nblocks;
    J.out = outs[i];
    J.out_pitch = out_pitch[i];
    J.width = I.width;
    J.height = I.height;
    J.ncomp = I.ncomp;
    J.hs0 = I.hs[0];
    J.vs0 = I.vs[0];
    J.block_begin = static_cast<int32_t>(blocks);
    J.pixel_tile_begin = static_cast<int32_t>(tiles);
    for (int c = 0; c < I.ncomp; ++c) {
      J.blocks_w[c] = I.blocks_w[c];
      J.blocks_h[c] = I.blocks_h[c];
      J.comp_w[c] = I.comp_w[c];
      J.comp_h[c] = I.comp_h[c];
      J.coef_offset[c] = I.coef_offset[c];
      J.plane_offset[c] = static_cast<int64_t>(off);
      off += a256(static_cast<size_t>(I.blocks_w[c]) * I.blocks_h[c] * 64);
      memcpy(J.qt[c], I.qt[c], sizeof(J.qt[c]));
    }
    blocks += (static_cast<long long>(I.nblocks) + 31) & ~31ll;  // every image starts a fresh group of 32 blocks
    tiles += static_cast<long long>((I.width + kPixTileW - 1) / kPixTileW) * ((I.height + kPixTileH - 1) / kPixTileH);
    B2C_REQUIRE(blocks < (1ll << 31) && tiles < (1ll << 31), "b2c_jpeg_reconstruct: batch too large");
  }
  JpegJobDev* jd = static_cast<JpegJobDev*>(ws);
  uint8_t* planes = static_cast<uint8_t*>(ws);
  B2C_TRY(upload_async(jd, jobs.data(), static_cast<size_t>(n) * sizeof(JpegJobDev), stream));  // no stream synchronisation
  ProfScope ps(B2C_PROF_OTHER, stream);
  jpeg_idct_kernel<<<static_cast<unsigned>(blocks / 32), kIdctThreads, 0, stream>>>(jd, n, static_cast<int>(blocks), planes);
  B2C_POST_LAUNCH("jpeg_idct_kernel");
  jpeg_color_kernel<<<static_cast<unsigned>(tiles), kPixTileW * kPixTileH, 0, stream>>>(jd, n, planes);
  B2C_POST_LAUNCH("jpeg_color_kernel");
  return 0;
}
}  // namespace
}  // namespace b2c

extern "C" int b2c_jpeg_reconstruct(const b2c_jpeg_info* infos, const int16_t* const* coefs, uint8_t* const* outs,
                                    const int* out_pitch, int n, void* ws, size_t ws_bytes, b2c_stream stream) {
  return b2c::reconstruct_impl(infos, coefs, nullptr, outs, out_pitch, n, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int b2c_jpeg_reconstruct_packed(const b2c_jpeg_info* infos, const uint8_t* const* packed, uint8_t* const* outs,
                                           const int* out_pitch, int n, void* ws, size_t ws_bytes, b2c_stream stream) {
  return b2c::reconstruct_impl(infos, nullptr, packed, outs, out_pitch, n, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
