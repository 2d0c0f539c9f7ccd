// b2c_preprocess.cu — K0: the reference's 4-crop extraction and per-crop preprocessing, fused and
// bit-exact on the GPU.
//   crop geometry ........ CustomImageDataset.extract_crops, utils/embedder.py:184-251 (host, integers)
//   per-crop transform ... utils/embedder.py:90-92,173 = open_clip val transform: torchvision
//                          Resize(R, BICUBIC) on a PIL image -> CenterCrop(R) -> ToTensor -> Normalize
//   resize arithmetic .... Pillow's ImagingResample: a=-0.5 bicubic, antialias support, 22-bit fixed
//                          point coefficients, horizontal pass -> uint8 -> vertical pass -> uint8
// Two kernels:
//   resample_plan_kernel : per (crop, axis) the Pillow coefficient tables for the R kept outputs,
//                          in fp64 with explicitly rounded operations (no FMA contraction) so the
//                          quantised coefficients equal Pillow's bit for bit;
//   resample_kernel      : per (crop, band of output rows): stage the needed source rows in shared
//                          memory (coalesced word copies), horizontal pass into a planar uint8 tile in
//                          shared memory, vertical pass + ToTensor/Normalize, write either
//                          f32 NCHW (parity layout) or bf16 patch-major (the patch-embed GEMM's A operand).
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "b2c_launch.h"

namespace b2c {

struct CropPlan {
  const uint8_t* img;
  int32_t H, W, pitch;
  int32_t cw, ch, dx, dy, out_w, out_h, off_x, off_y;
  int32_t pad;
};

constexpr int kPrecisionBits = 32 - 8 - 2;

// ------------------------------------------------------------------------------------------------
// host: crop geometry
// ------------------------------------------------------------------------------------------------
static int round_half_even_div2(int a) {  // int(round(a / 2.0)), a >= 0 (Python banker's rounding)
  int h = a / 2;
  if (a & 1) h += (h & 1);
  return h;
}

static void resized_size(int w, int h, int R, int* ow, int* oh) {
  // torchvision Resize(int): shorter side -> R, longer side -> int(R * long / short)
  const int s = w <= h ? w : h, l = w <= h ? h : w;
  const int nl = static_cast<int>(static_cast<double>(static_cast<long long>(R) * l) / static_cast<double>(s));
  if (w <= h) { *ow = R; *oh = nl; } else { *ow = nl; *oh = R; }
}

static void crop_geometry_host(int W, int H, int R, b2c_crop out[4]) {
  memset(out, 0, 4 * sizeof(b2c_crop));
  {  // centre_crop (embedder.py:196-202): torchvision CenterCrop(min(W,H))
    const int s = W < H ? W : H;
    out[0].cw = s; out[0].ch = s;
    out[0].dx = round_half_even_div2(W - s);
    out[0].dy = round_half_even_div2(H - s);
  }
  {  // square_padded_crop (embedder.py:204-212)
    const int S = W > H ? W : H;
    out[1].cw = S; out[1].ch = S;
    out[1].dx = -((S - W) / 2);
    out[1].dy = -((S - H) / 2);
  }
  {  // subcrop1 / subcrop2 (embedder.py:215-247); Python: int((W*H*f) ** 0.5)
    const double area = static_cast<double>(static_cast<long long>(W) * H);
    const int sz[2] = {static_cast<int>(pow(area * 0.15, 0.5)), static_cast<int>(pow(area * 0.1, 0.5))};
    int cx[2], cy[2];
    if (W >= H) { cx[0] = W / 4; cy[0] = H / 2; cx[1] = W / 4 * 3; cy[1] = H / 2; }
    else        { cx[0] = W / 2; cy[0] = H / 4; cx[1] = W / 2; cy[1] = H / 4 * 3; }
    for (int i = 0; i < 2; ++i) {
      int left = cx[i] - sz[i] / 2; if (left < 0) left = 0;
      int top = cy[i] - sz[i] / 2;  if (top < 0) top = 0;
      int right = left + sz[i];     if (right > W) right = W;
      int bottom = top + sz[i];     if (bottom > H) bottom = H;
      b2c_crop& c = out[2 + i];
      if (right - left > 0 && bottom - top > 0) { c.cw = right - left; c.ch = bottom - top; c.dx = left; c.dy = top; }
    }
  }
  for (int i = 0; i < 4; ++i) {
    b2c_crop& c = out[i];
    if (c.cw <= 0) continue;
    resized_size(c.cw, c.ch, R, &c.out_w, &c.out_h);
    c.off_x = round_half_even_div2(c.out_w - R);
    c.off_y = round_half_even_div2(c.out_h - R);
  }
}

static int ksize_for(int in_size, int out_size) {
  double scale = static_cast<double>(in_size) / out_size;
  if (scale < 1.0) scale = 1.0;
  return static_cast<int>(ceil(2.0 * scale)) * 2 + 1;
}

// ------------------------------------------------------------------------------------------------
// device: Pillow coefficient tables
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bicubic_filter(double x) {
  x = fabs(x);
  if (x < 1.0) {
    double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);  // ((a + 2) x - (a + 3)), a = -0.5
    t = __dmul_rn(__dmul_rn(t, x), x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    double t = __dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0);
    t = __dsub_rn(__dmul_rn(t, x), 4.0);
    return __dmul_rn(t, -0.5);
  }
  return 0.0;
}

// Coefficient tables in the form the resample kernel consumes.  A 22-bit Pillow coefficient k is split into three
// byte limbs, k = l0 + 256 l1 + 65536 l2 (l0, l1 unsigned, l2 = k >> 16 signed; |k| < 2^23 because no normalised bicubic
// weight reaches 2), and four consecutive taps share one uint4 {l0 x4, l1 x4, l2 x4, 0}: one `dp4a` per limb then covers
// four taps of four source bytes, exactly (integers).
// grid (n_crops, 2 axes), block R threads. tables: bounds int2[n_crops][2][R] = (xmin, count),
// coefs uint4[n_crops][2][R][KS4].  Block (0, 0) also writes the ToTensor/Normalize table lut f32[3][256].
struct NormParams {
  float mean[3], stdv[3];
};
__global__ void resample_plan_kernel(const CropPlan* __restrict__ plans, int2* __restrict__ bounds,
                                     uint4* __restrict__ coefs, float* __restrict__ lut, const NormParams norm, int R,
                                     int KS4) {
  const int crop = blockIdx.x, axis = blockIdx.y, xx = threadIdx.x;
  if (crop == 0 && axis == 0) {
    for (int i = xx; i < 768; i += blockDim.x) {
      const int c = i >> 8, v = i & 255;
      // ToTensor: float(v) / 255 ; Normalize: (x - mean) / std, each op rounded to fp32 like torch
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), norm.mean[c]), norm.stdv[c]);
    }
  }
  if (xx >= R) return;
  const CropPlan p = plans[crop];
  const int slot = xx;
  int2* b = bounds + (static_cast<size_t>(crop) * 2 + axis) * R + slot;
  uint4* k = coefs + ((static_cast<size_t>(crop) * 2 + axis) * R + slot) * KS4;
  if (p.cw <= 0) { *b = make_int2(0, 0); return; }
  const int in_size = axis ? p.ch : p.cw;
  const int out_size = axis ? p.out_h : p.out_w;
  const int o = (axis ? p.off_y : p.off_x) + xx;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(static_cast<double>(o) + 0.5, scale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    ww = __dadd_rn(ww, bicubic_filter(arg));
  }
  for (int g = 0; g < KS4; ++g) {
    uint32_t l0 = 0, l1 = 0, l2 = 0;
    for (int j = 0; j < 4; ++j) {
      const int x = 4 * g + j;
      if (x >= n) break;
      const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      double w = bicubic_filter(arg);
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double scaled = __dmul_rn(w, static_cast<double>(1 << kPrecisionBits));
      const int kv = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
      l0 |= static_cast<uint32_t>(kv & 255) << (8 * j);
      l1 |= static_cast<uint32_t>((kv >> 8) & 255) << (8 * j);
      l2 |= static_cast<uint32_t>((kv >> 16) & 255) << (8 * j);
    }
    k[g] = make_uint4(l0, l1, l2, 0u);
  }
  *b = make_int2(xmin, n);
}

// ------------------------------------------------------------------------------------------------
// device: fused two-pass resample + normalise
// ------------------------------------------------------------------------------------------------
struct ResampleParams {
  const CropPlan* plans;
  const int2* bounds;
  const uint4* coefs;
  const float* lut;   // f32[3][256]: (v / 255 - mean) / std
  void* out;
  int R, KS4, TR;     // outputs per side, tap groups of four, output rows per band
  int HP;             // bytes per (channel, output column) of the transposed horizontal tile (4 * odd)
  int SR;             // source rows staged per sub-batch (multiple of 4)
  int SPP;            // bytes per staged (channel, row) (multiple of 4)
  int nch;            // colour channels per CTA: 3, or 1 for very large sources (grid.z = 3, tiles a third the size)
  int hsm;            // horizontal coefficient table in shared memory (few taps) or read from global memory / L2
  int layout, patch, Kp;
  int g;              // patches per side (patch-major layout)
  uint32_t magicP;    // floor(2^32 / patch) + 1: n / patch == umulhi(n, magicP) for n < 65536
  uint32_t magicPad;  // same for (Kp - 3 patch^2) / 2
};

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}
// four u8 (a) x four u8 / s8 (b) + c
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// the limb-0 accumulator starts at Pillow's rounding constant 1 << (PRECISION_BITS - 1)
constexpr int kRound = 1 << (kPrecisionBits - 1);
__device__ __forceinline__ int limbs_to_u8(int a0, int a1, int a2) { return clip8(a0 + (a1 << 8) + (a2 << 16)); }

// shared-memory carve-up of resample_kernel (every region 16-byte aligned)
struct PreSmem {
  size_t lut, hcoef, hb, vcoef, vb, hbuf, stage, total;
};
__host__ __device__ inline size_t pre_a16(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline int pre_odd(int v) { return v | 1; }  // uint4 pitch of a coefficient row (bank spread)
__host__ __device__ inline PreSmem pre_smem(int R, int KS4, int TR, int HP, int SR, int SPP, int nch, int hsm) {
  PreSmem s;
  size_t off = 0;
  s.lut = off;   off += pre_a16(768 * sizeof(float));
  s.hcoef = off; off += hsm ? pre_a16(static_cast<size_t>(R) * pre_odd(KS4) * 16) : 0;
  s.hb = off;    off += pre_a16(static_cast<size_t>(R) * 8);
  s.vcoef = off; off += pre_a16(static_cast<size_t>(TR) * KS4 * 16);
  s.vb = off;    off += pre_a16(static_cast<size_t>(TR) * 8);
  s.hbuf = off;  off += pre_a16(static_cast<size_t>(nch) * R * HP);
  s.stage = off; off += pre_a16(static_cast<size_t>(nch) * SR * SPP);
  s.total = off;
  return s;
}

constexpr int kResThreads = 256;

// Horizontal pass over the staged rows [0, 4 * rgroups): work item = (4 source rows, local channel, output column).
// G > 0: tap-group count known at compile time (unrolled, every thread runs the CTA's maximum; the tiles carry the slack
// for reading zero-coefficient groups past a window's end); G == 0: each thread loops over its own window's groups.
// hcoef: coefficient rows of CP uint4 each, in shared memory or (many taps) global memory.
template <int G, int NCH>
__device__ __forceinline__ void resample_hpass(const uint8_t* __restrict__ stage, uint8_t* __restrict__ hbuf,
                                               const uint4* __restrict__ hcoef, const int2* __restrict__ hb, int R,
                                               int CP, int HP, int SRc, int SPP, int rgroups, int y0, int tid) {
  constexpr int nch = NCH;
  int rg = 0, cs = tid;
  const int per = nch * R;
  while (cs >= per) { cs -= per; ++rg; }
  while (rg < rgroups) {
    const int c = cs >= 2 * R ? 2 : (cs >= R ? 1 : 0);  // local channel (0 when the CTA holds one)
    const int x = cs - c * R;  // lanes walk x in natural order: their source windows overlap (few banks, no conflicts)
    const int slot = (x >> 1) + (x & 1) * (R >> 1);  // column of the transposed tile: the pair (2xp, 2xp+1) sits R/2 apart
    const int2 bx = hb[x];     // (first source column relative to the staged span, taps)
    const uint4* kc = hcoef + static_cast<size_t>(x) * CP;
    const int ng = G ? G : (bx.y + 3) >> 2;
    const uint8_t* s0 = stage + (static_cast<size_t>(c) * SRc + 4 * rg) * SPP + (bx.x & ~3);
    const uint32_t sh = (bx.x & 3) * 8;
    int a[4][3];
    uint32_t prev[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j][0] = kRound;
      a[j][1] = a[j][2] = 0;
      prev[j] = *reinterpret_cast<const uint32_t*>(s0 + j * SPP);
    }
#pragma unroll
    for (int g = 0; g < ng; ++g) {
      const uint4 k = kc[g];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t nxt = *reinterpret_cast<const uint32_t*>(s0 + j * SPP + 4 * (g + 1));
        const uint32_t d = __funnelshift_r(prev[j], nxt, sh);
        prev[j] = nxt;
        a[j][0] = dp4a_uu(d, k.x, a[j][0]);
        a[j][1] = dp4a_uu(d, k.y, a[j][1]);
        a[j][2] = dp4a_us(d, k.z, a[j][2]);
      }
    }
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) w |= static_cast<uint32_t>(limbs_to_u8(a[j][0], a[j][1], a[j][2])) << (8 * j);
    *reinterpret_cast<uint32_t*>(hbuf + (static_cast<size_t>(c) * R + slot) * HP + y0 + 4 * rg) = w;
    cs += kResThreads;
    while (cs >= per) { cs -= per; ++rg; }
  }
}

// Vertical pass + ToTensor/Normalize: work item = (output row of the band, channel, output column pair 2xp, 2xp+1).
template <int G, int NCH>
__device__ __forceinline__ void resample_vpass(const ResampleParams& p, const uint8_t* __restrict__ hbuf,
                                               const uint4* __restrict__ vcoef, const int2* __restrict__ vb,
                                               const float* __restrict__ lut, int crop, int r0, int nr, int cbeg,
                                               int tid) {
  const int R = p.R, R2 = R >> 1, HP = p.HP, KS4 = p.KS4;
  const int per = NCH * R2;
  // 32-bit element offsets below one crop's output (< 2^31 elements)
  float* out_f32 = reinterpret_cast<float*>(p.out) + static_cast<size_t>(crop) * 3 * R * R;
  __nv_bfloat16* out_bf16 = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(crop) * p.g * p.g * p.Kp;
  int r = 0, cx = tid;
  while (cx >= per) { cx -= per; ++r; }
  while (r < nr) {
    const int cl = cx >= 2 * R2 ? 2 : (cx >= R2 ? 1 : 0);  // local channel
    const int c = cbeg + cl;
    const int xp = cx - cl * R2;
    const int2 by = vb[r];  // (first tile row, taps)
    const uint4* kc = vcoef + r * KS4;
    const int ng = G ? G : (by.y + 3) >> 2;
    const uint8_t* h0 = hbuf + (static_cast<size_t>(cl) * R + xp) * HP + (by.x & ~3);
    const uint8_t* h1 = h0 + static_cast<size_t>(R2) * HP;
    const uint32_t sh = (by.x & 3) * 8;
    int a0 = kRound, a1 = 0, a2 = 0, b0 = kRound, b1 = 0, b2 = 0;
    uint32_t pa = *reinterpret_cast<const uint32_t*>(h0), pb = *reinterpret_cast<const uint32_t*>(h1);
#pragma unroll
    for (int g = 0; g < ng; ++g) {
      const uint4 k = kc[g];
      const uint32_t na = *reinterpret_cast<const uint32_t*>(h0 + 4 * (g + 1));
      const uint32_t nb = *reinterpret_cast<const uint32_t*>(h1 + 4 * (g + 1));
      const uint32_t da = __funnelshift_r(pa, na, sh), db = __funnelshift_r(pb, nb, sh);
      pa = na; pb = nb;
      a0 = dp4a_uu(da, k.x, a0); a1 = dp4a_uu(da, k.y, a1); a2 = dp4a_us(da, k.z, a2);
      b0 = dp4a_uu(db, k.x, b0); b1 = dp4a_uu(db, k.y, b1); b2 = dp4a_us(db, k.z, b2);
    }
    const float* l = lut + c * 256;
    const float f0 = l[limbs_to_u8(a0, a1, a2)], f1 = l[limbs_to_u8(b0, b1, b2)];
    const int row = r0 + r, x = 2 * xp;
    if (p.layout == B2C_OUT_NCHW_F32) {
      *reinterpret_cast<float2*>(out_f32 + (c * R + row) * R + x) = make_float2(f0, f1);
    } else {
      // patch-major: element (patch (gy,gx), k = c*p*p + py*p + px); p is even so a px pair never straddles patches
      const int P = p.patch;
      const int gy = __umulhi(row, p.magicP), py = row - gy * P;
      const int gx = __umulhi(x, p.magicP), px = x - gx * P;
      *reinterpret_cast<__nv_bfloat162*>(out_bf16 + (gy * p.g + gx) * p.Kp + (c * P + py) * P + px) =
          __floats2bfloat162_rn(f0, f1);
    }
    cx += kResThreads;
    while (cx >= per) { cx -= per; ++r; }
  }
}

template <int NCH>
__global__ void __launch_bounds__(kResThreads) resample_kernel(const ResampleParams p) {
  extern __shared__ __align__(16) uint8_t smem_pre[];
  const int R = p.R, KS4 = p.KS4, TR = p.TR, CP = pre_odd(KS4);
  constexpr int nch = NCH;  // colour channels this CTA holds
  const int cbeg = NCH == 3 ? 0 : static_cast<int>(blockIdx.z);
  const PreSmem sl = pre_smem(R, KS4, TR, p.HP, p.SR, p.SPP, nch, p.hsm);
  float* lut = reinterpret_cast<float*>(smem_pre + sl.lut);      // [3][256]
  uint4* hcoef = reinterpret_cast<uint4*>(smem_pre + sl.hcoef);  // [R slots][CP]
  int2* hb = reinterpret_cast<int2*>(smem_pre + sl.hb);          // [R slots]
  uint4* vcoef = reinterpret_cast<uint4*>(smem_pre + sl.vcoef);  // [TR][KS4]
  int2* vb = reinterpret_cast<int2*>(smem_pre + sl.vb);          // [TR]
  uint8_t* hbuf = smem_pre + sl.hbuf;                            // [3][R slots][HP]: horizontal results, rows contiguous
  uint8_t* stage = smem_pre + sl.stage;                          // [3][SR][SPP]: planar source rows
  __shared__ int s_nmax[2];

  const int crop = blockIdx.y;
  const int r0 = blockIdx.x * TR;
  const int nr = min(TR, R - r0);
  const CropPlan cp = p.plans[crop];
  const int tid = threadIdx.x;

  if (cp.cw <= 0) {  // crop dropped by the reference (zero area): emit zeros
    if (cbeg != 0) return;  // (one CTA of a channel-split triple does it)
    if (p.layout == B2C_OUT_NCHW_F32) {
      float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(crop) * 3 * R * R;
      for (int i = tid; i < 3 * nr * R; i += kResThreads) {
        const int c = i / (nr * R), rem = i - c * nr * R;
        o[(static_cast<size_t>(c) * R + r0 + rem / R) * R + rem % R] = 0.f;
      }
    } else {
      // zero whole patch rows (including the K padding): the band holding a patch row's first pixel row owns it
      const int P = p.patch, g = R / P;
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(crop) * g * g * p.Kp;
      for (int r = 0; r < nr; ++r) {
        const int row = r0 + r;
        if (row % P != 0) continue;
        __nv_bfloat16* prow = o + static_cast<size_t>(row / P) * g * p.Kp;
        for (int i = tid; i < g * p.Kp; i += kResThreads) prow[i] = __float2bfloat16_rn(0.f);
      }
    }
    return;
  }

  // ---- tables -> smem
  if (tid < 2) s_nmax[tid] = 0;
  for (int i = tid; i < 768; i += kResThreads) lut[i] = p.lut[i];
  const uint4* gh = p.coefs + (static_cast<size_t>(crop) * 2 + 0) * R * KS4;
  const uint4* gv = p.coefs + ((static_cast<size_t>(crop) * 2 + 1) * R + r0) * KS4;
  const int2* gbh = p.bounds + (static_cast<size_t>(crop) * 2 + 0) * R;
  const int2* gbv = p.bounds + (static_cast<size_t>(crop) * 2 + 1) * R + r0;
  // canvas columns [x_lo, x_hi) feed the kept outputs (output 0 is slot 0, output R-1 is slot R-1);
  // canvas rows [y_lo, y_hi) feed this band
  const int x_lo = gbh[0].x;
  const int2 hlast = gbh[R - 1];
  const int x_hi = hlast.x + hlast.y;
  const int y_lo = gbv[0].x;
  const int2 vlast = gbv[nr - 1];
  const int y_hi = vlast.x + vlast.y;
  __syncthreads();
  {
    int nh = 0, nv = 0;
    if (p.hsm) {
      for (int i = tid; i < R * KS4; i += kResThreads) {
        const int e = i / KS4, g = i - e * KS4;
        hcoef[e * CP + g] = gh[i];
      }
    }
    for (int i = tid; i < nr * KS4; i += kResThreads) vcoef[i] = gv[i];
    for (int i = tid; i < R; i += kResThreads) {
      const int2 b = gbh[i];
      hb[i] = make_int2(b.x - x_lo, b.y);
      nh = max(nh, b.y);
    }
    for (int i = tid; i < nr; i += kResThreads) {
      const int2 b = gbv[i];
      vb[i] = make_int2(b.x - y_lo, b.y);
      nv = max(nv, b.y);
    }
    if (nh) atomicMax(&s_nmax[0], nh);
    if (nv) atomicMax(&s_nmax[1], nv);
  }
  __syncthreads();
  const int gh4 = (s_nmax[0] + 3) >> 2, gv4 = (s_nmax[1] + 3) >> 2;  // tap groups any output of this CTA needs

  const int nrows = y_hi - y_lo;
  const int span_bytes = (x_hi - x_lo) * 3;
  // canvas columns that map inside the image
  const int xa = max(x_lo, -cp.dx), xb = min(x_hi, cp.W - cp.dx);
  const int va = (xa - x_lo) * 3, vbnd = (xb - x_lo) * 3;  // valid byte range within a staged row
  const int quads = (x_hi - x_lo + 3) >> 2;
  // an aligned word load may touch up to three bytes past the last pixel; allocations are word-granular
  const uintptr_t img_end = (reinterpret_cast<uintptr_t>(cp.img) + static_cast<size_t>(cp.H - 1) * cp.pitch +
                             static_cast<size_t>(cp.W) * 3 + 3) & ~uintptr_t(3);

  // staging items are (row i, pixel quad q) = 12 source bytes -> one word per colour plane, dealt round-robin: thread t
  // starts at item t and advances kResThreads items.  (Prefetching the source words into registers across the
  // horizontal pass was measured: same time, twice the registers — the kernel is issue-bound, not latency-bound.)
  const int st_i0 = tid / quads, st_q0 = tid - st_i0 * quads;
  const int st_di = kResThreads / quads, st_dq = kResThreads - st_di * quads;
  const size_t plane = static_cast<size_t>(p.SR) * p.SPP / 4;

  for (int y0 = 0; y0 < nrows; y0 += p.SR) {
    const int sr = min(p.SR, nrows - y0);
    // ---- stage sr source rows as three planes
    for (int i = st_i0, q = st_q0; i < sr;) {
      const int iy = y_lo + y0 + i + cp.dy;  // image row
      uint32_t pr = 0, pg = 0, pb = 0;
      if (iy >= 0 && iy < cp.H && xb > xa) {
        const uint8_t* row = cp.img + static_cast<size_t>(iy) * cp.pitch + static_cast<long long>(x_lo + cp.dx) * 3;
        const int b0 = q * 12;
        const uintptr_t src = reinterpret_cast<uintptr_t>(row + b0);
        const uintptr_t sa = src & ~uintptr_t(3);
        const uint32_t sh = static_cast<uint32_t>(src & 3) * 8;
        if (b0 >= va && b0 + 12 <= vbnd && sa >= reinterpret_cast<uintptr_t>(cp.img) && sa + 16 <= img_end) {
          const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(sa);
          const uint32_t w0 = __ldg(wsrc), w1 = __ldg(wsrc + 1), w2 = __ldg(wsrc + 2), w3 = __ldg(wsrc + 3);
          const uint32_t u0 = __funnelshift_r(w0, w1, sh), u1 = __funnelshift_r(w1, w2, sh),
                         u2 = __funnelshift_r(w2, w3, sh);
          // bytes of (u0,u1,u2) = R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
          pr = __byte_perm(__byte_perm(u0, u1, 0x0630), u2, 0x5210);
          pg = __byte_perm(__byte_perm(u0, u1, 0x0741), u2, 0x6210);
          pb = __byte_perm(__byte_perm(u0, u1, 0x0052), u2, 0x7410);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int b = b0 + 3 * e;
            if (b >= va && b + 3 <= vbnd) {
              pr |= static_cast<uint32_t>(__ldg(row + b)) << (8 * e);
              pg |= static_cast<uint32_t>(__ldg(row + b + 1)) << (8 * e);
              pb |= static_cast<uint32_t>(__ldg(row + b + 2)) << (8 * e);
            }
          }
        }
      }
      uint32_t* dst = reinterpret_cast<uint32_t*>(stage + static_cast<size_t>(i) * p.SPP) + q;
      if (NCH == 3) {
        dst[0] = pr; dst[plane] = pg; dst[2 * plane] = pb;
      } else {
        dst[0] = cbeg == 0 ? pr : (cbeg == 1 ? pg : pb);
      }
      i += st_di; q += st_dq;
      if (q >= quads) { q -= quads; ++i; }
    }
    __syncthreads();
    // ---- horizontal pass for the staged rows -> hbuf[c][slot][y0 + row]
    const int rgroups = (sr + 3) >> 2;
    // few taps (hsm): unrolled loops over the shared-memory table; many taps: per-thread loops, table read from global
    // memory (L2-resident, one uint4 per 12 dp4a)
    switch (p.hsm ? gh4 : 0) {
      case 1: resample_hpass<1, NCH>(stage, hbuf, hcoef, hb, R, CP, p.HP, p.SR, p.SPP, rgroups, y0, tid); break;
      case 2: resample_hpass<2, NCH>(stage, hbuf, hcoef, hb, R, CP, p.HP, p.SR, p.SPP, rgroups, y0, tid); break;
      case 3: resample_hpass<3, NCH>(stage, hbuf, hcoef, hb, R, CP, p.HP, p.SR, p.SPP, rgroups, y0, tid); break;
      case 4: resample_hpass<4, NCH>(stage, hbuf, hcoef, hb, R, CP, p.HP, p.SR, p.SPP, rgroups, y0, tid); break;
      default: resample_hpass<0, NCH>(stage, hbuf, gh, hb, R, KS4, p.HP, p.SR, p.SPP, rgroups, y0, tid); break;
    }
    __syncthreads();
  }

  // ---- vertical pass + ToTensor/Normalize
  switch (p.hsm ? gv4 : 0) {  // (hsm = few taps on both axes: the unrolled forms and their tile slack apply)
    case 1: resample_vpass<1, NCH>(p, hbuf, vcoef, vb, lut, crop, r0, nr, cbeg, tid); break;
    case 2: resample_vpass<2, NCH>(p, hbuf, vcoef, vb, lut, crop, r0, nr, cbeg, tid); break;
    case 3: resample_vpass<3, NCH>(p, hbuf, vcoef, vb, lut, crop, r0, nr, cbeg, tid); break;
    case 4: resample_vpass<4, NCH>(p, hbuf, vcoef, vb, lut, crop, r0, nr, cbeg, tid); break;
    default: resample_vpass<0, NCH>(p, hbuf, vcoef, vb, lut, crop, r0, nr, cbeg, tid); break;
  }
  // zero the K padding columns of the patch rows this band owns (Kp > 3*p*p): the band holding a patch row's first
  // pixel row owns it.  3 p^2 and Kp are even, so the padding is whole bf16 pairs.
  if (p.layout == B2C_OUT_PATCH_BF16 && p.Kp > 3 * p.patch * p.patch && cbeg == 0) {
    const int P = p.patch, g = p.g, pad2 = (p.Kp - 3 * P * P) >> 1;
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(crop) * g * g * p.Kp;
    int gy = __umulhi(r0 + P - 1, p.magicP);  // first patch row starting at or after r0
    for (; gy * P < r0 + nr; ++gy) {
      for (int i = tid; i < g * pad2; i += kResThreads) {
        const int gx = __umulhi(i, p.magicPad), j = i - gx * pad2;
        *reinterpret_cast<uint32_t*>(ob + (gy * g + gx) * p.Kp + 3 * P * P + 2 * j) = 0u;
      }
    }
  }
}

struct PreLayout {
  size_t plans, bounds, coefs, lut, total;
};
static PreLayout pre_layout(int B, int R, int KS) {
  const int KS4 = (KS + 3) / 4;
  PreLayout l;
  size_t off = 0;
  l.plans = off;  off += (static_cast<size_t>(B) * 4 * sizeof(CropPlan) + 255) & ~size_t(255);
  l.bounds = off; off += (static_cast<size_t>(B) * 4 * 2 * R * sizeof(int2) + 255) & ~size_t(255);
  l.coefs = off;  off += (static_cast<size_t>(B) * 4 * 2 * R * KS4 * sizeof(uint4) + 255) & ~size_t(255);
  l.lut = off;    off += 768 * sizeof(float);
  l.total = off;
  return l;
}

}  // namespace b2c

extern "C" int b2c_crop_geometry(int W, int H, int R, b2c_crop out[4]) {
  using namespace b2c;
  B2C_REQUIRE(out, "b2c_crop_geometry: null output");
  B2C_REQUIRE(W > 0 && H > 0 && R > 0, "b2c_crop_geometry: W=%d H=%d R=%d must be positive", W, H, R);
  B2C_REQUIRE(static_cast<long long>(W) * H < (1ll << 40), "b2c_crop_geometry: image too large");
  crop_geometry_host(W, H, R, out);
  return 0;
}

extern "C" int b2c_preprocess_workspace_bytes(int B, int max_side, int R, size_t* bytes) {
  using namespace b2c;
  B2C_REQUIRE(bytes, "b2c_preprocess_workspace_bytes: null output");
  B2C_REQUIRE(B > 0 && max_side > 0 && R > 0, "b2c_preprocess_workspace_bytes: bad arguments");
  // the largest tap count arises for the longest canvas side scaled by the shortest possible ratio;
  // the crop's shorter side maps to R, so scale <= max_side / R on both axes.
  *bytes = pre_layout(B, R, ksize_for(max_side, R)).total;
  return 0;
}

extern "C" int b2c_preprocess_4crop(const uint8_t* const* img_ptrs, const int* H, const int* W, const int* pitch, int B,
                                    int R, int patch, const float* mean, const float* stdv, int out_layout, void* out,
                                    void* ws, size_t ws_bytes, b2c_stream stream_) {
  using namespace b2c;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B2C_REQUIRE(img_ptrs && H && W && pitch && mean && stdv && out && ws, "b2c_preprocess_4crop: null pointer");
  B2C_REQUIRE(B > 0 && B <= 16383, "b2c_preprocess_4crop: B=%d must be in [1, 16383]", B);
  B2C_REQUIRE(R > 0 && R % 4 == 0 && R <= 1024, "b2c_preprocess_4crop: R=%d must be a multiple of 4, <= 1024", R);
  B2C_REQUIRE(out_layout == B2C_OUT_NCHW_F32 || out_layout == B2C_OUT_PATCH_BF16, "b2c_preprocess_4crop: layout %d",
              out_layout);
  if (out_layout == B2C_OUT_PATCH_BF16)
    B2C_REQUIRE(patch > 0 && patch % 2 == 0 && R % patch == 0, "b2c_preprocess_4crop: patch=%d must be even and divide R=%d",
                patch, R);

  std::vector<CropPlan> plans(static_cast<size_t>(B) * 4);
  int KS = 5;
  double scale_max = 1.0;
  int span_max = 0;
  for (int i = 0; i < B; ++i) {
    B2C_REQUIRE(img_ptrs[i] && H[i] > 0 && W[i] > 0 && pitch[i] >= 3 * W[i], "b2c_preprocess_4crop: bad image %d", i);
    b2c_crop c4[4];
    crop_geometry_host(W[i], H[i], R, c4);
    for (int c = 0; c < 4; ++c) {
      CropPlan& p = plans[static_cast<size_t>(i) * 4 + c];
      p.img = img_ptrs[i]; p.H = H[i]; p.W = W[i]; p.pitch = pitch[i];
      p.cw = c4[c].cw; p.ch = c4[c].ch; p.dx = c4[c].dx; p.dy = c4[c].dy;
      p.out_w = c4[c].out_w; p.out_h = c4[c].out_h; p.off_x = c4[c].off_x; p.off_y = c4[c].off_y; p.pad = 0;
      if (p.cw <= 0) continue;
      const int kx = ksize_for(p.cw, p.out_w), ky = ksize_for(p.ch, p.out_h);
      if (kx > KS) KS = kx;
      if (ky > KS) KS = ky;
      const double sx = static_cast<double>(p.cw) / p.out_w, sy = static_cast<double>(p.ch) / p.out_h;
      if (sx > scale_max) scale_max = sx;
      if (sy > scale_max) scale_max = sy;
      if (p.cw > span_max) span_max = p.cw;
    }
  }
  const PreLayout lay = pre_layout(B, R, KS);
  if (ws_bytes < lay.total)
    return set_error(B2C_ERR_WORKSPACE, "b2c_preprocess_4crop: workspace %zu B < required %zu B", ws_bytes, lay.total);
  uint8_t* wsb = static_cast<uint8_t*>(ws);
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(wsb) & 255) == 0, "b2c_preprocess_4crop: workspace must be 256-byte aligned");
  B2C_TRY(upload_async(wsb + lay.plans, plans.data(), plans.size() * sizeof(CropPlan), stream));  // no stream synchronisation

  ResampleParams rp;
  rp.plans = reinterpret_cast<const CropPlan*>(wsb + lay.plans);
  rp.bounds = reinterpret_cast<const int2*>(wsb + lay.bounds);
  rp.coefs = reinterpret_cast<const uint4*>(wsb + lay.coefs);
  rp.lut = reinterpret_cast<const float*>(wsb + lay.lut);
  rp.out = out;
  rp.R = R; rp.KS4 = (KS + 3) / 4;
  rp.layout = out_layout; rp.patch = patch > 0 ? patch : 2;
  rp.Kp = out_layout == B2C_OUT_PATCH_BF16 ? (3 * patch * patch + 63) / 64 * 64 : 0;
  NormParams norm;
  for (int c = 0; c < 3; ++c) { norm.mean[c] = mean[c]; norm.stdv[c] = stdv[c]; }
  rp.g = R / rp.patch;
  rp.magicP = static_cast<uint32_t>((1ull << 32) / rp.patch) + 1u;
  {
    const int pad2 = (rp.Kp - 3 * rp.patch * rp.patch) / 2;
    rp.magicPad = pad2 > 0 ? static_cast<uint32_t>((1ull << 32) / pad2) + 1u : 0u;
  }

  // Tile geometry.  Few taps (KS <= 16, sources up to ~1700 px per 224 outputs): coefficient table in shared memory,
  // unrolled tap loops that run the CTA's maximum group count, hence KS + 8 bytes of slack per staged row / tile column
  // (zero coefficients there).  Many taps: every thread loops over its own window (8 bytes of slack for the aligned
  // word reads), the horizontal table is read from global memory.  Bands as tall as fit two CTAs per SM; a source too
  // large for that (24 MP photographs and up) gets one colour channel per CTA (grid.z = 3) and, last, one CTA per SM.
  static const int tr_env = [] { const char* e = getenv("B2C_PRE_TR"); return e ? atoi(e) : 32; }();
  static const int budget_kb = [] { const char* e = getenv("B2C_PRE_SMEM_KB"); return e ? atoi(e) : 113; }();
  const size_t budget = static_cast<size_t>(budget_kb) * 1024;  // 113 KB: two CTAs per SM
  rp.hsm = rp.KS4 <= 4 ? 1 : 0;
  const int slack = rp.hsm ? KS + 8 : 8;
  rp.SPP = (span_max + slack + 3) & ~3;
  int TR = 0, SR = 4;
  size_t smem = 0;
  auto fit = [&](int nch, size_t limit) -> bool {
    for (int tr = std::min(tr_env > 0 ? tr_env : 32, R); tr >= 1; tr = tr > 8 ? tr - 8 : tr >> 1) {
      const int rows_cap = static_cast<int>(tr * scale_max) + KS + 2;
      const int hp = ((rows_cap + slack + 3) & ~3) | 4;  // 4 * odd: consecutive columns fall into different banks
      int sr = static_cast<int>((24 * 1024) / (nch * rp.SPP)) & ~3;
      sr = sr < 4 ? 4 : (sr > 32 ? 32 : sr);
      const size_t need = pre_smem(R, rp.KS4, tr, hp, sr, rp.SPP, nch, rp.hsm).total;
      if (need <= limit) {
        TR = tr; SR = sr; rp.HP = hp; rp.nch = nch; smem = need;
        return true;
      }
      if (tr <= 8 && nch == 3 && limit == budget) break;  // rather split the channels than shrink the band further
    }
    return false;
  };
  if (!fit(3, budget) && !fit(1, budget) && !fit(1, 220 * 1024))
    return set_error(B2C_ERR_ARG, "b2c_preprocess_4crop: images too large for the resample tile (longest crop side %d px)", span_max);
  rp.TR = TR; rp.SR = SR;

  ProfScope ps(B2C_PROF_PREPROCESS, stream);
  resample_plan_kernel<<<dim3(B * 4, 2), R, 0, stream>>>(rp.plans, const_cast<int2*>(rp.bounds),
                                                         const_cast<uint4*>(rp.coefs), const_cast<float*>(rp.lut), norm,
                                                         R, rp.KS4);
  B2C_POST_LAUNCH("resample_plan_kernel");
  const dim3 grid((R + TR - 1) / TR, B * 4, rp.nch == 3 ? 1 : 3);
  if (rp.nch == 3) {
    static PerDeviceMax smem_set;
    if (smem_set.raise(static_cast<long long>(smem)))
      B2C_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resample_kernel<3><<<grid, kResThreads, smem, stream>>>(rp);
  } else {
    static PerDeviceMax smem_set1;
    if (smem_set1.raise(static_cast<long long>(smem)))
      B2C_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resample_kernel<1><<<grid, kResThreads, smem, stream>>>(rp);
  }
  B2C_POST_LAUNCH("resample_kernel");
  return 0;
}
