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
#include <string.h>

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

// grid (n_crops, 2 axes), block R threads. tables: bounds int2[n_crops][2][R] = (xmin, count),
// coefs int32[n_crops][2][R][KS].
__global__ void resample_plan_kernel(const CropPlan* __restrict__ plans, int2* __restrict__ bounds,
                                     int32_t* __restrict__ coefs, int R, int KS) {
  const int crop = blockIdx.x, axis = blockIdx.y, xx = threadIdx.x;
  if (xx >= R) return;
  const CropPlan p = plans[crop];
  int2* b = bounds + (static_cast<size_t>(crop) * 2 + axis) * R + xx;
  int32_t* k = coefs + ((static_cast<size_t>(crop) * 2 + axis) * R + xx) * KS;
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
  for (int x = 0; x < n; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    double w = bicubic_filter(arg);
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    const double scaled = __dmul_rn(w, static_cast<double>(1 << kPrecisionBits));
    k[x] = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
  }
  for (int x = n; x < KS; ++x) k[x] = 0;
  *b = make_int2(xmin, n);
}

// ------------------------------------------------------------------------------------------------
// device: fused two-pass resample + normalise
// ------------------------------------------------------------------------------------------------
struct ResampleParams {
  const CropPlan* plans;
  const int2* bounds;
  const int32_t* coefs;
  void* out;
  int R, KS, TR;      // outputs per side, taps, output rows per band
  int rows_cap;       // rows of the horizontal tile held in smem
  int SR;             // source rows staged per sub-batch
  int span_cap;       // bytes per staged row (multiple of 4)
  int layout, patch, Kp;
  float mean[3], stdv[3];
};

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// shared-memory carve-up of resample_kernel (every region 16-byte aligned)
struct PreSmem {
  size_t lut, hcoef, hb, vcoef, vb, hbuf, stage, total;
};
__host__ __device__ inline size_t pre_a16(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline PreSmem pre_smem(int R, int KS, int TR, int rows_cap, int SR, int span_cap) {
  PreSmem s;
  size_t off = 0;
  s.lut = off;   off += pre_a16(768 * sizeof(float));
  s.hcoef = off; off += pre_a16(static_cast<size_t>(R) * KS * 4);
  s.hb = off;    off += pre_a16(static_cast<size_t>(R) * 8);
  s.vcoef = off; off += pre_a16(static_cast<size_t>(TR) * KS * 4);
  s.vb = off;    off += pre_a16(static_cast<size_t>(TR) * 8);
  s.hbuf = off;  off += pre_a16(static_cast<size_t>(rows_cap) * 3 * R);
  s.stage = off; off += pre_a16(static_cast<size_t>(SR) * span_cap);
  s.total = off;
  return s;
}

constexpr int kResThreads = 256;

__global__ void __launch_bounds__(kResThreads) resample_kernel(const ResampleParams p) {
  extern __shared__ __align__(16) uint8_t smem_pre[];
  const int R = p.R, KS = p.KS, TR = p.TR;
  const PreSmem sl = pre_smem(R, KS, TR, p.rows_cap, p.SR, p.span_cap);
  float* lut = reinterpret_cast<float*>(smem_pre + sl.lut);          // [3][256]
  int32_t* hcoef = reinterpret_cast<int32_t*>(smem_pre + sl.hcoef);  // [R][KS]
  int2* hb = reinterpret_cast<int2*>(smem_pre + sl.hb);              // [R]
  int32_t* vcoef = reinterpret_cast<int32_t*>(smem_pre + sl.vcoef);  // [TR][KS]
  int2* vb = reinterpret_cast<int2*>(smem_pre + sl.vb);              // [TR]
  uint8_t* hbuf = smem_pre + sl.hbuf;                                // [rows_cap][3][R]
  uint8_t* stage = smem_pre + sl.stage;                              // [SR][span_cap]

  const int crop = blockIdx.y;
  const int r0 = blockIdx.x * TR;
  const int nr = min(TR, R - r0);
  const CropPlan cp = p.plans[crop];
  const int tid = threadIdx.x;

  if (cp.cw <= 0) {  // crop dropped by the reference (zero area): emit zeros
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
  for (int i = tid; i < 768; i += kResThreads) {
    const int c = i >> 8, v = i & 255;
    // ToTensor: float(v) / 255 ; Normalize: (x - mean) / std, each op rounded to fp32 like torch
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), p.mean[c]), p.stdv[c]);
  }
  const int32_t* gh = p.coefs + (static_cast<size_t>(crop) * 2 + 0) * R * KS;
  const int32_t* gv = p.coefs + ((static_cast<size_t>(crop) * 2 + 1) * R + r0) * KS;
  for (int i = tid; i < R * KS; i += kResThreads) hcoef[i] = gh[i];
  for (int i = tid; i < nr * KS; i += kResThreads) vcoef[i] = gv[i];
  const int2* gbh = p.bounds + (static_cast<size_t>(crop) * 2 + 0) * R;
  const int2* gbv = p.bounds + (static_cast<size_t>(crop) * 2 + 1) * R + r0;
  for (int i = tid; i < R; i += kResThreads) hb[i] = gbh[i];
  for (int i = tid; i < nr; i += kResThreads) vb[i] = gbv[i];
  __syncthreads();

  const int x_lo = hb[0].x;
  const int x_hi = hb[R - 1].x + hb[R - 1].y;           // canvas columns [x_lo, x_hi) feed the kept outputs
  const int y_lo = vb[0].x;
  const int y_hi = vb[nr - 1].x + vb[nr - 1].y;         // canvas rows [y_lo, y_hi) feed this band
  const int nrows = y_hi - y_lo;
  const int span_bytes = (x_hi - x_lo) * 3;
  // canvas columns that map inside the image
  const int xa = max(x_lo, -cp.dx), xb = min(x_hi, cp.W - cp.dx);
  const int va = (xa - x_lo) * 3, vbnd = (xb - x_lo) * 3;  // valid byte range within a staged row

  for (int y0 = 0; y0 < nrows; y0 += p.SR) {
    const int sr = min(p.SR, nrows - y0);
    // ---- stage sr source rows: dst word w of row i <- 4 source bytes (zero outside the image)
    const int words = (span_bytes + 3) >> 2;
    for (int idx = tid; idx < sr * words; idx += kResThreads) {
      const int i = idx / words, w = idx - i * words;
      const int iy = y_lo + y0 + i + cp.dy;  // image row
      uint32_t val = 0;
      if (iy >= 0 && iy < cp.H && xb > xa) {
        const uint8_t* row = cp.img + static_cast<size_t>(iy) * cp.pitch + static_cast<long long>(x_lo + cp.dx) * 3;
        const int b0 = w * 4;
        const uint8_t* src = row + b0;  // may be unaligned
        const uintptr_t sa = reinterpret_cast<uintptr_t>(src) & ~uintptr_t(3);
        const uint8_t* vbeg = row + va;
        const uint8_t* vend = row + vbnd;
        if (b0 >= va && b0 + 4 <= vbnd && sa >= reinterpret_cast<uintptr_t>(vbeg) &&
            sa + 8 <= reinterpret_cast<uintptr_t>(vend)) {
          const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(sa));
          const uint32_t hi = __ldg(reinterpret_cast<const uint32_t*>(sa + 4));
          val = __funnelshift_r(lo, hi, static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) & 3) * 8);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int b = b0 + e;
            if (b >= va && b < vbnd) val |= static_cast<uint32_t>(__ldg(row + b)) << (8 * e);
          }
        }
      }
      reinterpret_cast<uint32_t*>(stage + static_cast<size_t>(i) * p.span_cap)[w] = val;
    }
    __syncthreads();
    // ---- horizontal pass for the staged rows -> hbuf[row][c][x]
    for (int idx = tid; idx < sr * R; idx += kResThreads) {
      const int i = idx / R, x = idx - i * R;
      const int2 bx = hb[x];
      const uint8_t* s = stage + static_cast<size_t>(i) * p.span_cap + (bx.x - x_lo) * 3;
      const int32_t* k = hcoef + x * KS;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
      for (int t = 0; t < bx.y; ++t) {
        const int kv = k[t];
        a0 += s[3 * t + 0] * kv;
        a1 += s[3 * t + 1] * kv;
        a2 += s[3 * t + 2] * kv;
      }
      uint8_t* h = hbuf + static_cast<size_t>(y0 + i) * 3 * R + x;
      h[0] = static_cast<uint8_t>(clip8(a0));
      h[R] = static_cast<uint8_t>(clip8(a1));
      h[2 * R] = static_cast<uint8_t>(clip8(a2));
    }
    __syncthreads();
  }

  // ---- vertical pass + ToTensor/Normalize; 4 adjacent outputs per thread (R % 4 == 0)
  const int R4 = R >> 2;
  for (int idx = tid; idx < nr * 3 * R4; idx += kResThreads) {
    const int r = idx / (3 * R4);
    const int rem = idx - r * 3 * R4;
    const int c = rem / R4, x4 = (rem - c * R4) * 4;
    const int2 by = vb[r];
    const int32_t* k = vcoef + r * KS;
    const uint8_t* h = hbuf + (static_cast<size_t>(by.x - y_lo) * 3 + c) * R + x4;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
    for (int t = 0; t < by.y; ++t) {
      const uint32_t px = *reinterpret_cast<const uint32_t*>(h + static_cast<size_t>(t) * 3 * R);
      const int kv = k[t];
      a0 += static_cast<int>(px & 0xff) * kv;
      a1 += static_cast<int>((px >> 8) & 0xff) * kv;
      a2 += static_cast<int>((px >> 16) & 0xff) * kv;
      a3 += static_cast<int>(px >> 24) * kv;
    }
    const float* l = lut + c * 256;
    const float f0 = l[clip8(a0)], f1 = l[clip8(a1)], f2 = l[clip8(a2)], f3 = l[clip8(a3)];
    const int row = r0 + r;
    if (p.layout == B2C_OUT_NCHW_F32) {
      float* o = reinterpret_cast<float*>(p.out) + ((static_cast<size_t>(crop) * 3 + c) * R + row) * R + x4;
      *reinterpret_cast<float4*>(o) = make_float4(f0, f1, f2, f3);
    } else {
      // patch-major: element (patch (gy,gx), k = c*p*p + py*p + px); p is even so px pairs never straddle patches
      const int P = p.patch, g = R / P;
      const int gy = row / P, py = row - gy * P;
      __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(crop) * g * g * p.Kp;
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const int x = x4 + e;
        const int gx = x / P, px = x - gx * P;
        __nv_bfloat162 v = __floats2bfloat162_rn(e == 0 ? f0 : f2, e == 0 ? f1 : f3);
        *reinterpret_cast<__nv_bfloat162*>(ob + static_cast<size_t>(gy * g + gx) * p.Kp + c * P * P + py * P + px) = v;
      }
    }
  }
  // zero the K padding columns of the patch rows this band owns (Kp > 3*p*p), once per patch row
  if (p.layout == B2C_OUT_PATCH_BF16 && p.Kp > 3 * p.patch * p.patch) {
    const int P = p.patch, g = R / P, padn = p.Kp - 3 * P * P;
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(crop) * g * g * p.Kp;
    for (int r = 0; r < nr; ++r) {
      const int row = r0 + r;
      if (row % P != 0) continue;  // first pixel row of a patch row owns the padding
      const int gy = row / P;
      for (int i = tid; i < g * padn; i += kResThreads) {
        const int gx = i / padn, j = i - gx * padn;
        ob[static_cast<size_t>(gy * g + gx) * p.Kp + 3 * P * P + j] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

struct PreLayout {
  size_t plans, bounds, coefs, total;
};
static PreLayout pre_layout(int B, int R, int KS) {
  PreLayout l;
  size_t off = 0;
  l.plans = off;  off += (static_cast<size_t>(B) * 4 * sizeof(CropPlan) + 255) & ~size_t(255);
  l.bounds = off; off += (static_cast<size_t>(B) * 4 * 2 * R * sizeof(int2) + 255) & ~size_t(255);
  l.coefs = off;  off += (static_cast<size_t>(B) * 4 * 2 * R * KS * sizeof(int32_t) + 255) & ~size_t(255);
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
  rp.coefs = reinterpret_cast<const int32_t*>(wsb + lay.coefs);
  rp.out = out;
  rp.R = R; rp.KS = KS;
  rp.layout = out_layout; rp.patch = patch > 0 ? patch : 2;
  rp.Kp = out_layout == B2C_OUT_PATCH_BF16 ? (3 * patch * patch + 63) / 64 * 64 : 0;
  for (int c = 0; c < 3; ++c) { rp.mean[c] = mean[c]; rp.stdv[c] = stdv[c]; }

  // band height / smem budget
  const size_t budget = 160 * 1024;
  rp.span_cap = (span_max * 3 + 3 + 15) & ~15;
  int TR = 16, SR = 8;
  size_t smem = 0;
  for (;;) {
    rp.rows_cap = static_cast<int>(TR * scale_max) + KS + 2;
    SR = static_cast<int>((24 * 1024) / rp.span_cap);
    if (SR < 1) SR = 1;
    if (SR > 16) SR = 16;
    smem = pre_smem(R, KS, TR, rp.rows_cap, SR, rp.span_cap).total;
    if (smem <= budget || TR == 1) break;
    TR >>= 1;
  }
  B2C_REQUIRE(smem <= 220 * 1024, "b2c_preprocess_4crop: images too large for the resample tile (need %zu B smem)", smem);
  rp.TR = TR; rp.SR = SR;

  ProfScope ps(B2C_PROF_PREPROCESS, stream);
  resample_plan_kernel<<<dim3(B * 4, 2), R, 0, stream>>>(rp.plans, const_cast<int2*>(rp.bounds),
                                                         const_cast<int32_t*>(rp.coefs), R, KS);
  B2C_POST_LAUNCH("resample_plan_kernel");
  static PerDeviceMax smem_set;
  if (smem_set.raise(static_cast<long long>(smem))) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  resample_kernel<<<dim3((R + TR - 1) / TR, B * 4), kResThreads, smem, stream>>>(rp);
  B2C_POST_LAUNCH("resample_kernel");
  return 0;
}
