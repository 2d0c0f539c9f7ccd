// b2c_imgstats.cu — K13: the 22 img_stat_* scalars of ImageFeaturizer.process (utils/image_features.py:52-94; called per
// image at utils/embedder.py:170) computed from the SAME device-resident uint8 image the 4-crop preprocess reads
// (SURVEY.md §8f row 2).  On the host this stage is 88 ms per image (70 % of the reference's per-image CPU time).
//
//   1. resize to ~768^2 pixels exactly like cv::resize(INTER_AREA) for 8-bit images (OpenCV 4.13, resize.cpp): integer
//      box filter for integer factors, float DecimateAlpha table + row accumulation (unfused float multiply / add, table
//      order) for fractional down-scales, 11-bit fixed-point 2-tap "area mode" interpolation when an axis is enlarged;
//      the tables are built on the host with the library's own double/float arithmetic, one plan per distinct size;
//      resized tiles (plus a one-pixel halo) are produced straight into shared memory — the resized image is never
//      written to HBM;
//   2. one pass over the resized tile: BGR2GRAY / BGR2HSV in OpenCV's fixed point (applied, like the reference, to an
//      RGB array), |R-G| and |R+G-2B|, the 3x3 cross Laplacian with BORDER_REFLECT_101, the 256-bin grey histogram —
//      all as exact integer sums (per-thread 32-bit partials -> 64-bit shared/global atomics: deterministic);
//   3. a finishing kernel turns the sums into the 22 float64 statistics with numpy's formulas.
// Byte/integer work; HBM traffic = the source image, read about once (0.79 MB per 512^2 image).  No tensor cores.
#include <math.h>

#include <algorithm>
#include <map>
#include <vector>

#include "b2c_launch.h"

namespace b2c {

constexpr int kStatsMaxPixels = 768 * 768;
constexpr int kNumSums = 24;      // see SumIdx
constexpr int kStatsOut = 22;

enum SumIdx {
  kC1 = 0,   // 3: sum of channel values
  kC2 = 3,   // 3: sum of squares
  kG1 = 6, kG2 = 7,
  kH1 = 8,   // 3: H, S, V sums
  kH2 = 11,  // 3
  kRG1 = 14, kRG2 = 15, kYB1 = 16, kYB2 = 17,
  kL1 = 18,  // sum of Laplacian (signed, two's complement in the unsigned accumulator)
  kL2 = 19,
  kN = 20,   // pixel count
};

enum ResizeMode { kModeFast = 0, kModeArea = 1, kModeLinear = 2 };

struct PlanHeader {   // one per distinct source size, followed in the table buffer by its int/float arrays
  int sw, sh, dw, dh, mode;
  int ix, iy;         // fast
  float fast_scale;   // fast (other than 2x2): 1.f / (ix*iy)
  int xmax;           // linear
  // offsets (in 4-byte words from the start of the table buffer)
  int x_start, x_si, x_al;   // area: prefix [dw+1], source index, float weight | linear: x_si = offset [dw], x_al = short weights [2*dw]
  int y_start, y_si, y_al;
};

struct ImageJob {
  const unsigned char* src;
  int pitch;
  int plan;           // word offset of the PlanHeader in the table buffer
  int pad;
};

// ---------------------------------------------------------------------------------------------- resize
// The three channels of resized pixel (dx, dy), bit-identical to cv::resize(src, (dw, dh), INTER_AREA).
__device__ __forceinline__ void resize_px(const PlanHeader* __restrict__ p, const int* __restrict__ tab,
                                          const unsigned char* __restrict__ S, long long pitch, int dx, int dy, int out[3]) {
  if (p->mode == kModeFast) {
    int s0 = 0, s1 = 0, s2 = 0;
    for (int yy = 0; yy < p->iy; ++yy) {
      const unsigned char* row = S + (static_cast<long long>(dy) * p->iy + yy) * pitch + (dx * p->ix) * 3;
      for (int xx = 0; xx < p->ix; ++xx) {
        s0 += row[xx * 3];
        s1 += row[xx * 3 + 1];
        s2 += row[xx * 3 + 2];
      }
    }
    if (p->ix == 2 && p->iy == 2) {
      out[0] = (s0 + 2) >> 2; out[1] = (s1 + 2) >> 2; out[2] = (s2 + 2) >> 2;
    } else {
      out[0] = __float2int_rn(__fmul_rn(static_cast<float>(s0), p->fast_scale));
      out[1] = __float2int_rn(__fmul_rn(static_cast<float>(s1), p->fast_scale));
      out[2] = __float2int_rn(__fmul_rn(static_cast<float>(s2), p->fast_scale));
    }
  } else if (p->mode == kModeArea) {
    const int* xs = tab + p->x_start;
    const int* xsi = tab + p->x_si;
    const float* xal = reinterpret_cast<const float*>(tab + p->x_al);
    const int* ys = tab + p->y_start;
    const int* ysi = tab + p->y_si;
    const float* yal = reinterpret_cast<const float*>(tab + p->y_al);
    const int x0 = xs[dx], x1 = xs[dx + 1], y0 = ys[dy], y1 = ys[dy + 1];
    float sum[3] = {0.f, 0.f, 0.f};
    for (int j = y0; j < y1; ++j) {
      const unsigned char* row = S + static_cast<long long>(ysi[j]) * pitch;
      float buf[3] = {0.f, 0.f, 0.f};
      for (int i = x0; i < x1; ++i) {  // buf += S * alpha: float multiply and add, unfused, in table order
        const unsigned char* px = row + xsi[i] * 3;
        const float a = xal[i];
        buf[0] = __fadd_rn(buf[0], __fmul_rn(static_cast<float>(px[0]), a));
        buf[1] = __fadd_rn(buf[1], __fmul_rn(static_cast<float>(px[1]), a));
        buf[2] = __fadd_rn(buf[2], __fmul_rn(static_cast<float>(px[2]), a));
      }
      const float beta = yal[j];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float term = __fmul_rn(beta, buf[c]);
        sum[c] = j == y0 ? term : __fadd_rn(sum[c], term);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = __float2int_rn(sum[c]);
  } else {
    const int* xo = tab + p->x_si;
    const int* xa = tab + p->x_al;
    const int* yo = tab + p->y_si;
    const int* ya = tab + p->y_al;
    const int sy = yo[dy];
    const int s0 = min(max(sy, 0), p->sh - 1), s1 = min(max(sy + 1, 0), p->sh - 1);
    const unsigned char* r0 = S + static_cast<long long>(s0) * pitch + xo[dx] * 3;
    const unsigned char* r1 = S + static_cast<long long>(s1) * pitch + xo[dx] * 3;
    const int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
    const bool two = dx < p->xmax;
    const int a0 = two ? xa[2 * dx] : 2048, a1 = two ? xa[2 * dx + 1] : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = r0[c] * a0 + (two ? r0[c + 3] * a1 : 0);
      const int h1 = r1[c] * a0 + (two ? r1[c + 3] * a1 : 0);
      out[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c] = min(max(out[c], 0), 255);
}

// ---------------------------------------------------------------------------------------------- fused resize + statistics
__device__ __forceinline__ int gray_of(const unsigned char* px) {
  return (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;  // cv::COLOR_BGR2GRAY, 15-bit fixed point
}

constexpr int kTileW = 128, kTileH = 16;                 // resized pixels per block (plus a one-pixel halo for the Laplacian)
constexpr int kTilePitch = (kTileW + 2) * 3;

// grid (ceil(dw / kTileW), ceil(dh / kTileH), images).  Phase 1: the block computes its tile of the RESIZED image (and the
// halo, at BORDER_REFLECT_101 coordinates) straight into shared memory — the resized image never exists in HBM, so the
// pass reads each source byte about once and writes 1.2 KB of sums per image.  Phase 2: thread = (column, half of the
// rows); 32-bit partial sums -> warp reduction (REDUX) -> one 64-bit shared atomic per warp and quantity.
__global__ void __launch_bounds__(256)
stats_fused_kernel(const ImageJob* __restrict__ jobs, const int* __restrict__ tab, const int* __restrict__ sdiv,
                   const int* __restrict__ hdiv, unsigned long long* __restrict__ sums, unsigned int* __restrict__ hist) {
  __shared__ unsigned char s_px[(kTileH + 2) * kTilePitch];
  __shared__ unsigned long long s_sum[kNumSums];
  __shared__ unsigned int s_hist[8][256];  // one histogram per warp
  const ImageJob job = jobs[blockIdx.z];
  const PlanHeader* p = reinterpret_cast<const PlanHeader*>(tab + job.plan);
  const int W = p->dw, H = p->dh;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  if (x0 >= W || y0 >= H) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kNumSums) s_sum[threadIdx.x] = 0ull;
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0u;
  const int tw = min(kTileW, W - x0), th = min(kTileH, H - y0);
  for (int t = threadIdx.x; t < (th + 2) * (tw + 2); t += 256) {
    const int ty = t / (tw + 2), tx = t - ty * (tw + 2);
    int gx = x0 + tx - 1, gy = y0 + ty - 1;
    gx = gx < 0 ? (W > 1 ? 1 : 0) : (gx >= W ? (W > 1 ? W - 2 : 0) : gx);  // BORDER_REFLECT_101
    gy = gy < 0 ? (H > 1 ? 1 : 0) : (gy >= H ? (H > 1 ? H - 2 : 0) : gy);
    int v[3];
    resize_px(p, tab, job.src, job.pitch, gx, gy, v);
    unsigned char* d = s_px + ty * kTilePitch + tx * 3;
    d[0] = static_cast<unsigned char>(v[0]);
    d[1] = static_cast<unsigned char>(v[1]);
    d[2] = static_cast<unsigned char>(v[2]);
  }
  __syncthreads();
  unsigned int acc[kNumSums];
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) acc[i] = 0u;
  int lap1 = 0;
  const int col = threadIdx.x & (kTileW - 1), half = threadIdx.x / kTileW;  // 2 threads per column, 8 rows each
  if (col < tw) {
    for (int r = half * (kTileH / 2); r < min((half + 1) * (kTileH / 2), th); ++r) {
      const unsigned char* px = s_px + (r + 1) * kTilePitch + (col + 1) * 3;
      const int c0 = px[0], c1 = px[1], c2 = px[2];
      acc[kC1 + 0] += c0; acc[kC1 + 1] += c1; acc[kC1 + 2] += c2;
      acc[kC2 + 0] += c0 * c0; acc[kC2 + 1] += c1 * c1; acc[kC2 + 2] += c2 * c2;
      const int g = gray_of(px);
      acc[kG1] += g; acc[kG2] += g * g;
      atomicAdd(&s_hist[warp][g], 1u);
      // cv::COLOR_BGR2HSV (8-bit): b = c0, g = c1, r = c2
      const int v = max(max(c0, c1), c2), vmin = min(min(c0, c1), c2), diff = v - vmin;
      const int sat = (diff * sdiv[v] + (1 << 11)) >> 12;
      int h = v == c2 ? c1 - c0 : (v == c1 ? c0 - c2 + 2 * diff : c2 - c1 + 4 * diff);
      h = (h * hdiv[diff] + (1 << 11)) >> 12;
      h += h < 0 ? 180 : 0;
      acc[kH1 + 0] += h; acc[kH1 + 1] += sat; acc[kH1 + 2] += v;
      acc[kH2 + 0] += h * h; acc[kH2 + 1] += sat * sat; acc[kH2 + 2] += v * v;
      const int rg = abs(c2 - c1), yb = abs(c2 + c1 - 2 * c0);
      acc[kRG1] += rg; acc[kRG2] += rg * rg; acc[kYB1] += yb; acc[kYB2] += yb * yb;
      const int lap = gray_of(px - kTilePitch) + gray_of(px + kTilePitch) + gray_of(px - 3) + gray_of(px + 3) - 4 * g;
      lap1 += lap;
      acc[kL2] += lap * lap;
      acc[kN] += 1;
    }
  }
  // per-thread partials stay below 2^27, a warp's total below 2^32: reduce in 32 bits, accumulate in 64
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) {
    if (i == kL1 || i > kN) continue;
    const unsigned int tot = __reduce_add_sync(0xffffffffu, acc[i]);
    if (lane == 0 && tot) atomicAdd(&s_sum[i], static_cast<unsigned long long>(tot));
  }
  const int ltot = __reduce_add_sync(0xffffffffu, lap1);
  if (lane == 0 && ltot) atomicAdd(&s_sum[kL1], static_cast<unsigned long long>(static_cast<long long>(ltot)));
  __syncthreads();
  unsigned long long* gs = sums + static_cast<size_t>(blockIdx.z) * kNumSums;
  unsigned int* gh = hist + static_cast<size_t>(blockIdx.z) * 256;
  if (threadIdx.x < kNumSums && s_sum[threadIdx.x]) atomicAdd(&gs[threadIdx.x], s_sum[threadIdx.x]);
  unsigned int hb = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) hb += s_hist[w][threadIdx.x];
  if (hb) atomicAdd(&gh[threadIdx.x], hb);
}

__device__ __forceinline__ void mean_std(double s1, double s2, double n, double& mean, double& sd) {
  mean = s1 / n;
  const double var = s2 / n - mean * mean;
  sd = sqrt(var > 0.0 ? var : 0.0);
}

// one thread per image
__global__ void stats_finish_kernel(const ImageJob* __restrict__ jobs, const int* __restrict__ tab,
                                    const unsigned long long* __restrict__ sums, const unsigned int* __restrict__ hist,
                                    double* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const PlanHeader* p = reinterpret_cast<const PlanHeader*>(tab + jobs[b].plan);
  const unsigned long long* s = sums + static_cast<size_t>(b) * kNumSums;
  double* o = out + static_cast<size_t>(b) * kStatsOut;
  const double n = static_cast<double>(s[kN]);
  o[0] = p->dw / 768.0;
  o[1] = p->dh / 768.0;
  o[2] = static_cast<double>(p->dw) / static_cast<double>(p->dh);
  double m, sd;
  mean_std(static_cast<double>(s[kC1] + s[kC1 + 1] + s[kC1 + 2]), static_cast<double>(s[kC2] + s[kC2 + 1] + s[kC2 + 2]), 3.0 * n, m, sd);
  o[3] = m / 255.0;
  o[4] = sd / 255.0;
  for (int i = 0; i < 3; ++i) {
    mean_std(static_cast<double>(s[kC1 + i]), static_cast<double>(s[kC2 + i]), n, m, sd);
    o[5 + i] = m / 255.0;
    o[8 + i] = sd / 255.0;
  }
  mean_std(static_cast<double>(s[kG1]), static_cast<double>(s[kG2]), n, m, sd);
  o[11] = m / 255.0;
  o[12] = sd / 255.0;
  for (int i = 0; i < 3; ++i) {
    mean_std(static_cast<double>(s[kH1 + i]), static_cast<double>(s[kH2 + i]), n, m, sd);
    o[13 + i] = m / 255.0;
    o[16 + i] = sd / 255.0;
  }
  double rgm, rgs, ybm, ybs;
  mean_std(static_cast<double>(s[kRG1]), static_cast<double>(s[kRG2]), n, rgm, rgs);
  mean_std(static_cast<double>(s[kYB1]), static_cast<double>(s[kYB2]), n, ybm, ybs);
  ybm *= 0.5;  // the accumulators hold 2*|0.5(R+G) - B|
  ybs *= 0.5;
  o[19] = (sqrt(rgs * rgs + ybs * ybs) + 0.3 * sqrt(rgm * rgm + ybm * ybm)) / 100.0;
  // calcHist is float32 and `histogram /= histogram.sum()` stays float32; the log term is evaluated in float64
  double ent = 0.0;
  const float nf = static_cast<float>(s[kN]);
  const unsigned int* h = hist + static_cast<size_t>(b) * 256;
  for (int i = 0; i < 256; ++i) {
    const float hf = __fdiv_rn(static_cast<float>(h[i]), nf);
    ent += static_cast<double>(hf) * log2(static_cast<double>(hf) + 2.220446049250313e-16);
  }
  o[20] = -ent / 8.0;
  const double lm = static_cast<double>(static_cast<long long>(s[kL1])) / n;
  const double lv = static_cast<double>(s[kL2]) / n - lm * lm;
  o[21] = tanh((lv > 0.0 ? lv : 0.0) * 1e-4);
}

// ---------------------------------------------------------------------------------------------- host: plans
static void append_words(std::vector<int>& buf, const void* p, size_t words) {
  const int* q = static_cast<const int*>(p);
  buf.insert(buf.end(), q, q + words);
}

// DecimateAlpha table of cv::resize's computeResizeAreaTab (resize.cpp), grouped per destination index
static void area_tab(int ssize, int dsize, double scale, std::vector<int>& start, std::vector<int>& si, std::vector<float>& al) {
  start.assign(dsize + 1, 0);
  for (int dx = 0; dx < dsize; ++dx) {
    start[dx] = static_cast<int>(si.size());
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = static_cast<int>(ceil(fsx1)), sx2 = static_cast<int>(floor(fsx2));
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) {
      si.push_back(sx1 - 1);
      al.push_back(static_cast<float>((sx1 - fsx1) / cell));
    }
    for (int sx = sx1; sx < sx2; ++sx) {
      si.push_back(sx);
      al.push_back(static_cast<float>(1.0 / cell));
    }
    if (fsx2 - sx2 > 1e-3) {
      si.push_back(sx2);
      al.push_back(static_cast<float>(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell));
    }
  }
  start[dsize] = static_cast<int>(si.size());
}

// offsets + 11-bit weights of the "area mode" 2-tap interpolation (resize.cpp, INTER_AREA with an enlarged axis)
static void linear_tab(int ssize, int dsize, bool clamp, std::vector<int>& ofs, std::vector<int>& al, int& dmax) {
  const double inv = static_cast<double>(dsize) / ssize, scale = 1.0 / inv;
  ofs.assign(dsize, 0);
  al.assign(2 * dsize, 0);
  dmax = dsize;
  for (int d = 0; d < dsize; ++d) {
    int s = static_cast<int>(floor(d * scale));
    float f = static_cast<float>((d + 1) - (s + 1) * inv);
    f = f <= 0 ? 0.f : f - floorf(f);
    if (clamp) {
      if (s < 0) { f = 0.f; s = 0; }
      if (s + 1 >= ssize) {
        dmax = std::min(dmax, d);
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
      }
    }
    ofs[d] = s;
    al[2 * d] = static_cast<int>(lrintf((1.f - f) * 2048.f));
    al[2 * d + 1] = static_cast<int>(lrintf(f * 2048.f));
  }
}

static int build_plan(int sw, int sh, std::vector<int>& buf) {
  PlanHeader h = {};
  h.sw = sw;
  h.sh = sh;
  // utils/image_features.py:57-58 — `w, h = rgb_image.shape[:2]` names the ROW count w: the swap is the reference's
  const double w = sh, hh = sw;
  h.dw = static_cast<int>(sqrt(static_cast<double>(kStatsMaxPixels) * w / hh));
  h.dh = static_cast<int>(sqrt(static_cast<double>(kStatsMaxPixels) * hh / w));
  if (h.dw < 1 || h.dh < 1) return -1;
  const double sx = 1.0 / (static_cast<double>(h.dw) / sw), sy = 1.0 / (static_cast<double>(h.dh) / sh);
  const int at = static_cast<int>(buf.size());
  buf.resize(buf.size() + (sizeof(PlanHeader) + 3) / 4);
  if (sx >= 1 && sy >= 1) {
    const int ix = static_cast<int>(lrint(sx)), iy = static_cast<int>(lrint(sy));
    if (fabs(sx - ix) < 2.220446049250313e-16 && fabs(sy - iy) < 2.220446049250313e-16) {
      h.mode = kModeFast;
      h.ix = ix;
      h.iy = iy;
      h.fast_scale = 1.f / static_cast<float>(ix * iy);
    } else {
      h.mode = kModeArea;
      std::vector<int> st, si;
      std::vector<float> al;
      area_tab(sw, h.dw, sx, st, si, al);
      h.x_start = static_cast<int>(buf.size()); append_words(buf, st.data(), st.size());
      h.x_si = static_cast<int>(buf.size()); append_words(buf, si.data(), si.size());
      h.x_al = static_cast<int>(buf.size()); append_words(buf, al.data(), al.size());
      st.clear(); si.clear(); al.clear();
      area_tab(sh, h.dh, sy, st, si, al);
      h.y_start = static_cast<int>(buf.size()); append_words(buf, st.data(), st.size());
      h.y_si = static_cast<int>(buf.size()); append_words(buf, si.data(), si.size());
      h.y_al = static_cast<int>(buf.size()); append_words(buf, al.data(), al.size());
    }
  } else {
    h.mode = kModeLinear;
    std::vector<int> ofs, al;
    int dmax;
    linear_tab(sw, h.dw, true, ofs, al, dmax);
    h.xmax = dmax;
    h.x_si = static_cast<int>(buf.size()); append_words(buf, ofs.data(), ofs.size());
    h.x_al = static_cast<int>(buf.size()); append_words(buf, al.data(), al.size());
    linear_tab(sh, h.dh, false, ofs, al, dmax);
    h.y_si = static_cast<int>(buf.size()); append_words(buf, ofs.data(), ofs.size());
    h.y_al = static_cast<int>(buf.size()); append_words(buf, al.data(), al.size());
  }
  memcpy(buf.data() + at, &h, sizeof(PlanHeader));
  return at;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace b2c

extern "C" int b2c_image_stats_target_size(int W, int H, int* new_w, int* new_h) {
  using namespace b2c;
  B2C_REQUIRE(W > 0 && H > 0 && new_w && new_h, "b2c_image_stats_target_size: bad arguments");
  const double w = H, h = W;
  *new_w = static_cast<int>(sqrt(static_cast<double>(kStatsMaxPixels) * w / h));
  *new_h = static_cast<int>(sqrt(static_cast<double>(kStatsMaxPixels) * h / w));
  return 0;
}

namespace b2c {
// plans for the distinct sizes of a batch (word offsets per image in `plan_of`), total table words in tab.size()
static int plan_batch(const int* H, const int* W, int B, std::vector<int>& tab, std::vector<int>& plan_of, int& max_dw, int& max_dh) {
  std::map<std::pair<int, int>, int> plans;
  plan_of.assign(B, 0);
  max_dw = max_dh = 1;
  for (int i = 0; i < B; ++i) {
    B2C_REQUIRE(H[i] > 0 && W[i] > 0, "b2c_image_stats: image %d has size %dx%d", i, W[i], H[i]);
    auto key = std::make_pair(W[i], H[i]);
    auto it = plans.find(key);
    if (it == plans.end()) {
      const int at = build_plan(W[i], H[i], tab);
      B2C_REQUIRE(at >= 0, "b2c_image_stats: image %d (%dx%d) resizes to an empty image", i, W[i], H[i]);
      it = plans.emplace(key, at).first;
    }
    plan_of[i] = it->second;
    const PlanHeader* p = reinterpret_cast<const PlanHeader*>(tab.data() + it->second);
    max_dw = std::max(max_dw, p->dw);
    max_dh = std::max(max_dh, p->dh);
  }
  return 0;
}

static size_t stats_ws_bytes(int B, size_t tab_words) {
  size_t total = align_up(static_cast<size_t>(B) * sizeof(ImageJob), 256);         // jobs
  total += align_up(static_cast<size_t>(B) * kNumSums * sizeof(unsigned long long), 256);
  total += align_up(static_cast<size_t>(B) * 256 * sizeof(unsigned int), 256);
  total += align_up(2 * 256 * sizeof(int), 256);                                   // HSV reciprocal tables
  total += align_up(tab_words * sizeof(int), 256);                                 // resize plans
  return total;
}
}  // namespace b2c

extern "C" int b2c_image_stats_workspace_bytes(const int* H, const int* W, int B, size_t* bytes) {
  using namespace b2c;
  B2C_REQUIRE(B >= 0 && bytes && (B == 0 || (H && W)), "b2c_image_stats_workspace_bytes: bad arguments");
  std::vector<int> tab, plan_of;
  int mw, mh;
  B2C_TRY(plan_batch(H, W, B, tab, plan_of, mw, mh));
  *bytes = stats_ws_bytes(B, tab.size());
  return 0;
}

extern "C" int b2c_image_stats(const uint8_t* const* img_ptrs, const int* H, const int* W, const int* pitch, int B, double* out,
                               void* ws, size_t ws_bytes, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(B >= 0, "b2c_image_stats: B=%d", B);
  if (B == 0) return 0;
  B2C_REQUIRE(img_ptrs && H && W && pitch && out && ws, "b2c_image_stats: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < B; ++i)
    B2C_REQUIRE(img_ptrs[i] != nullptr && H[i] > 0 && W[i] > 0 && pitch[i] >= 3 * W[i], "b2c_image_stats: image %d is malformed", i);
  // host: one resize plan per distinct size
  std::vector<int> tab, plan_of;
  int max_dw, max_dh;
  B2C_TRY(plan_batch(H, W, B, tab, plan_of, max_dw, max_dh));
  const size_t need = stats_ws_bytes(B, tab.size());
  if (ws_bytes < need) return set_error(B2C_ERR_WORKSPACE, "b2c_image_stats: workspace %zu < %zu bytes", ws_bytes, need);
  B2C_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 256 == 0, "b2c_image_stats: workspace must be 256-byte aligned");
  std::vector<ImageJob> jobs(B);
  for (int i = 0; i < B; ++i) {
    jobs[i].src = img_ptrs[i];
    jobs[i].pitch = pitch[i];
    jobs[i].plan = plan_of[i];
    jobs[i].pad = 0;
  }
  int hsv[512];
  hsv[0] = hsv[256] = 0;
  for (int i = 1; i < 256; ++i) {  // color_hsv.simd.hpp: sdiv_table / hdiv_table180 with hsv_shift = 12
    hsv[i] = static_cast<int>(lrint((255 << 12) / (1.0 * i)));
    hsv[256 + i] = static_cast<int>(lrint((180 << 12) / (6.0 * i)));
  }
  // carve the workspace
  char* base = static_cast<char*>(ws);
  ImageJob* d_jobs = reinterpret_cast<ImageJob*>(base);
  base += align_up(static_cast<size_t>(B) * sizeof(ImageJob), 256);
  unsigned long long* d_sums = reinterpret_cast<unsigned long long*>(base);
  base += align_up(static_cast<size_t>(B) * kNumSums * sizeof(unsigned long long), 256);
  unsigned int* d_hist = reinterpret_cast<unsigned int*>(base);
  base += align_up(static_cast<size_t>(B) * 256 * sizeof(unsigned int), 256);
  int* d_hsv = reinterpret_cast<int*>(base);
  base += align_up(2 * 256 * sizeof(int), 256);
  int* d_tab = reinterpret_cast<int*>(base);
  // pageable-source async copies are staged before returning, so the host vectors may die at the end of this call
  B2C_TRY(upload_async(d_jobs, jobs.data(), jobs.size() * sizeof(ImageJob), st));
  B2C_TRY(upload_async(d_tab, tab.data(), tab.size() * sizeof(int), st));
  B2C_TRY(upload_async(d_hsv, hsv, sizeof(hsv), st));
  B2C_CHECK_CUDA(cudaMemsetAsync(d_sums, 0, static_cast<size_t>(B) * kNumSums * sizeof(unsigned long long), st));
  B2C_CHECK_CUDA(cudaMemsetAsync(d_hist, 0, static_cast<size_t>(B) * 256 * sizeof(unsigned int), st));
  for (int first = 0; first < B; first += 32768) {  // gridDim.z limit
    const int nb = std::min(32768, B - first);
    dim3 grid((max_dw + kTileW - 1) / kTileW, (max_dh + kTileH - 1) / kTileH, nb);
    stats_fused_kernel<<<grid, 256, 0, st>>>(d_jobs + first, d_tab, d_hsv, d_hsv + 256,
                                             d_sums + static_cast<size_t>(first) * kNumSums, d_hist + static_cast<size_t>(first) * 256);
    B2C_POST_LAUNCH("stats_fused_kernel");
  }
  stats_finish_kernel<<<(B + 63) / 64, 64, 0, st>>>(d_jobs, d_tab, d_sums, d_hist, out, B);
  B2C_POST_LAUNCH("stats_finish_kernel");
  return 0;
}
