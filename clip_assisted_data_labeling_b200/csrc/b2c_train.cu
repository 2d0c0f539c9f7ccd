// b2c_train.cu — K12: the SimpleFC regressor's training step on the device.
//
// Replaces the inner loop of _4_train_model.py:196-204 — optimizer.zero_grad(); model(features); MSELoss; backward();
// Adam.step() — for utils/nn_model.SimpleFC (Linear -> LeakyReLU -> Dropout per hidden layer, Linear -> Sigmoid).
// A step on the reference's default batch of 16 is ~55 MFLOP spread by PyTorch eager over ~60 kernel launches; here it
// is 3L-1 launches (L = number of Linear layers) with no host synchronisation inside an epoch:
//   forward   one launch per layer: one warp per output neuron, the batch's input rows staged in shared memory, the
//             weight row read once with 16-byte loads (the work is reading W: L2/HBM-bound fp32 SIMT, no tensor cores);
//             LeakyReLU, dropout (counter-based Philox4x32-10 stream, reproducible by the oracle) and, on the last
//             layer, sigmoid + MSE + dLoss/dz fused into the epilogue;
//   backward  one launch per hidden layer: delta_l = (delta_{l+1} · W_{l+1}) ∘ act'(z_l) ∘ dropout-scale;
//   update    one launch per layer: dW = delta_lᵀ · a_{l-1} is formed per weight in registers and consumed immediately by
//             the Adam update (torch.optim.Adam semantics: L2 weight decay added to the gradient, bias-corrected
//             moments), so gradients are never written to memory.
#include <math.h>

#include <algorithm>
#include <new>
#include <vector>

#include "b2c_launch.h"

namespace b2c {

constexpr int kTrMaxBatch = 64;
constexpr int kTrKT = 256;  // input columns staged per shared-memory tile

// ---- Philox4x32-10 (Salmon et al. 2011), the counter-based generator behind the dropout masks ---------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// keep-scale of hidden unit `e` (= b*H + j) of layer `layer` at optimiser step `step`: 1/(1-p) if kept, 0 if dropped
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long step, int layer, uint32_t e,
                                               float p, float inv_keep) {
  uint32_t r[4];
  philox4x32_10(e >> 2, static_cast<uint32_t>(layer), static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32),
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float u = static_cast<float>(r[e & 3] >> 8) * (1.0f / 16777216.0f);  // uniform in [0,1)
  return u >= p ? inv_keep : 0.0f;
}

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FwdArgs {
  const float* in;        // a_{l-1} [B, K] (row stride K) or, with idx != nullptr, the feature matrix
  long long in_stride;    // row stride of `in` in elements
  const int* idx;         // optional gather: row b of the batch is in[idx[b]]
  const float* W;         // [H, K]
  const float* bias;      // [H]
  float* a_out;           // [B, H] activation after dropout (hidden) / sigmoid output (last layer)
  float* d_out;           // [B, H] act'(z) * dropout scale (hidden) / unused (last layer)
  float* delta_out;       // last layer only: dLoss/dz [B, H]
  const float* labels;    // last layer only: labels gathered through idx_labels
  const int* idx_labels;
  float* loss_sum;        // last layer only: += mean squared error of the batch
  int B, K, H, last, layer;
  float slope, p, inv_keep;
  unsigned long long seed, step;
  float* zpart;           // split-K partial sums [ksplit][B][H] (ksplit > 1 only)
  int ksplit, tiles_per_split;
};

// bias + activation (+ dropout) of a hidden unit, or sigmoid + MSE + dLoss/dz of an output unit
__device__ __forceinline__ float fwd_epilogue(const FwdArgs& a, int b, int j, float z) {
  const long long o = static_cast<long long>(b) * a.H + j;
  if (!a.last) {
    const float scale = a.p > 0.f ? dropout_scale(a.seed, a.step, a.layer, static_cast<uint32_t>(o), a.p, a.inv_keep) : 1.0f;
    a.a_out[o] = (z > 0.f ? z : a.slope * z) * scale;
    a.d_out[o] = (z > 0.f ? 1.0f : a.slope) * scale;
    return 0.f;
  }
  const float y = 1.0f / (1.0f + expf(-z));
  a.a_out[o] = y;
  const float t = a.labels[a.idx_labels ? a.idx_labels[b] : b];
  const float diff = y - t;
  // MSELoss(reduction='mean') over the B*H outputs (_4_train_model.py:127,199)
  const float inv_n = 1.0f / static_cast<float>(a.B * a.H);
  a.delta_out[o] = 2.0f * diff * inv_n * y * (1.0f - y);
  return diff * diff * inv_n;
}

// one warp per output neuron j; BT = compile-time batch capacity (accumulators live in registers)
template <int BT>
__global__ void __launch_bounds__(256) train_fwd_kernel(FwdArgs a) {
  extern __shared__ float s_x[];  // [B][kTrKT]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  float acc[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) acc[b] = 0.f;
  const bool vec = (a.K % 4 == 0) && (a.in_stride % 4 == 0);
  // split-K: blockIdx.y owns tiles [y*tiles_per_split, (y+1)*tiles_per_split) so that a wide, short layer (4096 -> 264)
  // still fills the GPU; the partial sums are combined in a fixed order by train_finish_kernel (deterministic)
  const int k_begin = blockIdx.y * a.tiles_per_split * kTrKT;
  const int k_end = min(a.K, k_begin + a.tiles_per_split * kTrKT);
  for (int k0 = k_begin; k0 < k_end; k0 += kTrKT) {
    const int kt = min(kTrKT, a.K - k0);
    __syncthreads();
    for (int t = threadIdx.x; t < a.B * kTrKT; t += 256) {
      const int b = t / kTrKT, k = t - b * kTrKT;
      const long long row = a.idx ? a.idx[b] : b;
      s_x[t] = k < kt ? a.in[row * a.in_stride + k0 + k] : 0.f;
    }
    __syncthreads();
    if (j < a.H) {
      const float* wrow = a.W + static_cast<long long>(j) * a.K + k0;
      // lane owns columns [8*lane, 8*lane + 8) of the tile
      float w[8];
      const int kb = lane * 8;
      if (vec && kb + 8 <= kt) {
        const float4 w0 = *reinterpret_cast<const float4*>(wrow + kb);
        const float4 w1 = *reinterpret_cast<const float4*>(wrow + kb + 4);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = kb + e < kt ? wrow[kb + e] : 0.f;
      }
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        if (b < a.B) {
          const float4 x0 = *reinterpret_cast<const float4*>(s_x + b * kTrKT + kb);
          const float4 x1 = *reinterpret_cast<const float4*>(s_x + b * kTrKT + kb + 4);
          float s = acc[b];
          s = fmaf(w[0], x0.x, s); s = fmaf(w[1], x0.y, s); s = fmaf(w[2], x0.z, s); s = fmaf(w[3], x0.w, s);
          s = fmaf(w[4], x1.x, s); s = fmaf(w[5], x1.y, s); s = fmaf(w[6], x1.z, s); s = fmaf(w[7], x1.w, s);
          acc[b] = s;
        }
      }
    }
  }
  if (j >= a.H) return;
  float mine = 0.f;  // lane b (and b+32) finishes row b
  float mine_hi = 0.f;
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    const float s = warp_sum_t(acc[b]);
    if ((b & 31) == lane) {
      if (b < 32) mine = s; else mine_hi = s;
    }
  }
  const float bj = a.bias[j];
  float loss_part = 0.f;
  for (int half = 0; half < (BT + 31) / 32; ++half) {
    const int b = half * 32 + lane;
    if (b >= a.B) continue;
    const float z = half ? mine_hi : mine;
    if (a.ksplit > 1)
      a.zpart[(static_cast<long long>(blockIdx.y) * a.B + b) * a.H + j] = z;
    else
      loss_part += fwd_epilogue(a, b, j, z + bj);
  }
  if (a.last && a.ksplit == 1) {
    loss_part = warp_sum_t(loss_part);
    if (lane == 0 && a.loss_sum != nullptr) atomicAdd(a.loss_sum, loss_part);
  }
}

// split-K epilogue: z[b,j] = bias[j] + sum over splits (fixed order), then the same activation epilogue; thread per (b,j)
__global__ void __launch_bounds__(256) train_finish_kernel(FwdArgs a) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  float loss_part = 0.f;
  if (t < a.B * a.H) {
    const int b = t / a.H, j = t - b * a.H;
    float z = a.bias[j];
    for (int ks = 0; ks < a.ksplit; ++ks) z += a.zpart[(static_cast<long long>(ks) * a.B + b) * a.H + j];
    loss_part = fwd_epilogue(a, b, j, z);
  }
  if (a.last) {
    loss_part = warp_sum_t(loss_part);
    if ((threadIdx.x & 31) == 0 && a.loss_sum != nullptr && loss_part != 0.f) atomicAdd(a.loss_sum, loss_part);
  }
}

// delta_l[b,k] = (sum_j delta_{l+1}[b,j] * W_{l+1}[j,k]) * d_l[b,k];  thread per k, delta_{l+1} staged in shared memory
template <int BT>
__global__ void __launch_bounds__(256)
train_delta_kernel(const float* __restrict__ delta_next, const float* __restrict__ Wn, const float* __restrict__ d,
                   float* __restrict__ delta, int B, int K /*H_l*/, int Hn /*H_{l+1}*/) {
  extern __shared__ float s_d[];  // [128][BT] tile of delta_next, j-major
  const int k = blockIdx.x * 256 + threadIdx.x;
  float acc[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) acc[b] = 0.f;
  for (int j0 = 0; j0 < Hn; j0 += 128) {
    const int jt = min(128, Hn - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < 128 * BT; t += 256) {
      const int jj = t / BT, b = t - jj * BT;
      s_d[t] = (jj < jt && b < B) ? delta_next[static_cast<long long>(b) * Hn + j0 + jj] : 0.f;
    }
    __syncthreads();
    if (k < K) {
#pragma unroll 8
      for (int jj = 0; jj < jt; ++jj) {
        const float w = Wn[static_cast<long long>(j0 + jj) * K + k];
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = fmaf(s_d[jj * BT + b], w, acc[b]);
      }
    }
  }
  if (k >= K) return;
#pragma unroll
  for (int b = 0; b < BT; ++b)
    if (b < B) delta[static_cast<long long>(b) * K + k] = acc[b] * d[static_cast<long long>(b) * K + k];
}

struct UpdArgs {
  const float* delta;     // [B, H]
  const float* a_prev;    // [B, K] or the feature matrix with idx
  long long a_stride;
  const int* idx;
  float* W;               // [H, K]
  float* bias;            // [H]
  float* mW; float* vW; float* mb; float* vb;
  int B, K, H;
  float beta1, beta2, eps, wd, step_size, bc2_sqrt;
};

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const UpdArgs& a) {
  g = fmaf(a.wd, p, g);                                   // weight_decay: grad = grad + wd * param
  m = m + (g - m) * (1.0f - a.beta1);                     // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(v, a.beta2, (1.0f - a.beta2) * g * g);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

// grid (ceil(K/256), ceil(H/8)); thread = one input column k, 8 output rows j
template <int BT>
__global__ void __launch_bounds__(256) train_update_kernel(UpdArgs a) {
  __shared__ float s_dl[8][BT];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * 8;
  for (int t = threadIdx.x; t < 8 * BT; t += 256) {
    const int jj = t / BT, b = t - jj * BT;
    s_dl[jj][b] = (j0 + jj < a.H && b < a.B) ? a.delta[static_cast<long long>(b) * a.H + j0 + jj] : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 8 && j0 + threadIdx.x < a.H) {  // bias of the 8 rows
    const int j = j0 + threadIdx.x;
    float g = 0.f;
    for (int b = 0; b < a.B; ++b) g += s_dl[threadIdx.x][b];
    float p = a.bias[j], m = a.mb[j], v = a.vb[j];
    adam_update(p, m, v, g, a);
    a.bias[j] = p; a.mb[j] = m; a.vb[j] = v;
  }
  if (k >= a.K) return;
  float x[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    if (b < a.B) {
      const long long row = a.idx ? a.idx[b] : b;
      x[b] = a.a_prev[row * a.a_stride + k];
    } else {
      x[b] = 0.f;
    }
  }
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= a.H) break;
    float g = 0.f;
#pragma unroll
    for (int b = 0; b < BT; ++b) g = fmaf(s_dl[jj][b], x[b], g);
    const long long o = static_cast<long long>(j) * a.K + k;
    float p = a.W[o], m = a.mW[o], v = a.vW[o];
    adam_update(p, m, v, g, a);
    a.W[o] = p; a.mW[o] = m; a.vW[o] = v;
  }
}

}  // namespace b2c

struct b2c_trainer {
  b2c_trainer_cfg cfg;
  int L;
  float* W[B2C_MLP_MAX_LAYERS];
  float* b[B2C_MLP_MAX_LAYERS];
  float* mW[B2C_MLP_MAX_LAYERS];
  float* vW[B2C_MLP_MAX_LAYERS];
  float* mb[B2C_MLP_MAX_LAYERS];
  float* vb[B2C_MLP_MAX_LAYERS];
  float* act[B2C_MLP_MAX_LAYERS];    // a_l   [max_batch, dims[l+1]]
  float* der[B2C_MLP_MAX_LAYERS];    // d_l
  float* delta[B2C_MLP_MAX_LAYERS];  // delta_l
  float* zpart;                      // split-K partial sums of the widest forward layer
  int ksplit[B2C_MLP_MAX_LAYERS], tiles_per_split[B2C_MLP_MAX_LAYERS];
  unsigned long long step;           // optimiser steps taken (Adam's `step`, also the dropout stream position)
  std::vector<void*> allocs;
};

namespace b2c {

static int tr_alloc(b2c_trainer* t, float** p, size_t n) {
  void* q = nullptr;
  B2C_CHECK_CUDA(cudaMalloc(&q, n * sizeof(float)));
  B2C_CHECK_CUDA(cudaMemset(q, 0, n * sizeof(float)));
  t->allocs.push_back(q);
  *p = static_cast<float*>(q);
  return 0;
}

template <int BT>
static int trainer_step_bt(b2c_trainer* t, const float* feats, long long fstride, const float* labels, const int* idx, int B,
                           const b2c_adam* h, float step_size, float bc2_sqrt, float* loss_sum, cudaStream_t st) {
  const int L = t->L;
  const float p = t->cfg.dropout_p;
  static PerDeviceFlag attr_once;  // one flag per BT instantiation
  if (attr_once.first_use()) {
    B2C_CHECK_CUDA(cudaFuncSetAttribute(train_fwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        BT * kTrKT * static_cast<int>(sizeof(float))));
  }
  for (int l = 0; l < L; ++l) {
    FwdArgs a;
    a.in = l == 0 ? feats : t->act[l - 1];
    a.in_stride = l == 0 ? fstride : t->cfg.dims[l];
    a.idx = l == 0 ? idx : nullptr;
    a.W = t->W[l]; a.bias = t->b[l];
    a.a_out = t->act[l]; a.d_out = t->der[l]; a.delta_out = t->delta[l];
    a.labels = labels; a.idx_labels = idx; a.loss_sum = loss_sum;
    a.B = B; a.K = t->cfg.dims[l]; a.H = t->cfg.dims[l + 1]; a.last = l == L - 1; a.layer = l;
    a.slope = t->cfg.leaky_slope; a.p = p; a.inv_keep = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
    a.seed = t->cfg.seed; a.step = t->step;
    a.zpart = t->zpart; a.ksplit = t->ksplit[l]; a.tiles_per_split = t->tiles_per_split[l];
    dim3 grid((a.H + 7) / 8, a.ksplit);
    train_fwd_kernel<BT><<<grid, 256, static_cast<size_t>(B) * kTrKT * sizeof(float), st>>>(a);
    B2C_POST_LAUNCH("train_fwd_kernel");
    if (a.ksplit > 1) {
      train_finish_kernel<<<(B * a.H + 255) / 256, 256, 0, st>>>(a);
      B2C_POST_LAUNCH("train_finish_kernel");
    }
  }
  for (int l = L - 2; l >= 0; --l) {
    const int K = t->cfg.dims[l + 1], Hn = t->cfg.dims[l + 2];
    train_delta_kernel<BT><<<(K + 255) / 256, 256, 128 * BT * sizeof(float), st>>>(t->delta[l + 1], t->W[l + 1], t->der[l],
                                                                                  t->delta[l], B, K, Hn);
    B2C_POST_LAUNCH("train_delta_kernel");
  }
  for (int l = 0; l < L; ++l) {
    UpdArgs a;
    a.delta = t->delta[l];
    a.a_prev = l == 0 ? feats : t->act[l - 1];
    a.a_stride = l == 0 ? fstride : t->cfg.dims[l];
    a.idx = l == 0 ? idx : nullptr;
    a.W = t->W[l]; a.bias = t->b[l]; a.mW = t->mW[l]; a.vW = t->vW[l]; a.mb = t->mb[l]; a.vb = t->vb[l];
    a.B = B; a.K = t->cfg.dims[l]; a.H = t->cfg.dims[l + 1];
    a.beta1 = h->beta1; a.beta2 = h->beta2; a.eps = h->eps; a.wd = h->weight_decay;
    a.step_size = step_size; a.bc2_sqrt = bc2_sqrt;
    dim3 grid((a.K + 255) / 256, (a.H + 7) / 8);
    train_update_kernel<BT><<<grid, 256, 0, st>>>(a);
    B2C_POST_LAUNCH("train_update_kernel");
  }
  return 0;
}

}  // namespace b2c

extern "C" int b2c_trainer_create(const b2c_trainer_cfg* cfg, b2c_trainer** out) {
  using namespace b2c;
  B2C_REQUIRE(cfg && out, "b2c_trainer_create: null pointer");
  B2C_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= B2C_MLP_MAX_LAYERS, "b2c_trainer_create: n_layers=%d", cfg->n_layers);
  B2C_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= kTrMaxBatch, "b2c_trainer_create: max_batch=%d (1..%d)", cfg->max_batch,
              kTrMaxBatch);
  B2C_REQUIRE(cfg->dropout_p >= 0.f && cfg->dropout_p < 1.f, "b2c_trainer_create: dropout_p=%f", cfg->dropout_p);
  for (int l = 0; l <= cfg->n_layers; ++l)
    B2C_REQUIRE(cfg->dims[l] >= 1 && cfg->dims[l] <= (1 << 20), "b2c_trainer_create: dims[%d]=%d", l, cfg->dims[l]);
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    return set_error(B2C_ERR_CUDA, "b2c_trainer_create: no CUDA device (there is no CPU fallback)");
  b2c_trainer* t = new (std::nothrow) b2c_trainer();
  B2C_REQUIRE(t != nullptr, "b2c_trainer_create: out of memory");
  t->cfg = *cfg;
  t->L = cfg->n_layers;
  t->step = 0;
  t->zpart = nullptr;
  size_t zpart_elems = 0;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  for (int l = 0; l < t->L; ++l) {  // split K until the forward grid covers about two waves of SMs
    const int tiles = (cfg->dims[l] + kTrKT - 1) / kTrKT, jblocks = (cfg->dims[l + 1] + 7) / 8;
    int ks = 1;
    while (ks * 2 <= tiles && jblocks * ks < 2 * sms) ks *= 2;
    t->tiles_per_split[l] = (tiles + ks - 1) / ks;
    t->ksplit[l] = (tiles + t->tiles_per_split[l] - 1) / t->tiles_per_split[l];
    if (t->ksplit[l] > 1)
      zpart_elems = std::max(zpart_elems, static_cast<size_t>(t->ksplit[l]) * cfg->max_batch * cfg->dims[l + 1]);
  }
  if (zpart_elems && tr_alloc(t, &t->zpart, zpart_elems) != 0) {
    b2c_trainer_destroy(t);
    return B2C_ERR_CUDA;
  }
  for (int l = 0; l < t->L; ++l) {
    const size_t K = cfg->dims[l], H = cfg->dims[l + 1], B = cfg->max_batch;
    int rc = 0;
    rc |= tr_alloc(t, &t->W[l], H * K); rc |= tr_alloc(t, &t->mW[l], H * K); rc |= tr_alloc(t, &t->vW[l], H * K);
    rc |= tr_alloc(t, &t->b[l], H); rc |= tr_alloc(t, &t->mb[l], H); rc |= tr_alloc(t, &t->vb[l], H);
    rc |= tr_alloc(t, &t->act[l], B * H); rc |= tr_alloc(t, &t->der[l], B * H); rc |= tr_alloc(t, &t->delta[l], B * H);
    if (rc != 0) {
      b2c_trainer_destroy(t);
      return B2C_ERR_CUDA;
    }
  }
  *out = t;
  return 0;
}

extern "C" int b2c_trainer_destroy(b2c_trainer* t) {
  if (t == nullptr) return 0;
  for (void* p : t->allocs) cudaFree(p);
  delete t;
  return 0;
}

extern "C" int b2c_trainer_set_layer(b2c_trainer* t, int layer, const float* W, const float* bias, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(t && W && bias, "b2c_trainer_set_layer: null pointer");
  B2C_REQUIRE(layer >= 0 && layer < t->L, "b2c_trainer_set_layer: layer %d", layer);
  const size_t K = t->cfg.dims[layer], H = t->cfg.dims[layer + 1];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2C_CHECK_CUDA(cudaMemcpyAsync(t->W[layer], W, H * K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  B2C_CHECK_CUDA(cudaMemcpyAsync(t->b[layer], bias, H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int b2c_trainer_get_layer(b2c_trainer* t, int layer, float* W, float* bias, b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(t && W && bias, "b2c_trainer_get_layer: null pointer");
  B2C_REQUIRE(layer >= 0 && layer < t->L, "b2c_trainer_get_layer: layer %d", layer);
  const size_t K = t->cfg.dims[layer], H = t->cfg.dims[layer + 1];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2C_CHECK_CUDA(cudaMemcpyAsync(W, t->W[layer], H * K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  B2C_CHECK_CUDA(cudaMemcpyAsync(bias, t->b[layer], H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int b2c_trainer_weights(b2c_trainer* t, b2c_mlp_weights* out) {
  using namespace b2c;
  B2C_REQUIRE(t && out, "b2c_trainer_weights: null pointer");
  out->n_layers = t->L;
  for (int l = 0; l <= t->L; ++l) out->dims[l] = t->cfg.dims[l];
  for (int l = 0; l < t->L; ++l) {
    out->weight[l] = t->W[l];
    out->bias[l] = t->b[l];
  }
  out->leaky_slope = t->cfg.leaky_slope;
  return 0;
}

extern "C" unsigned long long b2c_trainer_steps(const b2c_trainer* t) { return t ? t->step : 0ull; }

extern "C" int b2c_trainer_epoch(b2c_trainer* t, const float* feats, int64_t feat_stride, const float* labels,
                                 const int32_t* order, int64_t n, int batch, const b2c_adam* h, float* loss_sum,
                                 b2c_stream stream) {
  using namespace b2c;
  B2C_REQUIRE(t && feats && labels && order && h, "b2c_trainer_epoch: null pointer");
  B2C_REQUIRE(batch >= 1 && batch <= t->cfg.max_batch, "b2c_trainer_epoch: batch=%d exceeds max_batch=%d", batch,
              t->cfg.max_batch);
  B2C_REQUIRE(feat_stride >= t->cfg.dims[0], "b2c_trainer_epoch: feat_stride %lld < input width %d", (long long)feat_stride,
              t->cfg.dims[0]);
  B2C_REQUIRE(n >= 0, "b2c_trainer_epoch: n=%lld", (long long)n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t off = 0; off < n; off += batch) {
    const int B = static_cast<int>(n - off < batch ? n - off : batch);  // DataLoader keeps the last partial batch
    t->step += 1;
    // torch.optim.Adam: bias corrections in double from the Python step count, cast to the parameter dtype when applied
    const double bc1 = 1.0 - pow(static_cast<double>(h->beta1), static_cast<double>(t->step));
    const double bc2 = 1.0 - pow(static_cast<double>(h->beta2), static_cast<double>(t->step));
    const float step_size = static_cast<float>(static_cast<double>(h->lr) / bc1);
    const float bc2_sqrt = static_cast<float>(sqrt(bc2));
    int rc;
    if (B <= 16)
      rc = trainer_step_bt<16>(t, feats, feat_stride, labels, order + off, B, h, step_size, bc2_sqrt, loss_sum, st);
    else if (B <= 32)
      rc = trainer_step_bt<32>(t, feats, feat_stride, labels, order + off, B, h, step_size, bc2_sqrt, loss_sum, st);
    else
      rc = trainer_step_bt<64>(t, feats, feat_stride, labels, order + off, B, h, step_size, bc2_sqrt, loss_sum, st);
    if (rc != 0) return rc;
  }
  return 0;
}
