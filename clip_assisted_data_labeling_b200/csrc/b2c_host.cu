// b2c_host.cu — error plumbing, driver entry points, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "b2c_host.h"
#include "b2c_launch.h"

namespace b2c {

static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

int make_tmap_2d_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_cols, int dtype) {
  return make_tmap_2d_sw(out, base, rows, cols, row_stride_bytes, box_rows, box_cols, dtype, 128);
}

int make_tmap_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_cols, int dtype, int swizzle_bytes) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return set_error(B2C_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  B2C_REQUIRE(row_stride_bytes % 16 == 0, "TMA row stride %llu B is not a multiple of 16",
              (unsigned long long)row_stride_bytes);
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer is not 16-byte aligned");
  CUtensorMapDataType dt;
  size_t esz;
  switch (dtype) {
    case B2C_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; esz = 4; break;
    case B2C_F16: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; esz = 2; break;
    case B2C_BF16: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; esz = 2; break;
    default: return set_error(B2C_ERR_ARG, "make_tmap: unsupported dtype %d", dtype);
  }
  B2C_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64 || swizzle_bytes == 32, "make_tmap: swizzle %d", swizzle_bytes);
  B2C_REQUIRE(box_cols * esz == static_cast<size_t>(swizzle_bytes), "make_tmap: box of %u columns is not one %d-byte swizzle row",
              box_cols, swizzle_bytes);
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B2C_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu stride=%llu)",
                     (int)r, (unsigned long long)rows, (unsigned long long)cols,
                     (unsigned long long)row_stride_bytes);
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, int elem) {
  return make_tmap_2d_ex(out, base, rows, cols, row_stride_bytes, box_rows, 64, elem == 1 ? B2C_BF16 : B2C_F16);
}

// ---- stage timer ------------------------------------------------------------------------------------
std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec {
  int kind;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_begin(int kind, cudaStream_t stream) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  ProfRec r{kind, prof_event(), prof_event()};
  cudaEventRecord(r.e0, stream);
  g_prof_recs.push_back(r);
}

void prof_end(cudaStream_t stream) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, stream);
}

namespace {
struct PinnedChunk {
  void* host = nullptr;
  size_t bytes = 0;
  cudaEvent_t done = nullptr;  // created on `dev`: only recorded / queried with that device current
  int dev = 0;
  bool in_flight = false;
};
std::mutex g_pin_mu;
std::vector<PinnedChunk> g_pin;
}  // namespace

int upload_async(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  std::lock_guard<std::mutex> g(g_pin_mu);
  const int dev = current_device();
  PinnedChunk* c = nullptr;
  for (PinnedChunk& k : g_pin) {
    if (k.bytes < bytes || k.dev != dev) continue;
    if (k.in_flight && cudaEventQuery(k.done) != cudaSuccess) continue;
    k.in_flight = false;
    if (!c || k.bytes < c->bytes) c = &k;
  }
  (void)cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error of ours
  if (!c) {
    PinnedChunk k;
    k.dev = dev;
    k.bytes = 1 << 16;
    while (k.bytes < bytes) k.bytes <<= 1;
    B2C_CHECK_CUDA(cudaHostAlloc(&k.host, k.bytes, cudaHostAllocDefault));
    B2C_CHECK_CUDA(cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
    g_pin.push_back(k);
    c = &g_pin.back();
  }
  memcpy(c->host, src_host, bytes);
  B2C_CHECK_CUDA(cudaMemcpyAsync(dst_dev, c->host, bytes, cudaMemcpyHostToDevice, stream));
  B2C_CHECK_CUDA(cudaEventRecord(c->done, stream));
  c->in_flight = true;
  return 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return dev;
}

int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::atomic<int>& slot = n[dev & (kMaxDevices - 1)];
  int v = slot.load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 0;
    slot.store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace b2c

extern "C" const char* b2c_last_error(void) { return b2c::g_err; }
extern "C" int b2c_version(void) { return 100; }
extern "C" unsigned long long b2c_launch_count(void) { return b2c::g_launches.load(); }

extern "C" int b2c_prof_enable(int on) {
  using namespace b2c;
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (on) {
    for (ProfRec& r : g_prof_recs) {
      g_prof_pool.push_back(r.e0);
      g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
  }
  g_prof_on.store(on ? 1 : 0);
  return 0;
}

extern "C" int b2c_prof_read(double* ms, unsigned long long* stages) {
  using namespace b2c;
  B2C_REQUIRE(ms && stages, "b2c_prof_read: null pointer");
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (int k = 0; k < B2C_PROF_KINDS; ++k) {
    ms[k] = 0.0;
    stages[k] = 0;
  }
  for (const ProfRec& r : g_prof_recs) {
    B2C_CHECK_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    B2C_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.kind] += t;
    stages[r.kind] += 1;
  }
  return 0;
}

extern "C" const char* b2c_prof_kind_name(int kind) {
  static const char* names[B2C_PROF_KINDS] = {"preprocess", "patch_embed", "layernorm", "in_proj", "attention", "out_proj",
                                              "c_fc",       "c_proj",      "head",      "dedup",   "other"};
  return kind >= 0 && kind < B2C_PROF_KINDS ? names[kind] : "?";
}
