// b2c_host.h — host-side helpers shared by the translation units of libb2c.so:
// thread-local error string, CUDA error checks that never throw, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b2c.h"

namespace b2c {

// Sets the thread-local message returned by b2c_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);

#define B2C_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::b2c::set_error(B2C_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,          \
                              cudaGetErrorString(_e));                                            \
  } while (0)

#define B2C_REQUIRE(cond, ...)                                                                    \
  do {                                                                                            \
    if (!(cond)) return ::b2c::set_error(B2C_ERR_ARG, __VA_ARGS__);                               \
  } while (0)

#define B2C_TRY(expr)                                                                             \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != 0) return _r;                                                                       \
  } while (0)

// 2-D row-major tensor [rows, cols] of 16-bit elements, box = [box_rows, 64 cols], 128-B swizzle.
// elem: 0 = fp16, 1 = bf16.  row_stride_bytes must be a multiple of 16.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, int elem);
// General form: dtype B2C_F32 / B2C_F16 / B2C_BF16, box = [box_rows, box_cols] with box_cols * elem_size == 128 B
// (one 128-B swizzle row).  Used for the epilogue's TMA stores.
int make_tmap_2d_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_cols, int dtype);

// Same with a 32-, 64- or 128-byte swizzle; box_cols * elem_size must equal the swizzle span.
int make_tmap_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                    uint32_t box_rows, uint32_t box_cols, int dtype, int swizzle_bytes);

// SM count of the CURRENT device (cached per device).
int num_sms();

// Per-device one-time state.  Function attributes (opt-in dynamic shared memory), the SM count and pinned-pool events
// belong to a device, and the Python surface accepts device= everywhere: a process that drives two GPUs must not
// reuse device 0's "already done" for device 1.
constexpr int kMaxDevices = 64;
int current_device();  // cudaGetDevice(), 0 on error
struct PerDeviceFlag {
  std::atomic<unsigned long long> mask{0};
  // true exactly once per device (thread-safe): the caller then performs the per-device initialisation
  bool first_use() {
    const unsigned long long bit = 1ull << (current_device() & (kMaxDevices - 1));
    return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
  }
};
struct PerDeviceMax {
  std::atomic<long long> v[kMaxDevices];
  PerDeviceMax() { for (auto& a : v) a.store(0); }
  // true when `want` exceeds what this device was last configured for (and records it)
  bool raise(long long want) {
    std::atomic<long long>& a = v[current_device() & (kMaxDevices - 1)];
    long long cur = a.load(std::memory_order_acquire);
    while (want > cur)
      if (a.compare_exchange_weak(cur, want, std::memory_order_acq_rel)) return true;
    return false;
  }
};

// Descriptor upload that never blocks the host: cudaMemcpyAsync from PAGEABLE memory first synchronises the stream (the
// caller would wait for every kernel already queued, i.e. lose all host/device overlap), so small host tables (crop
// plans, job lists) are copied into a chunk of a process-wide pinned pool and sent from there.  A chunk is reused once
// the event recorded after its copy has completed.  Thread-safe.
int upload_async(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream);

}  // namespace b2c
