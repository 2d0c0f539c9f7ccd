// Microbenchmark: issue throughput per SM of ex2.approx.ftz.f32 vs ex2.approx.f16x2 vs fma.rn.f32x2 vs fma.rn.f32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bw mufu_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int iters, long long* cycles, float* sink, float seed) {
  float x[8];
  uint32_t hx[8];
  unsigned long long px[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
    hx[i] = 0xb800b800u + threadIdx.x + i;  // f16x2 (-0.5, -0.5) + noise
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(px[i]) : "f"(x[i]), "f"(x[i] * 0.5f));
  }
  unsigned long long pc;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(0.999f), "f"(1.001f));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hx[i]));
      if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(px[i]) : "l"(pc));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(seed));
      if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(hx[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(hx[i]) : "f"(x[i]));
      if (MODE == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(hx[i]) : "f"(x[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(px[i]));
    acc += x[i] + __uint_as_float(hx[i]) + lo + hi;
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 4096;
  const char* names[] = {"ex2.f32", "ex2.f16x2", "fma.f32x2", "fma.f32", "ex2.bf16x2", "cvt.bf16x2.f32", "cvt.f16x2.f32"};
  for (int mode = 0; mode < 7; ++mode)
    for (int warps : {4, 16}) {
      switch (mode) {
        case 0: k<0><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 1: k<1><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 2: k<2><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 3: k<3><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 4: k<4><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 5: k<5><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
        case 6: k<6><<<148, warps * 32>>>(iters, cyc, sink, 1.0f); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double insts = (double)iters * 8 * warps;  // warp-instructions per SM
      printf("%-15s warps %2d: %9lld cycles, %.2f warp-inst/clk/SM  (%.1f lanes/clk/SM)  %s\n", names[mode], warps, h, insts / h,
             insts * 32 / h, cudaGetErrorString(e));
    }
  return 0;
}
