// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2) throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
      if (MODE == 1) { a[i] = __ffma2_rn(a[i], b, c); }
      if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
      if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(d, 100);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * (double)iters * 16;  // scalar flop-ops (fma counted once)
  printf("%-8s %8.3f ms  %7.2f T scalar-ops/s\n", name, ms, ops / ms / 1e9);
  cudaFree(d);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); run<2>("FADD"); run<3>("FADD2"); return 0; }
