// Micro-probe: latency of dependent FP64 sqrt / division / __syncwarp / shuffle steps on sm_100a (one warp).
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int WHAT>
__global__ void chain(double* out, long long* cyc, int iters, double seed) {
  __shared__ double sm[64];
  double x = seed + threadIdx.x * 1e-3;
  sm[threadIdx.x & 63] = x;
  __syncwarp();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (WHAT == 0) x = sqrt(x) + 1.5;
    if (WHAT == 1) x = 3.0 / x + 0.5;
    if (WHAT == 2) x = 1.0 / sqrt(x) + 2.0;
    if (WHAT == 3) { x = x * 1.0000001; __syncwarp(); }
    if (WHAT == 4) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) * 1.0000001;
    if (WHAT == 5) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] * 1.0000001; __syncwarp(); }
    if (WHAT == 6) x = __fma_rn(x, 1.0000001, 1e-9);
    if (WHAT == 8) { const double y = x + 3.0; x = y / x + 0.5; }
    if (WHAT == 9) { const double l = sqrt(x); x = x / l + 2.0; }
    if (WHAT == 10) { const double l = sqrt(x); x = (x + 1.0) * (1.0 / l) + 2.0; }
    if (WHAT == 11) { const double l = sqrt(x); sm[threadIdx.x] = x / l; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 2.0; }
    if (WHAT == 7) { const double l = sqrt(x); sm[threadIdx.x] = x / l; __syncwarp(); x = __fma_rn(-sm[(threadIdx.x + 1) & 31], 0.5, x + 2.0); __syncwarp(); }
  }
  long long t1 = clock64();
  if (x == 1.2345) out[0] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int WHAT> void run(const char* name) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  chain<WHAT><<<1, 32>>>(out, cyc, iters, 2.5); chain<WHAT><<<1, 32>>>(out, cyc, iters, 2.5);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-48s %.1f cycles per step\n", name, (double)c / iters);
}
int main() {
  run<6>("dfma (dependent)");
  run<0>("sqrt + add (dependent)");
  run<1>("div + add (dependent)");
  run<2>("1/sqrt + add (dependent)");
  run<3>("dmul + __syncwarp");
  run<4>("shfl + dmul (dependent)");
  run<5>("sts, syncwarp, lds, dmul, syncwarp");
  run<8>("variable / variable + add");
  run<9>("sqrt, div (x / sqrt(x)) + add");
  run<10>("sqrt, reciprocal, mul + add");
  run<11>("sqrt, div, sts, syncwarp, lds, add");
  run<7>("cholesky-like column step (sqrt, div, sts, sync, lds, fma, sync)");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
