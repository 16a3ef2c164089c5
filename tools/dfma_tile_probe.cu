// Micro-probe: DFMA issue rate of register-tiled outer products acc[r][c] = fma(x[r], m[c], acc[r][c]) on sm_100a,
// against the plain independent chains x = fma(x, a, b) used for the roofline denominator.
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/dfma_tile_probe tools/dfma_tile_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int RT, int CT>
__global__ void tile(double* out, long long* cyc, int iters, double s) {
  double acc[RT][CT], x[RT], m[CT];
  for (int r = 0; r < RT; ++r) { x[r] = s + r + threadIdx.x * 1e-3; for (int c = 0; c < CT; ++c) acc[r][c] = 0.0; }
  for (int c = 0; c < CT; ++c) m[c] = 1.0 / (s + c + 1);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = __fma_rn(x[r], m[c], acc[r][c]);
  }
  long long t1 = clock64();
  double sum = 0;
  for (int r = 0; r < RT; ++r) for (int c = 0; c < CT; ++c) sum += acc[r][c];
  if (sum == 1.2345) out[0] = sum;
  if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}
template <int RT, int CT> void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 64);
  const int iters = 4096;
  tile<RT, CT><<<1, 32 * warps>>>(out, cyc, iters, 2.5); tile<RT, CT><<<1, 32 * warps>>>(out, cyc, iters, 2.5);
  long long cs[64]; cudaMemcpy(cs, cyc, 8 * warps, cudaMemcpyDeviceToHost);
  long long c = 0; for (int w = 0; w < warps; ++w) c = cs[w] > c ? cs[w] : c;
  const double per_smsp = (double)c / ((double)iters * RT * CT * ((warps + 3) / 4));
  printf("tile %dx%d, %2d warps (%d per sub-partition): %.2f cycles per warp-DFMA per sub-partition\n", RT, CT, warps, (warps + 3) / 4, per_smsp);
}
int main() {
  run<1, 8>(4); run<2, 4>(4); run<2, 4>(8); run<2, 4>(16); run<4, 4>(4); run<4, 4>(8); run<4, 4>(16); run<4, 8>(4); run<4, 8>(8);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
