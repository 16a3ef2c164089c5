// Micro-probe: DFMA dependent-issue latency and single-warp throughput vs ILP on sm_100a; LDS.128 latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __fma_rn(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lds_chain(int* out, long long* cyc, int iters) {
  __shared__ int buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) p = buf[p];
  long long t1 = clock64();
  if (p == -1) out[0] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps, int blocks) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  int iters = 4096;
  dfma_chain<ILP><<<blocks, 32 * warps>>>(out, cyc, iters, 0.9999, 1e-9);
  dfma_chain<ILP><<<blocks, 32 * warps>>>(out, cyc, iters, 0.9999, 1e-9);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA ILP=%2d warps/CTA=%2d: %.2f cycles per iteration (%d dfma) -> %.2f cycles/dfma/warp\n", ILP, warps,
         (double)c / iters, ILP, (double)c / iters / ILP);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1, 1); run<2>(1, 1); run<4>(1, 1); run<8>(1, 1); run<16>(1, 1); run<32>(1, 1);
  run<8>(4, 1); run<8>(8, 1); run<8>(16, 1); run<16>(4, 1); run<16>(8, 1); run<32>(4, 1);
  int* o; long long* cyc; cudaMalloc(&o, 4); cudaMalloc(&cyc, 8);
  lds_chain<<<1, 32>>>(o, cyc, 4096); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent chain: %.1f cycles per load\n", (double)c / 4096);
  return 0;
}
