// ubm_peaks.cuh — roofline denominators the driver does not measure: FP64 FMA and INT32 throughput.
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; the kernels of this path are bound by the FP64
// vector pipe (C1, C2, C3, C5) and by integer issue (C4), so bench.py measures those two peaks live, once,
// with these kernels: independent dependency chains per thread, enough of them to cover the pipe latency.
#pragma once
#include "ubm_common.cuh"

namespace ubm {

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;   // keep the chains alive
}

__global__ void __launch_bounds__(256) peak_imad_kernel(unsigned* out, int iters, unsigned a, unsigned b) {
  unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  unsigned s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (s == 0x12345678u) out[0] = s;
}

// FP64 tensor pipe: eight independent m8n8k4 accumulator chains per warp (256 fma per instruction)
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a, double b) {
  double d[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) { d[u][0] = threadIdx.x + u; d[u][1] = 0.5 * u; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(d[u][0]), "+d"(d[u][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += d[u][0] + d[u][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace ubm
