// Micro-probe: shared-memory pipe cost (cycles per warp-level LDS instruction per SM) of the access patterns the MVN
// GEMM phases use, on sm_100a.  8 warps per CTA, one CTA, independent loads (throughput, not latency).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lds_probe tools/lds_probe.cu && tools/lds_probe
#include <cstdio>
#include <cuda_runtime.h>

// pattern -> address (in doubles) for a lane; `it` advances all lanes together
__device__ __forceinline__ int addr_of(int pat, int lane, int it) {
  const int k = (it & 15) * 2;
  switch (pat) {
    case 0: return k;                                                 // all lanes the same 16 bytes
    case 1: return ((lane >> 4) * 2 + (lane & 1)) * 512 + k;          // 4 distinct addresses, far apart
    case 2: return ((lane >> 1) & 7) * 102 + k;                       // 8 distinct rows (stride 102 doubles), same k
    case 3: return ((lane >> 1) & 7) * 102 + ((lane >> 4) * 2 + (lane & 1)) * 4 + k;   // 8 rows x 4 k-offsets (32 distinct)
    case 4: return lane * 2 + 64 * (it & 15);                         // 32 consecutive 16-byte words (512 B)
    case 5: return (lane & 15) * 102 + k;                             // 16 distinct rows, same k (2 lanes each)
    case 6: return (lane & 7) * 4 + 32 * (it & 15);                   // 8 distinct 32-byte-strided words, each 4 lanes
    case 7: return (lane & 7) * 2 + 16 * (it & 15);                   // 8 consecutive 16-byte words (128 B), 4 lanes each
    default: return 0;
  }
}

template <int W>   // W = 2 (LDS.128 as double2) or 1 (LDS.64)
__global__ void probe(int pat, int iters, double* out, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it += 4) {
    if (W == 2) {
      const double2 a = *reinterpret_cast<const double2*>(sm + addr_of(pat, lane, it));
      const double2 b = *reinterpret_cast<const double2*>(sm + addr_of(pat, lane, it + 1));
      const double2 c = *reinterpret_cast<const double2*>(sm + addr_of(pat, lane, it + 2));
      const double2 d = *reinterpret_cast<const double2*>(sm + addr_of(pat, lane, it + 3));
      acc0 += a.x + a.y; acc1 += b.x + b.y; acc2 += c.x + c.y; acc3 += d.x + d.y;
    } else {
      acc0 += sm[addr_of(pat, lane, it)]; acc1 += sm[addr_of(pat, lane, it + 1)];
      acc2 += sm[addr_of(pat, lane, it + 2)]; acc3 += sm[addr_of(pat, lane, it + 3)];
    }
  }
  long long t1 = clock64();
  const double s = acc0 + acc1 + acc2 + acc3;
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"all lanes same 16 B", "4 distinct addresses", "8 rows, same k", "8 rows x 4 k-offsets (32 distinct)",
                         "32 consecutive words (512 B)", "16 rows, same k", "8 words stride 32 B", "8 consecutive words (128 B)"};
  const int iters = 8192, warps = 8;
  for (int W = 2; W >= 1; --W)
    for (int pat = 0; pat < 8; ++pat) {
      for (int rep = 0; rep < 2; ++rep) {
        if (W == 2) probe<2><<<1, 32 * warps, 6144 * 8>>>(pat, iters, out, cyc);
        else probe<1><<<1, 32 * warps, 6144 * 8>>>(pat, iters, out, cyc);
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("LDS.%d  %-36s %.2f cycles per warp-load (SM-wide, %d warps)\n", W == 2 ? 128 : 64, names[pat],
             (double)c / ((double)iters * warps), warps);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
