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
    case 8: return (lane >> 3) * 512 + k;                             // 4 distinct addresses, same banks, 8 contiguous lanes each
    case 9: return (lane & 1) * 512 + k;                              // 2 distinct addresses, same banks
    case 10: return lane * 64 + k;                                    // 32 distinct addresses, same banks (32-way conflict)
    case 11: case 12: case 13: case 14: {                             // the MVN kernel's own patterns, warp 0 / warp 2 (d = 100)
      const int w = (pat == 11 || pat == 13) ? 0 : 2;
      const int pr = 2 * w + (lane >> 4), h = lane & 1, rg = (lane >> 1) & 7;
      const int g = h ? 24 - pr : pr, k0 = h ? 48 - 4 * pr : 0;
      if (pat <= 12) return (8 * g * (g + 1) + 2 * g + 4 * (k0 + k)) % 5600;        // matrix load
      return rg * 102 + (k0 + k) % 100;                                              // row load
    }
    default: return 0;
  }
}

template <int W>   // W = 2 (LDS.128) or 1 (LDS.64)
__global__ void probe(int pat, int iters, unsigned* out, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + 8u * (unsigned)addr_of(pat, lane, 0);
  const unsigned step = 8u * (unsigned)(addr_of(pat, lane, 1) - addr_of(pat, lane, 0));
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unsigned x, y, z, w;
      if (W == 2) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + j * step));
      else { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(base + j * step)); z = w = 0; }
      acc ^= x ^ y ^ z ^ w;
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  if (lane == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

int main() {
  unsigned* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 32);
  const char* names[] = {"all lanes same 16 B", "4 distinct addresses", "8 rows, same k", "8 rows x 4 k-offsets (32 distinct)",
                         "32 consecutive words (512 B)", "16 rows, same k", "8 words stride 32 B", "8 consecutive words (128 B)",
                         "4 distinct same banks (8 lanes each)", "2 distinct same banks", "32 distinct same banks", "MVN matrix load warp 0",
                         "MVN matrix load warp 2", "MVN row load warp 0", "MVN row load warp 2"};
  const int iters = 8192, warps = 8;
  for (int W = 2; W >= 1; --W)
    for (int pat = 0; pat < 15; ++pat) {
      for (int rep = 0; rep < 2; ++rep) {
        if (W == 2) probe<2><<<1, 32 * warps, 6144 * 8>>>(pat, iters, out, cyc);
        else probe<1><<<1, 32 * warps, 6144 * 8>>>(pat, iters, out, cyc);
      }
      long long cs[8]; cudaMemcpy(cs, cyc, 8 * warps, cudaMemcpyDeviceToHost);
      long long c = 0; for (int w = 0; w < warps; ++w) c = cs[w] > c ? cs[w] : c;
      printf("LDS.%d  %-36s %.2f cycles per warp-load (SM-wide, %d warps)\n", W == 2 ? 128 : 64, names[pat],
             (double)c / ((double)iters * warps), warps);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
