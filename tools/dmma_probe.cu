// Micro-probe: FP64 tensor-core MMA (DMMA) on sm_100a — (1) the ORDER in which one instruction accumulates its k products
// (needed to restate it bit for bit on the host), (2) sustained throughput against the DFMA pipe.
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

// D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c/d = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// m16n8k4 / k8 / k16 (sm_90+).  A: 16 x K row-major fragments, B: K x 8, C/D: 16 x 8 (4 doubles per lane)
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b, const double (&c)[4]) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};\n"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2], const double (&c)[4]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
               : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]));
}

// ---- semantics: one m8n8k4 per test, all 32 lanes
__global__ void sem884(const double* A, const double* B, const double* C, double* D, int ntests) {
  const int lane = threadIdx.x;
  for (int t = 0; t < ntests; ++t) {
    const double* a = A + t * 32; const double* b = B + t * 32; const double* c = C + t * 64;
    double d0, d1;
    dmma884(d0, d1, a[(lane / 4) * 4 + lane % 4], b[(lane % 4) * 8 + lane / 4], c[(lane / 4) * 8 + 2 * (lane % 4)], c[(lane / 4) * 8 + 2 * (lane % 4) + 1]);
    D[t * 64 + (lane / 4) * 8 + 2 * (lane % 4)] = d0;
    D[t * 64 + (lane / 4) * 8 + 2 * (lane % 4) + 1] = d1;
  }
}
// m16n8k8: A 16x8 row-major: a[0]=A[g][t], a[1]=A[g+8][t], a[2]=A[g][t+4], a[3]=A[g+8][t+4] (g = lane/4, t = lane%4);
// B 8x8: b[0]=B[t][g], b[1]=B[t+4][g]; C: c[0]=C[g][2t], c[1]=C[g][2t+1], c[2]=C[g+8][2t], c[3]=C[g+8][2t+1]
__global__ void sem1688(const double* A, const double* B, const double* C, double* D, int ntests) {
  const int lane = threadIdx.x, g = lane / 4, t4 = lane % 4;
  for (int t = 0; t < ntests; ++t) {
    const double* a = A + t * 128; const double* b = B + t * 64; const double* c = C + t * 128;
    double av[4] = {a[g * 8 + t4], a[(g + 8) * 8 + t4], a[g * 8 + t4 + 4], a[(g + 8) * 8 + t4 + 4]};
    double bv[2] = {b[t4 * 8 + g], b[(t4 + 4) * 8 + g]};
    double cv[4] = {c[g * 8 + 2 * t4], c[g * 8 + 2 * t4 + 1], c[(g + 8) * 8 + 2 * t4], c[(g + 8) * 8 + 2 * t4 + 1]};
    double dv[4];
    dmma1688(dv, av, bv, cv);
    D[t * 128 + g * 8 + 2 * t4] = dv[0]; D[t * 128 + g * 8 + 2 * t4 + 1] = dv[1];
    D[t * 128 + (g + 8) * 8 + 2 * t4] = dv[2]; D[t * 128 + (g + 8) * 8 + 2 * t4 + 1] = dv[3];
  }
}

// ---- throughput: NACC independent accumulator tiles per warp
template <int KIND, int NACC>
__global__ void thr(double* out, int iters, double x, double y) {
  double acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 1e-9 + i; }
  double a4[4] = {x, x + 1e-9, x + 2e-9, x + 3e-9}, b2[2] = {y, y + 1e-9};
  double a2[2] = {x, x + 1e-9};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0) dmma884(acc[i][0], acc[i][1], x, y, acc[i][0], acc[i][1]);
      if (KIND == 1) dmma1684(acc[i], a2, y, acc[i]);
      if (KIND == 2) dmma1688(acc[i], a4, b2, acc[i]);
      if (KIND == 3) {   // plain DFMA: 4 per "instruction slot" for comparison
        acc[i][0] = __fma_rn(acc[i][0], x, y); acc[i][1] = __fma_rn(acc[i][1], x, y);
        acc[i][2] = __fma_rn(acc[i][2], x, y); acc[i][3] = __fma_rn(acc[i][3], x, y);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 1.2345) out[0] = s;
}

template <int KIND, int NACC>
void run_thr(const char* name, double fma_per_instr, int nsm, int warps_per_sm) {
  double* out; cudaMalloc(&out, 8);
  const int iters = 2048;
  const int threads = 32 * warps_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    thr<KIND, NACC><<<nsm, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double fma = (double)nsm * warps_per_sm * iters * NACC * fma_per_instr;
  printf("%-28s warps/SM %2d  acc tiles %d : %8.2f TFLOP/s (fma = 2 flop)   %s\n", name, warps_per_sm, NACC, 2.0 * fma / (best * 1e-3) * 1e-12,
         cudaGetErrorString(cudaGetLastError()));
}

// ---- latency: one warp, NCH independent accumulator chains, dependent DMMAs; cycles per DMMA issued
template <int NCH>
__global__ void lat(double* out, long long* cyc, int iters, double x, double y) {
  double acc[NCH][2];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { acc[i][0] = threadIdx.x * 1e-9 + i; acc[i][1] = 1.0; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma884(acc[i][0], acc[i][1], x, y, acc[i][0], acc[i][1]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i][0] + acc[i][1];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NCH>
void run_lat(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  lat<NCH><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9); lat<NCH><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA m8n8k4: %d warp(s) on one SM, %d chains per warp: %.1f cycles per loop trip = %.1f cycles per DMMA per warp\n", warps, NCH, (double)c / iters,
         (double)c / iters / NCH);
}

static double rnd(int wide) {
  double u = (double)rand() / RAND_MAX - 0.5;
  if (!wide) return u;
  return u * std::pow(2.0, (rand() % 41) - 20);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  // ---------------- semantics, m8n8k4
  {
    const int nt = 2000;
    std::vector<double> A(nt * 32), B(nt * 32), C(nt * 64), D(nt * 64);
    for (int t = 0; t < nt; ++t) {
      for (int i = 0; i < 32; ++i) { A[t * 32 + i] = rnd(t & 1); B[t * 32 + i] = rnd(t & 1); }
      for (int i = 0; i < 64; ++i) C[t * 64 + i] = rnd(t & 1);
    }
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
    sem884<<<1, 32>>>(dA, dB, dC, dD, nt);
    cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
    printf("sem884: %s\n", cudaGetErrorString(cudaGetLastError()));
    long ok_fwd = 0, ok_rev = 0, ok_pair = 0, ok_exact = 0, tot = 0;
    for (int t = 0; t < nt; ++t)
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
          const double* a = &A[t * 32 + i * 4]; const double c = C[t * 64 + i * 8 + j];
          double bb[4]; for (int k = 0; k < 4; ++k) bb[k] = B[t * 32 + k * 8 + j];
          double f = c; for (int k = 0; k < 4; ++k) f = fma(a[k], bb[k], f);
          double r = c; for (int k = 3; k >= 0; --k) r = fma(a[k], bb[k], r);
          double p = fma(a[1], bb[1], a[0] * bb[0]) + fma(a[3], bb[3], a[2] * bb[2]) + c;
          long double e = c; for (int k = 0; k < 4; ++k) e += (long double)a[k] * bb[k];
          const double d = D[t * 64 + i * 8 + j];
          ok_fwd += (d == f); ok_rev += (d == r); ok_pair += (d == p); ok_exact += (d == (double)e); ++tot;
        }
    printf("m8n8k4  : of %ld outputs  fma chain k=0,1,2,3: %ld   k=3,2,1,0: %ld   pairwise: %ld   exact-sum-then-round: %ld\n", tot, ok_fwd, ok_rev, ok_pair, ok_exact);
  }
  // ---------------- semantics, m16n8k8
  {
    const int nt = 1000;
    std::vector<double> A(nt * 128), B(nt * 64), C(nt * 128), D(nt * 128);
    for (int t = 0; t < nt; ++t) {
      for (int i = 0; i < 128; ++i) { A[t * 128 + i] = rnd(t & 1); C[t * 128 + i] = rnd(t & 1); }
      for (int i = 0; i < 64; ++i) B[t * 64 + i] = rnd(t & 1);
    }
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
    sem1688<<<1, 32>>>(dA, dB, dC, dD, nt);
    cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
    printf("sem1688: %s\n", cudaGetErrorString(cudaGetLastError()));
    long ok_fwd = 0, ok_rev = 0, ok_48 = 0, tot = 0;
    for (int t = 0; t < nt; ++t)
      for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 8; ++j) {
          const double* a = &A[t * 128 + i * 8]; const double c = C[t * 128 + i * 8 + j];
          double bb[8]; for (int k = 0; k < 8; ++k) bb[k] = B[t * 64 + k * 8 + j];
          double f = c; for (int k = 0; k < 8; ++k) f = fma(a[k], bb[k], f);
          double r = c; for (int k = 7; k >= 0; --k) r = fma(a[k], bb[k], r);
          double q = c; for (int k = 4; k < 8; ++k) q = fma(a[k], bb[k], q); for (int k = 0; k < 4; ++k) q = fma(a[k], bb[k], q);
          const double d = D[t * 128 + i * 8 + j];
          ok_fwd += (d == f); ok_rev += (d == r); ok_48 += (d == q); ++tot;
        }
    printf("m16n8k8 : of %ld outputs  fma chain k=0..7: %ld   k=7..0: %ld   k=4..7 then 0..3: %ld\n", tot, ok_fwd, ok_rev, ok_48);
  }
  // ---------------- latency / chains needed
  run_lat<1>(1); run_lat<2>(1); run_lat<4>(1); run_lat<8>(1);
  run_lat<1>(4); run_lat<2>(4); run_lat<4>(4); run_lat<8>(4);
  run_lat<1>(8); run_lat<2>(8); run_lat<4>(8);
  // ---------------- throughput
  const int nsm = prop.multiProcessorCount;
  for (int w : {4, 8, 16, 32}) {
    run_thr<3, 8>("DFMA x4 (vector pipe)", 4.0 * 32, nsm, w);
    run_thr<0, 8>("DMMA m8n8k4", 256.0, nsm, w);
    run_thr<1, 8>("DMMA m16n8k4", 512.0, nsm, w);
    run_thr<2, 8>("DMMA m16n8k8", 1024.0, nsm, w);
  }
  run_thr<0, 2>("DMMA m8n8k4", 256.0, nsm, 16);
  run_thr<0, 4>("DMMA m8n8k4", 256.0, nsm, 16);
  run_thr<2, 2>("DMMA m16n8k8", 1024.0, nsm, 16);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
