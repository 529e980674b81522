// Latency microbenchmarks (single warp): dependent chains of the ops the search kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dadd(double *o, double a) { double x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = x + a; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DADD dep latency   %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_dmul(double *o, double a) { double x = a; long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = x * a; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DMUL dep latency   %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_dadd_ilp4(double *o, double a) { double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { x0 += a; x1 += a; x2 += a; x3 += a; } long long t1 = clock64(); o[threadIdx.x] = x0 + x1 + x2 + x3; if (!threadIdx.x) printf("DADD x4 indep      %.1f cyc per 4 ops\n", double(t1 - t0) / N); }
__global__ void k_sqrt(double *o, double a) { double x = a; long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 512; ++i) x = sqrt(x + a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DSQRT(+add) dep    %.1f cyc\n", double(t1 - t0) / 512); }
__global__ void k_div(double *o, double a) { double x = a; long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 512; ++i) x = a / (x + a); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DDIV(+add) dep     %.1f cyc\n", double(t1 - t0) / 512); }
__global__ void k_shfl(double *o, double a) { double x = a + threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1); long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("SHFL f64 dep       %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_ballot(unsigned *o, unsigned a) { unsigned x = a + threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __ballot_sync(0xffffffffu, x & 1) + i; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("BALLOT dep         %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_lds(int *o) { __shared__ int s[1024]; for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i * 7 + 1) & 1023; __syncwarp(); int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = s[x]; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("LDS dep (ptr chase) %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_ldg(int *chain, int n, int *o) { int x = threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < 2048; ++i) x = chain[x]; long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("LDG dep n=%-9d %.1f cyc\n", n, double(t1 - t0) / 2048); }
__global__ void k_syncwarp(int *o) { int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { __syncwarp(); x += i; } long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("SYNCWARP+iadd      %.1f cyc\n", double(t1 - t0) / N); }
__global__ void k_stld(double *g, double *o) { double x = threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < 1024; ++i) { g[threadIdx.x] = x; __syncwarp(); x = g[threadIdx.x ^ 1] + 1.0; } long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("global st->ld (lane^1) round trip %.1f cyc\n", double(t1 - t0) / 1024); }
int main() {
  double *o; cudaMalloc(&o, 4096); unsigned *u = (unsigned *)o; int *oi = (int *)o;
  k_dadd<<<1, 32>>>(o, 1.0000001); k_dmul<<<1, 32>>>(o, 1.0000001); k_dadd_ilp4<<<1, 32>>>(o, 1.0000001);
  k_sqrt<<<1, 32>>>(o, 1.5); k_div<<<1, 32>>>(o, 1.5); k_shfl<<<1, 32>>>(o, 1.0); k_ballot<<<1, 32>>>(u, 3); k_lds<<<1, 32>>>(oi); k_syncwarp<<<1, 32>>>(oi);
  for (int n : {1 << 10, 1 << 16, 1 << 22, 1 << 26}) {   // 4 KB (L1), 256 KB (L2), 16 MB (L2), 256 MB (HBM)
    int *h = (int *)malloc(sizeof(int) * n); for (int i = 0; i < n; ++i) h[i] = (int)(((long long)i * 1000003 + 12345) % n);
    int *d; cudaMalloc(&d, sizeof(int) * n); cudaMemcpy(d, h, sizeof(int) * n, cudaMemcpyHostToDevice);
    k_ldg<<<1, 32>>>(d, n, oi); k_ldg<<<1, 32>>>(d, n, oi); cudaDeviceSynchronize(); cudaFree(d); free(h); }
  double *g; cudaMalloc(&g, 4096); k_stld<<<1, 32>>>(g, o);
  cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
