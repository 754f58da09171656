// How long does the single-thread 8x8 signed Cholesky block take? (variants)
#include <cstdio>
#include <cuda_runtime.h>
#define NB 8
extern __shared__ double sm[];
template <int CHECKS>
__device__ __forceinline__ bool diag8(double* K, const int* eptr, const double* sgn, const double* diag0, double* invd, double* Ld, int kb) {
  double a[NB][NB], sg8[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int base = eptr[kb + r] + kb;
#pragma unroll
    for (int c = 0; c < NB; ++c) a[r][c] = (c <= r) ? K[base + c] : 0.0;
    sg8[r] = sgn[kb + r];
  }
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (!CHECKS || ok) {
      const double sj = sg8[j];
      const double d = sj * a[j][j];
      if (CHECKS) {
        const double thr = (sj > 0.0) ? 1e-12 * fmax(diag0[kb + j], 1e-300) : 0.0;
        if (!(d > thr) || !isfinite(d)) { ok = false; continue; }
      }
      const double inv = rsqrt(d);
      a[j][j] = d * inv; invd[kb + j] = inv;
      const double f = inv * sj;
#pragma unroll
      for (int r = j + 1; r < NB; ++r) a[r][j] *= f;
#pragma unroll
      for (int r = j + 1; r < NB; ++r) {
        const double lr = sj * a[r][j];
#pragma unroll
        for (int c = j + 1; c <= r; ++c) a[r][c] -= lr * a[c][j];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int base = eptr[kb + r] + kb;
#pragma unroll
    for (int c = 0; c < NB; ++c) if (c <= r) { K[base + c] = a[r][c]; Ld[r * NB + c] = a[r][c]; }
  }
  return ok;
}
template <int CHECKS>
__global__ void k(long long* cyc, double* out, int reps) {
  double* K = sm; int* eptr = (int*)(sm + 4096); double* sgn = sm + 4200; double* diag0 = sm + 4300; double* invd = sm + 4400; double* Ld = sm + 4500;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) K[i] = 0.0;
  __syncthreads();
  if (threadIdx.x < 64) { eptr[threadIdx.x] = threadIdx.x * 64; sgn[threadIdx.x] = 1.0; diag0[threadIdx.x] = 4.0; }
  __syncthreads();
  if (threadIdx.x < 64) for (int c = 0; c <= (int)threadIdx.x; ++c) K[threadIdx.x * 64 + c] = (c == (int)threadIdx.x) ? 4.0 + threadIdx.x : 0.01 * (c + 1);
  __syncthreads();
  long long tot = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int pb = 0; pb < 8; ++pb) {
      long long t0 = clock64();
      if (threadIdx.x == 0) diag8<CHECKS>(K, eptr, sgn, diag0, invd, Ld, pb * 8);
      __syncthreads();
      long long t1 = clock64();
      tot += t1 - t0;
      // refill block so that it stays SPD
      if (threadIdx.x < 8) for (int c = 0; c <= (int)threadIdx.x; ++c) K[(pb * 8 + threadIdx.x) * 64 + pb * 8 + c] = (c == (int)threadIdx.x) ? 4.0 + threadIdx.x : 0.01 * (c + 1);
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) { cyc[0] = tot; out[0] = K[5]; }
}
int main() {
  long long* cyc; double* out; cudaMalloc(&cyc, 64); cudaMalloc(&out, 64);
  long long c;
  const int reps = 50;
  for (int nt = 32; nt <= 512; nt *= 4) {
    k<1><<<1, nt, 40000>>>(cyc, out, reps); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("diag8 with checks, %3d threads: %.0f cycles per block\n", nt, (double)c / (reps * 8));
    k<0><<<1, nt, 40000>>>(cyc, out, reps); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("diag8 no checks,   %3d threads: %.0f cycles per block\n", nt, (double)c / (reps * 8));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
