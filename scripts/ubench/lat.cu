// Dependent-chain latencies on one SM (cycles per op), for sizing the latency-bound single-CTA kernels.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double *out, long long *cyc, double a, double b, float fa, int one) {
  __shared__ double sm[64];
  const int t = threadIdx.x;
  double x = a + t; float f = fa + t; long long t0, t1; int idx = 0;
  if (t < 64) sm[t] = (double)((t + one) & 63);
  __syncthreads();
#define TIME(slot, body) __syncthreads(); t0 = clock64(); _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; } t1 = clock64(); if (t == 0) cyc[slot] = t1 - t0;
  TIME(0, x = fma(x, b, a))
  TIME(1, x = x + a)
  TIME(2, x = x * b)
  TIME(3, f = fmaf(f, fa, fa))
  TIME(4, f = rsqrtf(f) + fa)
  TIME(5, f = (float)x; x = (double)f + a)
  TIME(6, f = __shfl_xor_sync(0xffffffffu, f, 1))
  TIME(7, x = __shfl_xor_sync(0xffffffffu, x, 1))
  TIME(8, idx = (int)sm[idx & 63])
  TIME(9, __syncthreads())
  if (blockDim.x == 256) { TIME(10, asm volatile("bar.sync 1, 256;" ::: "memory")) }
  TIME(11, f = __fdividef(fa, f) + fa)
  // 4 independent DFMA chains per thread: issue-rate view
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
  TIME(12, y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a))
  out[t] = x + f + idx + y0 + y1 + y2 + y3;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  const char *names[] = {"DFMA dep", "DADD dep", "DMUL dep", "FFMA dep", "MUFU.RSQ+FADD dep", "F2F 64->32->64 + DADD", "SHFL32 dep",
                         "SHFL64 dep", "LDS.64+cvt dep", "__syncthreads", "bar.sync 1,256", "fdividef+FADD dep", "4 indep DFMA (per 4)"};
  for (int threads : {32, 256, 1024}) {
    k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, 1.0001f, 1);
    k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, 1.0001f, 1);
    long long h[16]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("threads=%d (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 13; ++i) if (i != 10 || threads == 256) printf("  %-26s %7.1f cycles/op\n", names[i], (double)h[i] / N);
  }
  return 0;
}
