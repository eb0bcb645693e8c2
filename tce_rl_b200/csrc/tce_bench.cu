// Pipe-throughput micro-benchmarks used as roofline denominators for the compute-bound kernels
// (MEASURED_PEAKS.json carries no FP32 / FP64 FMA figure): register-resident FMA chains, 8 independent
// accumulators per thread, grid = 8 CTAs x 256 threads per SM.
#include "tce_common.cuh"

namespace {
template <typename T>
__global__ void __launch_bounds__(256) fma_kernel(T *out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
    x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
  }
  const T s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == (T)123456789) out[0] = s;      // keeps the chain alive without a store in the common case
}
}  // namespace

// launches one FMA-chain kernel; *flops receives the FLOPs it executes (2 per FMA)
extern "C" int tce_bench_fma(int fp64, int iters, void *scratch, double *flops, void *stream) {
  if (!scratch || iters < 1) return TCE_ERR_INVALID_ARGUMENT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = 8u * (unsigned)sms;
  if (fp64) fma_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double *)scratch, iters, 1.0000001, 1e-9);
  else fma_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float *)scratch, iters, 1.0000001f, 1e-9f);
  TCE_CHECK_LAUNCH("fma_kernel");
  if (flops) *flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
  return TCE_OK;
}
