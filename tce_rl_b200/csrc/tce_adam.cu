// Gradient norm + clipping + Adam for the policy network in two launches on ONE flat gradient buffer
// (row "next" of SURVEY 8(f): the optimiser step that closes every policy epoch).  Replaces, on the critical
// path of the epoch, torch's vector_norm reduction (6 us), the clip multiply and the multi-tensor fused Adam
// (11 us for ~150 k parameters in 8 tensors) -- temporal_correlated_agent.py:561-589 (clip_grad_norm_, step).
// Semantics: torch.optim.Adam (L2 weight decay, bias correction, no amsgrad) preceded by clip_grad_norm_.
#include <math.h>

#include "tce_common.cuh"

namespace {

constexpr int ADAM_MAX_TENSORS = 32;

struct ParamTable {                 // passed by value (baked into a captured graph with the launch)
  int count;
  long long offset[ADAM_MAX_TENSORS + 1];   // prefix sums of the tensor sizes = offsets into the flat buffers
  float *ptr[ADAM_MAX_TENSORS];
};

// state[0] += 1 (step counter, read by the Adam kernel that follows on the stream); state[1] += sum g^2
__global__ void __launch_bounds__(1024)
grad_sumsq_kernel(const float *__restrict__ g, long long n, double *__restrict__ state) {
  __shared__ double red[32];
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4 *>(g + i);
      s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (long long k = i; k < n; ++k) s = fma((double)g[k], (double)g[k], s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      atomicAdd(state + 1, s);
      if (blockIdx.x == 0) state[0] += 1.0;
    }
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(ParamTable tab, const float *__restrict__ grad, float *__restrict__ m, float *__restrict__ v,
            const double *__restrict__ state, double max_norm, double lr, double beta1, double beta2, double eps,
            double weight_decay) {
  const long long n = tab.offset[tab.count];
  const double t = state[0];
  if (!(state[1] < INFINITY)) return;        // NaN / Inf gradient: leave parameters and moments untouched
  double coef = 1.0;
  if (max_norm > 0.0) coef = fmin(1.0, max_norm / (sqrt(state[1]) + 1e-6));      // clip_grad_norm_
  const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
  const float step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float b1 = (float)beta1, b2 = (float)beta2, ep = (float)eps, wd = (float)weight_decay, cf = (float)coef;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= tab.offset[k + 1]) ++k;                       // <= 32 tensors: a short scan
    float *p = tab.ptr[k] + (i - tab.offset[k]);
    const float pv = *p;
    const float gi = fmaf(wd, pv, cf * grad[i]);               // L2 weight decay on the (clipped) gradient
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    *p = pv - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + ep));
  }
}

}  // namespace

extern "C" int tce_grad_sumsq(const float *grad_flat, int64_t n, double *state, void *stream) {
  if (n == 0) return TCE_OK;
  if (!grad_flat || !state || n < 0 || ((uintptr_t)grad_flat & 15)) return TCE_ERR_INVALID_ARGUMENT;
  long long blocks = (n + 4095) / 4096;
  if (blocks > 148) blocks = 148;
  grad_sumsq_kernel<<<(unsigned)blocks, 1024, 0, (cudaStream_t)stream>>>(grad_flat, n, state);
  TCE_CHECK_LAUNCH("grad_sumsq_kernel");
  return TCE_OK;
}

extern "C" int tce_adam_step(int count, float *const *params, const int64_t *sizes, const float *grad_flat, float *m,
                             float *v, const double *state, double max_norm, double lr, double beta1, double beta2,
                             double eps, double weight_decay, void *stream) {
  if (count == 0) return TCE_OK;
  if (count < 0 || count > ADAM_MAX_TENSORS || !params || !sizes || !grad_flat || !m || !v || !state)
    return TCE_ERR_INVALID_ARGUMENT;
  ParamTable tab;
  tab.count = count;
  tab.offset[0] = 0;
  for (int k = 0; k < count; ++k) {
    if (!params[k] || sizes[k] < 0) return TCE_ERR_INVALID_ARGUMENT;
    tab.ptr[k] = params[k];
    tab.offset[k + 1] = tab.offset[k] + sizes[k];
  }
  const long long n = tab.offset[count];
  if (n == 0) return TCE_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tab, grad_flat, m, v, state, max_norm, lr, beta1, beta2,
                                                                 eps, weight_decay);
  TCE_CHECK_LAUNCH("adam_kernel");
  return TCE_OK;
}
