// (4b) GAE / value-target reduction and TCE segment advantages.
// Replaces TemporalCorrelatedAgent.get_advantage_return (temporal_correlated_agent.py:118-181, a Python
// loop over T with ~8 tiny launches per step) and get_segment_advantage (:183-321).  HBM-bound
// (GAE: 18 T + 4 bytes / episode).  The recurrences are affine, so one warp handles an episode with a
// reversed shuffle scan over blocks of 32 time steps (coalesced loads and stores).
#include <math.h>

#include "tce_common.cuh"

namespace {

// x_t = a_t * x_{t+1} + b_t, scanned from t = T-1 down to 0
__global__ void __launch_bounds__(256)
gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values, const uint8_t *__restrict__ dones,
           const uint8_t *__restrict__ tl_dones, float gamma, float lam, int use_gae, float *__restrict__ adv,
           float *__restrict__ ret, long long B, int T) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float *r = rewards + b * T, *v = values + b * (T + 1);
  const uint8_t *dn = dones + b * T, *tl = tl_dones ? tl_dones + b * T : nullptr;
  float carry = use_gae ? 0.f : v[T];         // gae_{T} = 0 ; ret_{T} = V_T
  for (int base = ((T - 1) / 32) * 32; base >= 0; base -= 32) {
    const int t = base + lane;
    float a = 0.f, c = 0.f, vt = 0.f;
    if (t < T) {
      vt = v[t];
      const float disc = gamma * (dn[t] ? 0.f : 1.f);
      const float ntl = (tl && tl[t]) ? 0.f : 1.f;
      if (use_gae) {
        const float td = r[t] + disc * v[t + 1] - vt;
        a = disc * lam * ntl;
        c = td * ntl;
      } else {
        a = ntl * disc;
        c = ntl * r[t] + (1.f - ntl) * vt;
      }
    } else {
      a = 1.f;                                // identity map for the padding lanes of the last block
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float a2 = __shfl_down_sync(0xffffffffu, a, o), c2 = __shfl_down_sync(0xffffffffu, c, o);
      if (lane + o < 32) { c = fmaf(a, c2, c); a *= a2; }
    }
    const float x = fmaf(a, carry, c);        // gae_t (or ret_t)
    if (t < T) {
      const float rt = use_gae ? x + vt : x;
      ret[b * T + t] = rt;
      adv[b * T + t] = rt - vt;
    }
    carry = __shfl_sync(0xffffffffu, x, 0);
  }
}

// raw segment advantages + {count, sum, sum sq}
__global__ void __launch_bounds__(256)
segadv_kernel(int mode, const float *__restrict__ rewards, const float *__restrict__ values,
              const float *__restrict__ advantages, const int64_t *__restrict__ pairs, float gamma,
              float *__restrict__ seg, double *__restrict__ stats, long long B, int T, int P) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double x = 0.0;
  const bool active = g < B * P;
  if (active) {
    const long long b = g / P;
    const int p = (int)(g % P);
    const int s = (int)pairs[2 * p], e = (int)pairs[2 * p + 1];
    float out;
    if (mode == 0) {                          // sum of step advantages over [s, e]
      float acc = 0.f;
      for (int t = s; t <= e; ++t) acc += advantages[b * T + t];
      out = acc;
    } else {
      float acc = 0.f, w = 1.f;               // sum_{t in [s, e)} gamma^(t-s) r_t
      for (int t = s; t < e; ++t) { acc = fmaf(w, rewards[b * T + t], acc); w *= gamma; }
      out = (mode == 1) ? acc + w * values[b * (T + 1) + e] - values[b * (T + 1) + s] : acc;
    }
    seg[g] = out;
    x = (double)out;
  }
  if (stats) {
    double n = active ? 1.0 : 0.0, s1 = x, s2 = x * x;
    n = warp_sum(n); s1 = warp_sum(s1); s2 = warp_sum(s2);
    __shared__ double sh[8][3];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[w][0] = n; sh[w][1] = s1; sh[w][2] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b2 = 0, c = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[i][0]; b2 += sh[i][1]; c += sh[i][2]; }
      atomicAdd(stats, a); atomicAdd(stats + 1, b2); atomicAdd(stats + 2, c);
    }
  }
}

__global__ void __launch_bounds__(256)
sum_stats_kernel(const float *__restrict__ x, double *__restrict__ stats, long long N) {
  double n = 0, s1 = 0, s2 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    n += 1.0; s1 += v; s2 = fma(v, v, s2);
  }
  n = warp_sum(n); s1 = warp_sum(s1); s2 = warp_sum(s2);
  __shared__ double sh[8][3];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[w][0] = n; sh[w][1] = s1; sh[w][2] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[i][0]; b += sh[i][1]; c += sh[i][2]; }
    atomicAdd(stats, a); atomicAdd(stats + 1, b); atomicAdd(stats + 2, c);
  }
}

__global__ void __launch_bounds__(256)
normalize_kernel(float *__restrict__ x, const double *__restrict__ stats, long long N) {
  const double n = stats[0], mean = stats[1] / n;
  double var = (stats[2] - n * mean * mean) / (n - 1.0);     // torch.std: unbiased
  var = var > 0.0 ? var : 0.0;
  const float m = (float)mean, inv = (float)(1.0 / (sqrt(var) + 1e-8));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    x[i] = (x[i] - m) * inv;
}

}  // namespace

extern "C" int tce_gae(const float *rewards, const float *values, const uint8_t *dones, const uint8_t *tl_dones,
                       float gamma, float lam, int use_gae, float *adv, float *ret, int64_t B, int64_t T,
                       void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!rewards || !values || !dones || !adv || !ret || B < 0 || T < 1) return TCE_ERR_INVALID_ARGUMENT;
  const unsigned grid = (unsigned)((B * 32 + 255) / 256);
  gae_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, tl_dones, gamma, lam, use_gae, adv, ret, B, (int)T);
  TCE_CHECK_LAUNCH("gae_kernel");
  return TCE_OK;
}

extern "C" int tce_segment_advantage_raw(int mode, const float *rewards, const float *values, const float *advantages,
                                         const int64_t *pred_pairs, float gamma, float *seg, double *stats, int64_t B,
                                         int64_t T, int64_t P, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (mode < 0 || mode > 2 || !pred_pairs || !seg || B < 0 || T < 1 || P < 1) return TCE_ERR_INVALID_ARGUMENT;
  if (mode == 0 ? !advantages : (!rewards || !values)) return TCE_ERR_INVALID_ARGUMENT;
  const unsigned grid = (unsigned)((B * P + 255) / 256);
  segadv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mode, rewards, values, advantages, pred_pairs, gamma, seg, stats, B, (int)T, (int)P);
  TCE_CHECK_LAUNCH("segadv_kernel");
  return TCE_OK;
}

extern "C" int tce_sum_stats(const float *x, double *stats, int64_t N, void *stream) {
  if (!x || !stats || N < 0) return TCE_ERR_INVALID_ARGUMENT;
  if (N == 0) return TCE_OK;
  long long grid = (N + 255) / 256;
  if (grid > 1184) grid = 1184;
  sum_stats_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, stats, N);
  TCE_CHECK_LAUNCH("sum_stats_kernel");
  return TCE_OK;
}

extern "C" int tce_normalize_by_stats(float *x, const double *stats, int64_t N, void *stream) {
  if (!x || !stats || N < 0) return TCE_ERR_INVALID_ARGUMENT;
  if (N == 0) return TCE_OK;
  long long grid = (N + 255) / 256;
  if (grid > 1184) grid = 1184;
  normalize_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, stats, N);
  TCE_CHECK_LAUNCH("normalize_kernel");
  return TCE_OK;
}
