// Per-episode MEAN chain of one TCE policy epoch with ONE shared covariance (non-contextual policy: every shipped
// config) -- the batch-sized pieces that sit between the policy network and the segment likelihood, forward and
// backward, as two kernels instead of ~45 small launches (Mahalanobis kernels, ATen glue of the loss arithmetic,
// autograd bookkeeping):
//   forward : d = mean - mean_old, z = L_old^-1 d, maha_old = |z|^2, u_old = L_old^-T z,
//             mean projection (trust-region-layers mean_projection, reached from
//             mprl/rl/agent/temporal_correlated_agent.py:530-533):  proj = (mean + w mean_old) / (1 + w + 1e-16),
//             w = sqrt(maha_old / (2 eps)) - 1 where 1/2 maha_old > eps
//   backward: given g = d loss / d proj_mean (from the segment likelihood)
//             grad_mean = (mean projection)^T g + gradient of the trust-region regression loss
//             (get_trust_region_loss, temporal_correlated_agent.py:561-567):
//             tr_coeff / B * 1/2 maha(mean, proj_mean DETACHED; Sigma_out), whose precision is known in closed form
//             from the KL projection:  Sigma_out^-1 = (Sigma~^-1 + eta Sigma_old^-1) / (alpha^2 (1 + eta))
//             (identity step: eta = 0) -- no factor of Sigma_out is needed; mean - proj_mean = (1 - s) d.
// Both kernels also accumulate the batch sums that the logging decomposition of
// temporal_correlated_agent.py:641-686 needs (mean parts of KL(new||old), KL(new||proj), KL(proj||old)).
// tce_epoch_metrics assembles the metrics vector of the epoch (7 loss values + 12 KL values) in one tiny launch.
#include <math.h>

#include "tce_common.cuh"

namespace {

constexpr int MC_THREADS = 256;
constexpr int MC_E = 8;          // episodes per CTA iteration (one warp finishes one episode's reductions)

__device__ __forceinline__ double mc_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage a dense row-major [n, n] fp64 matrix into shared memory with row stride LD (odd: rows of consecutive lanes
// fall into different banks), four loads in flight per thread; the strict upper triangle is forced to zero
__device__ __forceinline__ void mc_stage_lower(double *sA, const double *__restrict__ A, int n, int LD) {
  const int total = n * n;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * MC_THREADS) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * MC_THREADS;
      v[u] = e < total ? A[e] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * MC_THREADS;
      if (e < total) {
        const int i = e / n, j = e - i * n;
        sA[i * LD + j] = j <= i ? v[u] : 0.0;
      }
    }
  }
}

// z[e][i] = sum_{j <= i} A[i][j] d[e][j]   (ne episodes, vectors [ne][n] in shared memory)
__device__ __forceinline__ void mc_lower_matvec(const double *sA, int LD, int n, const double *sd, double *sz, int ne) {
  for (int o = threadIdx.x; o < ne * n; o += MC_THREADS) {
    const int e = o / n, i = o - e * n;
    const double *row = sA + i * LD, *d = sd + e * n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 3 <= i; j += 4) {
      a0 = fma(row[j], d[j], a0); a1 = fma(row[j + 1], d[j + 1], a1);
      a2 = fma(row[j + 2], d[j + 2], a2); a3 = fma(row[j + 3], d[j + 3], a3);
    }
    for (; j <= i; ++j) a0 = fma(row[j], d[j], a0);
    sz[o] = (a0 + a1) + (a2 + a3);
  }
}

// u[e][j] = sum_{i >= j} A[i][j] z[e][i]
__device__ __forceinline__ double mc_upper_dot(const double *sA, int LD, int n, const double *z, int j) {
  const double *col = sA + j;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = j;
  for (; i + 3 < n; i += 4) {
    a0 = fma(col[i * LD], z[i], a0); a1 = fma(col[(i + 1) * LD], z[i + 1], a1);
    a2 = fma(col[(i + 2) * LD], z[i + 2], a2); a3 = fma(col[(i + 3) * LD], z[i + 3], a3);
  }
  for (; i < n; ++i) a0 = fma(col[i * LD], z[i], a0);
  return (a0 + a1) + (a2 + a3);
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MC_THREADS)
epoch_mean_fwd_kernel(const float *__restrict__ mean, const float *__restrict__ mean_old,
                      const double *__restrict__ Linv_old, double eps_mean, float *__restrict__ proj_mean,
                      double *__restrict__ maha_old, float *__restrict__ u_old, double *__restrict__ acc, long long B,
                      int n) {
  extern __shared__ double sm[];
  const int LD = n | 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *sA = sm, *sd = sA + n * LD, *sz = sd + MC_E * n, *sc = sz + MC_E * n;      // sc: [MC_E] maha
  mc_stage_lower(sA, Linv_old, n, LD);
  double part0 = 0.0, part1 = 0.0;                       // thread 0: sums over this CTA's episodes
  const long long tiles = (B + MC_E - 1) / MC_E;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b0 = tile * MC_E;
    const int ne = (int)((B - b0) < MC_E ? (B - b0) : MC_E);
    __syncthreads();                                     // sA staged / previous iteration done with sd, sz, sc
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS)
      sd[o] = (double)mean[b0 * n + o] - (double)mean_old[b0 * n + o];
    __syncthreads();
    mc_lower_matvec(sA, LD, n, sd, sz, ne);
    __syncthreads();
    if (warp < ne) {
      double m = 0.0;
      for (int i = lane; i < n; i += 32) m = fma(sz[warp * n + i], sz[warp * n + i], m);
      m = mc_warp_sum(m);
      if (lane == 0) { sc[warp] = m; maha_old[b0 + warp] = m; }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS) {
      const int e = o / n, j = o - e * n;
      const double u = mc_upper_dot(sA, LD, n, sz + e * n, j);
      u_old[b0 * n + o] = (float)u;
      const double mp = 0.5 * sc[e];
      const double x = (double)mean[b0 * n + o];
      float out = (float)x;
      if (mp > eps_mean) {
        const double om = fabs(sqrt(mp / eps_mean) - 1.0), a = 1.0 / (1.0 + om + 1e-16);
        out = (float)((x + om * (double)mean_old[b0 * n + o]) * a);
      }
      proj_mean[b0 * n + o] = out;
    }
    if (threadIdx.x == 0) {
      for (int e = 0; e < ne; ++e) {
        const double mp = 0.5 * sc[e];
        double s = 1.0;
        if (mp > eps_mean) s = 1.0 / (1.0 + fabs(sqrt(mp / eps_mean) - 1.0) + 1e-16);
        part0 += mp;
        part1 += mp * s * s;                             // proj - mean_old = s d
      }
    }
  }
  if (threadIdx.x == 0 && acc) {
    atomicAdd(acc + 0, part0);
    atomicAdd(acc + 1, part1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MC_THREADS)
epoch_mean_bwd_kernel(const float *__restrict__ g_pm, const float *__restrict__ mean, const float *__restrict__ mean_old,
                      const double *__restrict__ maha_old, const float *__restrict__ u_old,
                      const double *__restrict__ Linv_new, const double *__restrict__ kl_sc, double eps_mean,
                      double tr_coeff, float *__restrict__ grad_mean, double *__restrict__ acc, long long B, int n) {
  extern __shared__ double sm[];
  const int LD = n | 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *sA = sm, *sd = sA + n * LD, *sz = sd + MC_E * n, *sc = sz + MC_E * n;      // sc: [2][MC_E]: |z~|^2, g.(xo - proj)
  mc_stage_lower(sA, Linv_new, n, LD);
  const double eta = kl_sc[1] != 0.0 ? kl_sc[0] : 0.0, alpha2 = kl_sc[6];
  const double prec_scale = 1.0 / (alpha2 * (1.0 + eta));           // Sigma_out^-1 = prec_scale (Sigma~^-1 + eta Sigma_old^-1)
  double part2 = 0.0;
  const long long tiles = (B + MC_E - 1) / MC_E;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b0 = tile * MC_E;
    const int ne = (int)((B - b0) < MC_E ? (B - b0) : MC_E);
    __syncthreads();
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS)
      sd[o] = (double)mean[b0 * n + o] - (double)mean_old[b0 * n + o];
    __syncthreads();
    mc_lower_matvec(sA, LD, n, sd, sz, ne);                       // z~ = L~^-1 d
    __syncthreads();
    if (warp < ne) {
      const long long b = b0 + warp;
      const double mp = 0.5 * maha_old[b];
      const bool active = mp > eps_mean;
      const double om = active ? sqrt(mp / eps_mean) - 1.0 : 0.0, a = 1.0 / (1.0 + om + 1e-16);
      double m = 0.0, dot = 0.0;
      for (int i = lane; i < n; i += 32) {
        m = fma(sz[warp * n + i], sz[warp * n + i], m);
        if (active) {
          const double x = mean[b * n + i], xo = mean_old[b * n + i];
          dot = fma((double)g_pm[b * n + i], a * (xo - (x + om * xo) * a), dot);
        }
      }
      m = mc_warp_sum(m);
      dot = mc_warp_sum(dot);
      if (lane == 0) { sc[warp] = m; sc[MC_E + warp] = dot; }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS) {
      const int e = o / n, j = o - e * n;
      const long long b = b0 + e;
      const double ut = mc_upper_dot(sA, LD, n, sz + e * n, j);      // (Sigma~^-1 d)_j
      const double mp = 0.5 * maha_old[b];
      const bool active = mp > eps_mean;
      const double uo = (double)u_old[b0 * n + o], g = (double)g_pm[b0 * n + o];
      double gm = g, s = 1.0;
      if (active) {
        const double om = sqrt(mp / eps_mean) - 1.0;
        s = 1.0 / (1.0 + om + 1e-16);
        gm = g * s + sc[MC_E + e] / (2.0 * sqrt(mp * eps_mean)) * uo;        // d mean_part / d mean = u_old
      }
      // trust-region regression: tr_coeff / B * Sigma_out^-1 (mean - proj_mean), mean - proj_mean = (1 - s) d
      gm += tr_coeff / (double)B * (1.0 - s) * prec_scale * (ut + eta * uo);
      grad_mean[b0 * n + o] = (float)gm;
    }
    if (threadIdx.x == 0) {
      for (int e = 0; e < ne; ++e) {
        const double mo = maha_old[b0 + e], mp = 0.5 * mo;
        double s = 1.0;
        if (mp > eps_mean) s = 1.0 / (1.0 + (sqrt(mp / eps_mean) - 1.0) + 1e-16);
        part2 += 0.5 * (1.0 - s) * (1.0 - s) * prec_scale * (sc[e] + eta * mo);   // 1/2 maha(mean, proj_mean; Sigma_out)
      }
    }
  }
  if (threadIdx.x == 0 && acc) atomicAdd(acc + 2, part2);
}

// metrics [19] = {surrogate, entropy_loss, trust_region_loss, policy_loss, entropy, imp_smp_ratio, policy_grad_norm,
//                 new_old {mean, cov, shape, volume}, new_proj {...}, proj_old {...}}   (rl/agent.py _LOSS_KEYS, _KL_KEYS)
__global__ void epoch_metrics_kernel(const double *__restrict__ acc, const double *__restrict__ lik_stats,
                                     const double *__restrict__ kl_sc, const double *__restrict__ adam_stats, double B,
                                     double tr_coeff, int with_cov, double ent_coef, double *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double sur = lik_stats[0], ratio = lik_stats[1];
  const double tr_mean = acc[2] / B, tr_shape = kl_sc[7], tr_vol = kl_sc[8];
  const double tr_loss = tr_coeff * (tr_mean + (with_cov ? tr_shape + tr_vol : 0.0));
  const double entropy = kl_sc[9], ent_loss = -ent_coef * entropy;
  out[0] = sur; out[1] = ent_loss; out[2] = tr_loss; out[3] = sur + ent_loss + tr_loss; out[4] = entropy;
  out[5] = ratio; out[6] = adam_stats ? sqrt(adam_stats[1]) : 0.0;
  out[7] = acc[0] / B; out[8] = kl_sc[10] + kl_sc[11]; out[9] = kl_sc[10]; out[10] = kl_sc[11];
  out[11] = tr_mean; out[12] = tr_shape + tr_vol; out[13] = tr_shape; out[14] = tr_vol;
  out[15] = acc[1] / B; out[16] = kl_sc[12] + kl_sc[13]; out[17] = kl_sc[12]; out[18] = kl_sc[13];
}

size_t mc_smem(int n) { return sizeof(double) * ((size_t)n * (n | 1) + 2 * (size_t)MC_E * n + 2 * MC_E); }

template <typename K>
int mc_set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { tce_set_cuda_error(e, "epoch mean smem attr"); return TCE_ERR_CUDA; }
  }
  return TCE_OK;
}

unsigned mc_grid(int64_t B) {
  const long long tiles = (B + MC_E - 1) / MC_E;
  return (unsigned)(tiles < 148 * 2 ? tiles : 148 * 2);
}

}  // namespace

extern "C" int tce_epoch_mean_fwd(const float *mean, const float *mean_old, const double *Linv_old, double eps_mean,
                                  float *proj_mean, double *maha_old, float *u_old, double *acc, int64_t B, int n,
                                  void *stream) {
  if (B == 0) return TCE_OK;
  if (!mean || !mean_old || !Linv_old || !proj_mean || !maha_old || !u_old || B < 0 || n < 1 || n > 128 ||
      !(eps_mean > 0.0))
    return TCE_ERR_INVALID_ARGUMENT;
  const size_t smem = mc_smem(n);
  int rc = mc_set_smem(epoch_mean_fwd_kernel, smem);
  if (rc) return rc;
  epoch_mean_fwd_kernel<<<mc_grid(B), MC_THREADS, smem, (cudaStream_t)stream>>>(mean, mean_old, Linv_old, eps_mean,
                                                                               proj_mean, maha_old, u_old, acc, B, n);
  TCE_CHECK_LAUNCH("epoch_mean_fwd_kernel");
  return TCE_OK;
}

extern "C" int tce_epoch_mean_bwd(const float *g_proj_mean, const float *mean, const float *mean_old,
                                  const double *maha_old, const float *u_old, const double *Linv_new,
                                  const double *kl_scalars, double eps_mean, double tr_coeff, float *grad_mean,
                                  double *acc, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;
  if (!g_proj_mean || !mean || !mean_old || !maha_old || !u_old || !Linv_new || !kl_scalars || !grad_mean || B < 0 ||
      n < 1 || n > 128 || !(eps_mean > 0.0))
    return TCE_ERR_INVALID_ARGUMENT;
  const size_t smem = mc_smem(n);
  int rc = mc_set_smem(epoch_mean_bwd_kernel, smem);
  if (rc) return rc;
  epoch_mean_bwd_kernel<<<mc_grid(B), MC_THREADS, smem, (cudaStream_t)stream>>>(
      g_proj_mean, mean, mean_old, maha_old, u_old, Linv_new, kl_scalars, eps_mean, tr_coeff, grad_mean, acc, B, n);
  TCE_CHECK_LAUNCH("epoch_mean_bwd_kernel");
  return TCE_OK;
}

extern "C" int tce_epoch_metrics(const double *acc, const double *lik_stats, const double *kl_scalars,
                                 const double *adam_stats, int64_t B, double tr_coeff, int with_cov, double ent_coef,
                                 double *out19, void *stream) {
  if (!acc || !lik_stats || !kl_scalars || !out19 || B < 1) return TCE_ERR_INVALID_ARGUMENT;
  epoch_metrics_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, lik_stats, kl_scalars, adam_stats, (double)B, tr_coeff,
                                                          with_cov, ent_coef, out19);
  TCE_CHECK_LAUNCH("epoch_metrics_kernel");
  return TCE_OK;
}
