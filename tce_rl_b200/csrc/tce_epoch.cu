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
constexpr int MC_E = 8;          // episodes per CTA iteration
constexpr int MC_Q = 4;          // split of the contraction index over thread groups (MC_THREADS = 64 * MC_Q)

__device__ __forceinline__ double mc_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage a dense row-major [n, n] fp64 matrix into shared memory with row stride LD (odd: rows of consecutive lanes
// fall into different banks), eight loads in flight per thread; the strict upper triangle is forced to zero
__device__ __forceinline__ void mc_stage_lower(double *sA, const double *__restrict__ A, int n, int LD) {
  const int total = n * n;
  for (int e0 = threadIdx.x; e0 < total; e0 += 8 * MC_THREADS) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * MC_THREADS;
      v[u] = e < total ? A[e] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * MC_THREADS;
      if (e < total) {
        const int i = e / n, j = e - i * n;
        sA[i * LD + j] = j <= i ? v[u] : 0.0;
      }
    }
  }
}

// The two triangular matrix-vector products of MC_E episodes at once, register blocked over the episodes: thread
// (r = t % 64, q = t / 64) owns row r (lower product) / column r (transposed product) and the contraction indices
// congruent to q mod MC_Q; one shared-memory load of the matrix element (conflict-free: consecutive lanes = consecutive
// rows at odd stride, or consecutive columns) feeds MC_E FMAs whose vector operands are 128-bit broadcast loads from
// the episode-minor layout vec[k][MC_E].  ~6 shared-memory wavefronts per 8 DFMA per warp (an output-per-thread loop
// needs ~28).  part: [MC_Q][64][MC_E] partial sums, reduced by the caller after a barrier.
__device__ __forceinline__ void mc_lower_partial(const double *sA, int LD, int n, const double *vec, double *part) {
  const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
  double acc[MC_E];
#pragma unroll
  for (int e = 0; e < MC_E; ++e) acc[e] = 0.0;
  if (r < n) {
    const double *row = sA + r * LD;
    for (int j = q; j <= r; j += MC_Q) {
      const double a = row[j];
      const double2 *v2 = reinterpret_cast<const double2 *>(vec + j * MC_E);
#pragma unroll
      for (int e = 0; e < MC_E; e += 2) {
        const double2 d = v2[e >> 1];
        acc[e] = fma(a, d.x, acc[e]);
        acc[e + 1] = fma(a, d.y, acc[e + 1]);
      }
    }
  }
  double2 *p2 = reinterpret_cast<double2 *>(part + (q * 64 + r) * MC_E);
#pragma unroll
  for (int e = 0; e < MC_E; e += 2) p2[e >> 1] = make_double2(acc[e], acc[e + 1]);
}

__device__ __forceinline__ void mc_upper_partial(const double *sA, int LD, int n, const double *vec, double *part) {
  const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
  double acc[MC_E];
#pragma unroll
  for (int e = 0; e < MC_E; ++e) acc[e] = 0.0;
  if (c < n) {
    const double *col = sA + c;
    for (int i = c + q; i < n; i += MC_Q) {
      const double a = col[i * LD];
      const double2 *v2 = reinterpret_cast<const double2 *>(vec + i * MC_E);
#pragma unroll
      for (int e = 0; e < MC_E; e += 2) {
        const double2 d = v2[e >> 1];
        acc[e] = fma(a, d.x, acc[e]);
        acc[e + 1] = fma(a, d.y, acc[e + 1]);
      }
    }
  }
  double2 *p2 = reinterpret_cast<double2 *>(part + (q * 64 + c) * MC_E);
#pragma unroll
  for (int e = 0; e < MC_E; e += 2) p2[e >> 1] = make_double2(acc[e], acc[e + 1]);
}

// out[k][e] = sum_q part[q][k][e]  for k < n (episode-minor layout)
__device__ __forceinline__ void mc_reduce_partials(const double *part, double *out, int n) {
  for (int o = threadIdx.x; o < n * MC_E; o += MC_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < MC_Q; ++q) s += part[q * 64 * MC_E + o];
    out[o] = s;
  }
}

struct McSmem {
  double *sA, *d, *z, *part, *sc;
};
__device__ __forceinline__ McSmem mc_carve(double *sm, int n) {
  McSmem s;
  const int LD = n | 1;
  s.sA = sm;
  s.d = s.sA + ((n * LD + 1) & ~1);          // 16-byte aligned vectors
  s.z = s.d + 64 * MC_E;
  s.part = s.z + 64 * MC_E;
  s.sc = s.part + MC_Q * 64 * MC_E;
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MC_THREADS)
epoch_mean_fwd_kernel(const float *__restrict__ mean, const float *__restrict__ mean_old,
                      const double *__restrict__ Linv_old, double eps_mean, float *__restrict__ proj_mean,
                      double *__restrict__ maha_old, float *__restrict__ u_old, double *__restrict__ acc, long long B,
                      int n) {
  extern __shared__ __align__(16) double sm[];
  const int LD = n | 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const McSmem S = mc_carve(sm, n);
  mc_stage_lower(S.sA, Linv_old, n, LD);
  double part0 = 0.0, part1 = 0.0;                       // thread 0: sums over this CTA's episodes
  const long long tiles = (B + MC_E - 1) / MC_E;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b0 = tile * MC_E;
    const int ne = (int)((B - b0) < MC_E ? (B - b0) : MC_E);
    __syncthreads();                                     // sA staged / previous iteration done with the vectors
    for (int o = threadIdx.x; o < MC_E * n; o += MC_THREADS) {      // d[k][e] (episode-minor), global reads coalesced
      const int e = o / n, k = o - e * n;
      S.d[k * MC_E + e] = e < ne ? (double)mean[(b0 + e) * n + k] - (double)mean_old[(b0 + e) * n + k] : 0.0;
    }
    __syncthreads();
    mc_lower_partial(S.sA, LD, n, S.d, S.part);
    __syncthreads();
    mc_reduce_partials(S.part, S.z, n);                  // z = L_old^-1 d
    __syncthreads();
    if (warp < MC_E) {
      double m = 0.0;
      for (int i = lane; i < n; i += 32) m = fma(S.z[i * MC_E + warp], S.z[i * MC_E + warp], m);
      m = mc_warp_sum(m);
      if (lane == 0) {
        S.sc[warp] = m;
        if (warp < ne) maha_old[b0 + warp] = m;
      }
    }
    mc_upper_partial(S.sA, LD, n, S.z, S.part);
    __syncthreads();
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS) {
      const int e = o / n, k = o - e * n;
      double u = 0.0;
#pragma unroll
      for (int q = 0; q < MC_Q; ++q) u += S.part[(q * 64 + k) * MC_E + e];
      u_old[b0 * n + o] = (float)u;
      const double mp = 0.5 * S.sc[e];
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
        const double mp = 0.5 * S.sc[e];
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
// backward, part 1 (needs only the projection's state: runs beside the segment likelihood): gradient of the trust-region
// regression term  tr_coeff / B * 1/2 maha(mean, proj_mean DETACHED; Sigma_out)  w.r.t. mean, and its value
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MC_THREADS)
epoch_tr_mean_kernel(const float *__restrict__ mean, const float *__restrict__ mean_old,
                     const double *__restrict__ maha_old, const float *__restrict__ u_old,
                     const double *__restrict__ Linv_new, const double *__restrict__ kl_sc, double eps_mean,
                     double tr_coeff, float *__restrict__ tr_grad, double *__restrict__ acc, long long B, int n) {
  extern __shared__ __align__(16) double sm[];
  const int LD = n | 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const McSmem S = mc_carve(sm, n);
  mc_stage_lower(S.sA, Linv_new, n, LD);
  const double eta = kl_sc[1] != 0.0 ? kl_sc[0] : 0.0, alpha2 = kl_sc[6];
  const double prec_scale = 1.0 / (alpha2 * (1.0 + eta));           // Sigma_out^-1 = prec_scale (Sigma~^-1 + eta Sigma_old^-1)
  double part2 = 0.0;
  const long long tiles = (B + MC_E - 1) / MC_E;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b0 = tile * MC_E;
    const int ne = (int)((B - b0) < MC_E ? (B - b0) : MC_E);
    __syncthreads();
    for (int o = threadIdx.x; o < MC_E * n; o += MC_THREADS) {
      const int e = o / n, k = o - e * n;
      S.d[k * MC_E + e] = e < ne ? (double)mean[(b0 + e) * n + k] - (double)mean_old[(b0 + e) * n + k] : 0.0;
    }
    __syncthreads();
    mc_lower_partial(S.sA, LD, n, S.d, S.part);
    __syncthreads();
    mc_reduce_partials(S.part, S.z, n);                  // z~ = L~^-1 d
    __syncthreads();
    if (warp < MC_E) {
      double m = 0.0;
      for (int i = lane; i < n; i += 32) m = fma(S.z[i * MC_E + warp], S.z[i * MC_E + warp], m);
      m = mc_warp_sum(m);
      if (lane == 0) S.sc[warp] = m;
    }
    mc_upper_partial(S.sA, LD, n, S.z, S.part);
    __syncthreads();
    for (int o = threadIdx.x; o < ne * n; o += MC_THREADS) {
      const int e = o / n, k = o - e * n;
      double ut = 0.0;                                   // (Sigma~^-1 d)_k
#pragma unroll
      for (int q = 0; q < MC_Q; ++q) ut += S.part[(q * 64 + k) * MC_E + e];
      const double mp = 0.5 * maha_old[b0 + e];
      double s = 1.0;
      if (mp > eps_mean) s = 1.0 / (1.0 + (sqrt(mp / eps_mean) - 1.0) + 1e-16);
      // mean - proj_mean = (1 - s) d
      tr_grad[b0 * n + o] = (float)(tr_coeff / (double)B * (1.0 - s) * prec_scale * (ut + eta * (double)u_old[b0 * n + o]));
    }
    if (threadIdx.x == 0) {
      for (int e = 0; e < ne; ++e) {
        const double mo = maha_old[b0 + e], mp = 0.5 * mo;
        double s = 1.0;
        if (mp > eps_mean) s = 1.0 / (1.0 + (sqrt(mp / eps_mean) - 1.0) + 1e-16);
        part2 += 0.5 * (1.0 - s) * (1.0 - s) * prec_scale * (S.sc[e] + eta * mo);   // 1/2 maha(mean, proj_mean; Sigma_out)
      }
    }
  }
  if (threadIdx.x == 0 && acc) atomicAdd(acc + 2, part2);
}

// backward, part 2 (after the likelihood): adjoint of the mean projection applied to g = d loss / d proj_mean, plus the
// trust-region gradient of part 1.  One warp per episode.
__global__ void __launch_bounds__(256)
epoch_mean_combine_kernel(const float *__restrict__ g_pm, const float *__restrict__ mean, const float *__restrict__ mean_old,
                          const double *__restrict__ maha_old, const float *__restrict__ u_old,
                          const float *__restrict__ tr_grad, double eps_mean, float *__restrict__ grad_mean, long long B,
                          int n) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const double mp = 0.5 * maha_old[b];
  const bool active = mp > eps_mean;
  const float *g = g_pm + b * n, *tg = tr_grad ? tr_grad + b * n : nullptr;
  if (!active) {
    for (int i = lane; i < n; i += 32) grad_mean[b * n + i] = g[i] + (tg ? tg[i] : 0.f);
    return;
  }
  const double om = sqrt(mp / eps_mean) - 1.0, a = 1.0 / (1.0 + om + 1e-16);
  double dot = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double x = mean[b * n + i], xo = mean_old[b * n + i];
    dot = fma((double)g[i], a * (xo - (x + om * xo) * a), dot);
  }
  dot = mc_warp_sum(dot);
  const double c = dot / (2.0 * sqrt(mp * eps_mean));                // d loss / d mean_part; d mean_part / d mean = u_old
  for (int i = lane; i < n; i += 32)
    grad_mean[b * n + i] = (float)((double)g[i] * a + c * (double)u_old[b * n + i] + (tg ? (double)tg[i] : 0.0));
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

size_t mc_smem(int n) {
  return sizeof(double) * ((((size_t)n * (n | 1) + 1) & ~(size_t)1) + 2 * 64 * MC_E + MC_Q * 64 * MC_E + 2 * MC_E);
}

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
  if (!mean || !mean_old || !Linv_old || !proj_mean || !maha_old || !u_old || B < 0 || n < 1 || n > 64 ||
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

extern "C" int tce_epoch_tr_mean(const float *mean, const float *mean_old, const double *maha_old, const float *u_old,
                                 const double *Linv_new, const double *kl_scalars, double eps_mean, double tr_coeff,
                                 float *tr_grad, double *acc, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;
  if (!mean || !mean_old || !maha_old || !u_old || !Linv_new || !kl_scalars || !tr_grad || B < 0 || n < 1 || n > 64 ||
      !(eps_mean > 0.0))
    return TCE_ERR_INVALID_ARGUMENT;
  const size_t smem = mc_smem(n);
  int rc = mc_set_smem(epoch_tr_mean_kernel, smem);
  if (rc) return rc;
  epoch_tr_mean_kernel<<<mc_grid(B), MC_THREADS, smem, (cudaStream_t)stream>>>(
      mean, mean_old, maha_old, u_old, Linv_new, kl_scalars, eps_mean, tr_coeff, tr_grad, acc, B, n);
  TCE_CHECK_LAUNCH("epoch_tr_mean_kernel");
  return TCE_OK;
}

extern "C" int tce_epoch_mean_combine(const float *g_proj_mean, const float *mean, const float *mean_old,
                                      const double *maha_old, const float *u_old, const float *tr_grad,
                                      double eps_mean, float *grad_mean, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;
  if (!g_proj_mean || !mean || !mean_old || !maha_old || !u_old || !grad_mean || B < 0 || n < 1 || !(eps_mean > 0.0))
    return TCE_ERR_INVALID_ARGUMENT;
  epoch_mean_combine_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      g_proj_mean, mean, mean_old, maha_old, u_old, tr_grad, eps_mean, grad_mean, B, n);
  TCE_CHECK_LAUNCH("epoch_mean_combine_kernel");
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
