// (3) TCE segment-wise trajectory likelihood, forward and backward.
// Replaces TemporalCorrelatedPolicy.log_prob (mprl/rl/policy/temporal_correlated_policy.py:104-203):
//   mp.update_inputs(times=[B,P,2], params=[B,P,Dp], params_L=[B,P,Dp,Dp]) ; get_traj_pos(flat) ;
//   get_traj_pos_cov() ; MultivariateNormal(covariance_matrix).log_prob
// without ever expanding L over the P segments (the reference recomputes L L^T per segment, :158).
//
// Math, per episode b and pair p = (t0, t1), n = 2D, rows (d, k):
//   h_pk   = scaled position basis of the ProDMP with initial conditions at time t_k      [K1]
//   Sigma  = L L^T                                                                        [Dp, Dp]
//   C[(d,k),(d',k')] = h_pk^T Sigma_{dd'} h_pk'   (+ reg on the diagonal),  Sigma_{dd'} = K1 x K1 block
//   mu[(d,k)] = xi1 y0_d + xi2 tau v0_d + h_pk . theta_d ;  r = x - mu
//   lp = -1/2 (n ln 2pi + r^T C^-1 r) - 1/2 logdet C
//   dlp/dmu = alpha = C^-1 r ; G = dlp/dC = 1/2 (alpha alpha^T - C^-1)
//   dlp/dSigma_{dd'} = sum_{p,k,k'} G[(d,k),(d',k')] h_pk h_pk'^T ;  dlp/dL = 2 (dlp/dSigma) L
//
// Precision: the quadratic forms, the residual and the per-segment Cholesky are ill conditioned
// (cond(C) ~ 1e4 because neighbouring time points are almost perfectly correlated); fp32
// accumulation there costs ~3e-4 absolute on the log-prob, above the 1e-4 parity bound.  They run in
// fp64 (B200 DFMA = 1/2 FFMA rate); the two O(Dp^3) products Sigma = L L^T and (dSigma) L run in fp32.
#include <math.h>

#include "tce_common.cuh"

// profiling scaffolding (-DTCE_PROFILE only): SM-clock stamps of block 0 of the last gram [0..15] / bwd [16..31] kernel
#ifdef TCE_PROFILE
__device__ long long g_sl_prof[32];
#define SL_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_sl_prof[i] = clock64(); } while (0)
#else
#define SL_STAMP(i) do { } while (0)
#endif

namespace {

constexpr int SL_THREADS = 256;

__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__device__ __forceinline__ int tri_idx(int r, int c) { return r * (r + 1) / 2 + c; }  // r >= c

// ---- basis rows of the 2P time points of one episode (fp64) -----------------------------------------
// hs [2P][K1]  : scaled H_pos row of time point q = 2p + k
// xi [2P][2]   : xi1, xi2
template <int K1>
__device__ void basis_points(const TabDev &tb, double t_init, const float *__restrict__ times_b,
                             const int64_t *__restrict__ pairs, int P, double *hs, double *xi, double *init_row) {
  // init_row: [0..3] y1b,y2b,dy1b,dy2b ; [4] 1/det ; [5..5+K1) pos_b ; [5+K1 .. 5+2K1) vel_b
  if (threadIdx.x <= K1) {
    int i0; double w;
    time_to_index(tb, t_init, i0, w);
    if (threadIdx.x < K1) {
      const int j = threadIdx.x;
      init_row[5 + j] = lerp_t(tb.pos[(size_t)i0 * K1 + j], tb.pos[(size_t)(i0 + 1) * K1 + j], w);
      init_row[5 + K1 + j] = lerp_t(tb.vel[(size_t)i0 * K1 + j], tb.vel[(size_t)(i0 + 1) * K1 + j], w);
    } else {
      const double a = lerp_t(tb.y1[i0], tb.y1[i0 + 1], w), b = lerp_t(tb.y2[i0], tb.y2[i0 + 1], w);
      const double c = lerp_t(tb.dy1[i0], tb.dy1[i0 + 1], w), d = lerp_t(tb.dy2[i0], tb.dy2[i0 + 1], w);
      init_row[0] = a; init_row[1] = b; init_row[2] = c; init_row[3] = d;
      init_row[4] = 1.0 / (a * d - b * c);
    }
  }
  __syncthreads();
  const double y1b = init_row[0], y2b = init_row[1], dy1b = init_row[2], dy2b = init_row[3], idet = init_row[4];
  for (int it = threadIdx.x; it < 2 * P * (K1 + 1); it += blockDim.x) {
    const int q = it / (K1 + 1), j = it - q * (K1 + 1);
    int i0; double w;
    time_to_index(tb, (double)times_b[pairs[q]], i0, w);           // pairs is [P][2] row-major -> q = 2p + k
    const double y1 = lerp_t(tb.y1[i0], tb.y1[i0 + 1], w), y2 = lerp_t(tb.y2[i0], tb.y2[i0 + 1], w);
    const double xi1 = (dy2b * y1 - dy1b * y2) * idet, xi2 = (y1b * y2 - y2b * y1) * idet;
    if (j < K1) {
      const double pj = lerp_t(tb.pos[(size_t)i0 * K1 + j], tb.pos[(size_t)(i0 + 1) * K1 + j], w);
      hs[q * K1 + j] = (pj - xi1 * init_row[5 + j] - xi2 * init_row[5 + K1 + j]) * tb.scale[j];
    } else {
      xi[2 * q] = xi1;
      xi[2 * q + 1] = xi2;
    }
  }
  __syncthreads();
}

// decode a lower-triangular tile index t -> (I, J), I >= J
__device__ __forceinline__ void tri_decode(int t, int &I, int &J) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > t) --i;
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  I = i;
  J = t - i * (i + 1) / 2;
}

// load the lower triangle of a dense [n, n] matrix into padded shared memory (row stride LD), upper = 0,
// rows n..NR-1 = 0
__device__ void load_lower(const float *__restrict__ L, float *Ls, int n, int NR, int LD) {
  for (int e = threadIdx.x; e < NR * LD; e += blockDim.x) {
    const int i = e / LD, c = e % LD;
    Ls[e] = (i < n && c <= i) ? L[(size_t)i * n + c] : 0.0f;
  }
}

// =====================================================================================================
// Stage 1: gram.  One CTA per episode.
// =====================================================================================================
// SIGMA_IN: the covariance Sigma = scale * Sigma0 ([Dp, Dp] fp64, ONE matrix for the batch, e.g. straight out of the
// KL projection) is given instead of its factor: no L load and no L L^T per episode.
template <int D, int K1, bool SIGMA_IN>
__global__ void __launch_bounds__(SL_THREADS, 4)
seglik_gram_kernel(TabDev tb, const float *__restrict__ smp_traj, const float *__restrict__ mean,
                   const float *__restrict__ L, long long ldb_L, const double *__restrict__ Sigma0,
                   const double *__restrict__ sigma_scale, const float *__restrict__ times,
                   const float *__restrict__ init_time, const float *__restrict__ init_pos,
                   const float *__restrict__ init_vel, const int64_t *__restrict__ pairs, double *__restrict__ Cmat,
                   double *__restrict__ Rres, double *__restrict__ diag_max, int T, int P) {
  constexpr int Dp = D * K1, N = 2 * D, NT = tri(N);
  constexpr int NR = (Dp + 1) & ~1;          // Sg rows / cols (even)
  constexpr int NR4 = (Dp + 3) & ~3;         // Ls rows padded to a multiple of 4 for the 4x4 tiles
  constexpr int LD = NR4 + 1;                // padded row stride of Ls (floats)
  constexpr int SD = NR + 1;                 // row stride of Sg (doubles), odd, >= NR (tile padding)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *Sg = reinterpret_cast<double *>(smem_raw);              // [NR][SD]
  double *hs = Sg + NR * SD;                                      // [2P][K1]
  double *xi = hs + 2 * P * K1;                                   // [2P][2]
  double *init_row = xi + 4 * P;                                  // [5 + 2 K1]
  float *Ls = reinterpret_cast<float *>(init_row + 5 + 2 * K1 + 1);  // [NR4][LD]
  __shared__ double s_max[SL_THREADS / 32];

  const long long b = blockIdx.x;
  const float *times_b = times + b * T;

  SL_STAMP(0);
  if (SIGMA_IN) {
    const double sc = sigma_scale ? *sigma_scale : 1.0;
    for (int e0 = threadIdx.x; e0 < Dp * Dp; e0 += 4 * SL_THREADS) {     // four loads in flight per thread
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (e0 + u * SL_THREADS < Dp * Dp) ? Sigma0[e0 + u * SL_THREADS] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * SL_THREADS;
        if (e < Dp * Dp) Sg[(e / Dp) * SD + (e % Dp)] = sc * v[u];
      }
    }
    if (NR > Dp)                                                          // zero padding row / column
      for (int t = threadIdx.x; t < NR; t += blockDim.x) { Sg[Dp * SD + t] = 0.0; Sg[t * SD + Dp] = 0.0; }
  } else {
    load_lower(L + b * ldb_L, Ls, Dp, NR4, LD);
  }
  __syncthreads();
  SL_STAMP(1);
  basis_points<K1>(tb, (double)init_time[b], times_b, pairs, P, hs, xi, init_row);   // ends with a barrier
  SL_STAMP(2);

  // ---- Sigma = L L^T (fp32 FFMA, 4x4 register tiles over the lower triangle: 8 LDS per 16 FFMA) -> Sg (fp64,
  //      mirrored).  Rows/cols are padded to NR4 (multiple of 4) with zeros.
  if (!SIGMA_IN) {
    constexpr int NT4 = NR4 / 4;
    for (int t = threadIdx.x; t < tri(NT4); t += blockDim.x) {
      int I, J;
      tri_decode(t, I, J);
      const float *a = Ls + (4 * I) * LD, *bq = Ls + (4 * J) * LD;
      float c[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) c[x][y] = 0.f;
      const int kmax = 4 * J + 3;                                 // L[j][k] = 0 for k > j
      for (int k = 0; k <= kmax; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { av[x] = a[x * LD + k]; bv[x] = bq[x * LD + k]; }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) c[x][y] = fmaf(av[x], bv[y], c[x][y]);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int i = 4 * I + x, j = 4 * J + y;
          if (i < NR && j < NR) { Sg[i * SD + j] = c[x][y]; Sg[j * SD + i] = c[x][y]; }
        }
    }
  }
  __syncthreads();

  SL_STAMP(3);
  // ---- C blocks (fp64).  When the pairs form a chain (second time of pair p == first time of pair p+1: the
  //      fixed-interval selection of every TCE config) the P+1 distinct time points are processed once:
  //      task (blk, u): v = Sigma_dd' h_u, then K[u,u], K[u+1,u], K[u-1,u] -> 108 instead of 198 DFMA per pair.
  double my_max = 0.0;
  double *Cb = Cmat + (size_t)b * NT * P;
  int chain_ok = 1;
  for (int pp = threadIdx.x; pp + 1 < P; pp += blockDim.x) chain_ok &= (pairs[2 * pp + 1] == pairs[2 * pp + 2]);
  const bool chained = __syncthreads_and(chain_ok) != 0;
  if (chained) {
    const int U = P + 1;
    for (int task = threadIdx.x; task < tri(D) * U; task += blockDim.x) {
      const int blk = task / U, u = task - blk * U;
      int d, dd;
      tri_decode(blk, d, dd);
      const double *hu = hs + (u < P ? 2 * u : 2 * P - 1) * K1;
      const double *hn = hs + (u + 1 < P ? 2 * (u + 1) : 2 * P - 1) * K1;      // time u + 1 (valid when u < P)
      const double *hm = hs + (u >= 1 ? 2 * (u - 1) : 0) * K1;                  // time u - 1 (valid when u >= 1)
      double h[K1];
#pragma unroll
      for (int j = 0; j < K1; ++j) h[j] = hu[j];
      double kuu = 0.0, kup = 0.0, kum = 0.0;
      const double *S = Sg + (d * K1) * SD + dd * K1;
#pragma unroll
      for (int i = 0; i < K1; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < K1; ++j) v = fma(S[i * SD + j], h[j], v);
        kuu = fma(hu[i], v, kuu);
        kup = fma(hn[i], v, kup);
        kum = fma(hm[i], v, kum);
      }
      const int r0 = 2 * d, q0 = 2 * dd;
      if (u < P) {                      // pair p = u: time u is its first point
        Cb[(size_t)tri_idx(r0, q0) * P + u] = kuu;
        Cb[(size_t)tri_idx(r0 + 1, q0) * P + u] = kup;          // (d, t_{u+1}) x (d', t_u)
      }
      if (u >= 1) {                     // pair p = u - 1: time u is its second point
        Cb[(size_t)tri_idx(r0 + 1, q0 + 1) * P + (u - 1)] = kuu;
        if (d != dd) Cb[(size_t)tri_idx(r0, q0 + 1) * P + (u - 1)] = kum;   // (d, t_{u-1}) x (d', t_u)
      }
      if (d == dd) my_max = fmax(my_max, kuu);
    }
  } else
  for (int task = threadIdx.x; task < tri(D) * P; task += blockDim.x) {
    const int blk = task / P, p = task % P;
    int d, dd;
    tri_decode(blk, d, dd);
    double h0[K1], h1[K1];
#pragma unroll
    for (int j = 0; j < K1; ++j) { h0[j] = hs[(2 * p) * K1 + j]; h1[j] = hs[(2 * p + 1) * K1 + j]; }
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    const double *S = Sg + (d * K1) * SD + dd * K1;
#pragma unroll
    for (int i = 0; i < K1; ++i) {
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int j = 0; j < K1; ++j) {
        const double s = S[i * SD + j];
        t0 = fma(s, h0[j], t0);
        t1 = fma(s, h1[j], t1);
      }
      c00 = fma(h0[i], t0, c00); c01 = fma(h0[i], t1, c01);
      c10 = fma(h1[i], t0, c10); c11 = fma(h1[i], t1, c11);
    }
    const int r0 = 2 * d, q0 = 2 * dd;
    Cb[(size_t)tri_idx(r0, q0) * P + p] = c00;
    Cb[(size_t)tri_idx(r0 + 1, q0) * P + p] = c10;
    Cb[(size_t)tri_idx(r0 + 1, q0 + 1) * P + p] = c11;
    if (d != dd) {
      Cb[(size_t)tri_idx(r0, q0 + 1) * P + p] = c01;
    } else {
      my_max = fmax(my_max, fmax(c00, c11));
    }
  }

  SL_STAMP(4);
  // ---- residual r = x - mu (fp64): task (p, k, d)
  double *Rb = Rres + (size_t)b * N * P;
  const double tau = tb.tau;
  for (int task = threadIdx.x; task < 2 * P * D; task += blockDim.x) {
    const int d = task / (2 * P), q = task % (2 * P);          // q = 2p + k
    const int p = q >> 1, k = q & 1;
    const double y0 = (double)init_pos[b * D + d], v0 = (double)init_vel[b * D + d] * tau;
    double mu = xi[2 * q] * y0 + xi[2 * q + 1] * v0;
    const float *th = mean + b * Dp + d * K1;
#pragma unroll
    for (int j = 0; j < K1; ++j) mu = fma(hs[q * K1 + j], (double)th[j], mu);
    if (tb.relative_goal) {
      const double shift = tb.relative_goal_scaled ? y0 : y0 / tb.scale[K1 - 1];
      mu = fma(hs[q * K1 + K1 - 1], shift, mu);
    }
    const double x = (double)smp_traj[(b * T + pairs[q]) * (2 * D) + d];
    Rb[(size_t)(2 * d + k) * P + p] = x - mu;
  }

  SL_STAMP(5);
  // ---- batch-global max of the un-regularised diagonal
  my_max = warp_max(my_max);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = my_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < SL_THREADS / 32; ++w) m = fmax(m, s_max[w]);
    atomic_max_pos_double(diag_max, m);
  }
  SL_STAMP(6);
}

// =====================================================================================================
// Stage 2: per-segment Cholesky / log-prob / adjoints.  One thread per (b, p); the packed lower
// triangle lives in shared memory as [entry][thread] (conflict free, compile-time offsets).
// =====================================================================================================
constexpr int CH_THREADS = 32;   // NT * CH_THREADS doubles of static shared memory (27 KB at n = 14)

template <int N>
__global__ void __launch_bounds__(CH_THREADS)
seglik_chol_kernel(const double *Cmat, const double *Rres, double *Gout, double *Aout, const double *__restrict__ diag_max,
                   double reg_rel, const float *__restrict__ grad_logp, const float *__restrict__ logp_old,
                   const float *__restrict__ advantage, double grad_scale, double *__restrict__ loss_acc,
                   float *__restrict__ logp, int32_t *__restrict__ info, long long BP, int P, int want_grad) {
  constexpr int NT = tri(N);
#ifndef SEGLIK_CHOL_SMEM
  double c[NT];                      // fully unrolled below: every index is a compile-time constant -> registers
#else
  __shared__ double sm[NT * CH_THREADS];
#endif
  const long long gid = (long long)blockIdx.x * CH_THREADS + threadIdx.x;
  const bool active = gid < BP;
  double loss_part = 0.0, ratio_part = 0.0;
  if (active) {
    const long long b = gid / P;
    const int p = (int)(gid % P);
    const double *Cb = Cmat + (size_t)b * NT * P + p;
    const double *Rb = Rres + (size_t)b * N * P + p;
    double *Gb = Gout + (size_t)b * NT * P + p;
    double *Ab = Aout + (size_t)b * N * P + p;
#ifndef SEGLIK_CHOL_SMEM
#define CE(r, q) c[tri_idx(r, q)]
#define CLIN(e) c[e]
#else
    double *c = sm + threadIdx.x;
#define CE(r, q) c[(tri_idx(r, q)) * CH_THREADS]
#define CLIN(e) c[(e) * CH_THREADS]
#endif
    const double reg = reg_rel * (*diag_max);
#pragma unroll
    for (int e = 0; e < NT; ++e) CLIN(e) = Cb[(size_t)e * P];
    double z[N];
#pragma unroll
    for (int i = 0; i < N; ++i) z[i] = Rb[(size_t)i * P];
    // Cholesky (in place, lower), forward substitution and log-determinant
    int bad = 0;
    double half_logdet = 0.0, maha = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double dj = CE(j, j) + reg;
#pragma unroll
      for (int k = 0; k < j; ++k) dj = fma(-CE(j, k), CE(j, k), dj);
      if (!(dj > 0.0) && bad == 0) bad = j + 1;
      // 1/sqrt(dj): fp32 MUFU seed + two fp64 Newton steps (1e-7 -> 1e-14 -> rounding).  The DIAGONAL STORES THE
      // RECIPROCAL 1/S_jj: every later division (substitutions, inverse) becomes a multiplication -- fp64 divide
      // and sqrt are ~40-instruction subroutines and there were ~130 of them per segment.
      double inv = (double)rsqrtf((float)dj);
      inv = inv * fma(-0.5 * dj, inv * inv, 1.5);
      inv = inv * fma(-0.5 * dj, inv * inv, 1.5);
      CE(j, j) = inv;
      half_logdet += 0.5 * log(dj);
#pragma unroll
      for (int i = j + 1; i < N; ++i) {
        double v = CE(i, j);
#pragma unroll
        for (int k = 0; k < j; ++k) v = fma(-CE(i, k), CE(j, k), v);
        CE(i, j) = v * inv;
      }
      double zj = z[j];
#pragma unroll
      for (int k = 0; k < j; ++k) zj = fma(-CE(j, k), z[k], zj);
      zj *= inv;
      z[j] = zj;
      maha = fma(zj, zj, maha);
    }
    const double lp = -0.5 * ((double)N * 1.8378770664093453 + maha) - half_logdet;
    if (logp) logp[gid] = (float)lp;
    if (info) info[gid] = bad;
    if (want_grad) {
      double g;
      if (logp_old) {                      // fused surrogate: loss = -grad_scale * sum ratio * adv
        const double ratio = exp(lp - (double)logp_old[gid]);
        g = -ratio * (double)advantage[gid] * grad_scale;
        loss_part = g;
        ratio_part = ratio * grad_scale;
      } else {
        g = (double)grad_logp[gid];
      }
      // alpha = S^-T z (back substitution)
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        double v = z[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) v = fma(-CE(k, i), z[k], v);
        z[i] = v * CE(i, i);
      }
      // S <- S^-1 (lower, in place; the diagonal already holds 1/S_jj = X_jj)
#pragma unroll
      for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
          double v = 0.0;
#pragma unroll
          for (int k = j; k < i; ++k) v = fma(CE(i, k), CE(k, j), v);
          CE(i, j) = -v * CE(i, i);
        }
      }
      // (column j: S[i][k], j <= k < i, are still the Cholesky entries; X[k][j], k < i, are done)
      // C^-1 = X^T X (lower, in place, row by row: entry (i,j) only reads rows k >= i, and within row i
      // the diagonal X[i][i] is consumed last), then G = g/2 (alpha alpha^T - C^-1)
#pragma unroll
      for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double v = 0.0;
#pragma unroll
          for (int k = i; k < N; ++k) v = fma(CE(k, i), CE(k, j), v);
          CE(i, j) = 0.5 * g * (z[i] * z[j] - v);
        }
      }
#pragma unroll
      for (int e = 0; e < NT; ++e) Gb[(size_t)e * P] = CLIN(e);
#pragma unroll
      for (int i = 0; i < N; ++i) Ab[(size_t)i * P] = g * z[i];
    }
#undef CE
#undef CLIN
  }
  if (loss_acc) {
    loss_part = warp_sum(loss_part);
    ratio_part = warp_sum(ratio_part);
    if ((threadIdx.x & 31) == 0 && ratio_part != 0.0) {
      atomicAdd(loss_acc, loss_part);
      atomicAdd(loss_acc + 1, ratio_part);
    }
  }
}

// =====================================================================================================
// Stage 3: backward accumulation.  One CTA per episode.
// =====================================================================================================
// DSIGMA: write up * d logp / d Sigma (symmetric [Dp, Dp]) instead of grad_L = 2 up tril(dSigma L).  With ONE
// covariance for the batch the product with L is linear in dSigma, so it is applied once to the batch sum
// (dsigma_to_dl_kernel) instead of once per episode, and L is not loaded at all.
template <int D, int K1, bool DSIGMA>
__global__ void __launch_bounds__(SL_THREADS, 4)
seglik_bwd_kernel(TabDev tb, const double *__restrict__ Gmat, const double *__restrict__ Alpha,
                  const float *__restrict__ L, long long ldb_L, const float *__restrict__ times,
                  const float *__restrict__ init_time, const int64_t *__restrict__ pairs,
                  const float *__restrict__ upstream, float *__restrict__ grad_mean, float *__restrict__ grad_L,
                  int T, int P) {
  constexpr int Dp = D * K1, N = 2 * D, NT = tri(N);
  constexpr int NR = (Dp + 3) & ~3;          // rows padded to a multiple of 4 for the 4x4 tiles
  constexpr int LD = NR + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *Gs = reinterpret_cast<float *>(smem_raw);                // [NT][P] staged adjoints, fp32 (later: out tile)
  const int gs_floats = (NT * P > Dp * Dp ? NT * P : Dp * Dp + 1) & ~1;              // same rule as bwd_smem()
  double *hs = reinterpret_cast<double *>(Gs + gs_floats);       // [2P][K1]
  double *xi = hs + 2 * P * K1;                                   // [2P][2]
  double *init_row = xi + 4 * P;                                  // [5 + 2 K1]
  float *Ls = reinterpret_cast<float *>(init_row + 5 + 2 * K1 + 1);  // [NR][LD]
  float *Ms = Ls + NR * LD;                                       // [NR][LD]  dlp/dSigma (symmetric)

  const long long b = blockIdx.x;
  const double *Gb = Gmat + (size_t)b * NT * P;
  const double *Ab = Alpha + (size_t)b * N * P;
  const float up = upstream ? *upstream : 1.0f;       // scalar gradient of the fused surrogate loss

  SL_STAMP(16);
  if (!DSIGMA) load_lower(L + b * ldb_L, Ls, Dp, NR, LD);
  for (int e = threadIdx.x; e < NT * P; e += blockDim.x) Gs[e] = (float)Gb[e];
  for (int e = threadIdx.x; e < NR * LD; e += blockDim.x) Ms[e] = 0.f;
  basis_points<K1>(tb, (double)init_time[b], times + b * T, pairs, P, hs, xi, init_row);

  SL_STAMP(17);
  // ---- grad_mean[d*K1 + j] = sum_{p,k} h_pk[j] * (g alpha)[(d,k)]
  if (grad_mean) {
    for (int o = threadIdx.x; o < Dp; o += blockDim.x) {
      const int d = o / K1, j = o % K1;
      double acc = 0.0;
      for (int q = 0; q < 2 * P; ++q) acc = fma(hs[q * K1 + j], Ab[(size_t)(2 * d + (q & 1)) * P + (q >> 1)], acc);
      grad_mean[b * Dp + o] = up * (float)acc;
    }
  }
  if (!grad_L) return;

  SL_STAMP(18);
  // ---- M_{dd'}[i][:] = sum_p sum_{k,k'} G[(d,k),(d',k')] h_pk[i] h_pk'[:]   (fp64), task (blk, i)
  for (int task = threadIdx.x; task < tri(D) * K1; task += blockDim.x) {
    const int blk = task / K1, i = task % K1;
    int d, dd;
    tri_decode(blk, d, dd);
    double acc[K1];
#pragma unroll
    for (int j = 0; j < K1; ++j) acc[j] = 0.0;
    const int r0 = 2 * d, q0 = 2 * dd;
    const float *g00 = Gs + (size_t)tri_idx(r0, q0) * P;
    const float *g10 = Gs + (size_t)tri_idx(r0 + 1, q0) * P;
    const float *g11 = Gs + (size_t)tri_idx(r0 + 1, q0 + 1) * P;
    const float *g01 = (d != dd) ? Gs + (size_t)tri_idx(r0, q0 + 1) * P : g10;    // symmetric inside a diagonal block
    for (int p = 0; p < P; ++p) {
      const double a0 = hs[(2 * p) * K1 + i], a1 = hs[(2 * p + 1) * K1 + i];
      const double w0 = a0 * (double)g00[p] + a1 * (double)g10[p];      // coefficient of h_p0[:]
      const double w1 = a0 * (double)g01[p] + a1 * (double)g11[p];      // coefficient of h_p1[:]
#pragma unroll
      for (int j = 0; j < K1; ++j)
        acc[j] = fma(w0, hs[(2 * p) * K1 + j], fma(w1, hs[(2 * p + 1) * K1 + j], acc[j]));
    }
#pragma unroll
    for (int j = 0; j < K1; ++j) {
      const float v = (float)acc[j];
      Ms[(d * K1 + i) * LD + dd * K1 + j] = v;
      if (d != dd) Ms[(dd * K1 + j) * LD + d * K1 + i] = v;
    }
  }
  __syncthreads();

  SL_STAMP(19);
  if (DSIGMA) {                                   // up * dSigma, symmetric, dense [Dp][Dp]
    float *gS = grad_L + (size_t)b * Dp * Dp;
    for (int e = threadIdx.x; e < Dp * Dp; e += blockDim.x) gS[e] = up * Ms[(e / Dp) * LD + (e % Dp)];
    return;
  }
  // ---- grad_L = 2 * tril(M L)  (fp32 FFMA, 4x4 tiles over the lower triangle), staged in the Gs buffer
  float *out = Gs;                                // [Dp][Dp] dense (the Gs region is sized for it)
  for (int e = threadIdx.x; e < Dp * Dp; e += blockDim.x) out[e] = 0.f;
  __syncthreads();
  {
    constexpr int NT4 = NR / 4;
    for (int t = threadIdx.x; t < tri(NT4); t += blockDim.x) {
      int I, J;
      tri_decode(t, I, J);
      const float *mrow = Ms + (4 * I) * LD;
      float c[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) c[x][y] = 0.f;
      for (int k = NR - 1; k >= 4 * J; --k) {            // L[k][c] = 0 for k < c; descending: lanes share k
        float mv[4], lv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { mv[x] = mrow[x * LD + k]; lv[x] = Ls[k * LD + 4 * J + x]; }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) c[x][y] = fmaf(mv[x], lv[y], c[x][y]);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int i = 4 * I + x, j = 4 * J + y;
          if (i < Dp && j <= i) out[i * Dp + j] = 2.f * up * c[x][y];
        }
    }
  }
  __syncthreads();
  SL_STAMP(20);
  float *gL = grad_L + (size_t)b * Dp * Dp;
  for (int e = threadIdx.x; e < Dp * Dp; e += blockDim.x) gL[e] = out[e];
  SL_STAMP(21);
}

// grad_L = 2 tril((sum_b dSigma_b) L): CTA per row i.  Stage 1: the row of the batch sum (B x n values, four
// interleaved partial sums per column, fixed order -> deterministic); stage 2: thread per column j, fp32 as the
// per-episode product it replaces, L read coalesced straight from L2 with eight loads in flight.
constexpr int DS_PARTS = 16;                                      // 1024 threads = 16 batch slices x 64 columns
__global__ void __launch_bounds__(DS_PARTS * 64)
dsigma_to_dl_kernel(const float *__restrict__ dS, long long B, const float *__restrict__ L, float *__restrict__ out,
                    int n) {
  extern __shared__ float sm_f[];
  float *part = sm_f, *srow = sm_f + DS_PARTS * 64;               // [16][64] partial sums, [n] row
  const int i = blockIdx.x, k = threadIdx.x & 63, q = threadIdx.x >> 6;
  const size_t nn = (size_t)n * n;
  for (int k0 = 0; k0 < n; k0 += 64) {                              // n <= 128: at most two passes
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // eight loads in flight per thread
    if (k0 + k < n) {
      const float *src = dS + (size_t)i * n + k0 + k;
      long long b = q;
      for (; b + 7 * DS_PARTS < B; b += 8 * DS_PARTS) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += src[(size_t)(b + u * DS_PARTS) * nn];
      }
      for (; b < B; b += DS_PARTS) a[0] += src[(size_t)b * nn];
    }
    part[q * 64 + k] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    __syncthreads();
    if (threadIdx.x < 64 && k0 + k < n) {
      float t = 0.f;
#pragma unroll
      for (int u = 0; u < DS_PARTS; ++u) t += part[u * 64 + k];
      srow[k0 + k] = t;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (j <= i) {
      int kk = j;                                                   // L[k][j] = 0 for k < j
#pragma unroll 2
      for (; kk + 3 < n; kk += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(srow[kk + u], L[(size_t)(kk + u) * n + j], acc[u]);
      }
      for (; kk < n; ++kk) acc[0] = fmaf(srow[kk], L[(size_t)kk * n + j], acc[0]);
    }
    out[(size_t)i * n + j] = j <= i ? 2.f * ((acc[0] + acc[1]) + (acc[2] + acc[3])) : 0.f;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
template <int D, int K1>
size_t gram_smem(int P) {
  constexpr int Dp = D * K1, NR = (Dp + 1) & ~1, NR4 = (Dp + 3) & ~3, LD = NR4 + 1, SD = NR + 1;
  return sizeof(double) * ((size_t)NR * SD + 2 * P * K1 + 4 * P + 5 + 2 * K1 + 1) + sizeof(float) * NR4 * LD + 16;
}
template <int D, int K1>
size_t bwd_smem(int P) {
  constexpr int Dp = D * K1, N = 2 * D, NT = tri(N), NR = (Dp + 3) & ~3, LD = NR + 1;
  const size_t gs_floats = (size_t)((NT * P > Dp * Dp ? NT * P : Dp * Dp + 1) & ~1);
  const size_t g = sizeof(float) * gs_floats;
  return g + sizeof(double) * (2 * P * K1 + 4 * P + 5 + 2 * K1 + 1) + sizeof(float) * 2 * NR * LD + 16;
}

}  // namespace

// (D, K1) combinations with instantiated kernels
#define TCE_FOR_SHAPES(X) X(7, 9) X(4, 9) X(7, 4) X(3, 4) X(2, 3)

// debugging aid: SM-clock stamps of block 0 of the last gram ([0..6]) and bwd ([16..21]) launches (synchronises)
extern "C" int tce_debug_seglik_phase_cycles(long long *out32) {
  if (!out32) return TCE_ERR_INVALID_ARGUMENT;
#ifdef TCE_PROFILE
  TCE_CUDA(cudaDeviceSynchronize(), "seglik prof sync");
  TCE_CUDA(cudaMemcpyFromSymbol(out32, g_sl_prof, 32 * sizeof(long long)), "seglik prof copy");
  return TCE_OK;
#else
  return TCE_ERR_UNSUPPORTED_SHAPE;          /* library built without -DTCE_PROFILE */
#endif
}

extern "C" size_t tce_seglik_work_bytes(const tce_tables_t *t, int64_t B, int64_t P) {
  if (!t || B < 0 || P < 0) return 0;
  const size_t N = 2 * (size_t)t->D;
  return sizeof(double) * (size_t)B * (size_t)P * (N * (N + 1) / 2 + N);
}

static inline double *work_R(const tce_tables_t *t, void *work, int64_t B, int64_t P) {
  const size_t N = 2 * (size_t)t->D;
  return (double *)work + (size_t)B * (size_t)P * (N * (N + 1) / 2);
}

template <bool SIGMA_IN>
static int seglik_gram_launch(const tce_tables_t *t, const float *smp_traj, const float *mean, const float *L,
                              int64_t ldb_L, const double *Sigma0, const double *sigma_scale, const float *times,
                              const float *init_time, const float *init_pos, const float *init_vel,
                              const int64_t *pred_pairs, void *work, double *diag_max, int64_t B, int64_t T, int64_t P,
                              void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!t || !smp_traj || !mean || (SIGMA_IN ? !Sigma0 : !L) || !times || !init_time || !init_pos || !init_vel ||
      !pred_pairs || !work || !diag_max || B < 0 || T < 1 || P < 1 || P > 4096)
    return TCE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  double *Cmat = (double *)work, *R = work_R(t, work, B, P);
#define X(Dv, Kv)                                                                                                 \
  if (t->D == Dv && t->K1 == Kv) {                                                                                \
    const size_t smem = gram_smem<Dv, Kv>((int)P);                                                                \
    if (smem > 200 * 1024) return TCE_ERR_UNSUPPORTED_SHAPE;                                                      \
    TCE_CUDA(cudaFuncSetAttribute(seglik_gram_kernel<Dv, Kv, SIGMA_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem), "gram smem attr");                                                  \
    cudaFuncSetAttribute(seglik_gram_kernel<Dv, Kv, SIGMA_IN>, cudaFuncAttributePreferredSharedMemoryCarveout,    \
                         cudaSharedmemCarveoutMaxShared);                                                         \
    seglik_gram_kernel<Dv, Kv, SIGMA_IN><<<(unsigned)B, SL_THREADS, smem, st>>>(                                   \
        tab_dev(t), smp_traj, mean, L, ldb_L, Sigma0, sigma_scale, times, init_time, init_pos, init_vel, pred_pairs, \
        Cmat, R, diag_max, (int)T, (int)P);                                                                       \
    TCE_CHECK_LAUNCH("seglik_gram_kernel");                                                                       \
    return TCE_OK;                                                                                                \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_gram(const tce_tables_t *t, const float *smp_traj, const float *mean, const float *L,
                               int64_t ldb_L, const float *times, const float *init_time, const float *init_pos,
                               const float *init_vel, const int64_t *pred_pairs, void *work, double *diag_max,
                               int64_t B, int64_t T, int64_t P, void *stream) {
  return seglik_gram_launch<false>(t, smp_traj, mean, L, ldb_L, nullptr, nullptr, times, init_time, init_pos, init_vel,
                                   pred_pairs, work, diag_max, B, T, P, stream);
}

extern "C" int tce_seglik_gram_sigma(const tce_tables_t *t, const float *smp_traj, const float *mean,
                                     const double *Sigma0, const double *sigma_scale, const float *times,
                                     const float *init_time, const float *init_pos, const float *init_vel,
                                     const int64_t *pred_pairs, void *work, double *diag_max, int64_t B, int64_t T,
                                     int64_t P, void *stream) {
  return seglik_gram_launch<true>(t, smp_traj, mean, nullptr, 0, Sigma0, sigma_scale, times, init_time, init_pos,
                                  init_vel, pred_pairs, work, diag_max, B, T, P, stream);
}

extern "C" int tce_seglik_chol(const tce_tables_t *t, const void *work, void *adj, const double *diag_max, double reg_rel,
                               const float *grad_logp, const float *logp_old, const float *advantage,
                               double grad_scale, double *loss_acc, float *logp, int32_t *info, int64_t B,
                               int64_t P, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!t || !work || !diag_max || B < 0 || P < 1) return TCE_ERR_INVALID_ARGUMENT;
  if (logp_old && !advantage) return TCE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const double *Cmat = (const double *)work, *R = work_R(t, const_cast<void *>(work), B, P);
  double *G = (double *)adj, *A = adj ? work_R(t, adj, B, P) : nullptr;
  const long long BP = (long long)B * P;
  const unsigned grid = (unsigned)((BP + CH_THREADS - 1) / CH_THREADS);
  const int want_grad = (grad_logp != nullptr) || (logp_old != nullptr);
  if (want_grad && !adj) return TCE_ERR_INVALID_ARGUMENT;
  switch (t->D) {
#define Y(Dv)                                                                                                  \
  case Dv:                                                                                                     \
    seglik_chol_kernel<2 * Dv><<<grid, CH_THREADS, 0, st>>>(Cmat, R, G, A, diag_max, reg_rel, grad_logp, logp_old, \
                                                            advantage, grad_scale, loss_acc, logp, info, BP,   \
                                                            (int)P, want_grad);                                \
    break;
    Y(2) Y(3) Y(4) Y(7)
#undef Y
    default: return TCE_ERR_UNSUPPORTED_SHAPE;
  }
  TCE_CHECK_LAUNCH("seglik_chol_kernel");
  return TCE_OK;
}

template <bool DSIGMA>
static int seglik_bwd_launch(const tce_tables_t *t, const void *work, const float *L, int64_t ldb_L, const float *times,
                             const float *init_time, const int64_t *pred_pairs, const float *upstream,
                             float *grad_mean, float *grad_out, int64_t B, int64_t T, int64_t P, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!t || !work || (!DSIGMA && !L) || !times || !init_time || !pred_pairs || B < 0 || T < 1 || P < 1)
    return TCE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const double *G = (const double *)work, *A = work_R(t, const_cast<void *>(work), B, P);
#define X(Dv, Kv)                                                                                                \
  if (t->D == Dv && t->K1 == Kv) {                                                                               \
    const size_t smem = bwd_smem<Dv, Kv>((int)P);                                                                \
    if (smem > 200 * 1024) return TCE_ERR_UNSUPPORTED_SHAPE;                                                     \
    TCE_CUDA(cudaFuncSetAttribute(seglik_bwd_kernel<Dv, Kv, DSIGMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem), "bwd smem attr");                                                  \
    cudaFuncSetAttribute(seglik_bwd_kernel<Dv, Kv, DSIGMA>, cudaFuncAttributePreferredSharedMemoryCarveout,      \
                         cudaSharedmemCarveoutMaxShared);                                                        \
    seglik_bwd_kernel<Dv, Kv, DSIGMA><<<(unsigned)B, SL_THREADS, smem, st>>>(                                     \
        tab_dev(t), G, A, L, ldb_L, times, init_time, pred_pairs, upstream, grad_mean, grad_out, (int)T, (int)P); \
    TCE_CHECK_LAUNCH("seglik_bwd_kernel");                                                                       \
    return TCE_OK;                                                                                               \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_bwd(const tce_tables_t *t, const void *work, const float *L, int64_t ldb_L,
                              const float *times, const float *init_time, const int64_t *pred_pairs,
                              const float *upstream, float *grad_mean, float *grad_L, int64_t B, int64_t T,
                              int64_t P, void *stream) {
  return seglik_bwd_launch<false>(t, work, L, ldb_L, times, init_time, pred_pairs, upstream, grad_mean, grad_L, B, T,
                                  P, stream);
}

extern "C" int tce_seglik_bwd_dsigma(const tce_tables_t *t, const void *work, const float *times,
                                     const float *init_time, const int64_t *pred_pairs, const float *upstream,
                                     float *grad_mean, float *grad_sigma, int64_t B, int64_t T, int64_t P,
                                     void *stream) {
  if (B != 0 && !grad_sigma) return TCE_ERR_INVALID_ARGUMENT;
  return seglik_bwd_launch<true>(t, work, nullptr, 0, times, init_time, pred_pairs, upstream, grad_mean, grad_sigma,
                                 B, T, P, stream);
}

extern "C" int tce_dsigma_to_dl(const float *grad_sigma, int64_t B, const float *L, float *grad_L, int n,
                                void *stream) {
  if (!grad_sigma || !L || !grad_L || n < 1 || n > 128 || B < 1) return TCE_ERR_INVALID_ARGUMENT;
  dsigma_to_dl_kernel<<<(unsigned)n, DS_PARTS * 64, sizeof(float) * (DS_PARTS * 64 + 128), (cudaStream_t)stream>>>(
      grad_sigma, (long long)B, L, grad_L, n);
  TCE_CHECK_LAUNCH("dsigma_to_dl_kernel");
  return TCE_OK;
}
