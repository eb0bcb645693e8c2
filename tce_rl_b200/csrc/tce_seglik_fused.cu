// (3) TCE segment-wise trajectory likelihood, fused: ONE persistent kernel per call does
//     basis rows -> gram C = H Sigma H^T -> per-segment Cholesky / log-prob (+ fused surrogate) -> adjoints ->
//     grad_mean and d logp / d Sigma (accumulated in the CTA over its episodes) [-> grad_L per episode],
// with no HBM workspace.  Replaces the staged kernels of tce_seglik.cu (gram / chol / bwd / dsigma_to_dl, ~80 MB of
// fp64 round trips at B = 1024) on the product path; the staged entry points stay for cross-checks.
// Reference semantics: TemporalCorrelatedPolicy.log_prob, mprl/rl/policy/temporal_correlated_policy.py:104-203
// (mp.update_inputs + get_traj_pos(flat) + get_traj_pos_cov + MultivariateNormal(covariance_matrix).log_prob) and
// TemporalCorrelatedAgent.surrogate_loss, mprl/rl/agent/temporal_correlated_agent.py:718-739.
//
// The batch-global regulariser reg = reg_rel * max_{b,p,i} C_bp[i,i] (mp_pytorch get_traj_pos_cov) is produced by a
// diagonal-only pre-pass (tce_seglik_diagmax, ~20 % of the gram work) BEFORE the fused kernel, so the
// all-reduce(MAX) of a multi-GPU run sits between two launches and nothing else splits the computation.
//
// Work decomposition of the fused kernel (one CTA of 256 threads per SM, E episodes per iteration, S = E * P):
//   phase 0  basis rows of the distinct time points (fp64 table lerps) and residuals r = x - mu  -> smem
//   phase 1  gram: warp task = (DoF block (d, d'), 32 (episode, time point) items): v = Sigma_dd' h (Sigma block rows
//            are warp-uniform -> broadcast 128-bit shared loads), then h.v, h_next.v, h_prev.v  -> C [entry][slot] smem
//   phase 2  thread per segment: packed 14 x 14 triangle in REGISTERS (compile-time indices), Cholesky with
//            reciprocal pivots, forward substitution, log-prob, upstream gradient (given or fused surrogate),
//            C^-1 via the explicit inverse of the factor, G = g/2 (alpha alpha^T - C^-1) in place
//   phase 3  thread per (episode, DoF block): 81 register accumulators,
//            dSigma_dd' = sum_q w_q h_q^T, w_q = (G-weighted) combination of h_q, h_next, h_prev  (chained pairs: the
//            P + 1 distinct time points are visited once); grad_mean by thread per (episode, parameter)
//   shared covariance : dSigma summed over the CTA's episodes in smem (fp32), ONE partial per CTA -> global;
//                       tce_seglik_dsigma_reduce sums the <= 148 partials and applies grad_L = 2 tril(dSigma L) once
//   per-episode factor: Sigma_b = L_b L_b^T (fp32 4x4 register tiles) before phase 1 of each episode and
//                       grad_L_b = 2 tril(dSigma_b L_b) after phase 3, both in shared memory.
//
// Uniform time grid (all episodes share init_time and hence the basis rows) + shared covariance -- the situation of
// every shipped TCE config -- makes C_p, its factor and its inverse identical for all episodes:
// tce_seglik_uniform_* evaluate them once and reduce the per-episode work to a residual, two 14 x 14 triangular
// products and a rank-1 update (~50x fewer FLOPs).
#include <math.h>

#include "tce_common.cuh"

namespace {

constexpr int FT = 256;                      // threads of the fused kernel
constexpr double LN_2PI = 1.8378770664093453;

__host__ __device__ constexpr int tri(int n) { return n * (n + 1) / 2; }
__device__ __forceinline__ int tri_idx(int r, int c) { return r * (r + 1) / 2 + c; }   // r >= c
__device__ __forceinline__ void tri_decode(int t, int &I, int &J) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > t) --i;
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  I = i;
  J = t - i * (i + 1) / 2;
}

// ---- links between time points and pairs -------------------------------------------------------------
// chained pairs (second time of pair p == first time of pair p + 1: the fixed-interval selection of every TCE
// config): nq = P + 1 distinct points, point q is the first point of pair q (q < P) and the second of pair q - 1.
// otherwise: nq = 2 P points, q = 2 p + k.
struct Links { int pf, ps, nxt, prv; };    // pair where q is first / second (-1: none), the partner points
__device__ __forceinline__ Links links(int chained, int q, int P) {
  Links l;
  if (chained) { l.pf = q < P ? q : -1; l.ps = q >= 1 ? q - 1 : -1; l.nxt = q + 1; l.prv = q - 1; }
  else if (q & 1) { l.pf = -1; l.ps = q >> 1; l.nxt = q; l.prv = q - 1; }
  else { l.pf = q >> 1; l.ps = -1; l.nxt = q + 1; l.prv = q; }
  return l;
}
__device__ __forceinline__ long long point_time_index(const int64_t *__restrict__ pairs, int chained, int q, int P) {
  if (!chained) return pairs[q];
  return q < P ? pairs[2 * q] : pairs[2 * P - 1];
}
__device__ __forceinline__ int point_of(int chained, int p, int k) { return chained ? p + k : 2 * p + k; }

__device__ int block_chained(const int64_t *__restrict__ pairs, int P) {
  int ok = 1;
  for (int pp = threadIdx.x; pp + 1 < P; pp += blockDim.x) ok &= (pairs[2 * pp + 1] == pairs[2 * pp + 2]);
  return __syncthreads_and(ok) != 0;
}

// ---- ProDMP basis rows with initial conditions ---------------------------------------------
// init_row [5 + 2 K1]: y1b, y2b, dy1b, dy2b, 1/det, pos_b[K1], vel_b[K1] at the episode's initial time
template <int K1>
__device__ __forceinline__ void init_row_entry(const TabDev &tb, double t_init, int j, double *init_row) {
  int i0; double w;
  time_to_index(tb, t_init, i0, w);
  if (j < K1) {
    init_row[5 + j] = lerp_t(tb.pos[(size_t)i0 * K1 + j], tb.pos[(size_t)(i0 + 1) * K1 + j], w);
    init_row[5 + K1 + j] = lerp_t(tb.vel[(size_t)i0 * K1 + j], tb.vel[(size_t)(i0 + 1) * K1 + j], w);
  } else {
    const double a = lerp_t(tb.y1[i0], tb.y1[i0 + 1], w), b = lerp_t(tb.y2[i0], tb.y2[i0 + 1], w);
    const double c = lerp_t(tb.dy1[i0], tb.dy1[i0 + 1], w), d = lerp_t(tb.dy2[i0], tb.dy2[i0 + 1], w);
    init_row[0] = a; init_row[1] = b; init_row[2] = c; init_row[3] = d;
    init_row[4] = 1.0 / (a * d - b * c);
  }
}
// entry j of the scaled basis row at time t (j < K1) or the pair (xi1, xi2) (j == K1)
template <int K1>
__device__ __forceinline__ void basis_entry(const TabDev &tb, const double *init_row, double t, int j, double *h_row,
                                            double *xi_row) {
  int i0; double w;
  time_to_index(tb, t, i0, w);
  const double y1 = lerp_t(tb.y1[i0], tb.y1[i0 + 1], w), y2 = lerp_t(tb.y2[i0], tb.y2[i0 + 1], w);
  const double idet = init_row[4];
  const double xi1 = (init_row[3] * y1 - init_row[2] * y2) * idet, xi2 = (init_row[0] * y2 - init_row[1] * y1) * idet;
  if (j < K1) {
    const double pj = lerp_t(tb.pos[(size_t)i0 * K1 + j], tb.pos[(size_t)(i0 + 1) * K1 + j], w);
    h_row[j] = (pj - xi1 * init_row[5 + j] - xi2 * init_row[5 + K1 + j]) * tb.scale[j];
  } else {
    xi_row[0] = xi1;
    xi_row[1] = xi2;
  }
}

// ---- per-segment factorisation in registers -----------------------------------------------------------------
// c: packed lower triangle of C (no regulariser), z: residual.  On return c holds the Cholesky factor with the
// RECIPROCAL pivots on the diagonal, z = S^-1 r; returns log-prob, bad = index + 1 of the first non-positive pivot.
template <int N>
__device__ __forceinline__ double seg_factor(double (&c)[tri(N)], double (&z)[N], double reg, int &bad) {
#define CE(r, q) c[(r) * ((r) + 1) / 2 + (q)]
  bad = 0;
  double half_logdet = 0.0, maha = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double dj = CE(j, j) + reg;
#pragma unroll
    for (int k = 0; k < j; ++k) dj = fma(-CE(j, k), CE(j, k), dj);
    if (!(dj > 0.0) && bad == 0) bad = j + 1;
    double inv = (double)rsqrtf((float)dj);             // fp32 MUFU seed + two fp64 Newton steps
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
  return -0.5 * ((double)N * LN_2PI + maha) - half_logdet;
}
// after seg_factor: c <- g/2 (alpha alpha^T - C^-1) (lower), z <- g alpha
template <int N>
__device__ __forceinline__ void seg_adjoint(double (&c)[tri(N)], double (&z)[N], double g) {
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {                     // alpha = S^-T z
    double v = z[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) v = fma(-CE(k, i), z[k], v);
    z[i] = v * CE(i, i);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {                          // S <- S^-1 (diagonal already holds 1 / S_jj)
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) v = fma(CE(i, k), CE(k, j), v);
      CE(i, j) = -v * CE(i, i);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {                          // C^-1 = X^T X row by row, then G
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = i; k < N; ++k) v = fma(CE(k, i), CE(k, j), v);
      CE(i, j) = 0.5 * g * (z[i] * z[j] - v);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] *= g;
#undef CE
}

// phase 2 of the fused kernel for one segment: column `slot` of C [NT][S] / r [N][S] in shared memory -> log-prob,
// upstream gradient, adjoints written back in place.  Deliberately NOT inlined: the packed triangle needs ~240
// registers; as a separate function it gets the whole register file instead of competing with the kernel's
// loop state (inlined: 4-6 KB of spill traffic per thread).
struct SegIO {
  const float *grad_logp, *logp_old, *advantage;
  float *logp;
  int32_t *info;
  double grad_scale, reg;
  int grad_mode;
};
template <int N>
__device__ __noinline__ void segment_thread(double *Ccol, double *Rcol, int S, long long gid, const SegIO &io,
                                            double &loss_part, double &ratio_part) {
  constexpr int NT = tri(N);
  double c[NT], z[N];
#pragma unroll
  for (int t = 0; t < NT; ++t) c[t] = Ccol[(size_t)t * S];
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] = Rcol[(size_t)i * S];
  int bad;
  const double lp = seg_factor<N>(c, z, io.reg, bad);
  if (io.logp) io.logp[gid] = (float)lp;
  if (io.info) io.info[gid] = bad;
  if (io.grad_mode == 0) return;
  double g;
  if (io.grad_mode == 2) {
    const double ratio = exp(lp - (double)io.logp_old[gid]);
    g = -ratio * (double)io.advantage[gid] * io.grad_scale;
    loss_part += g;
    ratio_part += ratio * io.grad_scale;
  } else {
    g = (double)io.grad_logp[gid];
  }
  seg_adjoint<N>(c, z, g);
#pragma unroll
  for (int t = 0; t < NT; ++t) Ccol[(size_t)t * S] = c[t];
#pragma unroll
  for (int i = 0; i < N; ++i) Rcol[(size_t)i * S] = z[i];
}

// phase 3b of the fused kernel: dSigma_dd'[i][j] = sum_q w_q[i] h_q[j] for one (episode, DoF block); G = the episode's
// adjoints (rows of C [NT][S], first pair at G).  Not inlined for the same reason as segment_thread (81 accumulators).
template <int K1>
__device__ __noinline__ void dsigma_block(const double *G, int S, const double *hm_e, int bd, int bdd, int chained,
                                          int nq, int P, double (&acc)[K1 * K1]) {
#pragma unroll
  for (int t = 0; t < K1 * K1; ++t) acc[t] = 0.0;
  const int r0 = 2 * bd, q0 = 2 * bdd;
  const double *g00 = G + (size_t)tri_idx(r0, q0) * S, *g10 = G + (size_t)tri_idx(r0 + 1, q0) * S;
  const double *g11 = G + (size_t)tri_idx(r0 + 1, q0 + 1) * S;
  const double *g01 = (bd != bdd) ? G + (size_t)tri_idx(r0, q0 + 1) * S : g10;    // symmetric in a diagonal block
  for (int q = 0; q < nq; ++q) {
    const Links lk = links(chained, q, P);
    const double *hq = hm_e + q * K1;
    double cs = 0.0, cn = 0.0, cp = 0.0;                 // coefficients of h_q, h_next, h_prev on the row side
    if (lk.pf >= 0) { cs += g00[lk.pf]; cn = g10[lk.pf]; }
    if (lk.ps >= 0) { cs += g11[lk.ps]; cp = g01[lk.ps]; }
    const double *hn = hm_e + (lk.pf >= 0 ? lk.nxt : q) * K1, *hp = hm_e + (lk.ps >= 0 ? lk.prv : q) * K1;
    double h[K1], w[K1];
#pragma unroll
    for (int j = 0; j < K1; ++j) h[j] = hq[j];
#pragma unroll
    for (int i = 0; i < K1; ++i) w[i] = fma(cs, h[i], fma(cn, hn[i], cp * hp[i]));
#pragma unroll
    for (int i = 0; i < K1; ++i)
#pragma unroll
      for (int j = 0; j < K1; ++j) acc[i * K1 + j] = fma(w[i], h[j], acc[i * K1 + j]);
  }
}

// ---- shared-memory layout of the fused kernel (host + device) ------------------------------------------------
template <int D, int K1>
struct FusedLayout {
  static constexpr int Dp = D * K1, N = 2 * D, NT = tri(N), NB = tri(D);
  static constexpr int KP = (K1 + 1) & ~1;                  // padded row of a Sigma block (16-byte aligned rows)
  static constexpr int NR4 = (Dp + 3) & ~3, LD = NR4 + 1;   // fp32 staging of a factor: [NR4][LD]
  static constexpr int IR = (5 + 2 * K1 + 1) & ~1;          // init row doubles (even)
  size_t sblk, dsacc, cs, rs, hm, xi, ir, total;
  int E, S, SP, nq;        // SP: row stride (doubles) of the [entry][slot] arrays, odd -> rows fall into distinct banks
  __host__ __device__ FusedLayout(int E_, int P, int chained) {
    E = E_; S = E * P; SP = S | 1; nq = chained ? P + 1 : 2 * P;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
    const size_t sb = sizeof(double) * NB * K1 * KP, ls = sizeof(float) * NR4 * LD;
    sblk = take(sb > ls ? sb : ls);
    dsacc = take(sizeof(float) * NB * K1 * K1);
    // C [NT][SP]; after phase 3 the region is scratch: dSigma_b [E][NR4][LD] or block partials [E][NB][K1 K1] (fp32)
    const size_t c1 = sizeof(double) * NT * SP, c2 = sizeof(float) * (size_t)E * NR4 * LD;
    const size_t c3 = sizeof(float) * (size_t)E * NB * K1 * K1, c12 = c1 > c2 ? c1 : c2;
    cs = take(c12 > c3 ? c12 : c3);
    rs = take(sizeof(double) * N * SP);
    hm = take(sizeof(double) * (size_t)E * nq * K1);
    xi = take(sizeof(double) * (size_t)E * nq * 2);
    ir = take(sizeof(double) * (size_t)E * IR);
    total = o;
  }
};

// Sigma block (d, dd), d >= dd, entry (i, j): Sblk[tri_idx(d, dd)][i][j] with row stride KP
template <int K1, int KP>
__device__ __forceinline__ double *sblk_at(double *Sblk, int blk) { return Sblk + (size_t)blk * K1 * KP; }

// store an element (i >= j) of the symmetric Sigma into the block layout (diagonal blocks are stored full)
template <int K1, int KP>
__device__ __forceinline__ void sblk_store(double *Sblk, int i, int j, double v) {
  const int d = i / K1, dd = j / K1, ii = i - d * K1, jj = j - dd * K1;
  double *blk = sblk_at<K1, KP>(Sblk, tri_idx(d, dd));
  blk[ii * KP + jj] = v;
  if (d == dd) blk[jj * KP + ii] = v;
}

// load the lower triangle of a dense [n, n] fp32 matrix into padded shared memory, zero elsewhere
__device__ __forceinline__ void stage_lower(const float *__restrict__ L, float *Ls, int n, int NR, int LD) {
  for (int e0 = threadIdx.x; e0 < NR * NR; e0 += 4 * FT) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * FT, i = e / NR, c = e - i * NR;
      v[u] = (e < NR * NR && i < n && c <= i) ? L[(size_t)i * n + c] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * FT, i = e / NR, c = e - i * NR;
      if (e < NR * NR) Ls[i * LD + c] = v[u];
    }
  }
}

// Sigma = L L^T: one 4x4 register tile of the lower triangle per thread (tri(NR4 / 4) <= 256 tiles), then -> Sblk fp64
template <int D, int K1>
__device__ __forceinline__ void sigma_from_factor(const float *Ls, double *Sblk) {
  using FL = FusedLayout<D, K1>;
  constexpr int NT4 = FL::NR4 / 4, LD = FL::LD, Dp = FL::Dp;
  static_assert(tri(NT4) <= FT, "one tile per thread");
  float c[4][4];
  int I = 0, J = 0;
  const bool act = threadIdx.x < tri(NT4);
  if (act) {
    tri_decode(threadIdx.x, I, J);
    const float *a = Ls + (4 * I) * LD, *bq = Ls + (4 * J) * LD;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) c[x][y] = 0.f;
    const int kmax = 4 * J + 3;
    for (int k = 0; k <= kmax; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { av[x] = a[x * LD + k]; bv[x] = bq[x * LD + k]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) c[x][y] = fmaf(av[x], bv[y], c[x][y]);
    }
  }
  __syncthreads();                                         // everybody has read Ls: the region becomes Sblk
  if (act) {
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int i = 4 * I + x, j = 4 * J + y;
        if (i < Dp && j <= i) sblk_store<K1, FL::KP>(Sblk, i, j, (double)c[x][y]);
      }
  }
  __syncthreads();
}

// one gram item: DoF block (d, dd), point q of episode slot e -> up to four entries of C (chained: of two pairs)
template <int D, int K1>
__device__ __forceinline__ double gram_item(const double *__restrict__ Sb, const double *__restrict__ hm_e, int q,
                                            const Links &lk, int d, int dd, double *CS, int S, int slot0) {
  constexpr int KP = FusedLayout<D, K1>::KP;
  double h[K1], v[K1];
  const double *hq = hm_e + q * K1;
#pragma unroll
  for (int j = 0; j < K1; ++j) h[j] = hq[j];
#pragma unroll
  for (int i = 0; i < K1; ++i) {
    const double *row = Sb + i * KP;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int j = 0; j + 1 < K1; j += 2) {
      const double2 s = *reinterpret_cast<const double2 *>(row + j);
      a0 = fma(s.x, h[j], a0);
      a1 = fma(s.y, h[j + 1], a1);
    }
    if (K1 & 1) a0 = fma(row[K1 - 1], h[K1 - 1], a0);
    v[i] = a0 + a1;
  }
  double kuu = 0.0;
#pragma unroll
  for (int i = 0; i < K1; ++i) kuu = fma(h[i], v[i], kuu);
  const int r0 = 2 * d, q0 = 2 * dd;
  if (lk.pf >= 0) {                          // q is the first point of pair pf; partner = its second point
    const double *hn = hm_e + lk.nxt * K1;
    double kup = 0.0;
#pragma unroll
    for (int i = 0; i < K1; ++i) kup = fma(hn[i], v[i], kup);
    CS[(size_t)tri_idx(r0, q0) * S + slot0 + lk.pf] = kuu;
    CS[(size_t)tri_idx(r0 + 1, q0) * S + slot0 + lk.pf] = kup;          // (d, second) x (dd, first)
  }
  if (lk.ps >= 0) {                          // q is the second point of pair ps; partner = its first point
    CS[(size_t)tri_idx(r0 + 1, q0 + 1) * S + slot0 + lk.ps] = kuu;
    if (d != dd) {
      const double *hp = hm_e + lk.prv * K1;
      double kum = 0.0;
#pragma unroll
      for (int i = 0; i < K1; ++i) kum = fma(hp[i], v[i], kum);
      CS[(size_t)tri_idx(r0, q0 + 1) * S + slot0 + lk.ps] = kum;        // (d, first) x (dd, second)
    }
  }
  return kuu;
}

// =====================================================================================================
// diagonal pre-pass: max_{b, p, i} C_bp[i, i] -> *diag_max (atomic max on the bit pattern of a positive double)
// =====================================================================================================
template <int D, int K1, bool SIGMA_IN>
__global__ void __launch_bounds__(FT)
seglik_diagmax_kernel(TabDev tb, const float *__restrict__ L, long long ldb_L, const double *__restrict__ Sigma0,
                      const double *__restrict__ sigma_scale, const float *__restrict__ times,
                      const float *__restrict__ init_time, const int64_t *__restrict__ pairs,
                      double *__restrict__ diag_max, long long B, int T, int P) {
  using FL = FusedLayout<D, K1>;
  constexpr int Dp = FL::Dp, NR4 = FL::NR4, LD = FL::LD, IR = FL::IR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *Sd = reinterpret_cast<double *>(smem_raw);                   // [D][K1][K1] diagonal blocks
  double *hm = Sd + D * K1 * K1;                                       // [nq][K1]
  double *init_row = hm + (size_t)2 * P * K1;                          // [IR]
  float *Ls = reinterpret_cast<float *>(init_row + IR);                // [NR4][LD] (per-episode factor only)
  __shared__ double s_max[FT / 32];
  const int chained = block_chained(pairs, P);
  const int nq = chained ? P + 1 : 2 * P;
  const bool shared_cov = SIGMA_IN || ldb_L == 0;
  double my_max = 0.0;
  auto load_diag_blocks = [&](long long b) {
    if (SIGMA_IN) {
      const double sc = sigma_scale ? *sigma_scale : 1.0;
      for (int e = threadIdx.x; e < D * K1 * K1; e += FT) {
        const int d = e / (K1 * K1), r = e - d * K1 * K1, i = r / K1, j = r - i * K1;
        Sd[e] = sc * Sigma0[(size_t)(d * K1 + i) * Dp + d * K1 + j];
      }
    } else {
      stage_lower(L + b * ldb_L, Ls, Dp, NR4, LD);
      __syncthreads();
      for (int e = threadIdx.x; e < D * K1 * K1; e += FT) {
        const int d = e / (K1 * K1), r = e - d * K1 * K1, i = r / K1, j = r - i * K1;
        const float *a = Ls + (d * K1 + i) * LD, *bq = Ls + (d * K1 + j) * LD;
        const int kmax = d * K1 + (i < j ? i : j);
        float acc = 0.f;
        for (int k = 0; k <= kmax; ++k) acc = fmaf(a[k], bq[k], acc);
        Sd[e] = (double)acc;
      }
    }
    __syncthreads();
  };
  if (shared_cov) load_diag_blocks(0);
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    if (!shared_cov) load_diag_blocks(b);
    if (threadIdx.x <= K1) init_row_entry<K1>(tb, (double)init_time[b], threadIdx.x, init_row);
    __syncthreads();
    for (int it = threadIdx.x; it < nq * K1; it += FT) {
      const int q = it / K1, j = it - q * K1;
      double dummy[2];
      basis_entry<K1>(tb, init_row, (double)times[b * T + point_time_index(pairs, chained, q, P)], j, hm + q * K1,
                      dummy);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < D * nq; it += FT) {
      const int d = it / nq, q = it - d * nq;
      const double *S = Sd + d * K1 * K1, *h = hm + q * K1;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < K1; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < K1; ++j) v = fma(S[i * K1 + j], h[j], v);
        acc = fma(h[i], v, acc);
      }
      my_max = fmax(my_max, acc);
    }
    __syncthreads();
  }
  my_max = warp_max(my_max);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = my_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < FT / 32; ++w) m = fmax(m, s_max[w]);
    atomic_max_pos_double(diag_max, m);
  }
}

// =====================================================================================================
// fused kernel
// =====================================================================================================
// grad_mode: 0 = log-probs only, 1 = upstream gradient grad_logp [B, P], 2 = fused surrogate
//            (g = -exp(lp - lp_old) * adv * grad_scale, loss_acc[0] += sum g, loss_acc[1] += sum ratio * grad_scale)
template <int D, int K1, bool SIGMA_IN>
__global__ void __launch_bounds__(FT, 1)
seglik_fused_kernel(TabDev tb, const float *__restrict__ smp_traj, const float *__restrict__ mean,
                    const float *__restrict__ L, long long ldb_L, const double *__restrict__ Sigma0,
                    const double *__restrict__ sigma_scale, const float *__restrict__ times,
                    const float *__restrict__ init_time, const float *__restrict__ init_pos,
                    const float *__restrict__ init_vel, const int64_t *__restrict__ pairs,
                    const double *__restrict__ diag_max, double reg_rel, int grad_mode,
                    const float *__restrict__ grad_logp, const float *__restrict__ logp_old,
                    const float *__restrict__ advantage, double grad_scale, double *__restrict__ loss_acc,
                    float *__restrict__ logp, int32_t *__restrict__ info, float *__restrict__ grad_mean,
                    float *__restrict__ grad_L, float *__restrict__ dsigma_part, long long B, int T, int P, int E,
                    int chained) {
  using FL = FusedLayout<D, K1>;
  constexpr int Dp = FL::Dp, N = FL::N, NT = FL::NT, NB = FL::NB, KP = FL::KP, NR4 = FL::NR4, LD = FL::LD, IR = FL::IR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // `chained` is the caller's claim (the shared-memory layout was sized with it): refuse a wrong one
  if (chained && !block_chained(pairs, P)) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && info) info[0] = -7;
    return;
  }
  const FL lay(E, P, chained);
  const int S = lay.SP, nq = lay.nq;     // S: row stride of CS / RS
  double *Sblk = reinterpret_cast<double *>(smem_raw + lay.sblk);
  float *Ls = reinterpret_cast<float *>(smem_raw + lay.sblk);          // aliases Sblk (per-episode factor staging)
  float *dsacc = reinterpret_cast<float *>(smem_raw + lay.dsacc);      // [NB][K1][K1] running dSigma of this CTA
  double *CS = reinterpret_cast<double *>(smem_raw + lay.cs);          // [NT][S]
  float *dSb = reinterpret_cast<float *>(smem_raw + lay.cs);           // aliases CS after phase 3: [E][NR4][LD]
  double *RS = reinterpret_cast<double *>(smem_raw + lay.rs);          // [N][S]
  double *hm = reinterpret_cast<double *>(smem_raw + lay.hm);          // [E][nq][K1]
  double *xi = reinterpret_cast<double *>(smem_raw + lay.xi);          // [E][nq][2]
  double *irow = reinterpret_cast<double *>(smem_raw + lay.ir);        // [E][IR]
  __shared__ double s_red[2 * (FT / 32)];

  const bool shared_cov = SIGMA_IN || ldb_L == 0;
  const bool want_grad = grad_mode != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double reg = reg_rel * (*diag_max);
  const double tau = tb.tau;
  double loss_part = 0.0, ratio_part = 0.0;

  // ---- the shared covariance is staged once per CTA ----------------------------------------------------
  if (shared_cov) {
    if (SIGMA_IN) {
      const double sc = sigma_scale ? *sigma_scale : 1.0;
      for (int e0 = threadIdx.x; e0 < Dp * Dp; e0 += 4 * FT) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (e0 + u * FT < Dp * Dp) ? Sigma0[e0 + u * FT] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * FT, i = e / Dp, j = e - i * Dp;
          if (e < Dp * Dp && j <= i) sblk_store<K1, KP>(Sblk, i, j, sc * v[u]);
        }
      }
      __syncthreads();
    } else {
      stage_lower(L, Ls, Dp, NR4, LD);
      __syncthreads();
      sigma_from_factor<D, K1>(Ls, Sblk);
    }
    if (want_grad && dsigma_part)
      for (int e = threadIdx.x; e < NB * K1 * K1; e += FT) dsacc[e] = 0.f;
  }

  const long long n_groups = (B + E - 1) / E;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long b0 = grp * E;
    const int Ec = (int)((B - b0) < E ? (B - b0) : E);
    __syncthreads();
    // ---- phase 0: initial-condition rows, basis rows, residuals -----------------------------------------
    for (int it = threadIdx.x; it < Ec * (K1 + 1); it += FT) {
      const int e = it / (K1 + 1), j = it - e * (K1 + 1);
      init_row_entry<K1>(tb, (double)init_time[b0 + e], j, irow + e * IR);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < Ec * nq * (K1 + 1); it += FT) {
      const int e = it / (nq * (K1 + 1)), r = it - e * nq * (K1 + 1), q = r / (K1 + 1), j = r - q * (K1 + 1);
      const double t = (double)times[(b0 + e) * T + point_time_index(pairs, chained, q, P)];
      basis_entry<K1>(tb, irow + e * IR, t, j, hm + ((size_t)e * nq + q) * K1, xi + ((size_t)e * nq + q) * 2);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < Ec * P * N; it += FT) {          // residual r = x - mu, task (e, d, p, k)
      const int e = it / (P * N), r = it - e * P * N, d = r / (2 * P), pk = r - d * 2 * P, p = pk >> 1, k = pk & 1;
      const long long b = b0 + e;
      const int q = point_of(chained, p, k);
      const double *hq = hm + ((size_t)e * nq + q) * K1, *xq = xi + ((size_t)e * nq + q) * 2;
      const double y0 = (double)init_pos[b * D + d], v0 = (double)init_vel[b * D + d] * tau;
      double mu = xq[0] * y0 + xq[1] * v0;
      const float *th = mean + b * Dp + d * K1;
#pragma unroll
      for (int j = 0; j < K1; ++j) mu = fma(hq[j], (double)th[j], mu);
      if (tb.relative_goal) {
        const double shift = tb.relative_goal_scaled ? y0 : y0 / tb.scale[K1 - 1];
        mu = fma(hq[K1 - 1], shift, mu);
      }
      const double x = (double)smp_traj[(b * T + pairs[2 * p + k]) * (2 * D) + d];
      RS[(size_t)(2 * d + k) * S + e * P + p] = x - mu;
    }
    // ---- phase 1: gram ------------------------------------------------------------------------------------
    if (shared_cov) {
      const int items = Ec * nq, chunks = (items + 31) >> 5;
      for (int wt = warp; wt < NB * chunks; wt += FT / 32) {
        const int blk = wt / chunks, item = (wt - blk * chunks) * 32 + lane;
        if (item < items) {
          int d, dd;
          tri_decode(blk, d, dd);
          const int e = item / nq, q = item - e * nq;
          gram_item<D, K1>(sblk_at<K1, KP>(Sblk, blk), hm + (size_t)e * nq * K1, q, links(chained, q, P), d, dd, CS, S,
                           e * P);
        }
      }
    } else {
      for (int e = 0; e < Ec; ++e) {
        __syncthreads();
        stage_lower(L + (b0 + e) * ldb_L, Ls, Dp, NR4, LD);
        __syncthreads();
        sigma_from_factor<D, K1>(Ls, Sblk);
        for (int wt = warp; wt < NB * ((nq + 31) >> 5); wt += FT / 32) {
          const int chunks = (nq + 31) >> 5, blk = wt / chunks, q = (wt - blk * chunks) * 32 + lane;
          if (q < nq) {
            int d, dd;
            tri_decode(blk, d, dd);
            gram_item<D, K1>(sblk_at<K1, KP>(Sblk, blk), hm + (size_t)e * nq * K1, q, links(chained, q, P), d, dd, CS,
                             S, e * P);
          }
        }
      }
    }
    __syncthreads();
    // ---- phase 2: thread per segment ---------------------------------------------------------------------
    if ((int)threadIdx.x < Ec * P) {
      const int slot = threadIdx.x, e = slot / P, p = slot - e * P;
      SegIO io{grad_logp, logp_old, advantage, logp, info, grad_scale, reg, grad_mode};
      segment_thread<N>(CS + slot, RS + slot, S, (b0 + e) * P + p, io, loss_part, ratio_part);
    }
    if (!want_grad) continue;
    __syncthreads();
    // ---- phase 3a: grad_mean[d K1 + j] = sum_{p, k} h_{p,k}[j] (g alpha)[(d, k)] ---------------------------------
    if (grad_mean) {
      for (int it = threadIdx.x; it < Ec * Dp; it += FT) {
        const int e = it / Dp, o = it - e * Dp, d = o / K1, j = o - d * K1;
        const double *hm_e = hm + (size_t)e * nq * K1;
        double acc = 0.0;
        for (int p = 0; p < P; ++p) {
          acc = fma(hm_e[point_of(chained, p, 0) * K1 + j], RS[(size_t)(2 * d) * S + e * P + p], acc);
          acc = fma(hm_e[point_of(chained, p, 1) * K1 + j], RS[(size_t)(2 * d + 1) * S + e * P + p], acc);
        }
        grad_mean[(b0 + e) * Dp + o] = (float)acc;
      }
    }
    if (!grad_L && !dsigma_part) continue;
    // ---- phase 3b: thread per (episode, DoF block): dSigma_dd'[i][j] = sum_q w_q[i] h_q[j] ------------------------
    double acc[K1 * K1];
    const bool own = (int)threadIdx.x < Ec * NB;
    int be = 0, bblk = 0, bd = 0, bdd = 0;
    if (own) {
      be = threadIdx.x / NB; bblk = threadIdx.x - be * NB;
      tri_decode(bblk, bd, bdd);
      dsigma_block<K1>(CS + be * P, S, hm + (size_t)be * nq * K1, bd, bdd, chained, nq, P, acc);
    }
    __syncthreads();                                         // all adjoints consumed: the C region is free
    if (shared_cov) {
      // sum over the episodes of this iteration: scratch [E][NB][K1*K1] fp32 in the C region, then += dsacc
      float *scr = dSb;
      if (own) {
#pragma unroll
        for (int t = 0; t < K1 * K1; ++t) scr[((size_t)be * NB + bblk) * (K1 * K1) + t] = (float)acc[t];
      }
      __syncthreads();
      for (int e2 = threadIdx.x; e2 < NB * K1 * K1; e2 += FT) {
        float s = dsacc[e2];
        for (int e = 0; e < Ec; ++e) s += scr[(size_t)e * NB * K1 * K1 + e2];
        dsacc[e2] = s;
      }
    } else {
      // dSigma_b dense symmetric fp32 [E][NR4][LD] in the C region, then grad_L_b = 2 tril(dSigma_b L_b) per episode
      for (int e0 = threadIdx.x; e0 < Ec * NR4 * LD; e0 += FT) dSb[e0] = 0.f;
      __syncthreads();
      if (own) {
        float *M = dSb + (size_t)be * NR4 * LD;
#pragma unroll
        for (int i = 0; i < K1; ++i)
#pragma unroll
          for (int j = 0; j < K1; ++j) {
            const float v = (float)acc[i * K1 + j];
            M[(bd * K1 + i) * LD + bdd * K1 + j] = v;
            if (bd != bdd) M[(bdd * K1 + j) * LD + bd * K1 + i] = v;
          }
      }
      for (int e = 0; e < Ec; ++e) {
        __syncthreads();
        stage_lower(L + (b0 + e) * ldb_L, Ls, Dp, NR4, LD);
        __syncthreads();
        const float *M = dSb + (size_t)e * NR4 * LD;
        float *gL = grad_L + (size_t)(b0 + e) * Dp * Dp;
        constexpr int NT4 = NR4 / 4;
        if (threadIdx.x < tri(NT4)) {
          int I, J;
          tri_decode(threadIdx.x, I, J);
          const float *mrow = M + (4 * I) * LD;
          float c4[4][4];
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) c4[x][y] = 0.f;
          for (int k = NR4 - 1; k >= 4 * J; --k) {           // L[k][c] = 0 for k < c
            float mv[4], lv[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { mv[x] = mrow[x * LD + k]; lv[x] = Ls[k * LD + 4 * J + x]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) c4[x][y] = fmaf(mv[x], lv[y], c4[x][y]);
          }
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              const int i = 4 * I + x, j = 4 * J + y;
              if (i < Dp && j < Dp) gL[(size_t)i * Dp + j] = j <= i ? 2.f * c4[x][y] : 0.f;
            }
        } else if (NT4 * 4 > 0) {
          // threads without a tile zero the strictly-upper tiles (grad_L is dense [Dp, Dp], upper = 0)
          for (int t = threadIdx.x - tri(NT4); t < NT4 * NT4; t += FT - tri(NT4)) {
            const int I = t / NT4, J = t - I * NT4;
            if (J > I)
              for (int x = 0; x < 4; ++x)
                for (int y = 0; y < 4; ++y) {
                  const int i = 4 * I + x, j = 4 * J + y;
                  if (i < Dp && j < Dp) gL[(size_t)i * Dp + j] = 0.f;
                }
          }
        }
      }
    }
  }
  // ---- epilogue --------------------------------------------------------------------------------------------
  __syncthreads();
  if (shared_cov && want_grad && dsigma_part) {
    float *out = dsigma_part + (size_t)blockIdx.x * NB * K1 * K1;
    for (int e = threadIdx.x; e < NB * K1 * K1; e += FT) out[e] = dsacc[e];
  }
  if (grad_mode == 2 && loss_acc) {
    loss_part = warp_sum(loss_part);
    ratio_part = warp_sum(ratio_part);
    if (lane == 0) { s_red[2 * warp] = loss_part; s_red[2 * warp + 1] = ratio_part; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, r = 0.0;
      for (int w = 0; w < FT / 32; ++w) { a += s_red[2 * w]; r += s_red[2 * w + 1]; }
      if (r != 0.0) { atomicAdd(loss_acc, a); atomicAdd(loss_acc + 1, r); }
    }
  }
}

// =====================================================================================================
// reduce of the per-CTA dSigma partials (block layout) + grad_L = 2 tril(dSigma L) for ONE shared factor
// =====================================================================================================
// One CTA.  part [nparts][NB][K1][K1] fp32 -> dS dense symmetric fp32 in smem -> grad_L [Dp, Dp] (lower, upper 0);
// optionally also grad_sigma [Dp, Dp] (dense symmetric).  Fixed summation order: deterministic.
constexpr int RT = 1024;
template <int D, int K1>
__global__ void __launch_bounds__(RT)
dsigma_reduce_kernel(const float *__restrict__ part, int nparts, const float *__restrict__ L,
                     const float *__restrict__ upstream, float *__restrict__ grad_L, float *__restrict__ grad_sigma) {
  using FL = FusedLayout<D, K1>;
  constexpr int Dp = FL::Dp, NB = FL::NB, NR4 = FL::NR4, LD = FL::LD, NE = NB * K1 * K1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *M = reinterpret_cast<float *>(smem_raw);          // [NR4][LD]
  float *Ls = M + NR4 * LD;                                // [NR4][LD]
  const float up = upstream ? *upstream : 1.0f;
  for (int e = threadIdx.x; e < NR4 * LD; e += RT) M[e] = 0.f;
  for (int e0 = threadIdx.x; e0 < NR4 * NR4; e0 += RT) {
    const int i = e0 / NR4, c = e0 - i * NR4;
    Ls[i * LD + c] = (L && i < Dp && c <= i) ? L[(size_t)i * Dp + c] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NE; e += RT) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 3 < nparts; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] += part[(size_t)(k + u) * NE + e];
    }
    for (; k < nparts; ++k) a[0] += part[(size_t)k * NE + e];
    const float v = up * ((a[0] + a[1]) + (a[2] + a[3]));
    const int blk = e / (K1 * K1), r = e - blk * K1 * K1, i = r / K1, j = r - i * K1;
    int d, dd;
    tri_decode(blk, d, dd);
    M[(d * K1 + i) * LD + dd * K1 + j] = v;
    if (d != dd) M[(dd * K1 + j) * LD + d * K1 + i] = v;
  }
  __syncthreads();
  if (grad_sigma)
    for (int e = threadIdx.x; e < Dp * Dp; e += RT) grad_sigma[e] = M[(e / Dp) * LD + (e % Dp)];
  if (!grad_L) return;
  // 2 tril(M L): 2x2 register tiles over the lower triangle of NR4 x NR4 (tri(32) = 528 tiles <= 1024 threads)
  constexpr int NT2 = NR4 / 2;
  for (int t = threadIdx.x; t < tri(NT2); t += RT) {
    int I, J;
    tri_decode(t, I, J);
    const float *m0 = M + (2 * I) * LD, *m1 = m0 + LD;
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
    for (int k = 2 * J; k < NR4; ++k) {
      const float a0 = m0[k], a1 = m1[k], l0 = Ls[k * LD + 2 * J], l1 = Ls[k * LD + 2 * J + 1];
      c00 = fmaf(a0, l0, c00); c01 = fmaf(a0, l1, c01);
      c10 = fmaf(a1, l0, c10); c11 = fmaf(a1, l1, c11);
    }
    const int i0 = 2 * I, j0 = 2 * J;
    if (i0 < Dp && j0 < Dp) grad_L[(size_t)i0 * Dp + j0] = j0 <= i0 ? 2.f * c00 : 0.f;
    if (i0 < Dp && j0 + 1 < Dp) grad_L[(size_t)i0 * Dp + j0 + 1] = j0 + 1 <= i0 ? 2.f * c01 : 0.f;
    if (i0 + 1 < Dp && j0 < Dp) grad_L[(size_t)(i0 + 1) * Dp + j0] = 2.f * c10;
    if (i0 + 1 < Dp && j0 + 1 < Dp) grad_L[(size_t)(i0 + 1) * Dp + j0 + 1] = 2.f * c11;
  }
  for (int e = threadIdx.x; e < Dp * Dp; e += RT) {        // strictly upper part = 0
    const int i = e / Dp, j = e - i * Dp;
    if (j > i && (j / 2 > i / 2)) grad_L[e] = 0.f;
  }
}

// =====================================================================================================
// uniform time grid + shared covariance
// =====================================================================================================
// ws (doubles): hm [nq][K1] | xi [nq][2] | C / G [P][NT] | X = S^-1 (lower, packed, true diagonal) [P][NT] |
//               Cinv [P][NT] | hld [P] (half log-determinant) | flags {chained, nq}
template <int D, int K1>
struct UniLayout {
  static constexpr int N = 2 * D, NT = tri(N);
  size_t hm, xi, cg, xinv, cinv, hld, flags, total;
  __host__ __device__ UniLayout(int P) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n + 1) & ~(size_t)1; return at; };
    hm = take((size_t)2 * P * K1); xi = take((size_t)4 * P); cg = take((size_t)P * NT); xinv = take((size_t)P * NT);
    cinv = take((size_t)P * NT); hld = take(P); flags = take(4); total = o;   // flags: chained, nq, ticket (u32)
  }
};

// phase 1 (what & 1): basis rows of the common grid, C_p (no regulariser), max diag -> atomic max into *diag_max
// phase 2 (what & 2): per segment (one warp each, factor in shared memory): S, X = S^-1, C^-1, half logdet
constexpr int UT = 1024;
template <int D, int K1, bool SIGMA_IN>
__global__ void __launch_bounds__(UT)
uniform_prep_kernel(TabDev tb, const float *__restrict__ L, const double *__restrict__ Sigma0,
                    const double *__restrict__ sigma_scale, const float *__restrict__ times,
                    const float *__restrict__ init_time, const int64_t *__restrict__ pairs, double *__restrict__ ws,
                    double *__restrict__ diag_max, double reg_rel, int P, int what) {
  using FL = FusedLayout<D, K1>;
  using UL = UniLayout<D, K1>;
  constexpr int Dp = FL::Dp, N = FL::N, NT = FL::NT, NB = FL::NB, KP = FL::KP, NR4 = FL::NR4, LD = FL::LD, IR = FL::IR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const UL ul(P);
  double *Sblk = reinterpret_cast<double *>(smem_raw);
  float *Ls = reinterpret_cast<float *>(smem_raw);
  const size_t sb = sizeof(double) * NB * K1 * KP, ls = sizeof(float) * NR4 * LD;
  double *irow = reinterpret_cast<double *>(smem_raw + (((sb > ls ? sb : ls) + 15) & ~(size_t)15));   // [IR]
  double *fac = irow + IR;                                 // [warps][NT + N] per-warp factor scratch
  __shared__ double s_max[UT / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chained = block_chained(pairs, P);
  const int nq = chained ? P + 1 : 2 * P;
  double *g_hm = ws + ul.hm, *g_xi = ws + ul.xi, *g_c = ws + ul.cg;
  if (what & 1) {
    // ---- Sigma blocks ----
    if (SIGMA_IN) {
      const double sc = sigma_scale ? *sigma_scale : 1.0;
      for (int e = threadIdx.x; e < Dp * Dp; e += UT) {
        const int i = e / Dp, j = e - i * Dp;
        if (j <= i) sblk_store<K1, KP>(Sblk, i, j, sc * Sigma0[e]);
      }
      __syncthreads();
    } else {
      for (int e = threadIdx.x; e < NR4 * NR4; e += UT) {
        const int i = e / NR4, c = e - i * NR4;
        Ls[i * LD + c] = (i < Dp && c <= i) ? L[(size_t)i * Dp + c] : 0.f;
      }
      __syncthreads();
      // Sigma = L L^T, entry per thread (fp32 accumulation as in the per-episode path), lower triangle
      float vals[(Dp * (Dp + 1) / 2 + UT - 1) / UT];
      int cnt = 0;
      for (int t = threadIdx.x; t < tri(Dp); t += UT, ++cnt) {
        int i, j;
        tri_decode(t, i, j);
        const float *a = Ls + i * LD, *bq = Ls + j * LD;
        float acc = 0.f;
        for (int k = 0; k <= j; ++k) acc = fmaf(a[k], bq[k], acc);
        vals[cnt] = acc;
      }
      __syncthreads();
      cnt = 0;
      for (int t = threadIdx.x; t < tri(Dp); t += UT, ++cnt) {
        int i, j;
        tri_decode(t, i, j);
        sblk_store<K1, KP>(Sblk, i, j, (double)vals[cnt]);
      }
      __syncthreads();
    }
    // ---- basis rows of the common grid (episode 0) ----
    if (threadIdx.x <= K1) init_row_entry<K1>(tb, (double)init_time[0], threadIdx.x, irow);
    __syncthreads();
    for (int it = threadIdx.x; it < nq * (K1 + 1); it += UT) {
      const int q = it / (K1 + 1), j = it - q * (K1 + 1);
      basis_entry<K1>(tb, irow, (double)times[point_time_index(pairs, chained, q, P)], j, g_hm + (size_t)q * K1,
                      g_xi + (size_t)q * 2);
    }
    __threadfence_block();
    __syncthreads();
    // ---- gram: C [P][NT] straight to the workspace (entry-major per pair: S = 1 layout with slot stride NT) ----
    double my_max = 0.0;
    for (int it = threadIdx.x; it < NB * nq; it += UT) {
      const int blk = it / nq, q = it - blk * nq;
      int d, dd;
      tri_decode(blk, d, dd);
      // CS[(entry) * S + slot] with S = 1 and one "slot" per pair at offset pair * NT: pass base pointers per pair
      const Links lk = links(chained, q, P);
      // gram_item writes CS[entry * S + slot0 + pair]; emulate [P][NT] by S = 0 stride trick: write manually
      double h[K1], v[K1];
      const double *Sb = sblk_at<K1, KP>(Sblk, blk), *hq = g_hm + (size_t)q * K1;
#pragma unroll
      for (int j = 0; j < K1; ++j) h[j] = hq[j];
#pragma unroll
      for (int i = 0; i < K1; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < K1; ++j) a = fma(Sb[i * KP + j], h[j], a);
        v[i] = a;
      }
      double kuu = 0.0;
#pragma unroll
      for (int i = 0; i < K1; ++i) kuu = fma(h[i], v[i], kuu);
      const int r0 = 2 * d, q0 = 2 * dd;
      if (lk.pf >= 0) {
        const double *hn = g_hm + (size_t)lk.nxt * K1;
        double kup = 0.0;
#pragma unroll
        for (int i = 0; i < K1; ++i) kup = fma(hn[i], v[i], kup);
        g_c[(size_t)lk.pf * NT + tri_idx(r0, q0)] = kuu;
        g_c[(size_t)lk.pf * NT + tri_idx(r0 + 1, q0)] = kup;
      }
      if (lk.ps >= 0) {
        g_c[(size_t)lk.ps * NT + tri_idx(r0 + 1, q0 + 1)] = kuu;
        if (d != dd) {
          const double *hp = g_hm + (size_t)lk.prv * K1;
          double kum = 0.0;
#pragma unroll
          for (int i = 0; i < K1; ++i) kum = fma(hp[i], v[i], kum);
          g_c[(size_t)lk.ps * NT + tri_idx(r0, q0 + 1)] = kum;
        }
      }
      if (d == dd) my_max = fmax(my_max, kuu);
    }
    my_max = warp_max(my_max);
    if (lane == 0) s_max[warp] = my_max;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int w = 0; w < UT / 32; ++w) m = fmax(m, s_max[w]);
      atomic_max_pos_double(diag_max, m);
      ws[ul.flags] = (double)chained;
      ws[ul.flags + 1] = (double)nq;
      *reinterpret_cast<unsigned int *>(ws + ul.flags + 2) = 0u;      // ticket of uniform_finish_kernel
    }
    __threadfence();
    __syncthreads();
  }
  if (!(what & 2)) return;
  // ---- per-segment factorisation: warp per segment, matrix in shared memory (latency matters here, not throughput)
  const double reg = reg_rel * (*reinterpret_cast<volatile double *>(diag_max));
  for (int p = warp; p < P; p += UT / 32) {
    double *A = fac + (size_t)warp * (2 * NT + N);          // packed lower C -> S ; X after it
    double *X = A + NT, *dinv = X + NT;
    for (int t = lane; t < NT; t += 32) A[t] = g_c[(size_t)p * NT + t];
    __syncwarp();
    double hld = 0.0;
    for (int j = 0; j < N; ++j) {
      // pivot
      double dj = A[tri_idx(j, j)] + reg;
      for (int k = 0; k < j; ++k) dj = fma(-A[tri_idx(j, k)], A[tri_idx(j, k)], dj);
      const double inv = 1.0 / sqrt(dj);
      hld += 0.5 * log(dj);
      __syncwarp();
      if (lane == 0) { A[tri_idx(j, j)] = dj * inv; dinv[j] = inv; }
      const int i = j + 1 + lane;
      if (i < N) {
        double v = A[tri_idx(i, j)];
        for (int k = 0; k < j; ++k) v = fma(-A[tri_idx(i, k)], A[tri_idx(j, k)], v);
        A[tri_idx(i, j)] = v * inv;
      }
      __syncwarp();
    }
    // X = S^-1: lane j solves column j by forward substitution
    if (lane < N) {
      const int j = lane;
      for (int i = j; i < N; ++i) {
        double v = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) v = fma(-A[tri_idx(i, k)], X[tri_idx(k, j)], v);
        X[tri_idx(i, j)] = v * dinv[i];
      }
    }
    __syncwarp();
    // C^-1 = X^T X (lower), X and hld to the workspace
    for (int t = lane; t < NT; t += 32) {
      int i, j;
      tri_decode(t, i, j);
      double v = 0.0;
      for (int k = i; k < N; ++k) v = fma(X[tri_idx(k, i)], X[tri_idx(k, j)], v);
      ws[ul.cinv + (size_t)p * NT + t] = v;
      ws[ul.xinv + (size_t)p * NT + t] = X[t];
    }
    if (lane == 0) ws[ul.hld + p] = hld;
    __syncwarp();
  }
}

// main: thread per (episode, pair); a warp works on ONE pair (X_p is read by broadcast) for 32 episodes.
// Outputs: logp, info (0), grad_mean (accumulated over the pairs of an episode through shared memory),
// and for the batch reduction: ga [P][N][Bpad] = g alpha, al [P][N][Bpad] = alpha, gs [P][Bpad] = g (fp64).
constexpr int UM_EP = 32;                // episodes per CTA
constexpr int UM_THREADS = 512;          // 16 warps, each loops over pairs p = warp, warp + 16, ...
template <int D, int K1>
__global__ void __launch_bounds__(UM_THREADS)
uniform_main_kernel(TabDev tb, const double *__restrict__ ws, const float *__restrict__ smp_traj,
                    const float *__restrict__ mean, const float *__restrict__ init_pos,
                    const float *__restrict__ init_vel, const int64_t *__restrict__ pairs, int grad_mode,
                    const float *__restrict__ grad_logp, const float *__restrict__ logp_old,
                    const float *__restrict__ advantage, double grad_scale, double *__restrict__ loss_acc,
                    float *__restrict__ logp, int32_t *__restrict__ info, float *__restrict__ grad_mean,
                    double *red, long long B, long long Bpad, int T, int P) {
  using UL = UniLayout<D, K1>;
  constexpr int Dp = D * K1, N = 2 * D, NT = tri(N);
  const UL ul(P);
  __shared__ double s_red[2 * (UM_THREADS / 32)];
  const int chained = (int)ws[ul.flags];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = UM_THREADS / 32;
  const long long b = (long long)blockIdx.x * UM_EP + lane;
  const bool want_grad = grad_mode != 0;
  const double tau = tb.tau;
  const double *g_hm = ws + ul.hm, *g_xi = ws + ul.xi;
  double loss_part = 0.0, ratio_part = 0.0;
  for (int p = warp; p < P; p += nwarps) {
    if (b < B) {
      const double *X = ws + ul.xinv + (size_t)p * NT;
      double r[N], z[N];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double y0 = (double)init_pos[b * D + d], v0 = (double)init_vel[b * D + d] * tau;
        const float *th = mean + b * Dp + d * K1;
        double thd[K1];
#pragma unroll
        for (int j = 0; j < K1; ++j) thd[j] = (double)th[j];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = point_of(chained, p, k);
          const double *hq = g_hm + (size_t)q * K1;
          double mu = g_xi[2 * q] * y0 + g_xi[2 * q + 1] * v0;
#pragma unroll
          for (int j = 0; j < K1; ++j) mu = fma(hq[j], thd[j], mu);
          if (tb.relative_goal) {
            const double shift = tb.relative_goal_scaled ? y0 : y0 / tb.scale[K1 - 1];
            mu = fma(hq[K1 - 1], shift, mu);
          }
          r[2 * d + k] = (double)smp_traj[(b * T + pairs[2 * p + k]) * (2 * D) + d] - mu;
        }
      }
      double maha = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {                         // z = X r  (X = S^-1, lower)
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) v = fma(X[i * (i + 1) / 2 + j], r[j], v);
        z[i] = v;
        maha = fma(v, v, maha);
      }
      const double lp = -0.5 * ((double)N * LN_2PI + maha) - ws[ul.hld + p];
      const long long gid = b * P + p;
      if (logp) logp[gid] = (float)lp;
      if (info) info[gid] = 0;
      if (want_grad) {
        double g;
        if (grad_mode == 2) {
          const double ratio = exp(lp - (double)logp_old[gid]);
          g = -ratio * (double)advantage[gid] * grad_scale;
          loss_part += g;
          ratio_part += ratio * grad_scale;
        } else {
          g = (double)grad_logp[gid];
        }
        double *ga = red + ((size_t)p * (2 * N + 1)) * Bpad;     // [N] g alpha | [N] alpha | g
#pragma unroll
        for (int j = 0; j < N; ++j) {                       // alpha = X^T z
          double v = 0.0;
#pragma unroll
          for (int i = j; i < N; ++i) v = fma(X[i * (i + 1) / 2 + j], z[i], v);
          ga[(size_t)j * Bpad + b] = g * v;
          ga[(size_t)(N + j) * Bpad + b] = v;
        }
        ga[(size_t)(2 * N) * Bpad + b] = g;
      }
    }
  }
  if (want_grad && grad_mean) {
    // grad_mean[b][d K1 + j] = sum_{p, k} h_{p,k}[j] (g alpha)_bp[(d, k)], from the rows this CTA has just written
    __syncthreads();
    for (int it = threadIdx.x; it < UM_EP * Dp; it += UM_THREADS) {
      const int o = it / UM_EP, l = it - o * UM_EP, d = o / K1, j = o - d * K1;
      const long long bb = (long long)blockIdx.x * UM_EP + l;
      if (bb >= B) continue;
      double acc = 0.0;
      for (int p = 0; p < P; ++p) {
        const double *ga = red + ((size_t)p * (2 * N + 1)) * Bpad;
        acc = fma(g_hm[(size_t)point_of(chained, p, 0) * K1 + j], ga[(size_t)(2 * d) * Bpad + bb], acc);
        acc = fma(g_hm[(size_t)point_of(chained, p, 1) * K1 + j], ga[(size_t)(2 * d + 1) * Bpad + bb], acc);
      }
      grad_mean[bb * Dp + o] = (float)acc;
    }
  }
  if (grad_mode == 2 && loss_acc) {
    loss_part = warp_sum(loss_part);
    ratio_part = warp_sum(ratio_part);
    if (lane == 0) { s_red[2 * warp] = loss_part; s_red[2 * warp + 1] = ratio_part; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, rr = 0.0;
      for (int w = 0; w < nwarps; ++w) { a += s_red[2 * w]; rr += s_red[2 * w + 1]; }
      if (rr != 0.0) { atomicAdd(loss_acc, a); atomicAdd(loss_acc + 1, rr); }
    }
  }
}

// finish: CTA p forms G_p = 1/2 (sum_b g alpha alpha^T - (sum_b g) C_p^-1) (fixed order); the LAST CTA to finish
// (atomic ticket) turns the P adjoints into dSigma (block layout, one "partial") for dsigma_reduce_kernel.
template <int D, int K1>
__global__ void __launch_bounds__(512)
uniform_finish_kernel(double *__restrict__ ws, const double *__restrict__ red, long long B, long long Bpad, int P,
                      float *__restrict__ dsigma_part) {
  using UL = UniLayout<D, K1>;
  constexpr int N = 2 * D, NT = tri(N), NB = tri(D);
  const UL ul(P);
  __shared__ double s_part[4][NT + 1];
  __shared__ int s_last;
  unsigned int *ticket = reinterpret_cast<unsigned int *>(ws + ul.flags + 2);   // zeroed by uniform_prep_kernel
  const int p = blockIdx.x;
  const double *ga = red + ((size_t)p * (2 * N + 1)) * Bpad, *al = ga + (size_t)N * Bpad, *gs = ga + (size_t)(2 * N) * Bpad;
  // entry t (+ the g sum as entry NT), 4-way split over the batch
  const int t = threadIdx.x & 127, part = threadIdx.x >> 7;
  if (t <= NT) {
    int i = 0, j = 0;
    if (t < NT) tri_decode(t, i, j);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const long long chunk = (B + 3) / 4, lo = part * chunk, hi = (lo + chunk < B) ? lo + chunk : B;
    const double *x = t < NT ? ga + (size_t)i * Bpad : gs, *y = al + (size_t)j * Bpad;
    long long b = lo;
    if (t < NT) {
      for (; b + 3 < hi; b += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(x[b + u], y[b + u], acc[u]);
      }
      for (; b < hi; ++b) acc[0] = fma(x[b], y[b], acc[0]);
    } else {
      for (; b < hi; ++b) acc[0] += x[b];
    }
    s_part[part][t] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  __syncthreads();
  if (threadIdx.x < NT) {
    const double a = (s_part[0][t] + s_part[1][t]) + (s_part[2][t] + s_part[3][t]);
    const double g = (s_part[0][NT] + s_part[1][NT]) + (s_part[2][NT] + s_part[3][NT]);
    ws[ul.cg + (size_t)p * NT + t] = 0.5 * (a - g * ws[ul.cinv + (size_t)p * NT + t]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- dSigma blocks from the P adjoints: task (blk, i): row i of the K1 x K1 block
  const int chained = (int)ws[ul.flags], nq = (int)ws[ul.flags + 1];
  const double *g_hm = ws + ul.hm;
  const volatile double *G = ws + ul.cg;
  for (int task = threadIdx.x; task < NB * K1; task += blockDim.x) {
    const int blk = task / K1, i = task - blk * K1;
    int d, dd;
    tri_decode(blk, d, dd);
    const int r0 = 2 * d, q0 = 2 * dd;
    const int e00 = tri_idx(r0, q0), e10 = tri_idx(r0 + 1, q0), e11 = tri_idx(r0 + 1, q0 + 1);
    const int e01 = (d != dd) ? tri_idx(r0, q0 + 1) : e10;
    double acc[K1];
#pragma unroll
    for (int j = 0; j < K1; ++j) acc[j] = 0.0;
    for (int q = 0; q < nq; ++q) {
      const Links lk = links(chained, q, P);
      double w = 0.0;
      const double hqi = g_hm[(size_t)q * K1 + i];
      if (lk.pf >= 0) w += G[(size_t)lk.pf * NT + e00] * hqi + G[(size_t)lk.pf * NT + e10] * g_hm[(size_t)lk.nxt * K1 + i];
      if (lk.ps >= 0) w += G[(size_t)lk.ps * NT + e11] * hqi + G[(size_t)lk.ps * NT + e01] * g_hm[(size_t)lk.prv * K1 + i];
#pragma unroll
      for (int j = 0; j < K1; ++j) acc[j] = fma(w, g_hm[(size_t)q * K1 + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < K1; ++j) dsigma_part[((size_t)blk * K1 + i) * K1 + j] = (float)acc[j];
  }
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------
#define TCE_FOR_SHAPES(X) X(7, 9) X(4, 9) X(7, 4) X(3, 4) X(2, 3)

namespace {
constexpr size_t SMEM_LIMIT = 227 * 1024;

template <int D, int K1>
int fused_pick_E(int P, int chained_worst) {
  int E = FT / P;
  if (E > FT / tri(D)) E = FT / tri(D);              // phase 3b: one thread per (episode, DoF block)
  while (E > 1 && FusedLayout<D, K1>(E, P, chained_worst).total > SMEM_LIMIT) --E;
  if (E < 1 || FusedLayout<D, K1>(E, P, chained_worst).total > SMEM_LIMIT) return 0;
  return E;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
}  // namespace

extern "C" int tce_seglik_fused_config(const tce_tables_t *t, int64_t B, int64_t P, int chained, int32_t *E_out,
                                       int32_t *grid_out, int64_t *part_floats) {
  if (!t || B < 0 || P < 1) return TCE_ERR_INVALID_ARGUMENT;
#define X(Dv, Kv)                                                                              \
  if (t->D == Dv && t->K1 == Kv) {                                                             \
    if (P > FT) return TCE_ERR_UNSUPPORTED_SHAPE;                                              \
    const int E = fused_pick_E<Dv, Kv>((int)P, chained);                                       \
    if (!E) return TCE_ERR_UNSUPPORTED_SHAPE;                                                  \
    const int64_t groups = (B + E - 1) / E;                                                    \
    int grid = (int)(groups < sm_count() ? groups : sm_count());                               \
    if (grid < 1) grid = 1;                                                                    \
    if (E_out) *E_out = E;                                                                     \
    if (grid_out) *grid_out = grid;                                                            \
    if (part_floats) *part_floats = (int64_t)tri(Dv) * Kv * Kv;                                \
    return TCE_OK;                                                                             \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

template <bool SIGMA_IN>
static int diagmax_launch(const tce_tables_t *t, const float *L, int64_t ldb_L, const double *Sigma0,
                          const double *sigma_scale, const float *times, const float *init_time,
                          const int64_t *pred_pairs, double *diag_max, int64_t B, int64_t T, int64_t P, void *stream) {
  if (B == 0) return TCE_OK;
  if (!t || (SIGMA_IN ? !Sigma0 : !L) || !times || !init_time || !pred_pairs || !diag_max || B < 0 || T < 1 || P < 1)
    return TCE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
#define X(Dv, Kv)                                                                                                 \
  if (t->D == Dv && t->K1 == Kv) {                                                                                \
    using FL = FusedLayout<Dv, Kv>;                                                                               \
    const size_t smem = sizeof(double) * ((size_t)Dv * Kv * Kv + 2 * P * Kv + FL::IR) +                           \
                        sizeof(float) * FL::NR4 * FL::LD + 32;                                                    \
    if (smem > SMEM_LIMIT) return TCE_ERR_UNSUPPORTED_SHAPE;                                                      \
    auto kern = seglik_diagmax_kernel<Dv, Kv, SIGMA_IN>;                                                          \
    TCE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "diagmax smem"); \
    const bool shared_cov = SIGMA_IN || ldb_L == 0;                                                               \
    int64_t grid = shared_cov ? (B + 3) / 4 : B;                                                                  \
    const int64_t cap = (int64_t)sm_count() * 4;                                                                  \
    if (grid > cap) grid = cap;                                                                                   \
    if (grid < 1) grid = 1;                                                                                       \
    kern<<<(unsigned)grid, FT, smem, st>>>(tab_dev(t), L, ldb_L, Sigma0, sigma_scale, times, init_time, pred_pairs, \
                                           diag_max, (long long)B, (int)T, (int)P);                               \
    TCE_CHECK_LAUNCH("seglik_diagmax_kernel");                                                                    \
    return TCE_OK;                                                                                                \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_diagmax(const tce_tables_t *t, const float *L, int64_t ldb_L, const double *Sigma0,
                                  const double *sigma_scale, const float *times, const float *init_time,
                                  const int64_t *pred_pairs, double *diag_max, int64_t B, int64_t T, int64_t P,
                                  void *stream) {
  if (Sigma0)
    return diagmax_launch<true>(t, nullptr, 0, Sigma0, sigma_scale, times, init_time, pred_pairs, diag_max, B, T, P,
                                stream);
  return diagmax_launch<false>(t, L, ldb_L, nullptr, nullptr, times, init_time, pred_pairs, diag_max, B, T, P, stream);
}

template <bool SIGMA_IN>
static int fused_launch(const tce_tables_t *t, const float *smp_traj, const float *mean, const float *L, int64_t ldb_L,
                        const double *Sigma0, const double *sigma_scale, const float *times, const float *init_time,
                        const float *init_pos, const float *init_vel, const int64_t *pred_pairs,
                        const double *diag_max, double reg_rel, int grad_mode, const float *grad_logp,
                        const float *logp_old, const float *advantage, double grad_scale, double *loss_acc,
                        float *logp, int32_t *info, float *grad_mean, float *grad_L, float *dsigma_part, int chained,
                        int64_t B, int64_t T, int64_t P, void *stream) {
  if (B == 0) return TCE_OK;
  if (!t || !smp_traj || !mean || (SIGMA_IN ? !Sigma0 : !L) || !times || !init_time || !init_pos || !init_vel ||
      !pred_pairs || !diag_max || B < 0 || T < 1 || P < 1 || grad_mode < 0 || grad_mode > 2)
    return TCE_ERR_INVALID_ARGUMENT;
  if (grad_mode == 1 && !grad_logp) return TCE_ERR_INVALID_ARGUMENT;
  if (grad_mode == 2 && (!logp_old || !advantage)) return TCE_ERR_INVALID_ARGUMENT;
  const bool shared_cov = SIGMA_IN || ldb_L == 0;
  if (grad_mode && !shared_cov && dsigma_part) return TCE_ERR_INVALID_ARGUMENT;
  if (grad_mode && shared_cov && grad_L) return TCE_ERR_INVALID_ARGUMENT;   /* shared: partials + tce_seglik_dsigma_reduce */
  cudaStream_t st = (cudaStream_t)stream;
  int32_t E = 0, grid = 0;
  int rc = tce_seglik_fused_config(t, B, P, chained, &E, &grid, nullptr);
  if (rc != TCE_OK) return rc;
#define X(Dv, Kv)                                                                                                  \
  if (t->D == Dv && t->K1 == Kv) {                                                                                 \
    const size_t smem = FusedLayout<Dv, Kv>(E, (int)P, chained).total;                                             \
    auto kern = seglik_fused_kernel<Dv, Kv, SIGMA_IN>;                                                             \
    TCE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fused smem");    \
    kern<<<(unsigned)grid, FT, smem, st>>>(tab_dev(t), smp_traj, mean, L, ldb_L, Sigma0, sigma_scale, times,       \
                                           init_time, init_pos, init_vel, pred_pairs, diag_max, reg_rel, grad_mode, \
                                           grad_logp, logp_old, advantage, grad_scale, loss_acc, logp, info,       \
                                           grad_mean, grad_L, dsigma_part, (long long)B, (int)T, (int)P, (int)E,   \
                                           chained);                                                                \
    TCE_CHECK_LAUNCH("seglik_fused_kernel");                                                                       \
    return TCE_OK;                                                                                                 \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_fused(const tce_tables_t *t, const float *smp_traj, const float *mean, const float *L,
                                int64_t ldb_L, const double *Sigma0, const double *sigma_scale, const float *times,
                                const float *init_time, const float *init_pos, const float *init_vel,
                                const int64_t *pred_pairs, const double *diag_max, double reg_rel, int grad_mode,
                                const float *grad_logp, const float *logp_old, const float *advantage,
                                double grad_scale, double *loss_acc, float *logp, int32_t *info, float *grad_mean,
                                float *grad_L, float *dsigma_part, int chained, int64_t B, int64_t T, int64_t P,
                                void *stream) {
  if (Sigma0)
    return fused_launch<true>(t, smp_traj, mean, nullptr, 0, Sigma0, sigma_scale, times, init_time, init_pos, init_vel,
                              pred_pairs, diag_max, reg_rel, grad_mode, grad_logp, logp_old, advantage, grad_scale,
                              loss_acc, logp, info, grad_mean, grad_L, dsigma_part, chained, B, T, P, stream);
  return fused_launch<false>(t, smp_traj, mean, L, ldb_L, nullptr, nullptr, times, init_time, init_pos, init_vel,
                             pred_pairs, diag_max, reg_rel, grad_mode, grad_logp, logp_old, advantage, grad_scale,
                             loss_acc, logp, info, grad_mean, grad_L, dsigma_part, chained, B, T, P, stream);
}

extern "C" int tce_seglik_dsigma_reduce(const tce_tables_t *t, const float *dsigma_part, int nparts, const float *L,
                                        const float *upstream, float *grad_L, float *grad_sigma, void *stream) {
  if (!t || !dsigma_part || nparts < 1 || (!grad_L && !grad_sigma) || (grad_L && !L)) return TCE_ERR_INVALID_ARGUMENT;
#define X(Dv, Kv)                                                                                               \
  if (t->D == Dv && t->K1 == Kv) {                                                                              \
    using FL = FusedLayout<Dv, Kv>;                                                                             \
    const size_t smem = sizeof(float) * 2 * FL::NR4 * FL::LD + 16;                                              \
    auto kern = dsigma_reduce_kernel<Dv, Kv>;                                                                   \
    TCE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "reduce smem"); \
    kern<<<1, RT, smem, (cudaStream_t)stream>>>(dsigma_part, nparts, L, upstream, grad_L, grad_sigma);          \
    TCE_CHECK_LAUNCH("dsigma_reduce_kernel");                                                                   \
    return TCE_OK;                                                                                              \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

/* ---- uniform time grid + shared covariance ---------------------------------------------------------------- */
extern "C" size_t tce_seglik_uniform_ws_doubles(const tce_tables_t *t, int64_t P) {
  if (!t || P < 1) return 0;
#define X(Dv, Kv) if (t->D == Dv && t->K1 == Kv) return UniLayout<Dv, Kv>((int)P).total;
  TCE_FOR_SHAPES(X)
#undef X
  return 0;
}

extern "C" size_t tce_seglik_uniform_red_doubles(const tce_tables_t *t, int64_t B, int64_t P) {
  if (!t || P < 1 || B < 0) return 0;
  const size_t Bpad = ((size_t)B + 31) & ~(size_t)31;
  return (size_t)P * (4 * (size_t)t->D + 1) * Bpad;
}

extern "C" int tce_seglik_uniform_prep(const tce_tables_t *t, const float *L, const double *Sigma0,
                                       const double *sigma_scale, const float *times, const float *init_time,
                                       const int64_t *pred_pairs, double *ws, double *diag_max, double reg_rel,
                                       int what, int64_t P, void *stream) {
  if (!t || (!L && !Sigma0) || !times || !init_time || !pred_pairs || !ws || !diag_max || P < 1 || !(what & 3))
    return TCE_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
#define X(Dv, Kv)                                                                                                   \
  if (t->D == Dv && t->K1 == Kv) {                                                                                  \
    using FL = FusedLayout<Dv, Kv>;                                                                                 \
    const size_t sb = sizeof(double) * FL::NB * Kv * FL::KP, ls = sizeof(float) * FL::NR4 * FL::LD;                 \
    const size_t smem = (((sb > ls ? sb : ls) + 15) & ~(size_t)15) +                                                \
                        sizeof(double) * (FL::IR + (size_t)(UT / 32) * (2 * FL::NT + FL::N)) + 16;                  \
    if (smem > SMEM_LIMIT) return TCE_ERR_UNSUPPORTED_SHAPE;                                                        \
    if (Sigma0) {                                                                                                   \
      auto kern = uniform_prep_kernel<Dv, Kv, true>;                                                                \
      TCE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "uprep smem");   \
      kern<<<1, UT, smem, st>>>(tab_dev(t), nullptr, Sigma0, sigma_scale, times, init_time, pred_pairs, ws,         \
                                diag_max, reg_rel, (int)P, what);                                                   \
    } else {                                                                                                        \
      auto kern = uniform_prep_kernel<Dv, Kv, false>;                                                               \
      TCE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "uprep smem");   \
      kern<<<1, UT, smem, st>>>(tab_dev(t), L, nullptr, nullptr, times, init_time, pred_pairs, ws, diag_max,        \
                                reg_rel, (int)P, what);                                                             \
    }                                                                                                               \
    TCE_CHECK_LAUNCH("uniform_prep_kernel");                                                                        \
    return TCE_OK;                                                                                                  \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_uniform_main(const tce_tables_t *t, const double *ws, const float *smp_traj,
                                       const float *mean, const float *init_pos, const float *init_vel,
                                       const int64_t *pred_pairs, int grad_mode, const float *grad_logp,
                                       const float *logp_old, const float *advantage, double grad_scale,
                                       double *loss_acc, float *logp, int32_t *info, float *grad_mean, double *red,
                                       int64_t B, int64_t T, int64_t P, void *stream) {
  if (B == 0) return TCE_OK;
  if (!t || !ws || !smp_traj || !mean || !init_pos || !init_vel || !pred_pairs || B < 0 || T < 1 || P < 1 ||
      grad_mode < 0 || grad_mode > 2 || (grad_mode && !red))
    return TCE_ERR_INVALID_ARGUMENT;
  if (grad_mode == 1 && !grad_logp) return TCE_ERR_INVALID_ARGUMENT;
  if (grad_mode == 2 && (!logp_old || !advantage)) return TCE_ERR_INVALID_ARGUMENT;
  const long long Bpad = ((long long)B + 31) & ~31LL;
  const unsigned grid = (unsigned)((B + UM_EP - 1) / UM_EP);
#define X(Dv, Kv)                                                                                                  \
  if (t->D == Dv && t->K1 == Kv) {                                                                                 \
    uniform_main_kernel<Dv, Kv><<<grid, UM_THREADS, 0, (cudaStream_t)stream>>>(                                    \
        tab_dev(t), ws, smp_traj, mean, init_pos, init_vel, pred_pairs, grad_mode, grad_logp, logp_old, advantage, \
        grad_scale, loss_acc, logp, info, grad_mean, red, (long long)B, Bpad, (int)T, (int)P);                     \
    TCE_CHECK_LAUNCH("uniform_main_kernel");                                                                       \
    return TCE_OK;                                                                                                 \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}

extern "C" int tce_seglik_uniform_finish(const tce_tables_t *t, double *ws, const double *red, float *dsigma_part,
                                         int64_t B, int64_t P, void *stream) {
  if (!t || !ws || !red || !dsigma_part || B < 1 || P < 1) return TCE_ERR_INVALID_ARGUMENT;
  const long long Bpad = ((long long)B + 31) & ~31LL;
#define X(Dv, Kv)                                                                                          \
  if (t->D == Dv && t->K1 == Kv) {                                                                         \
    uniform_finish_kernel<Dv, Kv><<<(unsigned)P, 512, 0, (cudaStream_t)stream>>>(ws, red, (long long)B, Bpad, \
                                                                                 (int)P, dsigma_part);        \
    TCE_CHECK_LAUNCH("uniform_finish_kernel");                                                             \
    return TCE_OK;                                                                                         \
  }
  TCE_FOR_SHAPES(X)
#undef X
  return TCE_ERR_UNSUPPORTED_SHAPE;
}
