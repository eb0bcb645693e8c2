// Small dense linear algebra on shared-memory matrices (fp64), executed cooperatively by one CTA.
// Building blocks of the trust-region projection kernels (n <= 64: Dp = 63 / 36 / 28 / 12).
// Every routine is called by ALL threads of the CTA and ends with __syncthreads().
#pragma once
#include "tce_common.cuh"

struct Mat {            // strided view; a transpose is a stride swap
  double *p;
  int rs, cs;
  __device__ __forceinline__ double &operator()(int i, int j) const { return p[i * rs + j * cs]; }
  __device__ __forceinline__ Mat T() const { return Mat{p, cs, rs}; }
};

enum { TRI_FULL = 0, TRI_LOWER = 1, TRI_UPPER = 2 };

// C = beta * C + alpha * A B (m x n x k) with 2x2 register tiles (halves the shared-memory traffic per DFMA).
// a_tri / b_tri describe the structure of A and B and only clip the k range of a tile: the unused triangle
// of a triangular operand MUST hold zeros.  c_tri = TRI_LOWER / TRI_UPPER: tiles entirely on the other side
// of the diagonal are skipped; inside diagonal tiles both triangles are written (callers read one only).
__device__ inline void la_gemm(Mat C, Mat A, Mat B, int m, int n, int k, int a_tri, int b_tri, int c_tri,
                               double alpha, double beta) {
  const int tm = (m + 1) >> 1, tn = (n + 1) >> 1;
  for (int t = threadIdx.x; t < tm * tn; t += blockDim.x) {
    const int I = t / tn, J = t - I * tn;
    const int i0 = 2 * I, j0 = 2 * J;
    const int i1 = i0 + 1 < m ? i0 + 1 : i0, j1 = j0 + 1 < n ? j0 + 1 : j0;     // clamped (odd sizes)
    if ((c_tri == TRI_LOWER && j0 > i1) || (c_tri == TRI_UPPER && j1 < i0)) continue;
    int lo = 0, hi = k;
    if (a_tri == TRI_LOWER) hi = min(hi, i1 + 1);
    if (a_tri == TRI_UPPER) lo = max(lo, i0);
    if (b_tri == TRI_LOWER) lo = max(lo, j0);
    if (b_tri == TRI_UPPER) hi = min(hi, j1 + 1);
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    for (int q = lo; q < hi; ++q) {
      const double a0 = A(i0, q), a1 = A(i1, q), b0 = B(q, j0), b1 = B(q, j1);
      c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01);
      c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
    }
    if (beta == 0.0) {
      C(i0, j0) = alpha * c00;
      if (j1 != j0) C(i0, j1) = alpha * c01;
      if (i1 != i0) { C(i1, j0) = alpha * c10; if (j1 != j0) C(i1, j1) = alpha * c11; }
    } else {
      C(i0, j0) = fma(beta, C(i0, j0), alpha * c00);
      if (j1 != j0) C(i0, j1) = fma(beta, C(i0, j1), alpha * c01);
      if (i1 != i0) {
        C(i1, j0) = fma(beta, C(i1, j0), alpha * c10);
        if (j1 != j0) C(i1, j1) = fma(beta, C(i1, j1), alpha * c11);
      }
    }
  }
  __syncthreads();
}

constexpr int LA_NB = 8;                              // block size of the blocked triangular routines
constexpr int LA_DINV_DOUBLES = 8 * LA_NB * LA_NB;    // scratch for the inverted diagonal blocks (n <= 64)

// dinv[bk][i][j] = inverse of the bk-th LA_NB x LA_NB diagonal block of the lower-triangular L (zeros above
// the diagonal).  One warp per block, lane j < 8 builds column j.
__device__ inline void la_diag_block_inverses(Mat L, double *dinv, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (int)(blockDim.x >> 5);
  const int nblk = (n + LA_NB - 1) / LA_NB;
  for (int bk = warp; bk < nblk; bk += nwarp) {
    if (lane < LA_NB) {
      const int r0 = bk * LA_NB, j = lane;
      double x[LA_NB];
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) {
        const int gi = r0 + i;
        double v = (i == j) ? 1.0 : 0.0;
        if (gi < n) {
#pragma unroll
          for (int q = 0; q < i; ++q) v = fma(-L(gi, r0 + q), x[q], v);     // x[q] = 0 for q < j
          v = (i >= j) ? v / L(gi, gi) : 0.0;
        } else {
          v = 0.0;
        }
        x[i] = v;
        dinv[(bk * LA_NB + i) * LA_NB + j] = v;
      }
    }
  }
  __syncthreads();
}

// X <- L^-1 X, L lower triangular n x n, X n x ncols, dinv from la_diag_block_inverses(L).  Four lanes
// share one column (split-k over the already solved rows), one barrier per block row.
// x_lower: X(i, c) = 0 for i < c on entry (and on exit), those entries are skipped.
__device__ inline void la_trsm_lower(Mat L, const double *dinv, Mat X, int n, int ncols, bool x_lower) {
  constexpr int SUB = 4;
  const int sub = threadIdx.x % SUB;
  const unsigned gmask = ((1u << SUB) - 1u) << ((threadIdx.x & 31) & ~(SUB - 1));
  const int ncols_pad = ((ncols + 7) / 8) * 8;          // keep whole warps in the loop (shuffles below)
  for (int r0 = 0; r0 < n; r0 += LA_NB) {
    const int bk = r0 / LA_NB;
    for (int cc = threadIdx.x / SUB; cc < ncols_pad; cc += (int)blockDim.x / SUB) {
      const bool live = cc < ncols && !(x_lower && cc >= r0 + LA_NB);    // else structurally zero / padding
      const int c = cc < ncols ? cc : ncols - 1;
      double acc[LA_NB];
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) acc[i] = 0.0;
      if (live) {
        const int q_lo = x_lower ? (c / SUB) * SUB : 0;       // aligned so that the SUB lanes partition it
        for (int q = q_lo + sub; q < r0; q += SUB) {
          const double xq = X(q, c);
#pragma unroll
          for (int i = 0; i < LA_NB; ++i) acc[i] = fma(L(min(r0 + i, n - 1), q), xq, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) {
        acc[i] += __shfl_xor_sync(gmask, acc[i], 1);
        acc[i] += __shfl_xor_sync(gmask, acc[i], 2);
      }
      if (live && sub == 0) {
        double x[LA_NB];
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) x[i] = (r0 + i < n) ? X(r0 + i, c) - acc[i] : 0.0;
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) {
          if (r0 + i < n) {
            double y = 0.0;
#pragma unroll
            for (int q = 0; q <= i; ++q) y = fma(dinv[(bk * LA_NB + i) * LA_NB + q], x[q], y);
            if (!(x_lower && r0 + i < c)) X(r0 + i, c) = y;
          }
        }
      }
    }
    __syncthreads();
  }
}

// X <- L^-T X (back substitution with the transpose of the lower-triangular L)
__device__ inline void la_trsm_lower_t(Mat L, const double *dinv, Mat X, int n, int ncols) {
  constexpr int SUB = 4;
  const int sub = threadIdx.x % SUB;
  const unsigned gmask = ((1u << SUB) - 1u) << ((threadIdx.x & 31) & ~(SUB - 1));
  const int nblk = (n + LA_NB - 1) / LA_NB;
  const int ncols_pad = ((ncols + 7) / 8) * 8;
  for (int bk = nblk - 1; bk >= 0; --bk) {
    const int r0 = bk * LA_NB, r1 = min(r0 + LA_NB, n);
    for (int cc = threadIdx.x / SUB; cc < ncols_pad; cc += (int)blockDim.x / SUB) {
      const bool live = cc < ncols;
      const int c = live ? cc : ncols - 1;
      double acc[LA_NB];
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) acc[i] = 0.0;
      if (live) {
        for (int q = r1 + sub; q < n; q += SUB) {
          const double xq = X(q, c);
#pragma unroll
          for (int i = 0; i < LA_NB; ++i) acc[i] = fma(L(q, min(r0 + i, n - 1)), xq, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) {
        acc[i] += __shfl_xor_sync(gmask, acc[i], 1);
        acc[i] += __shfl_xor_sync(gmask, acc[i], 2);
      }
      if (live && sub == 0) {
        double x[LA_NB];
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) x[i] = (r0 + i < n) ? X(r0 + i, c) - acc[i] : 0.0;
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) {
          if (r0 + i < n) {
            double y = 0.0;                                   // (D^-1)^T : y_i = sum_{q >= i} dinv[q][i] x_q
#pragma unroll
            for (int q = i; q < LA_NB; ++q) y = fma(dinv[(bk * LA_NB + q) * LA_NB + i], x[q], y);
            X(r0 + i, c) = y;
          }
        }
      }
    }
    __syncthreads();
  }
}

// in-place blocked Cholesky of the lower triangle of A (upper part ignored); *bad (shared int, pre-set to 0)
// receives the 1-based index of the first non-positive pivot.  Per block column: diagonal block by one warp,
// panel rows by one thread each, rank-LA_NB trailing update by everybody: 3 barriers per 8 columns.
__device__ inline void la_chol(Mat A, int n, int *bad) {
  const int lane = threadIdx.x & 31;
  __shared__ double s_invd[LA_NB];                 // 1 / L_jj of the current diagonal block
  for (int r0 = 0; r0 < n; r0 += LA_NB) {
    const int nb = min(LA_NB, n - r0);
    if (threadIdx.x < 32) {                       // factor the diagonal block: lane i owns row r0 + i
      for (int j = 0; j < nb; ++j) {
        const double djj = A(r0 + j, r0 + j);
        if (lane == 0 && !(djj > 0.0) && *bad == 0) *bad = r0 + j + 1;
        const double inv = rsqrt(djj);           // one special-function op instead of sqrt + divide
        __syncwarp();
        if (lane == j) { A(r0 + j, r0 + j) = djj * inv; s_invd[j] = inv; }
        if (lane > j && lane < nb) A(r0 + lane, r0 + j) *= inv;
        __syncwarp();
        if (lane > j && lane < nb) {
          const double lij = A(r0 + lane, r0 + j);
          for (int q = j + 1; q <= lane; ++q) A(r0 + lane, r0 + q) = fma(-lij, A(r0 + q, r0 + j), A(r0 + lane, r0 + q));
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = r0 + nb + threadIdx.x; i < n; i += blockDim.x) {      // panel: row i of L21 = A21 L11^-T
      double x[LA_NB];
#pragma unroll
      for (int j = 0; j < LA_NB; ++j) {
        if (j < nb) {
          double v = A(i, r0 + j);
#pragma unroll
          for (int q = 0; q < j; ++q) v = fma(-x[q], A(r0 + j, r0 + q), v);
          x[j] = v * s_invd[j];
          A(i, r0 + j) = x[j];
        } else {
          x[j] = 0.0;
        }
      }
    }
    __syncthreads();
    const int t0 = r0 + nb, mrem = n - t0;                               // trailing update (lower part)
    for (int e = threadIdx.x; e < mrem * mrem; e += blockDim.x) {
      const int ii = e / mrem, qq = e - ii * mrem;
      if (qq > ii) continue;
      const int i = t0 + ii, q = t0 + qq;
      double v = A(i, q);
#pragma unroll
      for (int j = 0; j < LA_NB; ++j) if (j < nb) v = fma(-A(i, r0 + j), A(q, r0 + j), v);
      A(i, q) = v;
    }
    __syncthreads();
  }
}

// One-sided (Hestenes) Jacobi on the columns of W (n x n, in place): on exit the columns are mutually
// orthogonal, W_out = W_in V for an (implicit) orthogonal V, and lam[j] = |W_out(:, j)|^2 are the
// eigenvalues of W_in^T W_in.  V is never formed: callers only need W_out (see proj_kl_cov_fwd_kernel).
// Round-robin ordering, ONE WARP PER COLUMN PAIR (blockDim.x >= 32 * ceil(n/2)), one barrier per step.
// The step cost is shared-memory traffic (every pair streams its two columns), so nothing but W moves;
// squared column norms are cached in `nrm` (>= n doubles) and updated in closed form (one warp reduction
// per step); the rotation angle is evaluated in fp32 (it only steers convergence) while (c, s) are
// normalised in fp64 so that every rotation is orthogonal to ~1e-15.
__device__ inline int la_jacobi_onesided(Mat W, double *lam, double *scratch, int n) {
  // scratch: >= n/2 + 3 * 32 doubles: [n] squared norms as FLOATS, [32] dot products, [64] (c, s) per pair.
  // The dependent fp64 chain of a step is what costs time on one SM (~40 cycles per op), so everything that
  // only steers convergence (norms, thresholds, the rotation angle) is kept in fp32; fp64 is used for the
  // dot product, for normalising (c, s) and for applying the rotation.
  const int m = (n + 1) & ~1, half = m / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (int)(blockDim.x >> 5);
  float *nrm = reinterpret_cast<float *>(scratch);
  double *gbuf = scratch + (n + 1) / 2 + 1, *cs = gbuf + 32;
  int sweeps = 0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    for (int j = warp; j < n; j += nwarp) {      // norms once per sweep (bounds the closed-form drift)
      double a = 0.0;
      for (int r = lane; r < n; r += 32) a = fma(W(r, j), W(r, j), a);
      a = warp_sum(a);
      if (lane == 0) nrm[j] = (float)a;
    }
    __syncthreads();
    int big = 0;      // some pair was still correlated above 1e-4 when it was visited in this sweep
    for (int step = 0; step < m - 1; ++step) {
      auto pair_of = [&](int w, int &p, int &q) {
        if (w == 0) { p = m - 1; q = step; }
        else { p = step + w; if (p >= m - 1) p -= m - 1; q = step - w; if (q < 0) q += m - 1; }
        if (p > q) { const int t = p; p = q; q = t; }
      };
      int p = 0, q = n;
      double x0 = 0.0, y0 = 0.0, x1 = 0.0, y1 = 0.0;
      const int r0 = lane, r1 = lane + 32;
      // phase A: dot products of the column pairs (one warp per pair)
      if (warp < half) {
        pair_of(warp, p, q);
        if (q < n) {
          if (r0 < n) { x0 = W(r0, p); y0 = W(r0, q); }
          if (r1 < n) { x1 = W(r1, p); y1 = W(r1, q); }
          double g = fma(x0, y0, x1 * y1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
          if (lane == 0) reinterpret_cast<float *>(gbuf)[warp] = (float)g;
        }
      }
      __syncthreads();
      // phase B: rotation parameters of all pairs by ONE warp (lane = pair)
      if (warp == 0 && lane < half) {
        int pp, qq;
        pair_of(lane, pp, qq);
        double c = 1.0, sn = 0.0;
        if (qq < n) {
          const float g = reinterpret_cast<const float *>(gbuf)[lane], a = nrm[pp], b = nrm[qq];
          const float g2 = g * g, ab = a * b;
          if (g2 > 1e-8f * ab) big = 1;
          if (g2 > 1e-24f * ab) {
            const float zeta = __fdividef(b - a, 2.0f * g);
            const float tf = copysignf(__frcp_rn(fabsf(zeta) + __fsqrt_rn(fmaf(zeta, zeta, 1.0f))), zeta);
            const float hf = fmaf(tf, tf, 1.0f), c0 = rsqrtf(hf);
            const double t = (double)tf, h = fma(t, t, 1.0);
            c = (double)c0;
            c = c * fma(-0.5 * h, c * c, 1.5);       // one Newton step: c^2 h = 1 to ~1e-14
            sn = c * t;
            const float cf = c0, sf = c0 * tf, cs2 = 2.0f * cf * sf * g;   // |c x - s y|^2, |s x + c y|^2 (fp32)
            nrm[pp] = cf * cf * a + sf * sf * b - cs2;
            nrm[qq] = sf * sf * a + cf * cf * b + cs2;
          }
        }
        cs[2 * lane] = c; cs[2 * lane + 1] = sn;
      }
      __syncthreads();
      // phase C: apply the rotations (columns are disjoint between warps)
      if (warp < half && q < n) {
        const double c = cs[2 * warp], sn = cs[2 * warp + 1];
        if (sn != 0.0) {
          if (r0 < n) { W(r0, p) = c * x0 - sn * y0; W(r0, q) = sn * x0 + c * y0; }
          if (r1 < n) { W(r1, p) = c * x1 - sn * y1; W(r1, q) = sn * x1 + c * y1; }
        }
      }
      __syncthreads();
    }
    ++sweeps;
    // quadratic convergence: a sweep that only met cosines <= 1e-4 leaves them at ~1e-8 or below, i.e.
    // eigenvalues exact to ~1e-16 and vectors to ~1e-8 -- no separate verification sweep is needed
    if (!__syncthreads_or(big)) break;
  }
  for (int j = warp; j < n; j += nwarp) {
    double a = 0.0;
    for (int r = lane; r < n; r += 32) a = fma(W(r, j), W(r, j), a);
    a = warp_sum(a);
    if (lane == 0) lam[j] = a;
  }
  __syncthreads();
  return sweeps;
}
