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

// C = beta * C + alpha * A B over the index ranges allowed by the triangular structure of A (m x k) and
// B (k x n); c_tri restricts which entries of C are written (others untouched).
__device__ inline void la_gemm(Mat C, Mat A, Mat B, int m, int n, int k, int a_tri, int b_tri, int c_tri,
                               double alpha, double beta) {
  for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    if ((c_tri == TRI_LOWER && j > i) || (c_tri == TRI_UPPER && j < i)) continue;
    int lo = 0, hi = k;
    if (a_tri == TRI_LOWER) hi = min(hi, i + 1);
    if (a_tri == TRI_UPPER) lo = max(lo, i);
    if (b_tri == TRI_LOWER) lo = max(lo, j);
    if (b_tri == TRI_UPPER) hi = min(hi, j + 1);
    double s0 = 0.0, s1 = 0.0;
    int q = lo;
    for (; q + 1 < hi; q += 2) {
      s0 = fma(A(i, q), B(q, j), s0);
      s1 = fma(A(i, q + 1), B(q + 1, j), s1);
    }
    if (q < hi) s0 = fma(A(i, q), B(q, j), s0);
    const double v = alpha * (s0 + s1);
    C(i, j) = beta == 0.0 ? v : fma(beta, C(i, j), v);
  }
  __syncthreads();
}

// X <- L^-1 X, L lower triangular n x n (inv_diag[i] = 1 / L(i,i)), X n x ncols.  Blocked forward
// substitution (block 8): parallel GEMM update + short per-column solve.  x_lower: X(i,c) = 0 for i < c.
__device__ inline void la_trsm_lower(Mat L, const double *inv_diag, Mat X, int n, int ncols, bool x_lower) {
  constexpr int NB = 8;
  for (int r0 = 0; r0 < n; r0 += NB) {
    const int nb = min(NB, n - r0);
    if (r0 > 0) {
      for (int e = threadIdx.x; e < nb * ncols; e += blockDim.x) {
        const int c = e % ncols, i = r0 + e / ncols;
        if (x_lower && i < c) continue;
        const int lo = x_lower ? c : 0;
        double s0 = 0.0, s1 = 0.0;
        int q = lo;
        for (; q + 1 < r0; q += 2) {
          s0 = fma(L(i, q), X(q, c), s0);
          s1 = fma(L(i, q + 1), X(q + 1, c), s1);
        }
        if (q < r0) s0 = fma(L(i, q), X(q, c), s0);
        X(i, c) -= s0 + s1;
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
      for (int ii = 0; ii < nb; ++ii) {
        const int i = r0 + ii;
        if (x_lower && i < c) continue;
        double v = X(i, c);
        for (int kk = 0; kk < ii; ++kk) v = fma(-L(i, r0 + kk), X(r0 + kk, c), v);
        X(i, c) = v * inv_diag[i];
      }
    }
    __syncthreads();
  }
}

// X <- L^-T X (back substitution with the transpose of a lower-triangular L)
__device__ inline void la_trsm_lower_t(Mat L, const double *inv_diag, Mat X, int n, int ncols) {
  constexpr int NB = 8;
  for (int r1 = n; r1 > 0; r1 -= NB) {
    const int r0 = max(r1 - NB, 0), nb = r1 - r0;
    if (r1 < n) {
      for (int e = threadIdx.x; e < nb * ncols; e += blockDim.x) {
        const int c = e % ncols, i = r0 + e / ncols;
        double s0 = 0.0, s1 = 0.0;
        int q = r1;
        for (; q + 1 < n; q += 2) {
          s0 = fma(L(q, i), X(q, c), s0);
          s1 = fma(L(q + 1, i), X(q + 1, c), s1);
        }
        if (q < n) s0 = fma(L(q, i), X(q, c), s0);
        X(i, c) -= s0 + s1;
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
      for (int i = r1 - 1; i >= r0; --i) {
        double v = X(i, c);
        for (int q = i + 1; q < r1; ++q) v = fma(-L(q, i), X(q, c), v);
        X(i, c) = v * inv_diag[i];
      }
    }
    __syncthreads();
  }
}

__device__ inline void la_inv_diag(Mat L, double *inv_diag, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) inv_diag[i] = 1.0 / L(i, i);
  __syncthreads();
}

// in-place Cholesky of the lower triangle of A (upper part untouched / ignored); returns through *bad
// (shared int, pre-set to 0) the 1-based index of the first non-positive pivot
__device__ inline void la_chol(Mat A, int n, int *bad) {
  for (int j = 0; j < n; ++j) {
    const double djj = A(j, j);
    if (threadIdx.x == 0 && !(djj > 0.0) && *bad == 0) *bad = j + 1;
    const double d = sqrt(djj), inv = 1.0 / d;
    __syncthreads();                       // everyone has read A(j,j)
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) A(i, j) = (i == j) ? d : A(i, j) * inv;
    __syncthreads();
    const int m = n - j - 1;               // trailing update of the lower triangle
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) A(i, k) = fma(-A(i, j), A(k, j), A(i, k));
    }
    __syncthreads();
  }
}

// Cyclic Jacobi eigen-decomposition of a symmetric matrix (full storage) A = Q diag(lam) Q^T with the
// round-robin parallel ordering (m/2 disjoint rotations per step).  A is destroyed (its diagonal ends as
// lam), Q must hold the identity on entry.  m = n rounded up to even; A and Q need m rows/cols of storage
// with the padding row/col zero.  rot: scratch of 4 * (m/2) doubles.
__device__ inline void la_jacobi(Mat A, Mat Q, double *lam, int n, double *rot, double *red /*[33]*/) {
  const int m = (n + 1) & ~1, half = m / 2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: sum of squared off-diagonal entries vs squared diagonal
    double off = 0.0, dia = 0.0;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e % n;
      const double v = A(i, j);
      if (i == j) dia = fma(v, v, dia); else off = fma(v, v, off);
    }
    off = warp_sum(off); dia = warp_sum(dia);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = off; red[32 + (threadIdx.x >> 5)] = dia; }
    __syncthreads();
    double toff = 0.0, tdia = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { toff += red[w]; tdia += red[32 + w]; }
    __syncthreads();
    if (toff <= 1e-30 * tdia) break;
    for (int step = 0; step < m - 1; ++step) {
      if (threadIdx.x < half) {
        const int k = threadIdx.x;
        int p, q;
        if (k == 0) { p = m - 1; q = step % (m - 1); }
        else { p = (step + k) % (m - 1); q = (step - k + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A(p, q);
          if (fabs(apq) > 1e-300) {
            const double tau = (A(q, q) - A(p, p)) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        rot[4 * k] = c; rot[4 * k + 1] = s; rot[4 * k + 2] = (double)p; rot[4 * k + 3] = (double)q;
      }
      __syncthreads();
      // columns: A <- A J, Q <- Q J   (J(p,p)=c, J(p,q)=s, J(q,p)=-s, J(q,q)=c)
      for (int e = threadIdx.x; e < 2 * half * n; e += blockDim.x) {
        const int which = e / (half * n), r = (e % (half * n)) / half, k = e % half;
        const double c = rot[4 * k], s = rot[4 * k + 1];
        const int p = (int)rot[4 * k + 2], q = (int)rot[4 * k + 3];
        if (s == 0.0 || q >= n) continue;
        Mat M = which ? Q : A;
        const double x = M(r, p), y = M(r, q);
        M(r, p) = c * x - s * y;
        M(r, q) = s * x + c * y;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = threadIdx.x; e < half * n; e += blockDim.x) {
        const int col = e / half, k = e % half;
        const double c = rot[4 * k], s = rot[4 * k + 1];
        const int p = (int)rot[4 * k + 2], q = (int)rot[4 * k + 3];
        if (s == 0.0 || q >= n) continue;
        const double x = A(p, col), y = A(q, col);
        A(p, col) = c * x - s * y;
        A(q, col) = s * x + c * y;
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) lam[i] = A(i, i);
  __syncthreads();
}

// One-sided (Hestenes) Jacobi SVD of W (n x n, destroyed): W V = U diag(sigma), i.e. W^T W = V diag(sigma^2) V^T.
// V must hold the identity on entry; on exit lam[i] = sigma_i^2.  Round-robin ordering, ONE WARP PER COLUMN
// PAIR (needs blockDim.x >= 32 * ceil(n/2)), one barrier per step.  Works on W directly (no W^T W), which
// keeps the small generalised eigenvalues accurate.  Latency matters (a single matrix sits on the critical
// path of every policy epoch): squared column norms are cached in `nrm` (>= n doubles) and updated in closed
// form, so a step needs ONE warp reduction; the rotation angle is evaluated in fp32 (it only steers
// convergence) while (c, s) are normalised in fp64 so every rotation stays orthogonal to 1e-16.
__device__ inline void la_jacobi_onesided(Mat W, Mat V, double *lam, double *nrm, int n) {
  const int m = (n + 1) & ~1, half = m / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (int)(blockDim.x >> 5);
  for (int sweep = 0; sweep < 40; ++sweep) {
    // (re)compute the squared column norms once per sweep (also bounds the drift of the closed-form update)
    for (int j = warp; j < n; j += nwarp) {
      double a = 0.0;
      for (int r = lane; r < n; r += 32) a = fma(W(r, j), W(r, j), a);
      a = warp_sum(a);
      if (lane == 0) nrm[j] = a;
    }
    __syncthreads();
    int rotated = 0;
    for (int step = 0; step < m - 1; ++step) {
      if (warp < half) {
        int p, q;
        if (warp == 0) { p = m - 1; q = step % (m - 1); }
        else { p = (step + warp) % (m - 1); q = (step - warp + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        if (q < n) {
          const int r0 = lane, r1 = lane + 32;
          const double x0 = r0 < n ? W(r0, p) : 0.0, y0 = r0 < n ? W(r0, q) : 0.0;
          const double x1 = r1 < n ? W(r1, p) : 0.0, y1 = r1 < n ? W(r1, q) : 0.0;
          double g = fma(x0, y0, x1 * y1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
          const double a = nrm[p], b = nrm[q];
          if (g * g > 1e-22 * a * b) {
            rotated = 1;
            const float zeta = (float)((b - a) / (2.0 * g));
            const float tf = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
            const double t = (double)tf;
            const double c = rsqrt(fma(t, t, 1.0)), s = c * t;
            if (r0 < n) {
              W(r0, p) = c * x0 - s * y0; W(r0, q) = s * x0 + c * y0;
              const double u = V(r0, p), v = V(r0, q);
              V(r0, p) = c * u - s * v; V(r0, q) = s * u + c * v;
            }
            if (r1 < n) {
              W(r1, p) = c * x1 - s * y1; W(r1, q) = s * x1 + c * y1;
              const double u = V(r1, p), v = V(r1, q);
              V(r1, p) = c * u - s * v; V(r1, q) = s * u + c * v;
            }
            if (lane == 0) {                       // |c x - s y|^2 and |s x + c y|^2
              const double c2 = c * c, s2 = s * s, cs2 = 2.0 * c * s * g;
              nrm[p] = c2 * a + s2 * b - cs2;
              nrm[q] = s2 * a + c2 * b + cs2;
            }
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) break;
  }
  for (int j = warp; j < n; j += nwarp) {
    double a = 0.0;
    for (int r = lane; r < n; r += 32) a = fma(W(r, j), W(r, j), a);
    a = warp_sum(a);
    if (lane == 0) lam[j] = a;
  }
  __syncthreads();
}
