// (2) Gaussian policy over the MP parameters: rsample, batched small Cholesky (fwd/bwd), the
// vector -> Cholesky policy head and the per-episode Gaussian scalars (maha, trace, logdet, entropy).
// Replaces black_box_policy.py:58-224 (torch MultivariateNormal / cholesky_solve / solve_triangular),
// abstract_policy.py:166-187 + util_matrix.py:12-33 (head) and torch.linalg.cholesky (util_matrix.py:133).
// One CTA per matrix, matrices staged in shared memory, warp shuffles for the reductions.
#include <math.h>

#include "tce_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al.) -> one standard normal per element (Box-Muller)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u0 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u1 = (float)c[1] * 2.3283064365386963e-10f;            // [0, 1)
  return sqrtf(-2.0f * logf(u0)) * cospif(2.0f * u1);
}

// out[b] = mean[b] + L[b] eps[b];  one CTA (128 threads) per episode, warp per row, lanes over columns
__global__ void __launch_bounds__(128)
rsample_kernel(const float *__restrict__ mean, const float *__restrict__ L, long long ldb_L,
               const float *__restrict__ eps, uint64_t seed, uint64_t offset, float *__restrict__ out, int n) {
  extern __shared__ float s_eps[];
  const long long b = blockIdx.x;
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    s_eps[j] = eps ? eps[b * n + j] : philox_normal(seed, offset, (uint64_t)(b * n + j));
  __syncthreads();
  const float *Lb = L + b * ldb_L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < n; i += nw) {
    float acc = 0.f;
    for (int j = lane; j <= i; j += 32) acc = fmaf(Lb[(size_t)i * n + j], s_eps[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[b * n + i] = mean[b * n + i] + acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// Cholesky forward: left-looking, thread per row, fp64 accumulation, one barrier pair per column
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
chol_fwd_kernel(const float *__restrict__ A, float *__restrict__ Lout, int32_t *__restrict__ info, int n) {
  extern __shared__ float sL[];            // [n][n+1]
  __shared__ int s_bad;
  const int LD = n + 1;
  const long long b = blockIdx.x;
  const float *Ab = A + (size_t)b * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) sL[(e / n) * LD + e % n] = Ab[e];
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const int i = threadIdx.x;
  for (int j = 0; j < n; ++j) {
    if (i == j) {
      double d = (double)sL[j * LD + j];
      for (int k = 0; k < j; ++k) d -= (double)sL[j * LD + k] * (double)sL[j * LD + k];
      if (!(d > 0.0) && s_bad == 0) s_bad = j + 1;
      sL[j * LD + j] = (float)sqrt(d);
    }
    __syncthreads();
    if (i > j && i < n) {
      double v = (double)sL[i * LD + j];
      for (int k = 0; k < j; ++k) v -= (double)sL[i * LD + k] * (double)sL[j * LD + k];
      sL[i * LD + j] = (float)(v / (double)sL[j * LD + j]);
    }
    __syncthreads();
  }
  float *Lb = Lout + (size_t)b * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e % n;
    Lb[e] = c <= r ? sL[r * LD + c] : 0.f;
  }
  if (info && threadIdx.x == 0) info[b] = s_bad;
}

// Cholesky backward: grad_A = sym(S^-T Phi(S^T Gbar) S^-1), Phi = tril with halved diagonal.
// fp64 in shared memory, thread per column / row for the two triangular solves.
__global__ void __launch_bounds__(128)
chol_bwd_kernel(const float *__restrict__ S, const float *__restrict__ gS, float *__restrict__ gA, int n) {
  extern __shared__ double sd[];
  const int LD = n + 1;
  double *sS = sd, *sG = sd + n * LD, *sM = sG + n * LD;
  const long long b = blockIdx.x;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e % n;
    sS[r * LD + c] = c <= r ? (double)S[(size_t)b * n * n + e] : 0.0;
    sG[r * LD + c] = c <= r ? (double)gS[(size_t)b * n * n + e] : 0.0;
  }
  __syncthreads();
  // M = Phi(S^T G): M[i][j] = sum_{k >= i} S[k][i] G[k][j], i >= j
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    double v = 0.0;
    if (j <= i) {
      for (int k = i; k < n; ++k) v = fma(sS[k * LD + i], sG[k * LD + j], v);
      if (i == j) v *= 0.5;
    }
    sM[i * LD + j] = v;
  }
  __syncthreads();
  // Y = S^-T M : column c, back substitution (thread per column) -> sG
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    for (int i = n - 1; i >= 0; --i) {
      double v = sM[i * LD + c];
      for (int k = i + 1; k < n; ++k) v = fma(-sS[k * LD + i], sG[k * LD + c], v);
      sG[i * LD + c] = v / sS[i * LD + i];
    }
  }
  __syncthreads();
  // X = Y S^-1 : row r, X[r][c] = (Y[r][c] - sum_{k > c} X[r][k] S[k][c]) / S[c][c] (thread per row) -> sM
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    for (int c = n - 1; c >= 0; --c) {
      double v = sG[r * LD + c];
      for (int k = c + 1; k < n; ++k) v = fma(-sM[r * LD + k], sS[k * LD + c], v);
      sM[r * LD + c] = v / sS[c * LD + c];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e % n;
    gA[(size_t)b * n * n + e] = (float)(0.5 * (sM[r * LD + c] + sM[c * LD + r]));
  }
}

// ---------------------------------------------------------------------------------------------------
// policy head
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void __launch_bounds__(256)
head_fwd_kernel(const float *__restrict__ vec, long long ldb_vec, float min_std, float *__restrict__ L, int n) {
  const long long b = blockIdx.x;
  const float *v = vec + b * ldb_vec;
  float *Lb = L + (size_t)b * n * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < n; r += nw) {
    const float *off = v + n + r * (r - 1) / 2;
    for (int c = lane; c < n; c += 32) {
      float o = 0.f;
      if (c < r) o = off[c];
      else if (c == r) o = softplus_f(v[r]) + min_std;
      Lb[r * n + c] = o;
    }
  }
}

// grid (ceil(nvec/128), batch chunks); grad_vec zero-initialised by the launcher when reducing
__global__ void head_bwd_kernel(const float *__restrict__ vec, long long ldb_vec, const float *__restrict__ gL,
                                float *__restrict__ gvec, long long B, int n, int chunk) {
  const int nvec = n + n * (n - 1) / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nvec) return;
  int r, c;
  if (e < n) { r = c = e; }
  else {
    const int o = e - n;
    r = (int)((sqrtf(8.0f * (float)o + 1.0f) + 1.0f) * 0.5f);
    while (r * (r - 1) / 2 > o) --r;
    while ((r + 1) * r / 2 <= o) ++r;
    c = o - r * (r - 1) / 2;
  }
  const long long b0 = (long long)blockIdx.y * chunk;
  const long long b1 = b0 + chunk < B ? b0 + chunk : B;
  if (ldb_vec == 0) {
    float acc = 0.f;
    for (long long b = b0; b < b1; ++b) acc += gL[(size_t)b * n * n + r * n + c];
    if (r == c) acc *= 1.f / (1.f + expf(-vec[e]));
    atomicAdd(gvec + e, acc);
  } else {
    for (long long b = b0; b < b1; ++b) {
      float g = gL[(size_t)b * n * n + r * n + c];
      if (r == c) g *= 1.f / (1.f + expf(-vec[b * ldb_vec + e]));
      gvec[b * nvec + e] = g;
    }
  }
}

}  // namespace

extern "C" int tce_mvn_rsample(const float *mean, const float *L, int64_t ldb_L, const float *eps, uint64_t seed,
                               uint64_t offset, float *out, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!mean || !L || !out || B < 0 || n < 1 || n > 200) return TCE_ERR_INVALID_ARGUMENT;
  rsample_kernel<<<(unsigned)B, 128, n * sizeof(float), (cudaStream_t)stream>>>(mean, L, ldb_L, eps, seed, offset, out, n);
  TCE_CHECK_LAUNCH("rsample_kernel");
  return TCE_OK;
}

extern "C" int tce_chol_fwd(const float *A, float *L, int32_t *info, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!A || !L || B < 0 || n < 1 || n > 128) return TCE_ERR_INVALID_ARGUMENT;
  const size_t smem = (size_t)n * (n + 1) * sizeof(float);
  if (smem > 48 * 1024)
    TCE_CUDA(cudaFuncSetAttribute(chol_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "chol attr");
  chol_fwd_kernel<<<(unsigned)B, 128, smem, (cudaStream_t)stream>>>(A, L, info, n);
  TCE_CHECK_LAUNCH("chol_fwd_kernel");
  return TCE_OK;
}

extern "C" int tce_chol_bwd(const float *L, const float *grad_L, float *grad_A, int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!L || !grad_L || !grad_A || B < 0 || n < 1 || n > 96) return TCE_ERR_INVALID_ARGUMENT;
  const size_t smem = 3 * (size_t)n * (n + 1) * sizeof(double);
  TCE_CUDA(cudaFuncSetAttribute(chol_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cholb attr");
  chol_bwd_kernel<<<(unsigned)B, 128, smem, (cudaStream_t)stream>>>(L, grad_L, grad_A, n);
  TCE_CHECK_LAUNCH("chol_bwd_kernel");
  return TCE_OK;
}

extern "C" int tce_policy_head_fwd(const float *vec, int64_t ldb_vec, float min_std, float *L, int64_t B, int n,
                                   void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!vec || !L || B < 0 || n < 1) return TCE_ERR_INVALID_ARGUMENT;
  head_fwd_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(vec, ldb_vec, min_std, L, n);
  TCE_CHECK_LAUNCH("head_fwd_kernel");
  return TCE_OK;
}

extern "C" int tce_policy_head_bwd(const float *vec, int64_t ldb_vec, const float *grad_L, float *grad_vec,
                                   int64_t B, int n, void *stream) {
  if (B == 0) return TCE_OK;              /* empty shard: pointers may be NULL */
  if (!vec || !grad_L || !grad_vec || B < 0 || n < 1) return TCE_ERR_INVALID_ARGUMENT;
  const int nvec = n + n * (n - 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (ldb_vec == 0) TCE_CUDA(cudaMemsetAsync(grad_vec, 0, nvec * sizeof(float), st), "head bwd memset");
  const int chunk = 32;
  dim3 grid((nvec + 127) / 128, (unsigned)((B + chunk - 1) / chunk));
  head_bwd_kernel<<<grid, 128, 0, st>>>(vec, ldb_vec, grad_L, grad_vec, B, n, chunk);
  TCE_CHECK_LAUNCH("head_bwd_kernel");
  return TCE_OK;
}
