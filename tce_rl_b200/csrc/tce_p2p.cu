// Data-parallel gradient exchange of the policy / critic update fused with the gradient norm, over NVLink peer memory:
// ONE kernel per optimiser step replaces  ncclAllReduce(AVG, flat gradient)  +  the sum-of-squares kernel
// (temporal_correlated_agent.py:583-589: clip_grad_norm_ -> optimizer.step(); the reference is single process, the
// exchange is what data parallelism adds: SURVEY 8(e) collective (1)).
//
// Every rank's flat gradient buffer lives in symmetric memory (same size on all ranks, each rank's buffer mapped into
// every peer's address space through NVLink / NVSwitch).  One-shot scheme, no staging copies:
//   1. arrive : block 0 writes the launch sequence number into slot [rank] of every peer's signal pad (release, system
//               scope).  Stream order guarantees this rank's gradient is complete before the kernel starts.
//   2. wait   : every block spins (acquire loads of its OWN pad: local memory) until all peers have arrived.
//   3. reduce : grid-stride over the buffer; element i = (1/W) * sum over ranks r = 0..W-1 IN RANK ORDER of
//               peer_r[i] (128-bit loads through the peer mapping) -> bit-identical results on all ranks, so the
//               replicas never drift; written to a LOCAL output buffer that Adam consumes; sum of squares of the
//               averaged gradient accumulated for the norm / clipping (state[1]), step counter state[0] += 1.
//   4. depart : the last block to finish tells every peer "I have read your buffer" (slot [W + rank]) and waits until
//               every peer has read this rank's buffer -- only then may the next epoch clear the gradients.
// Latency: two NVLink signal round trips (~2-3 us each) + ~1 MB of peer reads at W = 8, instead of a ring / tree
// NCCL all-reduce of a 144 KB message (~20-30 us inside a CUDA graph) followed by a separate reduction kernel.
// All spins are bounded (~2 s of SM clocks): on expiry state[2] is set and the kernel gives up instead of hanging.
#include <math.h>

#include "tce_common.cuh"

namespace {

constexpr int P2P_THREADS = 512;
constexpr int P2P_MAX_WORLD = 16;
constexpr int TCE_P2P_MAX_PHASES = 2;

struct PeerTable {                  // passed by value (baked into a captured graph with the launch)
  const float *buf[P2P_MAX_WORLD];              // every rank's gradient buffer as seen from this rank
  unsigned long long *pad[P2P_MAX_WORLD];       // every rank's signal pad as seen from this rank
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float *p) {           // L1-bypassing 128-bit load (data changes per launch)
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// spin until pad[slot] >= seq for the `count` slots starting at `first`; false on timeout
__device__ bool wait_slots(const unsigned long long *pad, int first, int count, unsigned long long seq) {
  const long long t0 = clock64();
  for (int s = threadIdx.x; s < count; s += blockDim.x) {
    while (ld_acquire_sys(pad + first + s) < seq) {
      if (clock64() - t0 > 4000000000LL) return false;
      __nanosleep(64);
    }
  }
  return true;
}

// sum over the ranks IN RANK ORDER of the 128-bit word i; all W peer loads are in flight before the first add (a loop with a
// run-time trip count issues them one NVLink round trip after the other)
template <int WT>
__device__ __forceinline__ float4 gather_sum(const PeerTable &pt, long long off, long long i, int W) {
  if (WT > 0) {
    float4 v[WT > 0 ? WT : 1];
#pragma unroll
    for (int r = 0; r < WT; ++r) v[r] = ld_peer(pt.buf[r] + off + 4 * i);
    float4 a = v[0];
#pragma unroll
    for (int r = 1; r < WT; ++r) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
    return a;
  }
  float4 a = ld_peer(pt.buf[0] + off + 4 * i);
  for (int r = 1; r < W; ++r) {
    const float4 b = ld_peer(pt.buf[r] + off + 4 * i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  return a;
}

// local: {counter (sequence number of the last launch), blocks finished}; state: {step, sum g^2, error flag}.
// The exchange covers elements [off, off + n) of every rank's buffer; `slot0`: first of the 2 W signal-pad slots of this
// phase (an update may exchange the gradient in several ranges as they become final, each with its own slots and `local`).
template <int WT>
__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_sumsq_kernel(PeerTable pt, long long off, long long n, int slot0, int bump_step, float *__restrict__ avg_out,
                           unsigned long long *local, double *__restrict__ state) {
  __shared__ double red[32];
  __shared__ int s_ok, s_last;
  const int W = pt.world;
  const unsigned long long seq = local[0] + 1;           // written back by the last block (all blocks read the old value
  unsigned long long *mypad = pt.pad[pt.rank] + slot0;   //  first: the write happens after a grid-wide ticket)
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(pt.pad[threadIdx.x] + slot0 + pt.rank, seq);   // 1. arrive
  if (!wait_slots(mypad, 0, W, seq)) s_ok = 0;                                                       // 2. wait
  __syncthreads();
  double ss = 0.0;
  if (s_ok) {                                                                                        // 3. reduce
    const float inv = 1.0f / (float)W;
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 a = gather_sum<WT>(pt, off, i, W);
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      *reinterpret_cast<float4 *>(avg_out + off + 4 * i) = a;
      ss += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {            // tail (n is a multiple of 4 in practice)
      const long long i = off + (n4 << 2) + threadIdx.x;
      float a = 0.f;
      for (int r = 0; r < W; ++r) a += __ldcg(pt.buf[r] + i);
      a *= inv;
      avg_out[i] = a;
      ss += (double)a * a;
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    ss = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    ss = warp_sum(ss);
    if (threadIdx.x == 0) {
      atomicAdd(state + 1, ss);
      if (!s_ok) state[2] = 1.0;
      __threadfence();
      s_last = (atomicAdd(reinterpret_cast<unsigned int *>(local + 1), 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) {                                                                             // bookkeeping
    if (bump_step) state[0] += 1.0;
    *reinterpret_cast<unsigned int *>(local + 1) = 0u;
    local[0] = seq;
  }
  __threadfence_system();
  if (threadIdx.x < W) st_release_sys(pt.pad[threadIdx.x] + slot0 + W + pt.rank, seq);                // 4. depart
  if (!wait_slots(mypad, W, W, seq) && threadIdx.x == 0) state[2] = 1.0;
}

// ---- push variant: one NVLink one-way trip instead of three round trips -----------------------------------------------
// Every rank owns an exchange area in symmetric memory: [flags: 2 phases x 16 sources x 32 blocks uint64 | receive buffers:
// 2 parities x W sources x npad floats].  Block b of rank s
//   1. push : stores its slice of the LOCAL gradient into slot [parity][s] of every peer's receive buffer (128-bit stores
//             through the peer mapping: fire and forget), then -- release, system scope -- writes the launch sequence number
//             into flag [phase][s][b] of every peer;
//   2. wait : spins on its OWN flags [phase][r][b] of all peers r (local memory) -- per block, no grid-wide step;
//   3. reduce: sums slot r = 0..W-1 IN RANK ORDER (own slice from the gradient itself) from LOCAL memory, scales by 1/W,
//             writes avg_out and accumulates the squared norm: bit-identical averages on all ranks.
// No arrival barrier and no departure barrier: the receive slot alternates with the parity of the sequence number, and a
// rank can only be one launch ahead of a peer (its launch k + 1 needs the peer's flag k + 1, which the peer writes after
// it has finished reading launch k - 1's slot of the same parity in stream order).
constexpr int P2P_FLAG_BLOCKS = 32;
constexpr size_t P2P_FLAG_BYTES = 2 * P2P_MAX_WORLD * P2P_FLAG_BLOCKS * sizeof(unsigned long long);

struct PushTable {
  char *xchg[P2P_MAX_WORLD];                    // every rank's exchange area as seen from this rank
  int world, rank;
};

__device__ __forceinline__ void st_peer(float *p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int WT>
__global__ void __launch_bounds__(P2P_THREADS)
p2p_push_allreduce_sumsq_kernel(PushTable pt, const float *__restrict__ grad, long long off, long long n, long long npad,
                                int phase, int bump_step, float *__restrict__ avg_out, unsigned long long *local,
                                double *__restrict__ state) {
  __shared__ double red[32];
  __shared__ int s_ok;
  __shared__ unsigned long long s_seq;
  const int W = WT > 0 ? WT : pt.world, rank = pt.rank;
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // the gradient words of the first pass are requested BEFORE the sequence number is needed (two independent misses
  // in flight instead of one after the other: the caches are cold when this kernel starts)
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (first < n4) v0 = *reinterpret_cast<const float4 *>(grad + off + 4 * first);
  if (threadIdx.x == 0) {
    s_seq = *reinterpret_cast<volatile unsigned long long *>(local) + 1;      // ONE reader per block (see the ticket below)
    s_ok = 1;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const long long parity = (long long)(seq & 1ULL);
  const size_t my_slot = ((size_t)parity * W + rank) * (size_t)npad + (size_t)off;
  for (long long i = first; i < n4; i += stride) {                                                   // 1. push
    const float4 v = i == first ? v0 : *reinterpret_cast<const float4 *>(grad + off + 4 * i);
#pragma unroll
    for (int r = 0; r < (WT > 0 ? WT : P2P_MAX_WORLD); ++r)
      if (r < W && r != rank) st_peer(reinterpret_cast<float *>(pt.xchg[r] + P2P_FLAG_BYTES) + my_slot + 4 * i, v);
  }
  __syncthreads();
  if (threadIdx.x < W && threadIdx.x != rank) {
    // release at system scope after the barrier: cumulative over the whole block's stores (no separate fence)
    st_release_sys(reinterpret_cast<unsigned long long *>(pt.xchg[threadIdx.x]) +
                       ((size_t)phase * P2P_MAX_WORLD + rank) * P2P_FLAG_BLOCKS + blockIdx.x, seq);
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(pt.xchg[rank]) +       // 2. wait
                                     ((size_t)phase * P2P_MAX_WORLD + threadIdx.x) * P2P_FLAG_BLOCKS + blockIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < seq) {
      if (clock64() - t0 > 4000000000LL) { s_ok = 0; break; }
    }
  } else if (threadIdx.x == 32) {
    // bookkeeping while the flags travel: the last block to get here knows that every block has read the sequence number
    // (each block's only read precedes its first barrier), so it may advance it for the next launch
    __threadfence();
    if (atomicAdd(reinterpret_cast<unsigned int *>(local + 1), 1u) == gridDim.x - 1) {
      *reinterpret_cast<unsigned int *>(local + 1) = 0u;
      local[0] = seq;
      if (bump_step) state[0] += 1.0;
    }
  }
  __syncthreads();
  double ss = 0.0;
  if (s_ok) {                                                                                        // 3. reduce
    const float inv = 1.0f / (float)W;
    const float *rbase = reinterpret_cast<const float *>(pt.xchg[rank] + P2P_FLAG_BYTES) + (size_t)parity * W * (size_t)npad +
                         (size_t)off;
    for (long long i = first; i < n4; i += stride) {
      float4 v[WT > 0 ? WT : P2P_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < (WT > 0 ? WT : P2P_MAX_WORLD); ++r)
        if (r < W)
          v[r] = r == rank ? (i == first ? v0 : *reinterpret_cast<const float4 *>(grad + off + 4 * i))
                           : ld_peer(rbase + (size_t)r * npad + 4 * i);
      float4 a = v[0];
#pragma unroll
      for (int r = 1; r < (WT > 0 ? WT : P2P_MAX_WORLD); ++r)
        if (r < W) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      *reinterpret_cast<float4 *>(avg_out + off + 4 * i) = a;
      ss += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    ss = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    ss = warp_sum(ss);
    if (threadIdx.x == 0) {
      atomicAdd(state + 1, ss);                  // result unused: a reduction, nothing waits for it
      if (!s_ok) state[2] = 1.0;
    }
  }
}

}  // namespace

extern "C" int tce_p2p_allreduce_sumsq_range(int world, int rank, const void *const *peer_bufs, void *const *peer_pads,
                                             int64_t offset, int64_t n, int phase, int bump_step, float *avg_out,
                                             void *local2, double *state3, void *stream) {
  if (n == 0) return TCE_OK;
  if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || !peer_bufs || !peer_pads || !avg_out ||
      !local2 || !state3 || n < 0 || offset < 0 || (offset & 3) || phase < 0 || phase >= TCE_P2P_MAX_PHASES ||
      ((uintptr_t)avg_out & 15))
    return TCE_ERR_INVALID_ARGUMENT;
  PeerTable pt;
  pt.world = world;
  pt.rank = rank;
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r] || !peer_pads[r] || ((uintptr_t)peer_bufs[r] & 15)) return TCE_ERR_INVALID_ARGUMENT;
    pt.buf[r] = static_cast<const float *>(peer_bufs[r]);
    pt.pad[r] = static_cast<unsigned long long *>(peer_pads[r]);
  }
  long long blocks = (n / 4 + P2P_THREADS - 1) / P2P_THREADS;
  if (blocks > 32) blocks = 32;                 // a few CTAs: the message is small, the barrier traffic stays low
  if (blocks < 1) blocks = 1;
  unsigned long long *local = static_cast<unsigned long long *>(local2) + 2 * phase;
  const int slot0 = 2 * world * phase;
#define TCE_P2P_LAUNCH(WT)                                                                                  \
  p2p_allreduce_sumsq_kernel<WT><<<(unsigned)blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(               \
      pt, (long long)offset, (long long)n, slot0, bump_step, avg_out, local, state3)
  switch (world) {
    case 2: TCE_P2P_LAUNCH(2); break;
    case 4: TCE_P2P_LAUNCH(4); break;
    case 8: TCE_P2P_LAUNCH(8); break;
    default: TCE_P2P_LAUNCH(0); break;
  }
#undef TCE_P2P_LAUNCH
  TCE_CHECK_LAUNCH("p2p_allreduce_sumsq_kernel");
  return TCE_OK;
}

extern "C" int tce_p2p_allreduce_sumsq(int world, int rank, const void *const *peer_bufs, void *const *peer_pads,
                                       int64_t n, float *avg_out, void *local2, double *state3, void *stream) {
  return tce_p2p_allreduce_sumsq_range(world, rank, peer_bufs, peer_pads, 0, n, 0, 1, avg_out, local2, state3, stream);
}

extern "C" size_t tce_p2p_push_xchg_bytes(int world, int64_t n) {
  const size_t npad = ((size_t)n + 3) / 4 * 4;
  return P2P_FLAG_BYTES + 2 * (size_t)world * npad * sizeof(float);
}

extern "C" int tce_p2p_push_allreduce_sumsq(int world, int rank, void *const *peer_xchg, const float *grad, int64_t n_total,
                                            int64_t offset, int64_t n, int phase, int bump_step, float *avg_out,
                                            void *local2, double *state3, void *stream) {
  if (n == 0) return TCE_OK;
  if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || !peer_xchg || !grad || !avg_out || !local2 ||
      !state3 || n < 0 || (n & 3) || offset < 0 || (offset & 3) || offset + n > (n_total + 3) / 4 * 4 || phase < 0 ||
      phase >= TCE_P2P_MAX_PHASES || ((uintptr_t)avg_out & 15) || ((uintptr_t)grad & 15))
    return TCE_ERR_INVALID_ARGUMENT;
  PushTable pt;
  pt.world = world;
  pt.rank = rank;
  for (int r = 0; r < world; ++r) {
    if (!peer_xchg[r] || ((uintptr_t)peer_xchg[r] & 15)) return TCE_ERR_INVALID_ARGUMENT;
    pt.xchg[r] = static_cast<char *>(peer_xchg[r]);
  }
  long long blocks = (n / 4 + P2P_THREADS - 1) / P2P_THREADS;
  if (blocks > P2P_FLAG_BLOCKS) blocks = P2P_FLAG_BLOCKS;
  if (blocks < 1) blocks = 1;
  const long long npad = (n_total + 3) / 4 * 4;
  unsigned long long *local = static_cast<unsigned long long *>(local2) + 2 * phase;
#define TCE_P2P_PUSH(WT)                                                                                    \
  p2p_push_allreduce_sumsq_kernel<WT><<<(unsigned)blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(          \
      pt, grad, (long long)offset, (long long)n, npad, phase, bump_step, avg_out, local, state3)
  switch (world) {
    case 2: TCE_P2P_PUSH(2); break;
    case 4: TCE_P2P_PUSH(4); break;
    case 8: TCE_P2P_PUSH(8); break;
    default: TCE_P2P_PUSH(0); break;
  }
#undef TCE_P2P_PUSH
  TCE_CHECK_LAUNCH("p2p_push_allreduce_sumsq_kernel");
  return TCE_OK;
}
