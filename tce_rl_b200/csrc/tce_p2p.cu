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

// local: {counter (sequence number of the last launch), blocks finished}; state: {step, sum g^2, error flag}
__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_sumsq_kernel(PeerTable pt, long long n, float *__restrict__ avg_out, unsigned long long *local,
                           double *__restrict__ state) {
  __shared__ double red[32];
  __shared__ int s_ok, s_last;
  const int W = pt.world;
  const unsigned long long seq = local[0] + 1;           // written back by the last block (all blocks read the old value
  unsigned long long *mypad = pt.pad[pt.rank];           //  first: the write happens after a grid-wide ticket)
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(pt.pad[threadIdx.x] + pt.rank, seq);      // 1. arrive
  if (!wait_slots(mypad, 0, W, seq)) s_ok = 0;                                                       // 2. wait
  __syncthreads();
  double ss = 0.0;
  if (s_ok) {                                                                                        // 3. reduce
    const float inv = 1.0f / (float)W;
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 a = ld_peer(pt.buf[0] + 4 * i);
      for (int r = 1; r < W; ++r) {
        const float4 b = ld_peer(pt.buf[r] + 4 * i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      *reinterpret_cast<float4 *>(avg_out + 4 * i) = a;
      ss += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {            // tail (n is a multiple of 4 in practice)
      const long long i = (n4 << 2) + threadIdx.x;
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
    state[0] += 1.0;
    *reinterpret_cast<unsigned int *>(local + 1) = 0u;
    local[0] = seq;
  }
  __threadfence_system();
  if (threadIdx.x < W) st_release_sys(pt.pad[threadIdx.x] + W + pt.rank, seq);                        // 4. depart
  if (!wait_slots(mypad, W, W, seq) && threadIdx.x == 0) state[2] = 1.0;
}

}  // namespace

extern "C" int tce_p2p_allreduce_sumsq(int world, int rank, const void *const *peer_bufs, void *const *peer_pads,
                                       int64_t n, float *avg_out, void *local2, double *state3, void *stream) {
  if (n == 0) return TCE_OK;
  if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || !peer_bufs || !peer_pads || !avg_out ||
      !local2 || !state3 || n < 0 || ((uintptr_t)avg_out & 15))
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
  p2p_allreduce_sumsq_kernel<<<(unsigned)blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(
      pt, n, avg_out, static_cast<unsigned long long *>(local2), state3);
  TCE_CHECK_LAUNCH("p2p_allreduce_sumsq_kernel");
  return TCE_OK;
}
