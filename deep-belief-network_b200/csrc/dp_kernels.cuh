// dp_kernels.cuh -- the data-parallel exchange of the CD step over NVLink peer memory (no NCCL on the data path).
//
// The reference accumulates the deltas of a mini-batch sample by sample and updates once
// (dbn/models.py:112-119).  Across GPUs every rank holds a partial sum over ITS rows of the mini-batch; the sum
// over ranks is a reduce-scatter (rank r ends up with the rows of dW it owns), the update is sharded, and the
// refreshed bf16 weights go back to everybody (all-gather).  Here
//   * the reduce-scatter is fused into the dW contraction: its epilogue stores each accumulator tile into slot
//     `rank` of the owner's receive buffer over NVLink (umma_gemm.cuh, UmmaUpdEpi::apply_scatter), so the wire time
//     sits under the mainloop of the following tiles;
//   * dp_signal_kernel ships the (fixed-point, hence exactly additive) bias statistics and raises one flag per peer;
//   * dp_apply_kernel is the owner's side: wait for the flags, sum the slots in rank order (deterministic), update the
//     fp32 master rows, and store their bf16 rounding straight into EVERY rank's shadow -- the all-gather is the
//     update's own store; then raise the "weights in" flags;
//   * dp_wait_kernel holds the next step's first contraction until every owner's rows have arrived.
// Flags are monotonically increasing 64-bit step counters written with st.release.sys and polled with
// ld.acquire.sys on LOCAL memory (a rank never polls over the link).  Each rank runs on its own GPU, so the kernels
// that wait on one another are co-resident by construction; waits are bounded and trap instead of hanging.
#pragma once
#include "epilogue.cuh"

namespace dbn {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait until *p >= value.  A peer may legitimately be seconds late (first step of a fit, host jitter); after 60 s
// the protocol is considered broken and the kernel traps rather than occupying the GPU for ever.
__device__ __forceinline__ void dp_spin(const unsigned long long* p, unsigned long long value, int who) {
  const unsigned long long t0 = global_ns();
  unsigned backoff = 32;
  while (ld_acquire_sys(p) < value) {
    __nanosleep(backoff);
    if (backoff < 1024) backoff <<= 1;
    if (global_ns() - t0 > 60ull * 1000000000ull) {
      printf("dbn_b200: data-parallel flag wait timed out (peer %d, waiting for %llu, have %llu)\n", who, value,
             ld_acquire_sys(p));
      __trap();
    }
  }
}

__device__ __forceinline__ unsigned long long* dp_flag(const dbn_dp_exchange& x, int on_rank, int which, int from_rank) {
  return reinterpret_cast<unsigned long long*>(x.flags[on_rank]) + which * DBN_MAX_PEERS + from_rank;
}

// One block per peer: statistics -> slot `rank` of that peer, then that peer's "deltas out" flag.  The flag also
// covers the dW tiles the preceding contraction kernel stored into the peer (stream order: that kernel has completed).
__global__ void __launch_bounds__(512) dp_signal_kernel(dbn_dp_exchange x, const unsigned long long* __restrict__ local_stats,
                                                        unsigned long long value) {
  const int peer = (int)blockIdx.x;
  const int n_stats = x.H + x.V;
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(x.stats[peer]) + (size_t)x.rank * n_stats;
  for (int i = threadIdx.x; i < n_stats; i += blockDim.x) dst[i] = local_stats[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(dp_flag(x, peer, 0, x.rank), value);
}

__global__ void __launch_bounds__(256) dp_apply_kernel(dbn_dp_exchange x, float scale, float* __restrict__ W, int ldw,
                                                       float* __restrict__ b, float* __restrict__ c,
                                                       unsigned long long* __restrict__ local_stats, unsigned int* counter,
                                                       unsigned long long value) {
  const int n = x.n_peers, me = x.rank, hs = x.rows_per_peer, V = x.V, H = x.H;
  if ((int)threadIdx.x < n) dp_spin(dp_flag(x, me, 0, (int)threadIdx.x), value, (int)threadIdx.x);
  __syncthreads();
  const float* recv = reinterpret_cast<const float*>(x.recv[me]);
  const int ldr = x.ld_recv;
  const int groups = (V + 7) >> 3;               // 8 columns = one 16-byte bf16 vector per thread and iteration
  const long total = (long)hs * groups;
  const long stride = (long)gridDim.x * blockDim.x;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (ldw & 3) == 0 && (ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  for (long t = gtid; t < total; t += stride) {
    const int r = (int)(t / groups), col = (int)(t % groups) << 3;
    const int nv = min(8, V - col);
    const int row = me * hs + r;                 // row of W
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float w[8];
    if (vec && nv == 8) {
      for (int j = 0; j < n; ++j) {              // fixed order: every run and every replica sums identically
        const float4* p = reinterpret_cast<const float4*>(recv + ((size_t)(j * hs + r) * ldr + col));
        const float4 a0 = __ldcg(p), a1 = __ldcg(p + 1);
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
        acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
      }
      float4* wp = reinterpret_cast<float4*>(W + (size_t)row * ldw + col);
      const float4 w0 = wp[0], w1 = wp[1];
      w[0] = fmaf(scale, acc[0], w0.x); w[1] = fmaf(scale, acc[1], w0.y); w[2] = fmaf(scale, acc[2], w0.z);
      w[3] = fmaf(scale, acc[3], w0.w); w[4] = fmaf(scale, acc[4], w1.x); w[5] = fmaf(scale, acc[5], w1.y);
      w[6] = fmaf(scale, acc[6], w1.z); w[7] = fmaf(scale, acc[7], w1.w);
      wp[0] = make_float4(w[0], w[1], w[2], w[3]);
      wp[1] = make_float4(w[4], w[5], w[6], w[7]);
    } else {
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < nv; ++i) acc[i] += __ldcg(recv + ((size_t)(j * hs + r) * ldr + col + i));
      for (int i = 0; i < 8; ++i) w[i] = 0.0f;
      for (int i = 0; i < nv; ++i) {
        float* wp = W + (size_t)row * ldw + col + i;
        w[i] = fmaf(scale, acc[i], *wp);
        *wp = w[i];
      }
    }
    // all-gather fused into the update: the bf16 row segment goes into every rank's shadow
    const int sdt = x.Wb[me].dtype;      // every rank's shadow has the same 16-bit type
    const uint4 pk = make_uint4(pack16x2(sdt, w[0], w[1]), pack16x2(sdt, w[2], w[3]), pack16x2(sdt, w[4], w[5]),
                                pack16x2(sdt, w[6], w[7]));
    for (int j = 0; j < n; ++j) {
      const dbn_mat& S = x.Wb[j];
      if (nv == 8 && (S.ld & 7) == 0) {
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(S.ptr) + (size_t)row * S.ld + col) = pk;
      } else {
        for (int i = 0; i < nv; ++i) mat_store(S, row, col + i, w[i]);
      }
    }
  }
  // biases: replicated, every rank adds the same exact integer sum of the per-rank statistics (dc | db)
  const int n_stats = H + V;
  const unsigned long long* slots = reinterpret_cast<const unsigned long long*>(x.stats[me]);
  for (long i = gtid; i < n_stats; i += stride) {
    unsigned long long s = 0ull;
    for (int j = 0; j < n; ++j) s += __ldcg(slots + (size_t)j * n_stats + i);
    const float d = scale * stat_value(s);
    if (i < H) c[i] += d;
    else b[i - H] += d;
    local_stats[i] = 0ull;                        // this rank's accumulators: clean for the next step
  }
  // "weights in": raised by the last block, after every block's peer stores have been fenced
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {
      *counter = 0u;
      __threadfence_system();
      for (int j = 0; j < n; ++j) st_release_sys(dp_flag(x, j, 1, me), value);
    }
  }
}

__global__ void dp_wait_kernel(dbn_dp_exchange x, int which, unsigned long long value) {
  if ((int)threadIdx.x < x.n_peers) dp_spin(dp_flag(x, x.rank, which, (int)threadIdx.x), value, (int)threadIdx.x);
}

// dc / db statistics -> column V / row H of a raw delta buffer [(H+1), ld] (parity tests, NCCL fallback), cleared after.
__global__ void __launch_bounds__(256) stats_to_raw_kernel(unsigned long long* stats, int H, int V, dbn_mat raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H + V) return;
  const float v = stat_take(stats, i);
  float* R = reinterpret_cast<float*>(raw.ptr);
  if (i < H) R[(size_t)i * raw.ld + V] = v;
  else R[(size_t)H * raw.ld + (i - H)] = v;
}

// Row-block form: dc of rows [h0, h1) and, when h1 == H, db.
__global__ void __launch_bounds__(256) stats_rows_to_raw_kernel(unsigned long long* stats, int H, int V, int h0, int h1,
                                                                dbn_mat raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float* R = reinterpret_cast<float*>(raw.ptr);
  if (i < h1 - h0) {
    R[(size_t)(h0 + i) * raw.ld + V] = stat_take(stats, h0 + i);
  } else if (h1 == H && i - (h1 - h0) < V) {
    const int v = i - (h1 - h0);
    R[(size_t)H * raw.ld + v] = stat_take(stats, H + v);
  }
}

struct PeerSrcs {
  const void* p[DBN_MAX_PEERS];
};

// Epoch shuffle over NVLink: one warp per destination row, 512 contiguous bytes per instruction.
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(PeerSrcs srcs, int rows_per_src, int ld,
                                                               const int32_t* __restrict__ idx, int n_rows, dbn_mat dst) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = ld >> 3;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_rows; i += warps) {
    const int g = idx[i];
    const int s = g / rows_per_src, r = g - s * rows_per_src;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(srcs.p[s]) + (size_t)r * ld);
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(dst.ptr) + (size_t)i * dst.ld);
    for (int v = lane; v < vecs; v += 32) d[v] = src[v];
  }
}

}  // namespace dbn
