// rbm_persistent.cuh -- persistent CD-k kernel for layers that fit on chip (strict FP32).
//
// One thread-block cluster trains the RBM for a WHOLE EPOCH in a single launch: the weight matrix
// stays resident in (distributed) shared memory across all mini-batches, the hidden units are sharded
// over the cluster's CTAs, and every phase of every CD-k step -- gather of the shuffled rows, hidden
// means, Philox/injected Bernoulli sampling, visible means, deltas and the in-place update -- runs
// on chip.  Replaces the epoch loop of BinaryRBM._stochastic_gradient_descent (dbn/models.py:105-119)
// for the small configurations (README digits 64-[256,256], the 13-[100] regression net) where a
// step is a few hundred kFLOP and a launch per contraction costs more than the arithmetic.
//
// Sharding ("on-chip CD-k"): CTA r owns hidden units [r*Hc, (r+1)*Hc):
//   v -> h   local (each CTA holds the whole mini-batch V and its rows of W)
//   h -> v   split-K over the cluster: each CTA contracts its hidden slice into a partial [B,V] sum,
//            then a reduce-scatter + all-gather through distributed shared memory gives every CTA
//            the full visible means
//   update   local for W rows / c; the visible bias b is replicated and updated identically
// Arithmetic: FFMA, expf sigmoid, same Philox counters (global row, global unit) as the tiled path,
// so sampled states are identical to the multi-launch path given the same seed.
#pragma once
#include <cooperative_groups.h>

#include "epilogue.cuh"

namespace dbn {
namespace cg = cooperative_groups;

struct RbmPersistArgs {
  const float* X; int ldx;      // un-permuted fp32 dataset [N, ldx]
  const int32_t* idx;           // epoch order: row i of the epoch is X[idx[i]]
  int n_rows, batch_size;
  int V, H, k, act;
  float lr_over_bs;
  float* W; int ldw;
  float* b;
  float* c;
  dbn_rng rng;                  // seed / epoch / stream base / row0; injected = U [n_rows, k*H] or NULL
  int Hc;                       // hidden units per CTA (multiple of 4)
};

constexpr int RP_THREADS = 512;

struct RpSmem {
  int Vx;                       // row stride of the [B, V] buffers: V rounded up to 4 (float4 rows)
  int Vp;                       // row stride of Ws: a multiple of 4 with Vp/4 odd, so that the 8 lanes of a
                                // quarter-warp reading float4 from 8 consecutive rows hit all 32 banks
  size_t off_W, off_c, off_b, off_V0, off_Vt, off_part, off_P0, off_Pk, off_Hs, total;
};
__host__ __device__ inline RpSmem rp_layout(int V, int Hc, int B) {
  RpSmem s;
  s.Vx = (V + 3) & ~3;
  int q = s.Vx / 4;
  if ((q & 1) == 0) q += 1;
  s.Vp = 4 * q;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~size_t(3); return r; };
  s.off_W = take((size_t)Hc * s.Vp);
  s.off_c = take(Hc);
  s.off_b = take(s.Vx);
  s.off_V0 = take((size_t)B * s.Vx);
  s.off_Vt = take((size_t)B * s.Vx);
  s.off_part = take((size_t)B * s.Vx);
  s.off_P0 = take((size_t)B * Hc);
  s.off_Pk = take((size_t)B * Hc);
  s.off_Hs = take((size_t)B * Hc);
  s.total = o * sizeof(float);
  return s;
}

// hidden means of this CTA's units: P[b][hl] = act(c[hl] + sum_v Vin[b][v] W[hl][v])   (dbn/models.py:176-183)
// Register tile: one hidden unit x 4 batch rows per thread, float4 along v.  Lanes run over consecutive
// units (conflict-free W rows by the Vp padding); the 4 input rows are broadcast loads.
__device__ __forceinline__ void rp_hidden(const float* Vin, const float* Ws, const float* cs, float* P, int Bc, int Vx, int Vp,
                                          int Hc, int h_valid, int act) {
  const int groups = (Bc + 3) >> 2;
  for (int o = threadIdx.x; o < groups * Hc; o += RP_THREADS) {
    const int hl = o % Hc, b0 = (o / Hc) << 2;
    const float4* w = reinterpret_cast<const float4*>(Ws + (size_t)hl * Vp);
    const float4* x0 = reinterpret_cast<const float4*>(Vin + (size_t)min(b0, Bc - 1) * Vx);
    const float4* x1 = reinterpret_cast<const float4*>(Vin + (size_t)min(b0 + 1, Bc - 1) * Vx);
    const float4* x2 = reinterpret_cast<const float4*>(Vin + (size_t)min(b0 + 2, Bc - 1) * Vx);
    const float4* x3 = reinterpret_cast<const float4*>(Vin + (size_t)min(b0 + 3, Bc - 1) * Vx);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int v = 0; v < (Vx >> 2); ++v) {   // columns V..Vx-1 of W and of the inputs are zero
      const float4 wv = w[v];
      float4 t = x0[v];
      a0 = fmaf(t.x, wv.x, a0); a0 = fmaf(t.y, wv.y, a0); a0 = fmaf(t.z, wv.z, a0); a0 = fmaf(t.w, wv.w, a0);
      t = x1[v];
      a1 = fmaf(t.x, wv.x, a1); a1 = fmaf(t.y, wv.y, a1); a1 = fmaf(t.z, wv.z, a1); a1 = fmaf(t.w, wv.w, a1);
      t = x2[v];
      a2 = fmaf(t.x, wv.x, a2); a2 = fmaf(t.y, wv.y, a2); a2 = fmaf(t.z, wv.z, a2); a2 = fmaf(t.w, wv.w, a2);
      t = x3[v];
      a3 = fmaf(t.x, wv.x, a3); a3 = fmaf(t.y, wv.y, a3); a3 = fmaf(t.z, wv.z, a3); a3 = fmaf(t.w, wv.w, a3);
    }
    const float acc[4] = {a0, a1, a2, a3};
    const float cb = cs[hl];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (b0 + j < Bc) P[(size_t)(b0 + j) * Hc + hl] = (hl < h_valid) ? apply_act<false>(act, acc[j] + cb) : 0.0f;
  }
}

__global__ void __launch_bounds__(RP_THREADS, 1) rbm_persistent_kernel(RbmPersistArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), csz = (int)cluster.num_blocks();
  extern __shared__ float4 rp_smem_raw[];
  float* smem = reinterpret_cast<float*>(rp_smem_raw);
  const int V = a.V, Hc = a.Hc, B = a.batch_size;
  const RpSmem L = rp_layout(V, Hc, B);
  const int Vp = L.Vp, Vx = L.Vx;
  float* Ws = smem + L.off_W;  float* cs = smem + L.off_c;  float* bs = smem + L.off_b;
  float* V0 = smem + L.off_V0; float* Vt = smem + L.off_Vt; float* part = smem + L.off_part;
  float* P0 = smem + L.off_P0; float* Pk = smem + L.off_Pk; float* Hs = smem + L.off_Hs;
  const int h0 = rank * Hc;                                   // first global hidden unit of this CTA
  const int h_valid = max(0, min(Hc, a.H - h0));              // units that really exist
  const uint32_t epoch = a.rng.epoch + (a.rng.epoch_ptr ? *a.rng.epoch_ptr : 0u);

  // ---- parameters -> shared memory (stay there for the whole epoch) ----
  for (int o = threadIdx.x; o < Hc * Vp; o += RP_THREADS) {   // padding columns are zero and stay zero
    const int hl = o / Vp, v = o % Vp;
    Ws[o] = (hl < h_valid && v < V) ? a.W[(size_t)(h0 + hl) * a.ldw + v] : 0.0f;
  }
  for (int o = threadIdx.x; o < Hc; o += RP_THREADS) cs[o] = (o < h_valid) ? a.c[h0 + o] : 0.0f;
  for (int o = threadIdx.x; o < Vx; o += RP_THREADS) bs[o] = (o < V) ? a.b[o] : 0.0f;
  for (int o = threadIdx.x; o < B * Vx; o += RP_THREADS) { V0[o] = 0.0f; Vt[o] = 0.0f; }
  __syncthreads();
  cluster.sync();

  for (int r0 = 0; r0 < a.n_rows; r0 += B) {
    const int Bc = min(B, a.n_rows - r0);                     // short last batch (dbn/utils.py:14-22)
    // ---- gather the mini-batch: epoch row r0+b is X[idx[r0+b]] (double shuffle composed on the host) ----
    for (int o = threadIdx.x; o < Bc * V; o += RP_THREADS) {
      const int bi = o / V, v = o % V;
      V0[(size_t)bi * Vx + v] = a.X[(size_t)a.idx[r0 + bi] * a.ldx + v];
    }
    __syncthreads();

    for (int t = 0; t < a.k; ++t) {
      // h_t ~ P(h | v_t): means, then strict u < p sampling (dbn/models.py:135, :148-155)
      float* P = (t == 0) ? P0 : Pk;
      rp_hidden(t == 0 ? V0 : Vt, Ws, cs, P, Bc, Vx, Vp, Hc, h_valid, a.act);
      __syncthreads();
      for (int q = threadIdx.x; q < Bc * (Hc >> 2); q += RP_THREADS) {
        const int bi = q / (Hc >> 2), hq = (q % (Hc >> 2)) << 2;
        float u[4];
        if (a.rng.injected.ptr != nullptr) {
          const float* U = reinterpret_cast<const float*>(a.rng.injected.ptr) + (size_t)(r0 + bi) * a.rng.injected.ld +
                           (size_t)t * a.H + h0 + hq;
#pragma unroll
          for (int j = 0; j < 4; ++j) u[j] = (hq + j < h_valid) ? U[j] : 1.0f;
        } else {
          dbn_rng g = a.rng;
          g.stream = a.rng.stream + (uint32_t)t;
          g.row0 = a.rng.row0 + (uint32_t)r0;
          g.col0 = a.rng.col0 + (uint32_t)h0;
          rng_uniform4(g, epoch, bi, hq, u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Hs[(size_t)bi * Hc + hq + j] = (u[j] < P[(size_t)bi * Hc + hq + j]) ? 1.0f : 0.0f;
      }
      __syncthreads();
      // v_{t+1}: split-K partial over this CTA's hidden slice (dbn/models.py:136, :195-201).
      // Register tile: 4 visible units (one float4 of a W row) x 4 batch rows per thread.
      {
        const int vq = Vx >> 2, groups = (Bc + 3) >> 2;
        for (int o = threadIdx.x; o < groups * vq; o += RP_THREADS) {
          const int v4 = o % vq, b0 = (o / vq) << 2;
          const float* hr0 = Hs + (size_t)min(b0, Bc - 1) * Hc;
          const float* hr1 = Hs + (size_t)min(b0 + 1, Bc - 1) * Hc;
          const float* hr2 = Hs + (size_t)min(b0 + 2, Bc - 1) * Hc;
          const float* hr3 = Hs + (size_t)min(b0 + 3, Bc - 1) * Hc;
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
          for (int hl = 0; hl < Hc; ++hl) {
            const float4 wv = *reinterpret_cast<const float4*>(Ws + (size_t)hl * Vp + 4 * v4);
            const float g0 = hr0[hl], g1 = hr1[hl], g2 = hr2[hl], g3 = hr3[hl];
            s0.x = fmaf(g0, wv.x, s0.x); s0.y = fmaf(g0, wv.y, s0.y); s0.z = fmaf(g0, wv.z, s0.z); s0.w = fmaf(g0, wv.w, s0.w);
            s1.x = fmaf(g1, wv.x, s1.x); s1.y = fmaf(g1, wv.y, s1.y); s1.z = fmaf(g1, wv.z, s1.z); s1.w = fmaf(g1, wv.w, s1.w);
            s2.x = fmaf(g2, wv.x, s2.x); s2.y = fmaf(g2, wv.y, s2.y); s2.z = fmaf(g2, wv.z, s2.z); s2.w = fmaf(g2, wv.w, s2.w);
            s3.x = fmaf(g3, wv.x, s3.x); s3.y = fmaf(g3, wv.y, s3.y); s3.z = fmaf(g3, wv.z, s3.z); s3.w = fmaf(g3, wv.w, s3.w);
          }
          const float4 sv[4] = {s0, s1, s2, s3};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b0 + j < Bc) *reinterpret_cast<float4*>(part + (size_t)(b0 + j) * Vx + 4 * v4) = sv[j];
        }
      }
      cluster.sync();
      // reduce-scatter + all-gather through distributed shared memory: this CTA finishes elements
      // [e0, e1) of the [Bc, V] means (sum of all partials + bias, activation) and writes them to every CTA
      {
        const int E = Bc * Vx;
        const int per = (E + csz - 1) / csz;
        const int e0 = rank * per, e1 = min(E, e0 + per);
        for (int e = e0 + threadIdx.x; e < e1; e += RP_THREADS) {
          const int v = e % Vx;
          float s = bs[v];
          for (int r = 0; r < csz; ++r) s += cluster.map_shared_rank(part, r)[e];
          const float val = (v < V) ? apply_act<false>(a.act, s) : 0.0f;   // padding columns stay zero
          for (int r = 0; r < csz; ++r) cluster.map_shared_rank(Vt, r)[e] = val;
        }
      }
      cluster.sync();
    }
    // h_k = P(h | v_k) (dbn/models.py:141); k == 0 never reaches here (host guards it)
    rp_hidden(Vt, Ws, cs, Pk, Bc, Vx, Vp, Hc, h_valid, a.act);
    __syncthreads();
    // dW = h_0^T v_0 - h_k^T v_k; W += lr/bs dW and the two biases (dbn/models.py:142-144, :117-119)
    {   // register tile: one float4 of a W row per thread, batch rows streamed (P0/Pk are broadcast loads)
      const int vq = Vx >> 2;
      for (int o = threadIdx.x; o < Hc * vq; o += RP_THREADS) {
        const int v4 = o % vq, hl = o / vq;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int bi = 0; bi < Bc; ++bi) {
          const float p0 = P0[(size_t)bi * Hc + hl], pk = -Pk[(size_t)bi * Hc + hl];
          const float4 x0 = *reinterpret_cast<const float4*>(V0 + (size_t)bi * Vx + 4 * v4);
          const float4 xk = *reinterpret_cast<const float4*>(Vt + (size_t)bi * Vx + 4 * v4);
          acc.x = fmaf(p0, x0.x, acc.x); acc.y = fmaf(p0, x0.y, acc.y); acc.z = fmaf(p0, x0.z, acc.z); acc.w = fmaf(p0, x0.w, acc.w);
          acc.x = fmaf(pk, xk.x, acc.x); acc.y = fmaf(pk, xk.y, acc.y); acc.z = fmaf(pk, xk.z, acc.z); acc.w = fmaf(pk, xk.w, acc.w);
        }
        float4* wp = reinterpret_cast<float4*>(Ws + (size_t)hl * Vp + 4 * v4);
        float4 w = *wp;
        w.x = fmaf(a.lr_over_bs, acc.x, w.x); w.y = fmaf(a.lr_over_bs, acc.y, w.y);
        w.z = fmaf(a.lr_over_bs, acc.z, w.z); w.w = fmaf(a.lr_over_bs, acc.w, w.w);
        *wp = w;   // padding columns: inputs are zero there, so they stay zero
      }
    }
    for (int v = threadIdx.x; v < V; v += RP_THREADS) {   // every CTA keeps its own identical copy of b
      float dbv = 0.0f;
      for (int bi = 0; bi < Bc; ++bi) dbv += V0[(size_t)bi * Vx + v] - Vt[(size_t)bi * Vx + v];
      bs[v] = fmaf(a.lr_over_bs, dbv, bs[v]);
    }
    for (int hl = threadIdx.x; hl < Hc; hl += RP_THREADS) {   // strided: Hc may exceed the block (H up to 8192 over <= 8 CTAs)
      float dcv = 0.0f;
      for (int bi = 0; bi < Bc; ++bi) dcv += P0[(size_t)bi * Hc + hl] - Pk[(size_t)bi * Hc + hl];
      cs[hl] = fmaf(a.lr_over_bs, dcv, cs[hl]);
    }
    __syncthreads();
  }

  // ---- trained parameters back to HBM ----
  for (int o = threadIdx.x; o < Hc * V; o += RP_THREADS) {
    const int hl = o / V, v = o % V;
    if (hl < h_valid) a.W[(size_t)(h0 + hl) * a.ldw + v] = Ws[(size_t)hl * Vp + v];
  }
  for (int o = threadIdx.x; o < h_valid; o += RP_THREADS) a.c[h0 + o] = cs[o];
  if (rank == 0)
    for (int o = threadIdx.x; o < V; o += RP_THREADS) a.b[o] = bs[o];
  cluster.sync();   // nobody retires while a peer may still read its partial sums
}

// Cluster size / hidden slice for a shape, or 0 if the layer does not fit on chip.
inline int rp_plan(int V, int H, int B, int* Hc_out) {
  if (V <= 0 || H <= 0 || B <= 0 || H > 8 * 1024) return 0;
  for (int csz = 1; csz <= 8; csz *= 2) {
    int Hc = (ceil_div(H, csz) + 3) & ~3;
    const size_t bytes = rp_layout(V, Hc, B).total;
    const bool want_more = csz < 8 && Hc >= 64;    // keep splitting while each CTA would still have >= 32 units
    if (bytes <= 200 * 1024 && !want_more) {
      *Hc_out = Hc;
      return csz;
    }
  }
  return 0;
}

inline int rbm_persistent_launch(cudaStream_t st, const RbmPersistArgs& args_in, int V, int H, int B) {
  RbmPersistArgs a = args_in;
  int Hc = 0;
  const int csz = rp_plan(V, H, B, &Hc);
  if (csz == 0) return fail(DBN_ERR_UNSUPPORTED, "layer does not fit the on-chip persistent kernel%s%s");
  a.Hc = Hc;
  const size_t smem = rp_layout(V, Hc, B).total;
  DBN_CUDA_OK(cudaFuncSetAttribute(rbm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(csz);
  cfg.blockDim = dim3(RP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DBN_CUDA_OK(cudaLaunchKernelEx(&cfg, rbm_persistent_kernel, a));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

}  // namespace dbn
