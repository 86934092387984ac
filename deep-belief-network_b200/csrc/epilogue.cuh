// epilogue.cuh -- the two fused epilogues shared by the SIMT (strict FP32) and tcgen05 (BF16)
// contraction kernels.  Both kernels hand the epilogue groups of 4 consecutive accumulator
// columns (n4 % 4 == 0) of one row, so one Philox call serves 4 units and stores vectorise.
#pragma once
#include "dbn_common.cuh"

namespace dbn {

// Strict sigmoid: expf (not __expf / ex2.approx) so the FP32 mode stays within 1e-5 of the
// float64 oracle (reference: dbn/activations.py:29  1/(1+exp(-x))).
__device__ __forceinline__ float sigmoid_strict(float x) { return 1.0f / (1.0f + expf(-x)); }

// BF16 tensor-core mode: ex2.approx + rcp (2 MUFU ops, ~1e-6 relative) -- far inside the 2e-3 budget
// of that mode and ~6x fewer issue slots than expf + IEEE division, which matters because the
// epilogue warps must keep up with the tcgen05 mainloop.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + __expf(-x)); }

// ---- fixed-point column statistics (dbn_colsum) ----------------------------------------------------
// value * 2^32 as a two's-complement 64-bit integer: integer atomics commute, so the sums are deterministic
__device__ __forceinline__ unsigned long long stat_fixed(float x) {
  return (unsigned long long)__float2ll_rn(x * 4294967296.0f);     // exact scaling; |x| < 2^31
}
__device__ __forceinline__ float stat_value(unsigned long long v) {
  return (float)((double)(long long)v * (1.0 / 4294967296.0));
}
__device__ __forceinline__ void stat_add(unsigned long long* acc, int n, float x) {
  if (x != 0.0f) atomicAdd(acc + n, stat_fixed(x));
}
// Consume a statistic: the one thread that owns index i reads, clears and returns it.
__device__ __forceinline__ float stat_take(unsigned long long* acc, int i) {
  const unsigned long long v = acc[i];
  acc[i] = 0ull;
  return stat_value(v);
}

// Column sums of a 32 x 32 block held one ROW per lane (x[j] = element (lane, j)): a butterfly in which every
// step halves the columns a lane still carries -- 31 shuffles in all instead of 32 five-step reductions.
// Afterwards lane l holds the sum over the 32 rows of column l.
__device__ __forceinline__ float colsum32(float x[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int j = 0; j < step; ++j) {
      const float send = up ? x[j] : x[j + step];
      const float keep = up ? x[j + step] : x[j];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return x[0];
}

template <bool kFast>
__device__ __forceinline__ float apply_act(int act, float x) {
  if (act == DBN_ACT_SIGMOID) return kFast ? sigmoid_fast(x) : sigmoid_strict(x);
  if (act == DBN_ACT_RELU) return fmaxf(x, 0.0f);  // dbn/activations.py:49
  return x;
}

// Activation epilogue on 4 columns.  M, N = logical output extent (stores are masked to it).
// pre_bias: bias[n4..n4+3] if the caller already holds it (split-K kernel: requested before the exchange), else nullptr
template <bool kFast = false, bool kStats = true>
__device__ __forceinline__ void act_epilogue_quad(const dbn_act_epilogue& e, uint32_t epoch, int m, int n4,
                                                  const float acc[4], int M, int N, const float* pre_bias = nullptr) {
  if (m >= M || n4 >= N) return;
  const int nvalid = min(4, N - n4);
  float a[4];
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.bias != nullptr && pre_bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = pre_bias[j];
  } else if (e.bias != nullptr) {
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(e.bias + n4) & 15) == 0) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(e.bias + n4));
      bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) bv[j] = __ldg(e.bias + n4 + j);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = apply_act<kFast>(e.act, acc[j] + bv[j]);
  if (e.mul != DBN_MUL_NONE) {
    float g[4];
    mat_load4(e.aux, m, n4, g, nvalid);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (e.mul == DBN_MUL_PRIME_SIGMOID) a[j] *= g[j] * (1.0f - g[j]);  // dbn/activations.py:38
      else a[j] = (g[j] > 0.0f) ? a[j] : 0.0f;                            // dbn/activations.py:58
    }
  }
  float u[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.rng.mode != DBN_RNG_NONE) {
    if (e.rng.injected.ptr != nullptr) mat_load4(e.rng.injected, m, n4, u, nvalid);
    else rng_uniform4(e.rng, epoch, m, n4, u);
  }
  if (e.rng.mode == DBN_RNG_MASK) {
    const bool inj = e.rng.injected.ptr != nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool keep = inj ? (u[j] != 0.0f) : (u[j] < e.rng.keep_p);
      a[j] = keep ? a[j] : 0.0f;  // no 1/p rescale: dbn/models.py:409-411
    }
  }
  if (e.out.ptr != nullptr) mat_store4(e.out, m, n4, a, nvalid);
  if (e.out2.ptr != nullptr) mat_store4(e.out2, m, n4, a, nvalid);
  if (kStats && e.colsum.acc != nullptr) {   // slow path of the column statistics: one atomic per element
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.colsum.ref.ptr != nullptr) mat_load4(e.colsum.ref, e.colsum.ref_row0 + m, n4, r, nvalid);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = e.colsum.ref.ptr != nullptr ? r[j] - a[j] : a[j];
      if (j < nvalid) stat_add(e.colsum.acc, n4 + j, e.colsum.square ? x * x : e.colsum.sign * x);
    }
  }
  if (e.rng.mode == DBN_RNG_SAMPLE && e.sample.ptr != nullptr) {
    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = (u[j] < a[j]) ? 1.0f : 0.0f;  // strict '<': dbn/models.py:155
    mat_store4(e.sample, m, n4, s, nvalid);
  }
}

// Destination row of accumulator row m when the raw delta is scattered over peer GPUs (see dbn_update_epilogue).
__device__ __forceinline__ float* scatter_row_ptr(const dbn_update_epilogue& e, int m) {
  const int o = m / e.rows_per_peer;
  void* base = e.raw_peers[0];      // select chain, not an indexed read: the table stays in registers / the constant bank
#pragma unroll
  for (int j = 1; j < DBN_MAX_PEERS; ++j)
    if (o == j) base = e.raw_peers[j];
  return reinterpret_cast<float*>(base) +
         (size_t)(e.raw_slot * e.rows_per_peer + (m - o * e.rows_per_peer)) * (size_t)e.raw.ld;
}
// Bias statistics (dbn_colsum) consumed by the update: every index has exactly one owner thread in the launch.
__device__ __forceinline__ void upd_take_stats(const dbn_update_epilogue& e, int m, int n0, int ncols, int M, int N) {
  if (e.statM != nullptr && e.vecM != nullptr && n0 == 0 && m < M) e.vecM[m] += e.scale * stat_take(e.statM, m);
  if (e.statN != nullptr && e.vecN != nullptr && m == 0) {
    for (int j = 0; j < ncols; ++j)
      if (n0 + j < N) e.vecN[n0 + j] += e.scale * stat_take(e.statN, n0 + j);
  }
}

// Update epilogue on 4 columns of the (M+augM) x (N+augN) accumulator.
// pre_w: W[m, n4..n4+3] if the caller already holds it (whole quad inside the core block), else nullptr
template <bool kScatter = true>
__device__ __forceinline__ void upd_epilogue_quad(const dbn_update_epilogue& e, int m, int n4, const float acc[4],
                                                  int M, int N, int augM, int augN, const float* pre_w = nullptr) {
  const int Mt = M + augM, Nt = N + augN;
  if (m >= Mt || n4 >= Nt) return;
  const int nvalid = min(4, Nt - n4);
  if (kScatter && e.raw_peers[0] != nullptr) {   // scattered raw delta (no ones row / column on this path)
    if (m < M && n4 < N) {
      dbn_mat dst;
      dst.ptr = scatter_row_ptr(e, m); dst.ld = e.raw.ld; dst.dtype = DBN_F32;
      mat_store4(dst, 0, n4, acc, min(4, N - n4));
    }
    return;
  }
  if (e.raw.ptr != nullptr) {
    mat_store4(e.raw, m, n4, acc, nvalid);
    return;
  }
  upd_take_stats(e, m, n4, 4, M, N);
  if (m < M) {
    const int ncore = max(0, min(4, N - n4));
    if (ncore > 0) {
      float w[4];
      dbn_mat Wm;
      Wm.ptr = e.W; Wm.ld = e.ldw; Wm.dtype = DBN_F32;
      if (pre_w != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = pre_w[j];
      } else {
        mat_load4(Wm, m, n4, w, ncore);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = e.decay * w[j] + e.scale * acc[j];
      mat_store4(Wm, m, n4, w, ncore);
      if (e.Wb.ptr != nullptr) mat_store4(e.Wb, m, n4, w, ncore);
    }
    if (augN && e.vecM != nullptr) {
      const int j = N - n4;  // position of the ones column inside this quad
      if (j >= 0 && j < 4) e.vecM[m] += e.scale * acc[j];
    }
  } else if (m == M && e.vecN != nullptr) {  // ones row: visible-bias statistics
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n4 + j < N) e.vecN[n4 + j] += e.scale * acc[j];
  }
}

// Out-of-line copies of the quad epilogues for the ragged edges of the tensor-core kernels: ONE copy of the
// option tree per kernel instead of W/4 inlined ones.  The batch-256 launches run every instruction exactly once
// per CTA, straight from a cold instruction cache, so code bytes on and around the hot path are latency.
template <bool kFast>
__device__ __noinline__ void act_epilogue_quad_ool(const dbn_act_epilogue& e, uint32_t epoch, int m, int n4, float a0,
                                                   float a1, float a2, float a3, int M, int N) {
  const float acc[4] = {a0, a1, a2, a3};
  act_epilogue_quad<kFast>(e, epoch, m, n4, acc, M, N);
}
__device__ __noinline__ void upd_epilogue_quad_ool(const dbn_update_epilogue& e, int m, int n4, float a0, float a1, float a2,
                                                   float a3, int M, int N, int augM, int augN) {
  const float acc[4] = {a0, a1, a2, a3};
  upd_epilogue_quad(e, m, n4, acc, M, N, augM, augN);
}

// ---- W-column row segments (the tcgen05 epilogues: one thread = one accumulator row x W columns) --------
// W = 32: a thread drains 32 TMEM columns of its own lane.  W = 8: the split-K kernel's second phase, where
// the reduced tile sits in shared memory and 8 consecutive lanes cover 64 columns of one row.
// All loads are issued before any dependent math and all stores are 16-byte vectors, so the W independent
// element chains give the epilogue warps enough ILP; runtime switches are hoisted to one branch per segment.
// Falls back to the generic quad path on ragged edges.
__device__ __forceinline__ bool mat_vec_ok(const dbn_mat& m) {
  if (m.ptr == nullptr) return true;
  const int per16 = (m.dtype == DBN_F32) ? 4 : 8;
  return (m.ld % per16) == 0 && (reinterpret_cast<uintptr_t>(m.ptr) & 15) == 0;
}
template <int W>
__device__ __forceinline__ void row_load(const dbn_mat& m, int r, int c, float v[W]) {
  const size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (m.dtype == DBN_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(m.ptr) + i);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) {
      const float4 t = p[j];
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(m.ptr) + i);
#pragma unroll
    for (int j = 0; j < W / 8; ++j) {
      const uint4 t = p[j];
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack16x2(m.dtype, w[k], v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
    }
  }
}
template <int W>
__device__ __forceinline__ void row_store(const dbn_mat& m, int r, int c, const float v[W]) {
  const size_t i = (size_t)r * (size_t)m.ld + (size_t)c;
  if (m.dtype == DBN_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(m.ptr) + i);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) p[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(m.ptr) + i);
#pragma unroll
    for (int j = 0; j < W / 8; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = pack16x2(m.dtype, v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
      p[j] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// whole segment inside the logical output and every matrix it touches 16-byte addressable
template <int W>
__device__ __forceinline__ bool act_vec_ok(const dbn_act_epilogue& e, int m, int n0, int M, int N) {
  return (m < M) && (n0 + W <= N) && mat_vec_ok(e.out) && mat_vec_ok(e.out2) && mat_vec_ok(e.sample) &&
         (e.mul == DBN_MUL_NONE || mat_vec_ok(e.aux)) && (e.rng.injected.ptr == nullptr || mat_vec_ok(e.rng.injected)) &&
         (e.bias == nullptr || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0);
}

// pre_bias: the segment's bias values if the caller already loaded them (only honoured on the vector path,
// i.e. when act_vec_ok<W> held for the same arguments), else nullptr.
template <int W, bool kFast>
__device__ __forceinline__ void act_epilogue_row(const dbn_act_epilogue& e, uint32_t epoch, int m, int n0,
                                                 const float acc[W], int M, int N, const float* pre_bias = nullptr) {
  static_assert(W % 8 == 0, "segments are whole 16-byte bf16 vectors");
  const bool inj = e.rng.injected.ptr != nullptr;
  const bool vec = act_vec_ok<W>(e, m, n0, M, N);
  // column statistics: the butterfly needs the whole warp on the vector path (callers enter converged, one row
  // per lane); a warp on a ragged edge falls back to one atomic per element
  bool fast_stats = false;
  if (e.colsum.acc != nullptr) {
    if constexpr (W == 32) fast_stats = __all_sync(0xffffffffu, vec && mat_vec_ok(e.colsum.ref));
  }
  if (!vec) {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < W / 4; ++j)   // (unrolled: acc stays in registers; the callee is one shared copy)
      if (n0 + 4 * j < N) act_epilogue_quad_ool<kFast>(e, epoch, m, n0 + 4 * j, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3], M, N);
    return;
  }
  float a[W];
  if (e.bias != nullptr && pre_bias != nullptr) {
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = acc[j] + pre_bias[j];
  } else if (e.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
    for (int j = 0; j < W / 4; ++j) {
      const float4 t = __ldg(bp + j);
      a[4 * j] = acc[4 * j] + t.x; a[4 * j + 1] = acc[4 * j + 1] + t.y;
      a[4 * j + 2] = acc[4 * j + 2] + t.z; a[4 * j + 3] = acc[4 * j + 3] + t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = acc[j];
  }
  if (e.act == DBN_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = kFast ? sigmoid_fast(a[j]) : sigmoid_strict(a[j]);
  } else if (e.act == DBN_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = fmaxf(a[j], 0.0f);
  }
  if (e.mul != DBN_MUL_NONE) {
    float g[W];
    row_load<W>(e.aux, m, n0, g);
    if (e.mul == DBN_MUL_PRIME_SIGMOID) {
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] *= g[j] * (1.0f - g[j]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] = (g[j] > 0.0f) ? a[j] : 0.0f;
    }
  }
  float u[W];
  if (e.rng.mode != DBN_RNG_NONE) {
    if (inj) {
      row_load<W>(e.rng.injected, m, n0, u);
    } else {
#pragma unroll
      for (int j = 0; j < W / 4; ++j) rng_uniform4(e.rng, epoch, m, n0 + 4 * j, u + 4 * j);
    }
    if (e.rng.mode == DBN_RNG_MASK) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const bool keep = inj ? (u[j] != 0.0f) : (u[j] < e.rng.keep_p);
        a[j] = keep ? a[j] : 0.0f;
      }
    }
  }
  if (e.out.ptr != nullptr) row_store<W>(e.out, m, n0, a);
  if (e.out2.ptr != nullptr) row_store<W>(e.out2, m, n0, a);
  if (e.rng.mode == DBN_RNG_SAMPLE && e.sample.ptr != nullptr) {
#pragma unroll
    for (int j = 0; j < W; ++j) u[j] = (u[j] < a[j]) ? 1.0f : 0.0f;
    row_store<W>(e.sample, m, n0, u);
  }
  if (e.colsum.acc != nullptr) {   // bias statistics while the values are in registers (dbn/models.py:143-144)
    if (e.colsum.ref.ptr != nullptr) {
      float r[W];
      if (mat_vec_ok(e.colsum.ref)) {
        row_load<W>(e.colsum.ref, e.colsum.ref_row0 + m, n0, r);
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) r[j] = mat_load(e.colsum.ref, e.colsum.ref_row0 + m, n0 + j);
      }
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] = r[j] - a[j];
    }
    if (e.colsum.square) {
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] *= a[j];
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] *= e.colsum.sign;
    }
    if constexpr (W == 32) {
      if (fast_stats) {
        const int lane = (int)(threadIdx.x & 31);
        const float sum = colsum32(a, lane);
        stat_add(e.colsum.acc, n0 + lane, sum);
        return;
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) stat_add(e.colsum.acc, n0 + j, a[j]);
  }
}

template <int W>
__device__ __forceinline__ bool upd_vec_ok(const dbn_update_epilogue& e, int m, int n0, int M, int N) {
  if (e.raw_peers[0] != nullptr) return (m < M) && (n0 + W <= N) && (e.raw.ld & 3) == 0;   // peer buffers are 16-byte aligned
  return (m < M) && (n0 + W <= N) && mat_vec_ok(e.raw) && mat_vec_ok(e.Wb) &&
         (e.raw.ptr != nullptr || ((e.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(e.W) & 15) == 0));
}

// pre_w: the segment of the fp32 master weights if the caller already loaded it (vector path only), else nullptr.
template <int W>
__device__ __forceinline__ void upd_epilogue_row(const dbn_update_epilogue& e, int m, int n0, const float acc[W],
                                                 int M, int N, int augM, int augN, const float* pre_w = nullptr) {
  const bool vec = upd_vec_ok<W>(e, m, n0, M, N);
  if (!vec) {
    if (m >= M + augM) return;
#pragma unroll
    for (int j = 0; j < W / 4; ++j)
      if (n0 + 4 * j < N + augN) upd_epilogue_quad_ool(e, m, n0 + 4 * j, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3], M, N, augM, augN);
    return;
  }
  if (e.raw_peers[0] != nullptr) {
    dbn_mat dst;
    dst.ptr = scatter_row_ptr(e, m); dst.ld = e.raw.ld; dst.dtype = DBN_F32;
    row_store<W>(dst, 0, n0, acc);
    return;
  }
  if (e.raw.ptr != nullptr) {
    row_store<W>(e.raw, m, n0, acc);
    return;
  }
  upd_take_stats(e, m, n0, W, M, N);
  dbn_mat Wm;
  Wm.ptr = e.W; Wm.ld = e.ldw; Wm.dtype = DBN_F32;
  float w[W];
  if (pre_w != nullptr) {
#pragma unroll
    for (int j = 0; j < W; ++j) w[j] = pre_w[j];
  } else {
    row_load<W>(Wm, m, n0, w);   // all loads in flight before the first store
  }
#pragma unroll
  for (int j = 0; j < W; ++j) w[j] = fmaf(e.scale, acc[j], e.decay * w[j]);
  row_store<W>(Wm, m, n0, w);
  if (e.Wb.ptr != nullptr) row_store<W>(e.Wb, m, n0, w);
}

}  // namespace dbn
