// misc_kernels.cuh -- the memory-bound helpers around the contractions: row gather (epoch
// shuffle), scale/cast (p-rescale, bf16 shadows), ones columns, reconstruction error, output head.
#pragma once
#include "epilogue.cuh"

namespace dbn {

// dst[i, :] = src[idx ? idx[i] : i, :] (optionally * Bernoulli(keep_p) input dropout), dst[i, ones_col] = 1.
// One warp-row per block row: 256 threads stride over the columns of `rows_per_block` rows.
__global__ void __launch_bounds__(256) gather_rows_kernel(dbn_mat src, const int32_t* __restrict__ idx, int n_rows,
                                                          int n_cols, dbn_mat dst, int ones_col, dbn_rng rng) {
  const uint32_t epoch = rng.epoch + (rng.epoch_ptr ? *rng.epoch_ptr : 0u);
  const int quads = (n_cols + 3) >> 2;
  for (int i = blockIdx.x; i < n_rows; i += gridDim.x) {
    const int s = idx ? idx[i] : i;
    for (int q = threadIdx.x; q < quads; q += blockDim.x) {
      const int c = q << 2;
      const int nv = min(4, n_cols - c);
      float v[4];
      mat_load4(src, s, c, v, nv);
      if (rng.mode == DBN_RNG_MASK) {
        float u[4];
        if (rng.injected.ptr) {
          mat_load4(rng.injected, i, c, u, nv);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (u[j] != 0.0f) ? v[j] : 0.0f;
        } else {
          rng_uniform4(rng, epoch, i, c, u);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (u[j] < rng.keep_p) ? v[j] : 0.0f;
        }
      }
      mat_store4(dst, i, c, v, nv);
    }
    if (ones_col >= 0 && threadIdx.x == 0) mat_store(dst, i, ones_col, 1.0f);
  }
}

__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ src, int lds, int rows, int cols,
                                                         float scale, dbn_mat dst) {
  const int quads = (cols + 3) >> 2;
  const long total = (long)rows * quads;
  dbn_mat s;
  s.ptr = const_cast<float*>(src); s.ld = lds; s.dtype = DBN_F32;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int r = (int)(t / quads), c = (int)(t % quads) << 2;
    const int nv = min(4, cols - c);
    float v[4];
    mat_load4(s, r, c, v, nv);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= scale;
    mat_store4(dst, r, c, v, nv);
  }
}

__global__ void fill_ones_column_kernel(dbn_mat m, int rows, int col) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) mat_store(m, r, col, 1.0f);
}

// out[0] += inv_denom * sum (A[r,c] - B[row0_B + r, c])^2
__global__ void __launch_bounds__(256) sum_sq_diff_kernel(dbn_mat A, dbn_mat B, int row0_B, int rows, int cols,
                                                          float inv_denom, float* out) {
  const int quads = (cols + 3) >> 2;
  const long total = (long)rows * quads;
  float acc = 0.0f;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int r = (int)(t / quads), c = (int)(t % quads) << 2;
    const int nv = min(4, cols - c);
    float a[4], b[4];
    mat_load4(A, r, c, a, nv);
    mat_load4(B, row0_B + r, c, b, nv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = (j < nv) ? (a[j] - b[j]) : 0.0f;
      acc = fmaf(d, d, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sums[8];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += warp_sums[w];
    atomicAdd(out, s * inv_denom);
  }
}

// One warp per sample row.  softmax (max-shifted; equal to the reference's un-shifted form
// dbn/models.py:603-605 within rounding) or identity; delta = out - y (dbn/models.py:617-626,
// :713-720); loss = -log p[label] (:673-680) or sum_t (pred - y)^2 (:733-741).
__global__ void __launch_bounds__(128) head_delta_kernel(int head, float* Z, int ldz, const float* __restrict__ Y,
                                                         int ldy, int B, int C, float* delta, int ldd,
                                                         float* loss_rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  float* z = Z + (size_t)warp * ldz;
  const bool train = (Y != nullptr) && (delta != nullptr);  // inference: outputs only
  const float* y = train ? Y + (size_t)warp * ldy : nullptr;
  float* d = train ? delta + (size_t)warp * ldd : nullptr;
  float loss = 0.0f;
  if (head == DBN_HEAD_SOFTMAX) {
    float mx = -INFINITY;
    for (int j = lane; j < C; j += 32) mx = fmaxf(mx, z[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int j = lane; j < C; j += 32) sum += expf(z[j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float py = 0.0f;
    for (int j = lane; j < C; j += 32) {
      const float p = expf(z[j] - mx) * inv;
      z[j] = p;
      if (train) {
        d[j] = p - y[j];
        py = fmaf(p, y[j], py);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) py += __shfl_xor_sync(0xffffffffu, py, o);
    loss = -logf(py);
  } else if (train) {
    for (int j = lane; j < C; j += 32) {
      const float e = z[j] - y[j];
      d[j] = e;
      loss = fmaf(e, e, loss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  }
  if (train && loss_rows != nullptr && lane == 0) loss_rows[warp] = loss;
}


// Fused output head for narrow heads (C <= 32): one warp per sample row computes the C scores
// z = a W_o^T + b_o, then softmax / identity, delta = out - y and the per-row loss --
// dbn/models.py:594-626, :696-720 in one launch instead of a skinny GEMM (4 CTAs at C = 10) plus a row kernel.
// Each lane owns 8 consecutive k of every 256-wide slab of K: one 16-byte load of the activations and two
// float4 loads per class row, all independent, so a slab puts 1 + 2C loads in flight per lane.
__device__ __forceinline__ void head_load8(const dbn_mat& A, int row, int k, float v[8]) {
  if (A.dtype == DBN_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A.ptr) + (size_t)row * A.ld + k);
    const float4 t0 = p[0], t1 = p[1];
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
  } else {
    const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(A.ptr) + (size_t)row * A.ld + k);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack16x2(A.dtype, w[j], v[2 * j], v[2 * j + 1]);
  }
}

// Block = 8 warps = 2 sample rows x 4 quarters of K: a row's K = 2000 activations are four independent
// 2-iteration loops instead of one 8-iteration chain, and a 256-row mini-batch fills 128 SMs instead of 32.
// With `delta_last` set the kernel also back-propagates through the head,
//     delta_{L-1}[row, h] = (sum_c delta_out[row, c] W_o[c, h]) * f'(a_{L-1}[row, h])      (dbn/models.py:495-499)
// while W_o and the row's activations are still in L1: one launch less on the critical path of the fine-tune step.
__global__ void __launch_bounds__(256) head_fused_kernel(dbn_mat A, int K, const float* __restrict__ Wo, int ldwo,
                                                         const float* __restrict__ bo, int B, int C, int head,
                                                         const float* __restrict__ Y, int ldy, float* out, int ldo,
                                                         float* delta, int ldd, dbn_mat delta_b, float* loss_rows,
                                                         dbn_mat delta_last, int prime) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  __shared__ float partial[2][4][32];
  __shared__ float dout[2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_in = warp >> 2, part = warp & 3;
  const int row = blockIdx.x * 2 + r_in;
  const bool active = row < B;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
  const bool vec = (K % 8 == 0) && (ldwo % 4 == 0) && (A.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Wo) & 15) == 0);
  if (active && vec) {
    for (int k = (part * 32 + lane) * 8; k < K; k += 1024) {
      float a[8];
      head_load8(A, row, k, a);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c < C) {
          const float4* w = reinterpret_cast<const float4*>(Wo + (size_t)c * ldwo + k);
          const float4 w0 = __ldg(w), w1 = __ldg(w + 1);
          float s = acc[c];
          s = fmaf(a[0], w0.x, s); s = fmaf(a[1], w0.y, s); s = fmaf(a[2], w0.z, s); s = fmaf(a[3], w0.w, s);
          s = fmaf(a[4], w1.x, s); s = fmaf(a[5], w1.y, s); s = fmaf(a[6], w1.z, s); s = fmaf(a[7], w1.w, s);
          acc[c] = s;
        }
      }
    }
  } else if (active) {
    for (int k = part * 32 + lane; k < K; k += 128) {
      const float a = mat_load(A, row, k);
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) acc[c] = fmaf(a, __ldg(Wo + (size_t)c * ldwo + k), acc[c]);
    }
  }
  float z = 0.0f;  // lane c ends up holding this quarter's share of score c
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c < C) {   // warp-uniform
      float s = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == c) z = s;
    }
  }
  partial[r_in][part][lane] = z;
  __syncthreads();
  if (part == 0) {
    const bool valid = active && lane < C;
    z = valid ? ((partial[r_in][0][lane] + partial[r_in][1][lane]) + (partial[r_in][2][lane] + partial[r_in][3][lane])) + bo[lane] : 0.0f;
    float o_val = z, loss = 0.0f;
    const float y = (valid && Y != nullptr) ? Y[(size_t)row * ldy + lane] : 0.0f;
    if (head == DBN_HEAD_SOFTMAX) {
      float mx = valid ? z : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = valid ? expf(z - mx) : 0.0f;
      float sum = e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      o_val = active ? e / sum : 0.0f;
      float py = o_val * y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) py += __shfl_xor_sync(0xffffffffu, py, o);
      loss = active ? -logf(py) : 0.0f;
    } else {
      const float d = valid ? (z - y) : 0.0f;
      loss = d * d;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    }
    if (valid) {
      out[(size_t)row * ldo + lane] = o_val;
      if (delta != nullptr) delta[(size_t)row * ldd + lane] = o_val - y;
      if (delta_b.ptr != nullptr) mat_store(delta_b, row, lane, o_val - y);   // tensor-core operand copy (bf16 / fp16)
    }
    dout[r_in][lane] = valid ? (o_val - y) : 0.0f;
    if (active && loss_rows != nullptr && Y != nullptr && lane == 0) loss_rows[row] = loss;
  }
  __syncthreads();
  if (delta_last.ptr == nullptr || !active) return;
  // delta of the last hidden layer, this warp's quarter of the units
  float d[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) d[c] = (c < C) ? dout[r_in][c] : 0.0f;
  if (vec && (delta_last.ld % 8) == 0 && (reinterpret_cast<uintptr_t>(delta_last.ptr) & 15) == 0) {
    for (int k = (part * 32 + lane) * 8; k < K; k += 1024) {
      float a[8], s[8];
      head_load8(A, row, k, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.0f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c < C) {
          const float4* w = reinterpret_cast<const float4*>(Wo + (size_t)c * ldwo + k);
          const float4 w0 = __ldg(w), w1 = __ldg(w + 1);
          s[0] = fmaf(d[c], w0.x, s[0]); s[1] = fmaf(d[c], w0.y, s[1]); s[2] = fmaf(d[c], w0.z, s[2]); s[3] = fmaf(d[c], w0.w, s[3]);
          s[4] = fmaf(d[c], w1.x, s[4]); s[5] = fmaf(d[c], w1.y, s[5]); s[6] = fmaf(d[c], w1.z, s[6]); s[7] = fmaf(d[c], w1.w, s[7]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        s[j] = (prime == DBN_MUL_PRIME_SIGMOID) ? s[j] * (a[j] * (1.0f - a[j])) : ((a[j] > 0.0f) ? s[j] : 0.0f);
      mat_store4(delta_last, row, k, s, 4);
      mat_store4(delta_last, row, k + 4, s + 4, 4);
    }
  } else {
    for (int k = part * 32 + lane; k < K; k += 128) {
      const float a = mat_load(A, row, k);
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) s = fmaf(d[c], __ldg(Wo + (size_t)c * ldwo + k), s);
      s = (prime == DBN_MUL_PRIME_SIGMOID) ? s * (a * (1.0f - a)) : ((a > 0.0f) ? s : 0.0f);
      mat_store(delta_last, row, k, s);
    }
  }
}

// Gradient + decayed SGD update of a narrow head (C <= 32), dbn/models.py:508-513 with :462-465:
//   W_o[c][h] = decay * W_o[c][h] + scale * sum_b delta[b][c] * a[b][h],   b_o[c] += scale * sum_b delta[b][c].
// One thread per input unit h (plus one for the bias, whose "activation" is 1); the deltas of the mini-batch
// are staged in shared memory; the batch loop is unrolled so the activation loads overlap.  If raw != NULL
// the un-scaled sums are written to raw [C, ldr] (column H = bias) and nothing is updated.
__global__ void __launch_bounds__(128) head_grad_kernel(dbn_mat A, int H, const float* __restrict__ delta, int ldd, int B,
                                                        int C, float* Wo, int ldwo, float* bo, float decay, float scale,
                                                        float* raw, int ldr) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float hg_delta[];   // [B][C]
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) hg_delta[i] = delta[(size_t)(i / C) * ldd + (i % C)];
  __syncthreads();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h > H) return;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
  int b = 0;
  for (; b + 4 <= B; b += 4) {
    float a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = (h < H) ? mat_load(A, b + j, h) : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* d = hg_delta + (size_t)(b + j) * C;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) acc[c] = fmaf(d[c], a[j], acc[c]);
    }
  }
  for (; b < B; ++b) {
    const float a = (h < H) ? mat_load(A, b, h) : 1.0f;
    const float* d = hg_delta + (size_t)b * C;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < C) acc[c] = fmaf(d[c], a, acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c < C) {
      if (raw != nullptr) raw[(size_t)c * ldr + h] = acc[c];
      else if (h < H) Wo[(size_t)c * ldwo + h] = fmaf(scale, acc[c], decay * Wo[(size_t)c * ldwo + h]);
      else bo[c] = fmaf(scale, acc[c], bo[c]);
    }
  }
}

// W += s * dW, c += s * dc (last column) for rows [h0, h0 + n_w_rows) and, if with_db, b += s * db (row H)
// from a raw [(H+1), ld] delta buffer whose first stored row is raw_row0 (a reduce-scatter shard).
__global__ void __launch_bounds__(256) apply_delta_kernel(int V, int H, int h0, int n_w_rows, int with_db, float scale,
                                                          const float* __restrict__ raw, int ld, int raw_row0, float* W,
                                                          int ldw, float* b, float* c, dbn_mat Wb) {
  const int quads = (V + 1 + 3) >> 2;
  const int rows = n_w_rows + (with_db ? 1 : 0);
  const long total = (long)rows * quads;
  dbn_update_epilogue e;
  memset(&e, 0, sizeof(e));
  e.W = W; e.ldw = ldw; e.Wb = Wb; e.vecM = c; e.vecN = b; e.decay = 1.0f; e.scale = scale;
  e.raw.ptr = nullptr; e.raw.ld = 0; e.raw.dtype = DBN_F32;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int r = (int)(t / quads), n4 = (int)(t % quads) << 2;
    const int m = (r < n_w_rows) ? h0 + r : H;       // the extra row is the visible-bias statistics row
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = (n4 + j <= V) ? raw[(size_t)(m - raw_row0) * ld + n4 + j] : 0.0f;
    upd_epilogue_quad(e, m, n4, acc, H, V, 1, 1);
  }
}

}  // namespace dbn
