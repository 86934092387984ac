// epilogue.cuh -- the two fused epilogues shared by the SIMT (strict FP32) and tcgen05 (BF16)
// contraction kernels.  Both kernels hand the epilogue groups of 4 consecutive accumulator
// columns (n4 % 4 == 0) of one row, so one Philox call serves 4 units and stores vectorise.
#pragma once
#include "dbn_common.cuh"

namespace dbn {

// Strict sigmoid: expf (not __expf / ex2.approx) so the FP32 mode stays within 1e-5 of the
// float64 oracle (reference: dbn/activations.py:29  1/(1+exp(-x))).
__device__ __forceinline__ float sigmoid_strict(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(int act, float x) {
  if (act == DBN_ACT_SIGMOID) return sigmoid_strict(x);
  if (act == DBN_ACT_RELU) return fmaxf(x, 0.0f);  // dbn/activations.py:49
  return x;
}

// Activation epilogue on 4 columns.  M, N = logical output extent (stores are masked to it).
__device__ __forceinline__ void act_epilogue_quad(const dbn_act_epilogue& e, uint32_t epoch, int m, int n4,
                                                  const float acc[4], int M, int N) {
  if (m >= M || n4 >= N) return;
  const int nvalid = min(4, N - n4);
  float a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = acc[j];
    if (e.bias != nullptr && j < nvalid) x += e.bias[n4 + j];
    a[j] = apply_act(e.act, x);
  }
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
  if (e.rng.mode == DBN_RNG_SAMPLE && e.sample.ptr != nullptr) {
    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = (u[j] < a[j]) ? 1.0f : 0.0f;  // strict '<': dbn/models.py:155
    mat_store4(e.sample, m, n4, s, nvalid);
  }
}

// Update epilogue on 4 columns of the (M+augM) x (N+augN) accumulator.
__device__ __forceinline__ void upd_epilogue_quad(const dbn_update_epilogue& e, int m, int n4, const float acc[4],
                                                  int M, int N, int augM, int augN) {
  const int Mt = M + augM, Nt = N + augN;
  if (m >= Mt || n4 >= Nt) return;
  const int nvalid = min(4, Nt - n4);
  if (e.raw.ptr != nullptr) {
    mat_store4(e.raw, m, n4, acc, nvalid);
    return;
  }
  if (m < M) {
    const int ncore = max(0, min(4, N - n4));
    if (ncore > 0) {
      float w[4];
      dbn_mat Wm;
      Wm.ptr = e.W; Wm.ld = e.ldw; Wm.dtype = DBN_F32;
      mat_load4(Wm, m, n4, w, ncore);
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

}  // namespace dbn
