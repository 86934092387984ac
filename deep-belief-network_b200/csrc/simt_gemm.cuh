// simt_gemm.cuh -- strict-FP32 contraction (FFMA on CUDA cores) with the fused epilogues.
//
//   acc[m,n] = sum_k A0(m,k) B0(n,k)  -  sum_k A1(m,k) B1(n,k)          (second pair optional)
//
// This is the 1e-5-parity variant of every contraction on the CD-k / backprop path, and the
// kernel for shapes too small or too ragged for the tcgen05 tiles.  Operands may be fp32 or bf16
// (converted on load); accumulation is fp32 in registers; no tensor cores, no fast-math.
// 64x64x16 tiles, 256 threads, 4x4 accumulators per thread, smem double-buffered by register
// prefetch.  The optional ones row/column (augM/augN) is synthesised in the loader, never stored.
#pragma once
#include "epilogue.cuh"

namespace dbn {

struct SimtGemmArgs {
  dbn_operands ops[2];
  int n_ops;
  int M, N;        // core extent
  int augM, augN;  // 0/1: virtual ones row of A / ones row of B
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;
constexpr int SG_PAD = 4;

// element (i, k) of an operand whose "row" index i runs over M (or N) and k over K.
__device__ __forceinline__ float sg_load(const dbn_mat& X, int trans, int row0, int i, int k, int I, int K,
                                         int aug) {
  if (k >= K) return 0.0f;
  if (i >= I) return (aug && i == I) ? 1.0f : 0.0f;
  return trans ? mat_load(X, row0 + k, i) : mat_load(X, row0 + i, k);
}

template <class Epi>
__device__ __forceinline__ void sg_run(const SimtGemmArgs& g, const Epi& epi_fn) {
  __shared__ float As[SG_BK][SG_BM + SG_PAD];
  __shared__ float Bs[SG_BK][SG_BN + SG_PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 rows x 4 cols
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int p = 0; p < g.n_ops; ++p) {
    const dbn_operands& o = g.ops[p];
    const float sign = (p == 0) ? 1.0f : -1.0f;
    // thread -> tile element maps chosen so global reads run along the contiguous dimension
    int a_i[4], a_k[4], b_i[4], b_k[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * SG_THREADS;  // 0..1023
      if (o.transA) { a_i[r] = idx % SG_BM; a_k[r] = idx / SG_BM; }
      else          { a_k[r] = idx % SG_BK; a_i[r] = idx / SG_BK; }
      if (o.transB) { b_i[r] = idx % SG_BN; b_k[r] = idx / SG_BN; }
      else          { b_k[r] = idx % SG_BK; b_i[r] = idx / SG_BK; }
    }
    float ra[4], rb[4];
    const int nk = (o.K + SG_BK - 1) / SG_BK;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ra[r] = sign * sg_load(o.A, o.transA, o.row0_A, m0 + a_i[r], a_k[r], g.M, o.K, g.augM);
      rb[r] = sg_load(o.B, o.transB, o.row0_B, n0 + b_i[r], b_k[r], g.N, o.K, g.augN);
    }
    for (int kb = 0; kb < nk; ++kb) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        As[a_k[r]][a_i[r]] = ra[r];
        Bs[b_k[r]][b_i[r]] = rb[r];
      }
      __syncthreads();
      if (kb + 1 < nk) {
        const int k1 = (kb + 1) * SG_BK;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          ra[r] = sign * sg_load(o.A, o.transA, o.row0_A, m0 + a_i[r], k1 + a_k[r], g.M, o.K, g.augM);
          rb[r] = sg_load(o.B, o.transB, o.row0_B, n0 + b_i[r], k1 + b_k[r], g.N, o.K, g.augN);
        }
      }
#pragma unroll
      for (int k = 0; k < SG_BK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {av.x, av.y, av.z, av.w};
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epi_fn(m0 + ty * 4 + i, n0 + tx * 4, acc[i]);
}

__global__ void __launch_bounds__(SG_THREADS) simt_gemm_act_kernel(SimtGemmArgs g, dbn_act_epilogue e) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  const uint32_t epoch = e.rng.epoch + (e.rng.epoch_ptr ? *e.rng.epoch_ptr : 0u);
  sg_run(g, [&](int m, int n4, const float* acc) { act_epilogue_quad(e, epoch, m, n4, acc, g.M, g.N); });
}

__global__ void __launch_bounds__(SG_THREADS) simt_gemm_update_kernel(SimtGemmArgs g, dbn_update_epilogue e) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  sg_run(g, [&](int m, int n4, const float* acc) { upd_epilogue_quad(e, m, n4, acc, g.M, g.N, g.augM, g.augN); });
}

inline int simt_gemm_act(cudaStream_t st, int M, int N, const dbn_operands* ops, int n_ops,
                         const dbn_act_epilogue& e) {
  SimtGemmArgs g;
  g.n_ops = n_ops; g.M = M; g.N = N; g.augM = 0; g.augN = 0;
  for (int i = 0; i < n_ops; ++i) g.ops[i] = ops[i];
  if (n_ops == 1) g.ops[1] = ops[0];
  dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM));
  DBN_CUDA_OK(launch_pdl(simt_gemm_act_kernel, grid, dim3(SG_THREADS), 0, st, g, e));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

inline int simt_gemm_update(cudaStream_t st, int M, int N, int augM, int augN, const dbn_operands* ops,
                            int n_ops, const dbn_update_epilogue& e) {
  SimtGemmArgs g;
  g.n_ops = n_ops; g.M = M; g.N = N; g.augM = augM; g.augN = augN;
  for (int i = 0; i < n_ops; ++i) g.ops[i] = ops[i];
  if (n_ops == 1) g.ops[1] = ops[0];
  dim3 grid(ceil_div(N + augN, SG_BN), ceil_div(M + augM, SG_BM));
  DBN_CUDA_OK(launch_pdl(simt_gemm_update_kernel, grid, dim3(SG_THREADS), 0, st, g, e));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

}  // namespace dbn
