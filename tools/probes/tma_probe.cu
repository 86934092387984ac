// tma_probe.cu -- how fast can ONE CTA pull 128-byte-row boxes through TMA?  (debug micro-benchmark)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu && ./tma_probe
// For each box height R (rows of 128 B) one thread issues OPS bulk-tensor loads into a ring of DEPTH
// smem slots (each with its own mbarrier) and waits for slot i before re-using it; prints cycles per op.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
  uint32_t d = 0;
  while (!d) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

template <int DEPTH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int rows, int ops, int row_span, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[DEPTH];
  const uint32_t bytes = rows * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mb_init(s32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const int r0 = blockIdx.x * row_span;
    long long t0 = clock64();
    for (int i = 0; i < ops; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mb_wait(s32(&bars[s]), ((i / DEPTH) - 1) & 1);
      mb_expect(s32(&bars[s]), bytes);
      tma2d(s32(base + (size_t)s * 32768), &map, s32(&bars[s]), (i * 64) % 4096, r0 + ((i * rows) % row_span));
    }
    for (int i = ops - DEPTH; i < ops; ++i) if (i >= 0) mb_wait(s32(&bars[i % DEPTH]), (i / DEPTH) & 1);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int ROWS = 32768, COLS = 4160;   // bf16 matrix, 128-byte aligned pitch
  __nv_bfloat16* d;
  cudaMalloc(&d, (size_t)ROWS * COLS * 2);
  cudaMemset(d, 0, (size_t)ROWS * COLS * 2);
  long long* out;
  cudaMalloc(&out, 148 * 8);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncFn enc = (EncFn)fp;
  const int OPS = 64;
  for (int grid : {1, 16, 148}) {
    for (int rows : {32, 64, 128, 256}) {
      CUtensorMap map;
      cuuint64_t gd[2] = {(cuuint64_t)COLS, (cuuint64_t)ROWS};
      cuuint64_t gs[1] = {(cuuint64_t)COLS * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)rows};
      cuuint32_t es[2] = {1, 1};
      enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      long long h[148];
      auto run = [&](auto kern, int depth) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, depth * 32768 + 2048);
        for (int rep = 0; rep < 3; ++rep) kern<<<grid, 32, depth * 32768 + 2048>>>(map, rows, OPS, 1024 / 4 * 4, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid=%3d box=%3d rows (%5d B)  depth=%d : %6.0f cycles/op  %5.1f B/clk/SM  (%s)\n", grid, rows, rows * 128, depth,
               (double)mx / OPS, (double)rows * 128 * OPS / mx, cudaGetErrorString(cudaGetLastError()));
      };
      run(probe<2>, 2);
      run(probe<4>, 4);
      run(probe<6>, 6);
    }
  }
  return 0;
}
