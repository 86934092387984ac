// tma_burst_probe.cu -- what does ONE CTA pay for a short burst of bulk-tensor loads?  (debug micro-benchmark, round 2)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_burst_probe tma_burst_probe.cu && ./tma_burst_probe
//
// The batch-256 kernels (profiles/r01_notes.md) never reach a steady state: a split-K CTA loads 48-192 KB in 2-8
// boxes and is done.  Its trace shows ~2000 cycles until the first 48 KB stage has landed and ~1000 per further
// stage, i.e. 25-50 B/clk, against 93 B/clk in the long config-5 mainloop.  This probe separates latency from rate:
// NBOX boxes of ROWS x 128 B (3-D map, KS slabs per box like the kernels') are issued back to back by one or two
// threads into distinct shared-memory slots, and the clock is read when the k-th box has landed (k = 1..NBOX).
// Variables: box height, boxes per burst, one or two issuing threads, number of CTAs pulling at once, and whether
// the rows were touched before (L2-resident) or not (first touch comes from HBM).
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
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int MAX_BOX = 8;

// out[cta][k] = cycles from the first issue until box k has landed
__global__ void burst(const __grid_constant__ CUtensorMap map, int box_bytes, int rows, int nbox, int issuers, int row_stride_per_cta,
                      long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bars[MAX_BOX];
  __shared__ long long t0s;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    for (int i = 0; i < nbox; ++i) mb_init(s32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int r0 = blockIdx.x * row_stride_per_cta;
  if (threadIdx.x == 0) t0s = clock64();
  __syncthreads();
  // issuer w (warp w, lane 0) takes boxes w, w + issuers, ...
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < issuers) {
    for (int i = warp; i < nbox; i += issuers) {
      mb_expect(s32(&bars[i]), box_bytes);
      tma3d(s32(base + (size_t)i * box_bytes), &map, s32(&bars[i]), 0, r0 + (i & 1) * rows, (i >> 1) * 2);   // slabs 2*(i/2)..+1
    }
  }
  if (threadIdx.x == 64) {                  // a third warp watches the barriers in order
    for (int i = 0; i < nbox; ++i) {
      mb_wait(s32(&bars[i]), 0);
      out[(size_t)blockIdx.x * MAX_BOX + i] = clock64() - t0s;
    }
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void touch(const __nv_bfloat16* p, size_t n, float* sink) {
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += __bfloat162float(p[i]);
  if (s == 12345.f) *sink = s;
}

int main() {
  const int ROWS_TOTAL = 148 * 512, COLS = 832;   // like the staged config-2 rows: pitch 832 bf16 (13 slabs of 64)
  __nv_bfloat16* d;
  cudaMalloc(&d, (size_t)ROWS_TOTAL * COLS * 2);
  cudaMemset(d, 0, (size_t)ROWS_TOTAL * COLS * 2);
  long long* out;
  cudaMalloc(&out, 148 * MAX_BOX * 8);
  float* sink;
  cudaMalloc(&sink, 4);
  // a buffer larger than L2 to evict with
  const size_t FLUSH = (size_t)256 << 20;
  __nv_bfloat16* fl;
  cudaMalloc(&fl, FLUSH);
  cudaMemset(fl, 0, FLUSH);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncFn enc = (EncFn)fp;
  for (int warm : {1, 0}) {
    for (int grid : {1, 64, 128}) {
      for (int rows : {64, 128, 192}) {          // box = rows x 2 slabs x 128 B: 16 / 32 / 48 KB
        for (int nbox : {1, 2, 4}) {
          for (int issuers : {1, 2}) {
            if (issuers > nbox) continue;
            CUtensorMap map;
            cuuint64_t gd[3] = {64, (cuuint64_t)ROWS_TOTAL, (cuuint64_t)(COLS / 64)};
            cuuint64_t gs[2] = {(cuuint64_t)COLS * 2, 128};
            cuuint32_t box[3] = {64, (cuuint32_t)rows, 2};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const int box_bytes = rows * 2 * 128;
            const int smem = nbox * box_bytes + 2048;
            cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            long long h[148 * MAX_BOX];
            double best[MAX_BOX];
            for (int k = 0; k < MAX_BOX; ++k) best[k] = 1e30;
            for (int rep = 0; rep < 5; ++rep) {
              if (warm) touch<<<296, 256>>>(d, (size_t)ROWS_TOTAL * COLS, sink);
              else touch<<<296, 256>>>(fl, FLUSH / 2, sink);
              burst<<<grid, 96, smem>>>(map, box_bytes, rows, nbox, issuers, 512, out);
              cudaDeviceSynchronize();
              cudaMemcpy(h, out, (size_t)grid * MAX_BOX * 8, cudaMemcpyDeviceToHost);
              for (int k = 0; k < nbox; ++k) {      // median over CTAs would be nicer; the max is what a kernel waits for
                long long mx = 0;
                for (int c = 0; c < grid; ++c) mx = h[c * MAX_BOX + k] > mx ? h[c * MAX_BOX + k] : mx;
                if ((double)mx < best[k]) best[k] = (double)mx;
              }
            }
            printf("%s grid=%3d box=%2d KB x%d issuers=%d :", warm ? "L2-warm" : "cold   ", grid, box_bytes / 1024, nbox, issuers);
            for (int k = 0; k < nbox; ++k) printf(" box%d@%5.0f", k + 1, best[k]);
            printf("  -> %5.1f B/clk/SM over the burst  (%s)\n", (double)nbox * box_bytes / best[nbox - 1], cudaGetErrorString(cudaGetLastError()));
          }
        }
      }
    }
  }
  return 0;
}
