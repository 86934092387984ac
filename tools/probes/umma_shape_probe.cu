// umma_shape_probe.cu -- what does ONE tcgen05.mma cost as a function of its shape?  (debug micro-benchmark, round 2)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_shape_probe umma_shape_probe.cu && ./umma_shape_probe
//
// The persistent epoch kernel (csrc/rbm_tc_epoch.cuh) issues chains of 32-49 small MMAs (64 x 32..56 x 16) whose
// operands are already in shared memory; its trace shows ~66 cycles per MMA whatever N is, with one or three issuing
// warps and with one or four accumulators.  This probe times n back-to-back MMAs of one CTA on zero operands:
// M in {64, 128}, N in {32 .. 256}, K-major or MN-major, 1 / 2 / 4 issuer threads (separate accumulators).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
  uint32_t d = 0;
  while (!d) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t mk_idesc(int bm, int bn, int tA, int tB) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)tA << 15) | ((uint32_t)tB << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// out[0] = cycles from the first issue to completion of all MMAs, out[1] = cycles spent issuing (issuer 0)
__global__ void probe(int M, int N, int trans, int n_mma, int issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mb_init(s32(&bar), issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  const long long t0 = clock64();
  if (lane == 0 && warp < issuers) {
    const uint32_t idesc = mk_idesc(M, N, trans, trans);
    const uint32_t sa = s32(base), sb = s32(base + 64 * 1024);
    const int per = n_mma / issuers;
    const uint32_t cols = (uint32_t)(warp * (issuers > 1 ? 512 / issuers : 0));
    uint64_t da = trans ? mk_desc(sa, 32768, 1024) : mk_desc(sa, 16, 1024);
    uint64_t db = trans ? mk_desc(sb, 32768, 1024) : mk_desc(sb, 16, 1024);
    const uint32_t inc = trans ? 128 : 2;
#pragma unroll 4
    for (int i = 0; i < per; ++i) {
      const uint32_t off = (uint32_t)(i & 3) * inc + (uint32_t)((i >> 2) & 3) * (trans ? 512 : 512);
      mma(tmem + cols, da + off, db + off, idesc, i > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    if (warp == 0) out[1] = clock64() - t0;
  }
  if (threadIdx.x == 128) {
    mb_wait(s32(&bar), 0);
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("   M    N  major  n_mma issuers |  total cycles  per MMA | issue cycles per MMA\n");
  for (int trans = 0; trans < 2; ++trans)
    for (int M : {64, 128})
      for (int N : {32, 64, 128, 256})
        for (int issuers : {1, 2, 4}) {
          if (issuers * N > 512) continue;
          for (int n : {16, 64}) {
            long long h[2] = {0, 0};
            for (int rep = 0; rep < 3; ++rep) {
              probe<<<1, 160, 200 * 1024>>>(M, N, trans, n, issuers, out);
              cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            printf("%4d %4d  %s  %5d %5d   | %10lld %10.1f | %10.1f\n", M, N, trans ? "MN" : "K ", n, issuers, h[0], (double)h[0] / n,
                   (double)h[1] / (n / issuers));
          }
        }
  return 0;
}
