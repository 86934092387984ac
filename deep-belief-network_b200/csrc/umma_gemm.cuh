// umma_gemm.cuh -- BF16 contraction on the 5th-generation tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma with FP32 accumulation
// in TMEM -> tcgen05.ld -> fused epilogue (epilogue.cuh).
//
//   acc[m,n] = sum_k A0(m,k) B0(n,k)  -  sum_k A1(m,k) B1(n,k)        (second pair optional, negated
//                                                                        by the instruction descriptor)
//
// Every operand may be K-major (stored [M|N, K]) or MN-major (stored [K, M|N]); the same bf16 W
// shadow therefore serves v->h (K-major B), h->v (MN-major B) and, with the activations, the
// dW = P0^T V0 - Pk^T Vk contraction (MN-major A and B) without any transposed copy in HBM.
//
// Stages are FAT: one pipeline stage carries KS 64-wide k-slabs of both operands, each fetched by ONE
// 3-D bulk-tensor copy (dims: 64 elements of a slab row, rows, slab/chunk index at a 128-byte stride).
// Measured on B200 (tools/probes/tma_probe.cu): a bulk-tensor copy costs ~310 cycles whatever it
// carries (4 KB or 32 KB; >100 B/clk per SM for 32 KB boxes), so the mainloop is paced by the NUMBER
// of copies -- two per stage here, i.e. two per KS*64 of K.
//
// Kernel shape: persistent, one CTA per SM, 11 warps:
//   warp 0    TMA producer (one elected lane)         -- ring of S smem stages, mbarrier full/empty
//   warp 1    TMEM allocator + tcgen05.mma issuer      -- one elected lane issues 4 MMAs (K=16) per stage
//   warps 2-5 epilogue (one TMEM lane quarter each)    -- double-buffered accumulator (2 x BN columns)
// Tile 128 x BN x 64 (BN in {64,128,256}); ragged edges are zero-filled by TMA and masked in the
// epilogue, so no operand is ever padded to a tile multiple.
#pragma once
#include <cuda.h>

#include "epilogue.cuh"

namespace dbn {

constexpr int UG_BM = 128;
constexpr int UG_BK = 64;   // bf16 elements per stage along K = one 128-byte swizzle row
constexpr int UG_UK = 16;   // K of one tcgen05.mma.kind::f16
constexpr int UG_EPI_WARPS = 8;          // two warps per TMEM lane quarter, each takes half of the columns
constexpr int UG_THREADS = 64 + 32 * UG_EPI_WARPS + 32;   // + TMA producer A, MMA issuer, TMA producer B
constexpr int UG_WARP_PB = 2 + UG_EPI_WARPS;              // warp index of the second (B-operand) producer
constexpr int UG_SMEM_BUDGET = 196608;  // operand ring; + 1 KB barriers/alignment  (<= 227 KB per CTA)

struct UmmaGemmParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  int n_ops;
  int K[2];
  int nslab[2];   // ceil(K / 64): 64-wide k-slabs per pass
  int transA[2], transB[2];
  int rowA[2], rowB[2];
  int M, N;  // accumulator extent (incl. the ones row / column)
  int tiles_m, tiles_n, group_m;
  long long* trace;  // debug: per-CTA phase timestamps (dbn_debug_set_trace), normally NULL
};

// debug trace slots per CTA: 0 entry (globaltimer ns), 1..6 clock64 at: entry, setup done, first operands
// landed, mainloop of first tile issued, accumulator of first tile ready, first tile's epilogue done, 7 exit
constexpr int UG_TRACE_SLOTS = 16;
__device__ __forceinline__ void trace_mark(const UmmaGemmParams& p, int slot) {
  if (p.trace != nullptr) p.trace[(size_t)blockIdx.x * UG_TRACE_SLOTS + slot] = clock64();
}
__device__ __forceinline__ void trace_mark_ns(const UmmaGemmParams& p, int slot) {  // slots 0, 8, 9: globaltimer
  if (p.trace != nullptr) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[(size_t)blockIdx.x * UG_TRACE_SLOTS + slot] = (long long)gt;
  }
}
inline long long*& umma_trace_ptr() {
  static long long* t = nullptr;
  return t;
}

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 24)) {
      printf("dbn_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   bits [46,48) = 1 (Blackwell) | [61,64) layout type = 2 (SWIZZLE_128B)
// K-major operand  : rows of 128 B (64 bf16 along K); 8-row groups 1024 B apart (SBO); LBO unused (=1).
// MN-major operand : rows of 128 B (64 bf16 along M/N) indexed by k; 8-k groups 1024 B apart (SBO);
//                    64-wide M/N chunks 8192 B apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor of tcgen05.mma.kind::f16: D = F32, A = B = BF16.
//   [4,6) D format (1 = F32) | [7,10) A format (1 = BF16) | [10,13) B format | 13 negate A | 15 A MN-major
//   16 B MN-major | [17,23) N >> 3 | [24,29) M >> 4
__device__ __forceinline__ uint32_t make_idesc(int bm, int bn, int transA, int transB, int negA) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)negA << 13) | ((uint32_t)transA << 15) |
         ((uint32_t)transB << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(const UmmaGemmParams& p, int tile, int& mb, int& nb) {
  // grouped raster: `group_m` row-blocks x all column-blocks form an L2-resident super-tile
  const int per_group = p.group_m * p.tiles_n;
  const int g = tile / per_group;
  const int first_m = g * p.group_m;
  const int gm = min(p.group_m, p.tiles_m - first_m);
  const int r = tile - g * per_group;
  mb = first_m + r % gm;
  nb = r / gm;
}

template <int BN, int KS, class Epi>
__global__ void __launch_bounds__(UG_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ UmmaGemmParams p, const __grid_constant__ Epi epi_args, int coreM, int coreN,
                 int augM, int augN) {
  constexpr int BM = UG_BM;   // UMMA M = 128: accumulator row m lives in TMEM lane m
  constexpr int A_SLAB = BM * UG_BK * 2;                 // one 64-wide k-slab of the A tile
  constexpr int B_SLAB = BN * UG_BK * 2;
  constexpr int A_BYTES = KS * A_SLAB;                   // operand regions of one (fat) stage
  constexpr int B_BYTES = KS * B_SLAB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = (UG_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (UG_SMEM_BUDGET / STAGE_BYTES);
  constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  constexpr int EPI_GROUPS = (BN / 32) < 2 ? 1 : 2;      // column halves drained in parallel
  static_assert(STAGES >= 2, "pipeline too shallow");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.tiles_m * p.tiles_n;
  if (threadIdx.x == 0) {
    trace_mark_ns(p, 0);
    trace_mark(p, 1);
  }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_ops; ++i) {   // hide the descriptor fetch behind the barrier / TMEM setup
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB[i]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * EPI_GROUPS);  // one arrive per active epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();   // nothing above touches global memory (bar the debug trace); everything below may
  if (threadIdx.x == 0) {
    trace_mark(p, 2);
    trace_mark_ns(p, 8);
  }

  if (warp == 0 || warp == UG_WARP_PB) {
    // ===================== TMA producers: warp 0 feeds A, warp UG_WARP_PB feeds B =====================
    // (one bulk-tensor instruction costs its issuing thread a few hundred cycles, so two issuers
    //  double the rate at which stages fill when the tiles are small)
    const bool is_pa = warp == 0;
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mb, nb;
        tile_coords(p, tile, mb, nb);
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nfs = (p.nslab[pass] + KS - 1) / KS;   // fat stages of this pass
          for (int f = 0; f < nfs; ++f, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
            const int slab0 = f * KS;
            if (is_pa) {
              mbar_expect_tx(full_bar(s), STAGE_BYTES);   // A and B bytes; the B producer only adds transactions
              if (!p.transA[pass]) tma_load_3d(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + mb * BM, slab0);
              else                 tma_load_3d(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + slab0 * UG_BK, mb * (BM / 64));
            } else {
              if (!p.transB[pass]) tma_load_3d(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + nb * BN, slab0);
              else                 tma_load_3d(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + slab0 * UG_BK, nb * (BN / 64));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int as = tl & 1;
        const uint32_t aph = (tl >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        uint32_t accumulate = 0;
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nslab = p.nslab[pass];
          const int nfs = (nslab + KS - 1) / KS;
          const int tA = p.transA[pass], tB = p.transB[pass];
          const uint32_t idesc = make_idesc(BM, BN, tA, tB, pass == 1);
          for (int f = 0; f < nfs; ++f, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (it == 0) trace_mark(p, 3);
            const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
            const int slabs = min(KS, nslab - f * KS);   // the last stage of a pass may be partly beyond K (zero-filled)
#pragma unroll
            for (int sl = 0; sl < KS; ++sl) {
              if (sl < slabs) {
#pragma unroll
                for (int k4 = 0; k4 < UG_BK / UG_UK; ++k4) {
                  // K-major: slab sl = [rows][128 B] at sl * SLAB.  MN-major: [64-wide chunk][KS*64 k rows][128 B],
                  // slab sl = rows sl*64.. of every chunk, chunks KS*8192 B apart (LBO).
                  const uint64_t da = tA ? make_smem_desc(sa + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                         : make_smem_desc(sa + sl * A_SLAB + k4 * 32, 16, 1024);
                  const uint64_t db = tB ? make_smem_desc(sb + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                         : make_smem_desc(sb + sl * B_SLAB + k4 * 32, 16, 1024);
                  tc_mma_bf16(tmem_d, da, db, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            tc_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
          }
        }
        tc_commit(tfull_bar(as));  // accumulator complete -> epilogue
        if (tl == 0) trace_mark(p, 4);
      }
    }
  } else if (warp < UG_WARP_PB && (warp - 2) / 4 < EPI_GROUPS) {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int q = warp & 3;
    uint32_t epoch = 0;
    if constexpr (!Epi::kIsUpdate) epoch = epi_args.e.rng.epoch + (epi_args.e.rng.epoch_ptr ? *epi_args.e.rng.epoch_ptr : 0u);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      int mb, nb;
      tile_coords(p, tile, mb, nb);
      const int as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      if (tl == 0 && warp == 2 && lane == 0) trace_mark(p, 5);
      const int m = mb * BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      constexpr int CH_PER_WARP = (BN / 32) / EPI_GROUPS;
      const int ch0 = ((warp - 2) >> 2) * CH_PER_WARP;   // which column half this warp drains
#pragma unroll 1
      for (int ch = ch0; ch < ch0 + CH_PER_WARP; ++ch) {
        const int n0 = nb * BN + ch * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tc_ld32(trow + ch * 32, v);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        epi_args.apply(epoch, m, n0, acc, coreM, coreN, augM, augN);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (tl == 0 && warp == 2 && lane == 0) trace_mark(p, 6);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
  if (threadIdx.x == 32) {
    trace_mark(p, 7);
    trace_mark_ns(p, 9);
  }
}


// =====================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.  Each CTA
// stages its own 128 rows of A and HALF of the B tile (BN/2 columns); tcgen05.mma.cta_group::2, issued
// by the leader CTA only, reads both halves and writes each CTA's 128 x BN accumulator into that CTA's
// own TMEM.  Operand bytes fetched from L2 per MAC drop by a third (BN = 256: 32 KB instead of 48 KB
// per CTA and k-block), which is what bounds the single-CTA kernel on the 4096-wide layers.
// =====================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  // the transaction bytes are credited to the LEADER CTA's barrier (peer bit cleared)
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_2sm(int bn, int transA, int transB, int negA) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)negA << 13) | ((uint32_t)transA << 15) |
         ((uint32_t)transB << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int KS, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UG_THREADS, 1)
umma_gemm2_kernel(const __grid_constant__ UmmaGemmParams p, const __grid_constant__ Epi epi_args, int coreM, int coreN,
                  int augM, int augN) {
  constexpr int HB = BN / 2;                     // B columns staged by each CTA
  constexpr int A_SLAB = UG_BM * UG_BK * 2;
  constexpr int B_SLAB = HB * UG_BK * 2;
  constexpr int A_BYTES = KS * A_SLAB;
  constexpr int B_BYTES = KS * B_SLAB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = UG_SMEM_BUDGET / STAGE_BYTES;
  static_assert(STAGES >= 2, "pipeline too shallow");
  constexpr int TMEM_COLS = 2 * BN;
  static_assert(HB % 64 == 0, "MN-major B halves come in 64-column chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int n_tiles = p.tiles_m * p.tiles_n;     // tiles_m counts 256-row blocks here
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_ops; ++i) {   // hide the descriptor fetch behind the barrier / TMEM setup
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB[i]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * UG_EPI_WARPS);  // epilogue warps of both CTAs (used on the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        int mb, nb;
        tile_coords(p, tile, mb, nb);
        const int m0 = mb * 256 + (int)cta * UG_BM;
        const int n0 = nb * BN + (int)cta * HB;
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nfs = (p.nslab[pass] + KS - 1) / KS;
          for (int f = 0; f < nfs; ++f, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (leader) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
            const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
            const int slab0 = f * KS;
            if (!p.transA[pass]) tma_load_3d_2sm(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + m0, slab0);
            else                 tma_load_3d_2sm(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + slab0 * UG_BK, m0 / 64);
            if (!p.transB[pass]) tma_load_3d_2sm(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + n0, slab0);
            else                 tma_load_3d_2sm(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + slab0 * UG_BK, n0 / 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tl) {
        const int as = tl & 1;
        const uint32_t aph = (tl >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        uint32_t accumulate = 0;
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nslab = p.nslab[pass];
          const int nfs = (nslab + KS - 1) / KS;
          const int tA = p.transA[pass], tB = p.transB[pass];
          const uint32_t idesc = make_idesc_2sm(BN, tA, tB, pass == 1);
          for (int f = 0; f < nfs; ++f, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
            const int slabs = min(KS, nslab - f * KS);
#pragma unroll
            for (int sl = 0; sl < KS; ++sl) {
              if (sl < slabs) {
#pragma unroll
                for (int k4 = 0; k4 < UG_BK / UG_UK; ++k4) {
                  const uint64_t da = tA ? make_smem_desc(sa + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                         : make_smem_desc(sa + sl * A_SLAB + k4 * 32, 16, 1024);
                  const uint64_t db = tB ? make_smem_desc(sb + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                         : make_smem_desc(sb + sl * B_SLAB + k4 * 32, 16, 1024);
                  tc_mma_bf16_2sm(tmem_d, da, db, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            tc_commit_2sm(empty_bar(s));   // frees this stage in BOTH CTAs
          }
        }
        tc_commit_2sm(tfull_bar(as));      // both CTAs' epilogues may drain their half
      }
    }
  } else if (warp < UG_WARP_PB) {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    uint32_t epoch = 0;
    if constexpr (!Epi::kIsUpdate) epoch = epi_args.e.rng.epoch + (epi_args.e.rng.epoch_ptr ? *epi_args.e.rng.epoch_ptr : 0u);
    uint32_t tl = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tl) {
      int mb, nb;
      tile_coords(p, tile, mb, nb);
      const int as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const int m = mb * 256 + (int)cta * UG_BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      constexpr int CH_PER_WARP = (BN / 32) / (UG_EPI_WARPS / 4);
      const int ch0 = ((warp - 2) >> 2) * CH_PER_WARP;   // which column half this warp drains
#pragma unroll 1
      for (int ch = ch0; ch < ch0 + CH_PER_WARP; ++ch) {
        const int n0 = nb * BN + ch * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tc_ld32(trow + ch * 32, v);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        epi_args.apply(epoch, m, n0, acc, coreM, coreN, augM, augN);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(tempty_bar(as), 0);  // the leader's MMA thread owns the accumulator hand-off
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA may retire while its pair can still read its smem / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}


struct UmmaActEpi {
  static constexpr bool kIsUpdate = false;
  dbn_act_epilogue e;
  __device__ __forceinline__ void apply(uint32_t epoch, int m, int n0, const float acc[32], int M, int N, int, int) const {
    act_epilogue_row32<true>(e, epoch, m, n0, acc, M, N);
  }
};
struct UmmaUpdEpi {
  static constexpr bool kIsUpdate = true;
  dbn_update_epilogue e;
  __device__ __forceinline__ void apply(uint32_t, int m, int n0, const float acc[32], int M, int N, int augM, int augN) const {
    upd_epilogue_row32(e, m, n0, acc, M, N, augM, augN);
  }
};

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 3-D tensor map over a bf16 operand stored [rows, ld] (ld a multiple of 64 elements):
//   dim0 = 64 elements (one 128-byte slab row), dim1 = stored rows, dim2 = 64-column slab index (stride 128 B).
// K-major operand : dim1 = M/N rows (valid_rows), dim2 = k-slabs (ceil(K/64));   box {64, tile rows, KS}
// MN-major operand: dim1 = k rows (valid_rows),   dim2 = 64-wide M/N chunks;     box {64, KS*64, tile chunks}
// Out-of-range rows and slabs are zero-filled by TMA.  Columns between an extent and the next multiple of 64
// are READ: M/N tails only feed accumulator rows/columns the epilogue masks, but for a K tail the padding
// columns of at least one operand of the pair must be zero and the other finite (weight shadows are zero there).
inline int make_operand_map(CUtensorMap* map, const dbn_mat& X, uint64_t valid_rows, uint64_t n_slabs, uint32_t box_rows,
                            uint32_t box_slabs) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(DBN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available%s%s");
  cuuint64_t gdim[3] = {64, valid_rows, n_slabs};
  cuuint64_t gstride[2] = {(cuuint64_t)X.ld * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_slabs};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, X.ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return fail(DBN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %s%s", buf);
  }
  return DBN_OK;
}

inline bool umma_operand_ok(const dbn_mat& X, int cols) {
  // pitch: multiple of 64 elements and wide enough for every 64-column slab the 3-D map addresses; base 16-byte
  // aligned (TMA), 128-byte aligned in practice -- column-offset views (row blocks of P0 in the data-parallel
  // path) start at multiples of 64 columns
  return X.ptr != nullptr && X.dtype == DBN_BF16 && (X.ld % 64) == 0 && X.ld >= (int)align_up((size_t)cols, 64) &&
         (reinterpret_cast<uintptr_t>(X.ptr) % 16) == 0;
}

// The tensor-core path takes bf16 operands with 16-byte aligned rows and is worth its 128 x BN tiles
// only for contractions that fill at least part of one; everything else runs on the FP32 SIMT kernel.
inline bool umma_gemm_supported(int M, int N, const dbn_operands* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i) {
    if (!umma_operand_ok(ops[i].A, ops[i].transA ? M : ops[i].K)) return false;
    if (!umma_operand_ok(ops[i].B, ops[i].transB ? N : ops[i].K)) return false;
    if (ops[i].K < 32) return false;
  }
  // worth a 128-row tile: a real row block, or a skinny output whose contraction is still large
  // (the narrow-head gradient: 10 x 2001 with K = batch)
  return N >= 32 && (M >= 64 || (long)N * ops[0].K >= 64 * 1024);
}

// SMs the persistent contraction kernels may occupy.  A data-parallel caller reserves a few SMs
// (dbn_set_sm_reserve) so that the collective's own kernels can be co-scheduled with the contraction
// they are meant to overlap with: a 1-CTA-per-SM persistent grid would otherwise starve them.
inline int& umma_sm_reserve() {
  static int r = 0;
  return r;
}
inline int umma_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return std::max(2, (n - umma_sm_reserve()) & ~1);
}

// slabs per pipeline stage: as many as keep >= 2 stages in the 192 KB ring (and a 256-row TMA box)
constexpr int umma_ks(int bm, int bn) { return (bm + bn) * 128 * 4 * 2 <= UG_SMEM_BUDGET ? 4 : 2; }

template <int BN, class Epi>
inline int umma_launch(cudaStream_t st, UmmaGemmParams& p, const Epi& epi, int coreM, int coreN, int augM, int augN) {
  constexpr int BM = UG_BM;
  constexpr int KS = umma_ks(BM, BN);
  constexpr int STAGE_BYTES = KS * (BM + BN) * UG_BK * 2;
  constexpr int STAGES = (UG_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (UG_SMEM_BUDGET / STAGE_BYTES);
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + (2 * STAGES + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    DBN_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<BN, KS, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  p.tiles_m = ceil_div(p.M, BM);
  p.tiles_n = ceil_div(p.N, BN);
  // super-tile of group_m row blocks: A slab (group_m * BM rows) + all of B stay L2-resident
  p.group_m = std::max(1, std::min(p.tiles_m, 16));
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int grid = std::min(n_tiles, umma_num_sms());
  DBN_CUDA_OK(launch_pdl(umma_gemm_kernel<BN, KS, Epi>, dim3(grid), dim3(UG_THREADS), SMEM, st, p, epi, coreM, coreN,
                         augM, augN));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

constexpr int UG2_KS = 2;   // CTA-pair kernel: 2 slabs (K = 128) per stage, 64 KB per CTA, 3 stages

template <int BN, class Epi>
inline int umma_launch2(cudaStream_t st, UmmaGemmParams& p, const Epi& epi, int coreM, int coreN, int augM, int augN) {
  constexpr int STAGE_BYTES = UG2_KS * (UG_BM + BN / 2) * UG_BK * 2;
  constexpr int STAGES = UG_SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + (2 * STAGES + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    DBN_CUDA_OK(cudaFuncSetAttribute(umma_gemm2_kernel<BN, UG2_KS, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  p.tiles_m = ceil_div(p.M, 256);
  p.tiles_n = ceil_div(p.N, BN);
  p.group_m = std::max(1, std::min(p.tiles_m, 8));
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int clusters = std::min(n_tiles, umma_num_sms() / 2);
  DBN_CUDA_OK(launch_pdl(umma_gemm2_kernel<BN, UG2_KS, Epi>, dim3(clusters * 2), dim3(UG_THREADS), SMEM, st, p, epi, coreM,
                         coreN, augM, augN));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

inline bool umma_use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_UMMA_2CTA");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

template <class Epi>
inline int umma_gemm_run(cudaStream_t st, int M, int N, int augM, int augN, const dbn_operands* ops, int n_ops,
                         const Epi& epi) {
  UmmaGemmParams p;
  memset(&p, 0, sizeof(p));
  p.trace = umma_trace_ptr();
  p.n_ops = n_ops;
  p.M = M + augM;
  p.N = N + augN;
  // Tile width by a two-term cost model: waves over the SMs x cycles per K=16 slice of one tile, the
  // latter = max(tcgen05 floor 128*BN/256, smem operand read (128 + BN) * 32 B / 128 B/clk).
  const int sms = umma_num_sms();
  const int tm = ceil_div(p.M, UG_BM);
  int bn = 64;
  long best = -1;
  for (int cand : {256, 128, 64}) {
    const long tiles = (long)tm * ceil_div(p.N, cand);
    const long waves = (tiles + sms - 1) / sms;
    const long cyc = std::max(cand / 2, 32 + cand / 4);
    const long cost = waves * cyc;
    if (best < 0 || cost < best) { best = cost; bn = cand; }
  }
  // Large outputs (256 x 256 tiles for at least half of the SM pairs) run on CTA pairs: a third fewer
  // operand bytes per MAC out of L2.  B boxes then cover half a tile.
  const bool two_cta = umma_use_2cta() && (long)ceil_div(p.M, 256) * ceil_div(p.N, 256) >= sms / 4;
  const int bm = UG_BM;
  const int ks = two_cta ? UG2_KS : umma_ks(bm, bn);
  const int b_rows = two_cta ? 128 : bn;        // B rows (K-major) / columns (MN-major) staged per CTA
  for (int i = 0; i < n_ops; ++i) {
    const dbn_operands& o = ops[i];
    p.K[i] = o.K;
    p.nslab[i] = ceil_div(o.K, UG_BK);
    p.transA[i] = o.transA; p.transB[i] = o.transB;
    p.rowA[i] = o.row0_A; p.rowB[i] = o.row0_B;
    if (!o.transA) DBN_TRY(make_operand_map(&p.tmA[i], o.A, (uint64_t)o.row0_A + p.M, p.nslab[i], bm, ks));
    else           DBN_TRY(make_operand_map(&p.tmA[i], o.A, (uint64_t)o.row0_A + o.K, ceil_div(p.M, 64), ks * 64, bm / 64));
    if (!o.transB) DBN_TRY(make_operand_map(&p.tmB[i], o.B, (uint64_t)o.row0_B + p.N, p.nslab[i], b_rows, ks));
    else           DBN_TRY(make_operand_map(&p.tmB[i], o.B, (uint64_t)o.row0_B + o.K, ceil_div(p.N, 64), ks * 64, b_rows / 64));
  }
  if (two_cta) return umma_launch2<256, Epi>(st, p, epi, M, N, augM, augN);
  if (bn == 256) return umma_launch<256, Epi>(st, p, epi, M, N, augM, augN);
  if (bn == 128) return umma_launch<128, Epi>(st, p, epi, M, N, augM, augN);
  return umma_launch<64, Epi>(st, p, epi, M, N, augM, augN);
}

inline int umma_gemm_act(cudaStream_t st, int M, int N, const dbn_operands* ops, int n_ops, const dbn_act_epilogue& e) {
  UmmaActEpi epi;
  epi.e = e;
  return umma_gemm_run(st, M, N, 0, 0, ops, n_ops, epi);
}
inline int umma_gemm_update(cudaStream_t st, int M, int N, int augM, int augN, const dbn_operands* ops, int n_ops,
                            const dbn_update_epilogue& e) {
  UmmaUpdEpi epi;
  epi.e = e;
  return umma_gemm_run(st, M, N, augM, augN, ops, n_ops, epi);
}

}  // namespace dbn
