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
// Three kernels share the mainloop, the descriptors and the epilogues:
//   umma_gemm_kernel     whole K per CTA, persistent over 128 x BN tiles (BN in {64,128,256}), 11 warps:
//                          warp 0     TMA producer A (one elected lane)   -- ring of fat smem stages, mbarrier full/empty
//                          warp 1     TMEM allocator + tcgen05.mma issuer -- one elected lane, 4 MMAs (K=16) per k-slab
//                          warps 2-9  epilogue (two per TMEM lane quarter) -- double-buffered accumulator (2 x BN columns)
//                          warp 10    TMA producer B
//   umma_gemm2_kernel    cta_group::2: a CTA pair per 256 x 256 tile (the 4096-wide layers of config 5)
//   umma_gemm_sk_kernel  split K over a cluster of 2/4 CTAs per tile + DSMEM reduce-scatter (the batch-256 launches
//                        of configs 2 and 3, where the output is a handful of tiles and latency is everything)
// Ragged edges are zero-filled by TMA and masked in the epilogue, so no operand is ever padded to a tile
// multiple.  umma_gemm_run() picks kernel, tile width and split per launch from measured cycle models.
#pragma once
#include <cuda.h>
#include <cooperative_groups.h>
#include <mutex>

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
  int fmtA[2], fmtB[2];   // tcgen05 kind::f16 operand format per pass: 0 = fp16, 1 = bf16
  int rowA[2], rowB[2];
  int M, N;  // accumulator extent (incl. the ones row / column)
  int tiles_m, tiles_n, group_m;
  int own_start, own_rows;   // peer scatter: row blocks [own_start, own_start + own_rows) belong to this rank and are visited last
  int split, fs_per;   // split-K kernel: cluster size and fat stages per CTA
  long long* trace;  // debug: per-CTA phase timestamps (dbn_debug_set_trace), normally NULL
  // CTA-pair kernel, tail split: tiles [0, n_whole) are contracted whole; every later tile by TWO clusters, the fat
  // stages [0, ts_T0) and [ts_T0, T) of the pass list -- the second one parks its accumulator in ts_part and raises
  // ts_flag, the first one adds it before the fused epilogue (n_whole == tiles_m * tiles_n: no split)
  int n_whole, ts_T0;
  float* ts_part;           // [tile - n_whole][cta][128 rows][BN] fp32
  unsigned int* ts_flag;    // [tile - n_whole][cta][epilogue warp]
};

// debug trace slots per CTA: 0 entry (globaltimer ns), 1..6 clock64 at: entry, setup done, first operands
// landed, mainloop of first tile issued, accumulator of first tile ready, first tile's epilogue done, 7 exit
constexpr int UG_TRACE_SLOTS = 16;
__device__ __forceinline__ void trace_mark(const UmmaGemmParams& p, int slot) {
  if (p.trace != nullptr) p.trace[(size_t)(blockIdx.x + blockIdx.y * gridDim.x) * UG_TRACE_SLOTS + slot] = clock64();
}
__device__ __forceinline__ void trace_mark_ns(const UmmaGemmParams& p, int slot) {  // slots 0, 8, 9: globaltimer
  if (p.trace != nullptr) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[(size_t)(blockIdx.x + blockIdx.y * gridDim.x) * UG_TRACE_SLOTS + slot] = (long long)gt;
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

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
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
// fmtA / fmtB: 0 = F16, 1 = BF16 -- the two 16-bit formats may be mixed (one field per operand)
__device__ __forceinline__ uint32_t make_idesc(int bm, int bn, int transA, int transB, int negA, int fmtA, int fmtB) {
  return (1u << 4) | ((uint32_t)fmtA << 7) | ((uint32_t)fmtB << 10) | ((uint32_t)negA << 13) | ((uint32_t)transA << 15) |
         ((uint32_t)transB << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(const UmmaGemmParams& p, int tile, int& mb, int& nb) {
  if (p.own_rows > 0) {
    // peer scatter (see UmmaUpdEpi::scatter_order): the other owners' row blocks first, walked block by block for one
    // column block at a time, starting behind this rank's own rows; the own row blocks (local stores) last
    const int remote_m = p.tiles_m - p.own_rows;
    const int n_remote = remote_m * p.tiles_n;
    if (tile < n_remote) {
      mb = p.own_start + p.own_rows + tile % remote_m;
      if (mb >= p.tiles_m) mb -= p.tiles_m;
      nb = tile / remote_m;
    } else {
      const int t = tile - n_remote;
      mb = p.own_start + t % p.own_rows;
      nb = t / p.own_rows;
    }
    return;
  }
  // grouped raster: `group_m` row-blocks x all column-blocks form an L2-resident super-tile
  const int per_group = p.group_m * p.tiles_n;
  const int g = tile / per_group;
  const int first_m = g * p.group_m;
  const int gm = min(p.group_m, p.tiles_m - first_m);
  const int r = tile - g * per_group;
  mb = first_m + r % gm;
  nb = r / gm;
}

// Kernel parameters live in the constant bank; every use compiles to an LDC next to its consumer, and the
// epilogue's option tree needs ~40 of them.  In a one-tile-per-CTA launch those loads sit between "partials
// received" and the first store -- measured ~800 cycles before the first useful instruction.  Reading the whole
// struct into registers BEFORE the epilogue warps wait for the accumulator moves them off the critical path.
// The identity shuffle makes each word opaque to ptxas, which otherwise rematerialises a parameter load at its
// use (an empty inline asm only hides the value from the front end).  Call with all 32 lanes converged.
template <class T>
__device__ __forceinline__ T params_to_registers(const T& src) {
  static_assert(sizeof(T) % 4 == 0, "whole 32-bit words");
  constexpr int NW = (int)sizeof(T) / 4;
  uint32_t w[NW];
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(&src);
  const int self = (int)(threadIdx.x & 31);
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = __shfl_sync(0xffffffffu, s32[i], self);
  T out;
  memcpy(&out, w, sizeof(T));
  return out;
}
__device__ __forceinline__ int to_register(int v) { return __shfl_sync(0xffffffffu, v, (int)(threadIdx.x & 31)); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
          const uint32_t idesc = make_idesc(BM, BN, tA, tB, pass == 1, p.fmtA[pass], p.fmtB[pass]);
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
// relaxed: the closing cluster barrier of the split-K kernel only keeps CTAs alive, it orders no data
// (a .release arrive costs ~800 cycles of fence on the critical path)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
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
__device__ __forceinline__ uint32_t make_idesc_2sm(int bn, int transA, int transB, int negA, int fmtA, int fmtB) {
  return (1u << 4) | ((uint32_t)fmtA << 7) | ((uint32_t)fmtB << 10) | ((uint32_t)negA << 13) | ((uint32_t)transA << 15) |
         ((uint32_t)transB << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// Work unit u of the CTA-pair kernel -> tile, range of fat stages, role (0 whole tile, 1 first half + epilogue, 2 second
// half: partial only).  A grid of few tiles quantises badly over the 74 SM pairs (4096 x 4096: 256 tiles = 3 full waves
// and a fourth one 46 % full); when the last wave is at most half full, its tiles are split in two along K so that
// twice as many pairs share it: 3.5 tile times instead of 4.
struct TsWork {
  int tile, g0, g1, role;
};
__device__ __forceinline__ TsWork ts_work(const UmmaGemmParams& p, int u, int T) {
  TsWork w;
  if (u < p.n_whole) {
    w.tile = u; w.g0 = 0; w.g1 = T; w.role = 0;
  } else {
    const int j = u - p.n_whole;
    w.tile = p.n_whole + (j >> 1);
    if (j & 1) { w.g0 = p.ts_T0; w.g1 = T; w.role = 2; }
    else { w.g0 = 0; w.g1 = p.ts_T0; w.role = 1; }
  }
  return w;
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
  const int n_units = p.n_whole + 2 * (n_tiles - p.n_whole);   // (tail split: see ts_work)
  int T = 0;                                                   // fat stages of a whole tile
  for (int pass = 0; pass < p.n_ops; ++pass) T += (p.nslab[pass] + KS - 1) / KS;

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
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        const TsWork w = ts_work(p, u, T);
        int mb, nb;
        tile_coords(p, w.tile, mb, nb);
        const int m0 = mb * 256 + (int)cta * UG_BM;
        const int n0 = nb * BN + (int)cta * HB;
        int g = 0;
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nfs = (p.nslab[pass] + KS - 1) / KS;
          for (int f = 0; f < nfs; ++f, ++g) {
            if (g < w.g0 || g >= w.g1) continue;
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
            ++it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters, ++tl) {
        const TsWork w = ts_work(p, u, T);
        const int as = tl & 1;
        const uint32_t aph = (tl >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        uint32_t accumulate = 0;
        int g = 0;
        for (int pass = 0; pass < p.n_ops; ++pass) {
          const int nslab = p.nslab[pass];
          const int nfs = (nslab + KS - 1) / KS;
          const int tA = p.transA[pass], tB = p.transB[pass];
          const uint32_t idesc = make_idesc_2sm(BN, tA, tB, pass == 1, p.fmtA[pass], p.fmtB[pass]);
          for (int f = 0; f < nfs; ++f, ++g) {
            if (g < w.g0 || g >= w.g1) continue;
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
            ++it;
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
    for (int u = cluster_id; u < n_units; u += n_clusters, ++tl) {
      const TsWork w = ts_work(p, u, T);
      int mb, nb;
      tile_coords(p, w.tile, mb, nb);
      const int as = tl & 1;
      const uint32_t aph = (tl >> 1) & 1u;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      // tail split: this thread's row of the parked partial, this warp's flag (same cta / warp / lane in both clusters)
      // (the partial is stored the way the warp holds it -- [chunk][quad of columns][lane] float4 -- so that every store /
      // load instruction of a warp covers 512 contiguous bytes)
      float4* part = nullptr;
      unsigned int* flag = nullptr;
      bool waited = w.role != 1;
      if (w.role != 0) {
        const size_t ws = ((size_t)(w.tile - p.n_whole) * 2 + cta) * UG_EPI_WARPS + (size_t)(warp - 2);
        part = reinterpret_cast<float4*>(p.ts_part) + ws * (size_t)((BN / 32) / (UG_EPI_WARPS / 4) * 8 * 32) + lane;
        flag = p.ts_flag + ws;
      }
      const int m = mb * 256 + (int)cta * UG_BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      constexpr int CH_PER_WARP = (BN / 32) / (UG_EPI_WARPS / 4);
      const int ch0 = ((warp - 2) >> 2) * CH_PER_WARP;   // which column half this warp drains
      // peer scatter: this warp's 32 x 32 transposition buffer behind the barriers (only allocated for such launches)
      const uint32_t stage = smem_base + STAGES * STAGE_BYTES + 128 + (uint32_t)(warp - 2) * 4096u;
#pragma unroll 1
      for (int ch = ch0; ch < ch0 + CH_PER_WARP; ++ch) {
        const int n0 = nb * BN + ch * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tc_ld32(trow + ch * 32, v);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        if (w.role == 2) {         // second half of a split tile: park the partial, no epilogue
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(part + ((ch - ch0) * 8 + j) * 32, make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
          continue;
        }
        if (w.role == 1) {
          if (!waited) {           // the other half's accumulator must have been parked (first chunk only)
            if (lane == 0) {
              unsigned f = 0;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
              } while (f == 0);
            }
            __syncwarp();
            waited = true;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = __ldcg(part + ((ch - ch0) * 8 + j) * 32);
            acc[4 * j] += x.x; acc[4 * j + 1] += x.y; acc[4 * j + 2] += x.z; acc[4 * j + 3] += x.w;
          }
        }
        if (epi_args.scatter()) epi_args.apply_scatter(stage, m, n0, acc, coreM, coreN);
        else epi_args.apply(epoch, m, n0, acc, coreM, coreN, augM, augN);
      }
      if (w.role == 2) {           // publish the partial: every lane's stores, then the flag
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
      } else if (w.role == 1) {    // consumed: the flag is clear again for the next launch that uses this slot
        __syncwarp();
        if (lane == 0) {
          if (!waited) {           // (a warp whose columns all lie beyond N never looked: the flag must be up before it is cleared)
            unsigned f = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
            } while (f == 0);
          }
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");
        }
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


// =====================================================================================================
// Split-K variant for SMALL outputs (the batch-256 layers of configs 2 and 3): a 256 x 500 output is 16
// tiles, so the persistent kernel above leaves 130 SMs idle while 16 CTAs walk all of K at the pace of
// their own TMA feed.  Here a cluster of S = 2 or 4 CTAs shares ONE 128 x BN tile: CTA r contracts a
// contiguous K range (fat stages [r * fs_per, (r+1) * fs_per) of the pass list) into its own TMEM, then
// the partial tiles are reduce-scattered by rows through distributed shared memory -- each CTA stages its
// accumulator in its own shared memory and pushes the 128/S rows owned by CTA j into j's receive buffer with
// ONE bulk DSMEM copy that completes on j's mbarrier -- and every CTA sums the S partials of its rows and
// runs the fused epilogue on them quad by quad, consecutive lanes on consecutive columns (so the epilogue,
// too, is spread over S times the threads and stores whole 128-byte row segments).
// One tile per cluster; no L2 atomics, deterministic summation order (source rank 0..S-1).
// =====================================================================================================
constexpr int SK_KS = 2;                    // k-slabs per stage: 1-2 stages per CTA is the regime here
// 16 epilogue warps (four per TMEM lane quarter, a quarter of the columns each): the exchange and the epilogue of a
// one-tile launch are chains of dependent instructions run ONCE, i.e. ~4 cycles per instruction per warp whatever
// the issue width -- the way to shorten them is fewer instructions per thread and more warps per scheduler.
constexpr int SK_EPI_WARPS = 16;
constexpr int SK_THREADS = 64 + 32 * SK_EPI_WARPS + 32;
constexpr int SK_WARP_PB = 2 + SK_EPI_WARPS;
constexpr int SK_SMEM_MAX = 232448;         // 227 KB opt-in limit per CTA

template <int BN>
struct SkCfg {
  static constexpr int A_SLAB = UG_BM * UG_BK * 2;
  static constexpr int B_SLAB = BN * UG_BK * 2;
  static constexpr int A_BYTES = SK_KS * A_SLAB;
  static constexpr int B_BYTES = SK_KS * B_SLAB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RED_PITCH = BN * 4 + 16;          // +16 B: conflict-free 16-byte row-strided stores
  static constexpr int RED_BYTES = UG_BM * RED_PITCH;    // S sources x 128/S owned rows
  static constexpr int TAIL = 256;                       // barriers + TMEM slot
  static constexpr int STAGES_FIT = (SK_SMEM_MAX - 1024 - RED_BYTES - TAIL) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
  static constexpr int SMEM = STAGES * STAGE_BYTES + RED_BYTES + TAIL + 1024 /*align*/;
  static_assert(STAGES >= 2, "pipeline too shallow");
};

template <int BN, class Epi>
__global__ void __launch_bounds__(SK_THREADS, 1)
umma_gemm_sk_kernel(const __grid_constant__ UmmaGemmParams p, const __grid_constant__ Epi epi_args, int coreM, int coreN,
                    int augM, int augN) {
  using Cfg = SkCfg<BN>;
  constexpr int KS = SK_KS;
  constexpr int BM = UG_BM;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int A_BYTES = Cfg::A_BYTES;
  constexpr int A_SLAB = Cfg::A_SLAB, B_SLAB = Cfg::B_SLAB;
  constexpr int TMEM_COLS = BN;
  static_assert(BN == 64 || BN == 128, "two column halves of whole 32-column chunks; power-of-two TMEM allocation");
  static_assert(STAGES * STAGE_BYTES >= Cfg::RED_BYTES, "the staging tile aliases the operand ring");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* red = smem + STAGES * STAGE_BYTES;            // [S sources][128/S rows][RED_PITCH]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + Cfg::RED_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] partials received
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t red_bar = bar_base + 8u * (2 * STAGES + 1);
  const uint32_t red_base = smem_u32(red);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.split;                                  // cluster size (2 or 4), == %cluster_nctaid.x
  const int rank = (int)cluster_ctarank();
  // grid = (tiles_n * S, tiles_m): no integer division on the way to the first instruction that matters
  const int mb = (int)blockIdx.y, nb = (int)blockIdx.x >> (S == 4 ? 2 : 1);
  // this CTA's share of the fat-stage list (pass 0 stages, then pass 1 stages)
  const int nfs0 = (p.nslab[0] + KS - 1) / KS;
  const int nfs_all = nfs0 + (p.n_ops > 1 ? (p.nslab[1] + KS - 1) / KS : 0);
  const int g0 = rank * p.fs_per;
  const int g1 = min(g0 + p.fs_per, nfs_all);            // host guarantees g0 < g1 for every rank
  if (threadIdx.x == 0) {
    trace_mark_ns(p, 0);
    trace_mark(p, 1);
  }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_ops; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB[i]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(red_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(red_bar, (uint32_t)((S - 1) * (BM / S) * Cfg::RED_PITCH));   // S-1 peers x owned rows
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // every peer is resident and its barriers are initialised before anyone stores into its shared memory (hidden by PDL)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    trace_mark(p, 2);
    trace_mark_ns(p, 8);
  }

  if (warp == 0 || warp == SK_WARP_PB) {
    // ===================== TMA producers: warp 0 feeds A, warp SK_WARP_PB feeds B =====================
    const bool is_pa = warp == 0;
    if (lane == 0) {
      for (int g = g0; g < g1; ++g) {
        const int it = g - g0;
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int pass = g >= nfs0 ? 1 : 0;
        const int slab0 = (g - pass * nfs0) * KS;
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        if (is_pa) {
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          if (!p.transA[pass]) tma_load_3d(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + mb * BM, slab0);
          else                 tma_load_3d(sa, &p.tmA[pass], full_bar(s), 0, p.rowA[pass] + slab0 * UG_BK, mb * (BM / 64));
        } else {
          if (!p.transB[pass]) tma_load_3d(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + nb * BN, slab0);
          else                 tma_load_3d(sb, &p.tmB[pass], full_bar(s), 0, p.rowB[pass] + slab0 * UG_BK, nb * (BN / 64));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t accumulate = 0;
      for (int g = g0; g < g1; ++g) {
        const int it = g - g0;
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int pass = g >= nfs0 ? 1 : 0;
        const int f = g - pass * nfs0;
        const int tA = p.transA[pass], tB = p.transB[pass];
        const uint32_t idesc = make_idesc(BM, BN, tA, tB, pass == 1, p.fmtA[pass], p.fmtB[pass]);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (it == 0) trace_mark(p, 3);
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        const int slabs = min(KS, p.nslab[pass] - f * KS);
#pragma unroll
        for (int sl = 0; sl < KS; ++sl) {
          if (sl < slabs) {
#pragma unroll
            for (int k4 = 0; k4 < UG_BK / UG_UK; ++k4) {
              const uint64_t da = tA ? make_smem_desc(sa + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                     : make_smem_desc(sa + sl * A_SLAB + k4 * 32, 16, 1024);
              const uint64_t db = tB ? make_smem_desc(sb + sl * 8192 + k4 * 2048, KS * 8192, 1024)
                                     : make_smem_desc(sb + sl * B_SLAB + k4 * 32, 16, 1024);
              tc_mma_bf16(tmem_base, da, db, idesc, accumulate);
              accumulate = 1;
            }
          }
        }
        tc_commit(empty_bar(s));
      }
      tc_commit(tfull_bar);   // this CTA's partial tile is complete
      trace_mark(p, 4);
    }
  } else if (warp < SK_WARP_PB) {
    // ============ phase 1 (epilogue warps): partial tile TMEM -> own staging -> bulk DSMEM copies ============
    const int q = warp & 3;
    const Epi ea = params_to_registers(epi_args);         // ... and everything else phase 2 reads from the constant bank
    const int pM = to_register(p.M), pN = to_register(p.N);
    coreM = to_register(coreM); coreN = to_register(coreN); augM = to_register(augM); augN = to_register(augN);
    const int R = BM / S;                                 // rows owned per CTA
    constexpr int QPR = BN / 4;                           // 4-column quads per row
    constexpr int ETHREADS = 32 * SK_EPI_WARPS;
    uint32_t epoch = 0;
    if constexpr (!Epi::kIsUpdate) epoch = ea.e.rng.epoch + (ea.e.rng.epoch_ptr ? *ea.e.rng.epoch_ptr : 0u);
    // Phase-2 work list of this thread: quads tid, tid + 256, ... of the owned rows (16+ consecutive lanes cover a
    // row: coalesced stores).  What the first quad's epilogue reads from global memory (bias / master weights) is
    // requested NOW, so that its latency overlaps the mainloop and the exchange.
    const int quads = R * QPR;
    const int tid = (int)threadIdx.x - 64;
    float pre[4];
    bool has_pre = tid < quads && ea.preload4(mb * BM + rank * R + tid / QPR, nb * BN + (tid % QPR) * 4, pre, coreM, coreN);

    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) trace_mark(p, 5);
    // staging tile [128 rows][RED_PITCH] aliases the operand ring: every MMA of this CTA has completed
    const int row = q * 32 + lane;                        // accumulator row = TMEM lane
    const uint32_t stg = smem_base + (uint32_t)(row * Cfg::RED_PITCH);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int CW = BN / (SK_EPI_WARPS / 4);           // columns drained by this warp: 16 (BN = 64) or 32
    const int col0 = ((warp - 2) >> 2) * CW;
    if (nb * BN + col0 < pN) {                            // warp-uniform; phase 2 skips the same columns
      uint32_t v[CW];
      if constexpr (CW == 32) tc_ld32(trow + col0, v);
      else tc_ld16(trow + col0, v);
#pragma unroll
      for (int j = 0; j < CW / 4; ++j)
        sts128(stg + col0 * 4 + j * 16, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    }
    tc_fence_before();
    if (threadIdx.x == 64) trace_mark(p, 12);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the bulk-copy engine
    asm volatile("bar.sync 1, %0;" ::"n"(32 * SK_EPI_WARPS) : "memory");
    if (lane == 0 && warp - 2 < S - 1) {
      // rows owned by `peer` go to slot `rank` of peer's receive buffer; completion is counted on peer's barrier
      const uint32_t peer = (uint32_t)((rank + 1 + (warp - 2)) % S);
      const uint32_t bytes = (uint32_t)(R * Cfg::RED_PITCH);
      const uint32_t src = smem_base + peer * bytes;
      const uint32_t dst = mapa_shared(red_base + (uint32_t)rank * bytes, peer);
      const uint32_t rbar = mapa_shared(red_bar, peer);
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "r"(src), "r"(bytes), "r"(rbar)
                   : "memory");
    }
    if (threadIdx.x == 64) trace_mark(p, 13);
    mbar_wait(red_bar, 0);      // the S-1 partials of the owned rows have landed
    if (threadIdx.x == 64) trace_mark(p, 14);
    cluster_arrive();           // ... so every peer's copy INTO this CTA is done (peers may retire)
    if (threadIdx.x == 64) trace_mark(p, 10);

    // ============ phase 2: sum the S partials of the owned rows, fused epilogue quad by quad ============
    // A ROLLED loop on purpose: every CTA executes this code once, from a cold instruction cache, where straight-
    // line code runs at ~4 cycles per instruction (profiles/r01_notes.md); a second trip through a small loop body
    // is served from the instruction cache, a second unrolled copy is not.
#pragma unroll 1
    for (int q = tid; q < quads; q += ETHREADS) {
      const int rowl = q / QPR, c4 = q - rowl * QPR;
      const int m = mb * BM + rank * R + rowl;
      const int n4 = nb * BN + c4 * 4;
      float nxt[4];
      const int qn = q + ETHREADS;
      const bool has_nxt = qn < quads && ea.preload4(mb * BM + rank * R + qn / QPR, nb * BN + (qn % QPR) * 4, nxt, coreM, coreN);
      if (m < pM && n4 < pN) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int src = 0; src < 4; ++src) {                 // fixed order: deterministic sums
          if (src < S) {                                    // own partial: staging tile; the others: receive buffer
            const float4 x = lds128((src == rank ? smem_base : red_base) + (uint32_t)((src * R + rowl) * Cfg::RED_PITCH + c4 * 16));
            acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
          }
        }
        if (threadIdx.x == 64 && q == tid) trace_mark(p, 11);
        ea.apply4(epoch, m, n4, acc, has_pre ? pre : nullptr, coreM, coreN, augM, augN);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) pre[j] = nxt[j];
      has_pre = has_nxt;
    }
    if (threadIdx.x == 64) trace_mark(p, 6);
  }
  if (warp < 2 || warp >= SK_WARP_PB) cluster_arrive();

  tc_fence_before();
  __syncthreads();
  cluster_wait();         // this CTA's outgoing copies have been received: its staging tile may go away
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


struct UmmaActEpi {
  static constexpr bool kIsUpdate = false;
  dbn_act_epilogue e;
  bool scatter_host() const { return false; }
  void scatter_order(UmmaGemmParams&) const {}
  __device__ __forceinline__ void apply(uint32_t epoch, int m, int n0, const float acc[32], int M, int N, int, int) const {
    act_epilogue_row<32, true>(e, epoch, m, n0, acc, M, N);
  }
  __device__ __forceinline__ bool scatter() const { return false; }
  __device__ __forceinline__ void apply_scatter(uint32_t, int, int, const float*, int, int) const {}
  // split-K kernel: one quad (4 columns) per loop iteration; preload4 requests early what apply4 would wait for
  __device__ __forceinline__ bool preload4(int m, int n4, float pre[4], int M, int N) const {
    if (e.bias == nullptr || m >= M || n4 + 4 > N || (reinterpret_cast<uintptr_t>(e.bias + n4) & 15) != 0) return false;
    const float4 t = __ldg(reinterpret_cast<const float4*>(e.bias + n4));
    pre[0] = t.x; pre[1] = t.y; pre[2] = t.z; pre[3] = t.w;
    return true;
  }
  __device__ __forceinline__ void apply4(uint32_t epoch, int m, int n4, const float acc[4], const float* pre, int M, int N,
                                         int, int) const {
    act_epilogue_quad<true>(e, epoch, m, n4, acc, M, N, pre);
  }
};
struct UmmaUpdEpi {
  static constexpr bool kIsUpdate = true;
  dbn_update_epilogue e;
  bool scatter_host() const { return e.raw_peers[0] != nullptr; }
  // Tile order of a scattered dW (CTA-pair kernel, 256-row blocks).  The accumulator tiles leave over NVLink as
  // they are finished, so the ORDER of the tiles is the traffic pattern:
  //   * the other owners' row blocks are walked block by block for one column block at a time, so consecutive tiles
  //     go to consecutive owners -- every owner receives all the time, none is the target of a whole wave -- and rank
  //     r starts behind its own rows, so the ranks target different owners at the same moment;
  //   * the rows this rank owns itself come last: the final wave of tiles stores locally and nothing is left on the
  //     wire when the kernel ends.
  // (Measured on 8 x B200, profiles/r02_notes.md: the default raster -- 8 row blocks = 4 owners per super-tile -- hands
  //  the owners of the second half their rows late; one super-tile per owner is slower still.)
  void scatter_order(UmmaGemmParams& p) const {
    if (e.raw_peers[0] == nullptr || e.rows_per_peer <= 0 || (e.rows_per_peer % 256) != 0) return;
    const int per_owner = e.rows_per_peer / 256;
    if (p.tiles_m % per_owner != 0 || p.tiles_m == per_owner) return;
    p.own_start = e.raw_slot * per_owner;
    p.own_rows = per_owner;
  }
  __device__ __forceinline__ void apply(uint32_t, int m, int n0, const float acc[32], int M, int N, int augM, int augN) const {
    upd_epilogue_row<32>(e, m, n0, acc, M, N, augM, augN);
  }
  // Raw delta scattered over peer GPUs (reduce-scatter fused into the contraction).  A TMEM lane is an accumulator
  // ROW, so a warp's natural store is 32 rows x 16 bytes -- over NVLink that is 32 small write packets.  The warp
  // therefore transposes its 32 x 32 chunk through 4 KB of shared memory (XOR-swizzled, conflict-free both ways)
  // and stores 4 rows x 128 contiguous bytes per instruction: full-line packets on the wire.
  __device__ __forceinline__ bool scatter() const { return e.raw_peers[0] != nullptr; }
  __device__ __forceinline__ void apply_scatter(uint32_t stage, int m, int n0, const float acc[32], int M, int N) const {
    const int lane = (int)(threadIdx.x & 31);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(stage + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)), acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    __syncwarp();
    const int m0 = m - lane, slot = lane & 7;
    const int n = n0 + slot * 4;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const int r = i + (lane >> 3);
      const float4 x = lds128(stage + (uint32_t)(r * 128 + ((slot ^ (r & 7)) << 4)));
      const int mr = m0 + r;
      if (mr < M && n < N) {
        float* dst = scatter_row_ptr(e, mr) + n;
        if (n + 4 <= N) {
          asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
        } else {
          const float t[4] = {x.x, x.y, x.z, x.w};
          for (int j = 0; j < N - n; ++j) dst[j] = t[j];
        }
      }
    }
    __syncwarp();   // the buffer is rewritten by the next chunk
  }
  __device__ __forceinline__ bool preload4(int m, int n4, float pre[4], int M, int N) const {   // master weights
    if (e.raw.ptr != nullptr || m >= M || n4 + 4 > N || (e.ldw & 3) != 0 || (reinterpret_cast<uintptr_t>(e.W) & 15) != 0)
      return false;
    const float4 t = *reinterpret_cast<const float4*>(e.W + (size_t)m * (size_t)e.ldw + (size_t)n4);
    pre[0] = t.x; pre[1] = t.y; pre[2] = t.z; pre[3] = t.w;
    return true;
  }
  __device__ __forceinline__ void apply4(uint32_t, int m, int n4, const float acc[4], const float* pre, int M, int N,
                                         int augM, int augN) const {
    upd_epilogue_quad<false>(e, m, n4, acc, M, N, augM, augN, pre);   // the split-K plan is never used for a peer scatter
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
  const CUtensorMapDataType dt = X.dtype == DBN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(map, dt, 3, X.ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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
  return X.ptr != nullptr && (X.dtype == DBN_BF16 || X.dtype == DBN_F16) && (X.ld % 64) == 0 && X.ld >= (int)align_up((size_t)cols, 64) &&
         (reinterpret_cast<uintptr_t>(X.ptr) % 16) == 0;
}

// The tensor-core path takes bf16 operands with 16-byte aligned rows and is worth its 128 x BN tiles
// only for contractions that fill at least part of one; everything else runs on the FP32 SIMT kernel.
inline bool umma_gemm_supported(int M, int N, const dbn_operands* ops, int n_ops) {
  for (int i = 0; i < n_ops; ++i) {
    if (!umma_operand_ok(ops[i].A, ops[i].transA ? M : ops[i].K)) return false;
    if (!umma_operand_ok(ops[i].B, ops[i].transB ? N : ops[i].K)) return false;
    if (ops[i].K < 32) return false;
    if (ops[i].A.dtype != ops[i].B.dtype) return false;   // one operand format per MMA (mixed bf16 / fp16 traps on sm_100a)
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
  static int cache[DBN_MAX_DEVICES] = {0};
  const int dev = current_device();
  int n = cache[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev] = n;
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
  static std::atomic<uint64_t> attr_set{0};   // per device: a second GPU in the same process needs its own opt-in
  if (first_use_on_device(attr_set))
    DBN_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<BN, KS, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
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

// ---- split-K launch -------------------------------------------------------------------------------------
inline bool umma_use_splitk() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_SPLITK");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

template <int BN, class Epi>
inline void sk_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute attr[2], int grid, int S, cudaStream_t st) {
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = SkCfg<BN>::SMEM;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
}

// Clusters of S CTAs of the split-K kernel the device can hold at once (GPC granularity makes this less than
// SMs / S); 0 when the query fails, which disables the split plan.
template <int BN, class Epi>
inline int sk_max_clusters(int S) {
  static int cache_all[DBN_MAX_DEVICES][9];   // per device; stores the answer + 1 (0 = not asked yet)
  if (S < 1 || S > 8) return 0;
  int* cache = cache_all[current_device()];
  if (cache[S] == 0) {
    int n = 0;
    if (cudaFuncSetAttribute(umma_gemm_sk_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<BN>::SMEM) ==
        cudaSuccess) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[2];
      sk_launch_config<BN, Epi>(cfg, attr, S, S, nullptr);
      cfg.numAttrs = 1;   // the occupancy query only wants the cluster shape
      if (cudaOccupancyMaxActiveClusters(&n, umma_gemm_sk_kernel<BN, Epi>, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    const int cap = umma_num_sms() / S;
    cache[S] = (n < cap ? n : cap) + 1;
  }
  return cache[S] - 1;
}

template <int BN, class Epi>
inline int umma_launch_sk(cudaStream_t st, UmmaGemmParams& p, const Epi& epi, int coreM, int coreN, int augM, int augN) {
  p.tiles_m = ceil_div(p.M, UG_BM);
  p.tiles_n = ceil_div(p.N, BN);
  p.group_m = std::max(1, std::min(p.tiles_m, 16));
  const int n_tiles = p.tiles_m * p.tiles_n;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  sk_launch_config<BN, Epi>(cfg, attr, p.tiles_n * p.split, p.split, st);
  cfg.gridDim = dim3(p.tiles_n * p.split, p.tiles_m);
  (void)n_tiles;
  DBN_CUDA_OK(cudaLaunchKernelEx(&cfg, umma_gemm_sk_kernel<BN, Epi>, p, epi, coreM, coreN, augM, augN));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

// Cycle model of one batch-256 launch after its prologue, fitted to the phase traces in profiles/r01_notes.md
// (B200, cycles at ~1.9 GHz; "act" = activation epilogue, "upd" = update epilogue):
//   whole-K kernel : 1000 + 600 per 64-wide k-slab (BN = 64; the TMA feed of ONE SM sustains ~30 B/clk on these
//                    row-strided boxes) + epilogue 3500 (act) / 2300 (upd) per 64 columns
//   split-K kernel : 1350 + 990 per 2-slab stage (BN = 64), exchange 700 + outgoing partial bytes / 16 (DSMEM),
//                    epilogue 400 + 1270 per quad a thread finalises (act) / 430 + 500 (upd)
inline long umma_whole_k_cycles(int bn, long slabs, bool upd) {
  return 1000 + slabs * 600 * (UG_BM + bn) / 192 + (upd ? 2300L : 3500L) * std::max(1, bn / 64);
}
inline long sk_cycles(int bn, int S, int per, bool upd) {
  const long out_bytes = (long)(S - 1) * (UG_BM / S) * (bn * 4 + 16);
  const long quads = std::max(1, (UG_BM / S) * (bn / 4) / (32 * SK_EPI_WARPS));
  return 1350 + (long)per * 990 * (UG_BM + bn) / 192 + 700 + out_bytes / 16 + (upd ? 430 + 500 * quads : 400 + 1270 * quads);
}

constexpr int UG2_KS = 2;   // CTA-pair kernel: 2 slabs (K = 128) per stage, 64 KB per CTA, 3 stages

// ---- tail split of the CTA-pair kernel (see ts_work) ----------------------------------------------------------
// Scratch for the parked partials: TS_SLOTS launches may be in flight on different streams (a fine-tune step forks its
// updates onto a second stream); a launch takes the next slot round-robin.  Allocated once per device, outside any
// stream capture (a launch that meets a capturing stream before the first allocation simply does not split).
constexpr int TS_SLOTS = 4, TS_MAX_TILES = 40;
struct TsScratch {
  float* part;
  unsigned int* flag;
  std::atomic<unsigned> next;
};
inline int& ts_enabled_flag() {   // DBN_B200_TAIL_SPLIT=0 contracts every tile whole; dbn_set_tail_split() overrides (tests, A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_TAIL_SPLIT");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v;
}
inline std::atomic<uint64_t>& ts_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}
inline TsScratch* ts_scratch(cudaStream_t st) {
  static TsScratch sc[DBN_MAX_DEVICES];
  static std::atomic<int> state[DBN_MAX_DEVICES];   // 0 not tried, 1 ready, 2 unavailable
  const int dev = current_device();
  int stt = state[dev].load(std::memory_order_acquire);
  if (stt == 1) return &sc[dev];
  if (stt == 2) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return nullptr;               // not now: allocate at the next eager launch
  }
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  stt = state[dev].load(std::memory_order_acquire);
  if (stt == 1) return &sc[dev];
  if (stt == 2) return nullptr;
  const size_t part_bytes = (size_t)TS_SLOTS * TS_MAX_TILES * 2 * UG_BM * 256 * sizeof(float);
  const size_t flag_bytes = (size_t)TS_SLOTS * TS_MAX_TILES * 2 * UG_EPI_WARPS * sizeof(unsigned int);
  void *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, part_bytes) != cudaSuccess || cudaMalloc(&b, flag_bytes) != cudaSuccess ||
      cudaMemset(b, 0, flag_bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    if (a) cudaFree(a);
    if (b) cudaFree(b);
    state[dev].store(2, std::memory_order_release);
    return nullptr;
  }
  sc[dev].part = static_cast<float*>(a);
  sc[dev].flag = static_cast<unsigned int*>(b);
  sc[dev].next.store(0);
  state[dev].store(1, std::memory_order_release);
  return &sc[dev];
}

template <int BN, class Epi>
inline int umma_launch2(cudaStream_t st, UmmaGemmParams& p, const Epi& epi, int coreM, int coreN, int augM, int augN) {
  constexpr int STAGE_BYTES = UG2_KS * (UG_BM + BN / 2) * UG_BK * 2;
  constexpr int STAGES = UG_SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_PLAIN = STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;
  constexpr int SMEM_MAX = SMEM_PLAIN + UG_EPI_WARPS * 4096;     // + one 32 x 32 fp32 transposition buffer per epilogue warp
  static_assert((2 * STAGES + 4) * 8 + 16 <= 128 && SMEM_MAX <= 232448, "shared-memory budget");
  const int SMEM = epi.scatter_host() ? SMEM_MAX : SMEM_PLAIN;
  static std::atomic<uint64_t> attr_set{0};
  if (first_use_on_device(attr_set))
    DBN_CUDA_OK(cudaFuncSetAttribute(umma_gemm2_kernel<BN, UG2_KS, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
  p.tiles_m = ceil_div(p.M, 256);
  p.tiles_n = ceil_div(p.N, BN);
  p.group_m = std::max(1, std::min(p.tiles_m, 8));
  epi.scatter_order(p);    // peer scatter: the tile order is the traffic pattern
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int clusters = std::min(n_tiles, umma_num_sms() / 2);
  // tail split: the last wave of tiles, when at most half full, is contracted by two clusters per tile (half of K each)
  p.n_whole = n_tiles; p.ts_T0 = 0; p.ts_part = nullptr; p.ts_flag = nullptr;
  int T = 0;
  for (int i = 0; i < p.n_ops; ++i) T += ceil_div(p.nslab[i], UG2_KS);
  const int tail = n_tiles % clusters;
  if (ts_enabled_flag() == 1 && n_tiles > clusters && tail > 0 && 2 * tail <= clusters && tail <= TS_MAX_TILES && T >= 24) {   // (K >= 3072: below that the hand-over of the partial costs more than half a tile)
    if (TsScratch* sc = ts_scratch(st)) {
      const unsigned slot = sc->next.fetch_add(1) % TS_SLOTS;
      p.n_whole = n_tiles - tail;
      p.ts_T0 = T / 2;
      p.ts_part = sc->part + (size_t)slot * TS_MAX_TILES * 2 * UG_BM * 256;
      p.ts_flag = sc->flag + (size_t)slot * TS_MAX_TILES * 2 * UG_EPI_WARPS;
      ts_counter().fetch_add(1);
    }
  }
  DBN_CUDA_OK(launch_pdl(umma_gemm2_kernel<BN, UG2_KS, Epi>, dim3(clusters * 2), dim3(UG_THREADS), SMEM, st, p, epi, coreM,
                         coreN, augM, augN));
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

// relative cost of a 256 x 128 CTA-pair tile against half a 256 x 256 one (more operand bytes per MAC out of L2)
constexpr double UG2_BN128_COST = 1.50;   // measured on B200: 1040 vs 1600 TFLOP/s at 4096-wide layers (L2 operand traffic), so 256 x 128 never pays there
inline int umma_2cta_bn_override() {   // DBN_B200_2CTA_BN=128|256 forces the CTA-pair tile width (experiments)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_2CTA_BN");
    v = e ? atoi(e) : 0;
  }
  return v;
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
  // Small outputs: split K over a cluster of 2 or 4 CTAs per tile if the cycle model says the shorter K walk
  // and the S-times wider epilogue beat the reduction (every CTA must get at least one stage of 2 slabs).
  int sk_bn = 0, sk_S = 0, sk_per = 0;
  if (!two_cta && umma_use_splitk() && !epi.scatter_host()) {
    long total_slabs = 0, fs_all = 0;
    for (int i = 0; i < n_ops; ++i) {
      total_slabs += ceil_div(ops[i].K, UG_BK);
      fs_all += ceil_div(ceil_div(ops[i].K, UG_BK), SK_KS);
    }
    const long tiles_ns = (long)tm * ceil_div(p.N, bn);
    long best_cost = ((tiles_ns + sms - 1) / sms) * umma_whole_k_cycles(bn, total_slabs, Epi::kIsUpdate);
    best_cost -= best_cost / 10;   // the split has to win by 10 %
    for (int cand : {64, 128}) {
      for (int S : {4, 2}) {
        const int per = (int)((fs_all + S - 1) / S);
        if ((fs_all + per - 1) / per != S) continue;             // some CTA would have nothing to contract
        const long tiles = (long)tm * ceil_div(p.N, cand);
        const int room = cand == 64 ? sk_max_clusters<64, Epi>(S) : sk_max_clusters<128, Epi>(S);
        if (tiles > room) continue;
        const long cost = sk_cycles(cand, S, per, Epi::kIsUpdate);
        if (cost < best_cost) { best_cost = cost; sk_bn = cand; sk_S = S; sk_per = per; }
      }
    }
  }
  if (sk_S) bn = sk_bn;
  // CTA-pair tile width: 256 x 256 has the best operand reuse, but a grid of few tiles quantises badly over the
  // 74 SM pairs (4096 x 4096: 256 tiles = 3.46 waves, the last one 46 % full); 256 x 128 tiles halve the tail.
  int bn2 = 256;
  if (two_cta) {
    const int pairs = sms / 2;
    const long t256 = (long)ceil_div(p.M, 256) * ceil_div(p.N, 256), t128 = (long)ceil_div(p.M, 256) * ceil_div(p.N, 128);
    const double c256 = (double)((t256 + pairs - 1) / pairs), c128 = 0.5 * UG2_BN128_COST * (double)((t128 + pairs - 1) / pairs);
    if (c128 < c256) bn2 = 128;
    const int forced = umma_2cta_bn_override();
    if (forced == 128 || forced == 256) bn2 = forced;
  }
  const int ks = two_cta ? UG2_KS : sk_S ? SK_KS : umma_ks(bm, bn);
  const int b_rows = two_cta ? bn2 / 2 : bn;    // B rows (K-major) / columns (MN-major) staged per CTA
  for (int i = 0; i < n_ops; ++i) {
    const dbn_operands& o = ops[i];
    p.K[i] = o.K;
    p.nslab[i] = ceil_div(o.K, UG_BK);
    p.transA[i] = o.transA; p.transB[i] = o.transB;
    p.fmtA[i] = o.A.dtype == DBN_F16 ? 0 : 1;
    p.fmtB[i] = o.B.dtype == DBN_F16 ? 0 : 1;
    p.rowA[i] = o.row0_A; p.rowB[i] = o.row0_B;
    if (!o.transA) DBN_TRY(make_operand_map(&p.tmA[i], o.A, (uint64_t)o.row0_A + p.M, p.nslab[i], bm, ks));
    else           DBN_TRY(make_operand_map(&p.tmA[i], o.A, (uint64_t)o.row0_A + o.K, ceil_div(p.M, 64), ks * 64, bm / 64));
    if (!o.transB) DBN_TRY(make_operand_map(&p.tmB[i], o.B, (uint64_t)o.row0_B + p.N, p.nslab[i], b_rows, ks));
    else           DBN_TRY(make_operand_map(&p.tmB[i], o.B, (uint64_t)o.row0_B + o.K, ceil_div(p.N, 64), ks * 64, b_rows / 64));
  }
  if (sk_S) {
    p.split = sk_S;
    p.fs_per = sk_per;
    if (bn == 128) return umma_launch_sk<128, Epi>(st, p, epi, M, N, augM, augN);
    return umma_launch_sk<64, Epi>(st, p, epi, M, N, augM, augN);
  }
  if (two_cta && bn2 == 128) return umma_launch2<128, Epi>(st, p, epi, M, N, augM, augN);
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
