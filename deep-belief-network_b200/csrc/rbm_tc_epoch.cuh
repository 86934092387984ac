// rbm_tc_epoch.cuh -- one PERSISTENT tcgen05 kernel per layer-epoch of CD-k at batch 256 (configs 2 / 3).
//
// The tiled path spends a mini-batch of 256 rows in four dependent launches (v->h, h->v, v->h, dW), each of which
// pays a prologue, a cold first-operand fetch and a teardown: 22-34 us per step for < 1 us of tensor-core time.
// Here ONE cooperative launch trains the whole epoch (dbn/models.py:105-119): <= 148 co-resident CTAs walk the
// phases of every mini-batch and meet at a grid barrier (one L2 counter) between phases instead of at a launch
// boundary.  TMEM, mbarriers and tensor maps are set up once per epoch; per mini-batch
//
//   * the W operands of a CTA are fetched ONCE and stay in shared memory for the step: the row slice
//     W[h-slice, :] serves both hidden phases (G1, G3), the column slice W[:, v-slice] the visible phase (G2);
//     both are requested before the barrier that releases the phase, as is the next mini-batch's input tile,
//     so only activations produced by the previous phase are fetched on the critical path;
//   * P0, the hidden sample, V1 and Pk are written once (16-bit) and read back out of L2 -- they never reach HBM;
//   * every phase is a full-K tile on UMMA M = 64 (rows r of a tile in TMEM lanes 32*(r/16) + r%16), except the
//     visible phase of very wide hidden layers, whose K is split over 2/4 CTAs: partial tiles go through an L2
//     scratch and the LAST arriver sums them in a fixed order (deterministic);
//   * dW = P0^T V0 - Pk^T Vk is contracted per 64 x n4 tile (K = batch) and applied to the fp32 master by the
//     tile's owner, which also refreshes the 16-bit shadow; db / dc are fixed-point column sums taken in the
//     phase epilogues (dbn_colsum) and folded into b / c by their owner threads.
//
// Warp roles (20 warps): warp 0 = TMA producer (one lane), warps 1-3 = MMA issuers (one lane each; warp 1 also owns
// the TMEM allocation), warps 4-19 = epilogue (four warps per TMEM lane quarter, 8-column chunks dealt round robin)
// + grid barrier.
// Results match the tiled path to fp32 round-off (other summation order inside a dot product), samples are
// drawn from the same Philox counters.  Every wait is bounded: a protocol bug traps instead of hanging.
#pragma once
#include "umma_gemm.cuh"

namespace dbn {

constexpr int TE_EPI_WARPS = 16;
constexpr int TE_EPI_THREADS = 32 * TE_EPI_WARPS;
constexpr int TE_NISS = 3;                 // MMA issuer warps
constexpr int TE_WARP_EPI = 1 + TE_NISS;   // first epilogue warp (a multiple of 4: warp % 4 = TMEM lane quarter)
constexpr int TE_THREADS = 32 * TE_WARP_EPI + TE_EPI_THREADS;   // 20 warps (registers are allotted per 4 warps: same budget as 18)
constexpr int TE_MAX_TILE_CTRS = 480;
constexpr int TE_PLAN_CTAS = 148;          // SMs of a B200: the plan never asks for more CTAs than can be co-resident
constexpr int TE_SMEM_MAX = 232448;        // 227 KB opt-in limit per CTA
constexpr int TE_BOX_BYTES = 256 * 128;    // one MN-major operand chunk: 256 k rows x 64 elements
// One thread issues a 64 x N x 16 MMA every ~70 cycles (descriptor arithmetic + the election wrapper, one warp's
// dependent-issue latency) while the tensor pipe needs N/2 cycles for it: at N = 32..56 the ISSUE is the bound of a
// phase.  The k-steps of a tile are therefore dealt to TE_NISS issuer warps in contiguous blocks of 4 (one 64-wide
// slab); issuer w accumulates into its own TMEM columns [w * width, (w + 1) * width) and the epilogue adds the
// partial accumulators up.  Every phase starts at column 0: a phase's MMAs cannot start before this CTA's epilogue
// of the previous phase has drained TMEM (its operands arrive behind the barrier that epilogue arrives at).
constexpr int TE_TRACE_STEPS = 8, TE_TRACE_SLOTS = 8 * 16;        // debug trace: steps x (phases x 16 stamps)

struct TeCtl {                 // device control block, zeroed before every launch
  unsigned int arrived;        // grid barrier: arrivals so far (monotonic)
  unsigned int error;
  unsigned int pad[30];
  unsigned int tile_done[TE_MAX_TILE_CTRS];   // split-K visible phase: partials stored per output tile (monotonic)
};

struct TePlan {
  int grid;
  int rb_h, hs, nt_h_rows, nt_h_cols, kv_slabs;            // hidden phases: rb_h rows x hs units, K = V
  int rb_v, vs, nt_v_rows, nt_v_cols, ksplit, kc_slabs;    // visible phase: rb_v rows x vs units x K chunk of kc_slabs*64
  int n4, n4_chunks, nt_u_rows, nt_u_cols;                 // update: 64 hidden x n4 visible, K = batch
  uint32_t off_wr, off_act, off_wc, off_g4p, off_v0c, off_v1c, off_p1, off_bars;
  int v1c_early, p1_in_act;
  uint32_t smem_bytes;
  size_t scratch_bytes;
};

struct TeArgs {
  CUtensorMap tm_x3, tm_vt3, tm_hs3, tm_wb3;             // K-major operands {64, rows, 64-wide slabs}
  CUtensorMap tm_wb2, tm_p02, tm_pk2, tm_x2, tm_vt2;     // MN-major operands {columns, rows}
  TePlan pl;
  int V, H, B, k, act, dtype;
  int n_rows, row0;
  float scale;
  dbn_mat X, P0, Hs, Vt, Pk, Wb;
  float* W;
  int ldw;
  float* b;
  float* c;
  unsigned long long* stats;   // [H] dc | [V] db, fixed point
  dbn_rng rng;
  TeCtl* ctl;
  float* scratch;
  int exp;                     // debug (DBN_B200_TE_EXP): experiment switches, 0 in production
  long long* trace;            // debug (dbn_debug_set_trace): [cta][step < 8][phase < 8][16] clock64 stamps: 0 operands landed
                               // (issuer 0), 1 accumulator ready (epilogue thread 0), 2 its epilogue done, 3 CTA done, 4 arrived
                               // at the grid barrier, 5 barrier passed, 6 MMAs complete (issuer 0), 7 prefetch done, 8 first
                               // accumulator chunk in registers, 9 first chunk stored, 10 MMAs issued (issuer 0), 11 loads issued
};

enum { TE_B_WR = 0, TE_B_WC, TE_B_ACT, TE_B_G4P, TE_B_V0C, TE_B_P1, TE_B_ACC, TE_NBARS };
enum { TE_HID0 = 0, TE_VIS = 1, TE_HID = 2, TE_HIDK = 3, TE_UPD = 4 };

// blocks of 4 k-steps per issuer, and the number of issuers that get any (= partial accumulators to add up)
__device__ __forceinline__ int te_units_per_issuer(int nunits) { return (nunits + TE_NISS - 1) / TE_NISS; }
__device__ __forceinline__ int te_nacc(int nunits) {
  const int per = te_units_per_issuer(nunits);
  return (nunits + per - 1) / per;
}

__device__ __forceinline__ int te_phase_kind(int p, int np) {
  if (p == 0) return TE_HID0;
  if (p == np - 1) return TE_UPD;
  if (p == np - 2) return TE_HIDK;
  return (p & 1) ? TE_VIS : TE_HID;
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// one lane of a converged warp (the compiler then keeps warp-uniform values -- descriptors, coordinates -- in
// uniform registers: no per-instruction R2UR shuffle in front of every UTCHMMA / UTMALDG)
__device__ __forceinline__ bool te_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// A tile row's 8-column chunk sits in lanes 0-15 of a TMEM quarter (UMMA M = 64).  All TE_NISS partial accumulators
// are requested in one go (one wait; those of issuers without k-steps hold stale but allocated columns and are
// dropped) and their sum is split over the whole warp: lane l < 16 keeps columns 0-3 of row l, lane l + 16 takes
// columns 4-7, so all 32 lanes run the element-wise tail on 4 columns each.
__device__ __forceinline__ void te_ld_acc4(uint32_t taddr, int nacc, int width, int lane, float x[4]) {
  static_assert(TE_NISS == 3, "three partial accumulators");
  uint32_t v0[8], v1[8], v2[8];
  tc_ld8_nowait(taddr, v0);
  tc_ld8_nowait(taddr + (uint32_t)width, v1);
  tc_ld8_nowait(taddr + (uint32_t)(2 * width), v2);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const bool hi = lane >= 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // (a stale accumulator may hold anything, NaN included: select, do not multiply)
    const float lo_s = __uint_as_float(v0[j]) + (nacc > 1 ? __uint_as_float(v1[j]) : 0.0f) + (nacc > 2 ? __uint_as_float(v2[j]) : 0.0f);
    const float hi_s = __uint_as_float(v0[4 + j]) + (nacc > 1 ? __uint_as_float(v1[4 + j]) : 0.0f) + (nacc > 2 ? __uint_as_float(v2[4 + j]) : 0.0f);
    const float got = __shfl_sync(0xffffffffu, hi_s, lane & 15);
    x[j] = hi ? got : lo_s;
  }
}

__device__ __noinline__ void te_fail(TeCtl* ctl, int code) {
  atomicExch(&ctl->error, (unsigned)code);
  printf("dbn_b200: persistent epoch kernel gave up (code %d, block %d thread %d)\n", code, blockIdx.x, threadIdx.x);
  __trap();
}

// bounded wait (a protocol bug traps instead of hanging the GPU); the failure path is out of line
__device__ __forceinline__ void te_mbar_wait(TeCtl* ctl, uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 24)) te_fail(ctl, 3);
  }
}

// ---- grid barrier: one monotonic arrival counter in L2 ---------------------------------------------------
// Writers: every epilogue thread has stored its outputs and met the others at a CTA barrier; one thread then
// publishes them (release) and arrives.  The activations are read back by OTHER CTAs through TMA, i.e. through the
// async proxy: a proxy fence orders the generic-proxy stores against those reads.
__device__ __forceinline__ void te_grid_arrive(TeCtl* ctl, int exp) {
  if (!(exp & 2)) asm volatile("fence.proxy.async.global;" ::: "memory");
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ctl->arrived) : "memory");   // the release publishes the CTA's stores (bar.sync made them this thread's)
}
// Everything a CTA reads of another CTA's output after the barrier comes straight from L2 (TMA, ld.global.cg), so the
// poll is a relaxed L2 load: no per-iteration acquire (which would also flush L1 every time round the loop).
__device__ __forceinline__ void te_grid_wait(TeCtl* ctl, unsigned target) {
  unsigned v = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&ctl->arrived) : "memory");
    if (v >= target) break;
    if (spin > (1u << 22)) te_fail(ctl, 1);
  }
}
// progress counter in shared memory (epilogue thread -> producer): grid barriers passed so far
__device__ __forceinline__ void te_publish(volatile unsigned* slot, unsigned v) {
  __threadfence_block();
  *slot = v;
}
__device__ __forceinline__ void te_await(TeCtl* ctl, volatile unsigned* slot, unsigned v) {
  for (unsigned spin = 0; *slot < v; ++spin) {
    __nanosleep(20);      // a tight shared-memory spin would take issue slots from the epilogue warps on this scheduler
    if (spin > (1u << 24)) te_fail(ctl, 2);
  }
  __threadfence_block();
}

// ---- small vector helpers (4 consecutive columns, n0 % 4 == 0) --------------------------------------------
// 4 floats of a vector that is rewritten inside the kernel (b, c, master W): L2 only, never a stale L1 line
__device__ __forceinline__ void te_load4_cg(const float* p, int n0, int N, float v[4]) {
  if (n0 + 4 <= N && (reinterpret_cast<uintptr_t>(p + n0) & 15) == 0) {
    const float4 x = __ldcg(reinterpret_cast<const float4*>(p + n0));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (n0 + j < N) ? __ldcg(p + n0 + j) : 0.0f;
  }
}
__device__ __forceinline__ void te_store4_f32(float* p, int n0, int N, const float v[4]) {
  if (n0 + 4 <= N && (reinterpret_cast<uintptr_t>(p + n0) & 15) == 0) {
    *reinterpret_cast<float4*>(p + n0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + j < N) p[n0 + j] = v[j];
  }
}
// 16-bit rows: pitch a multiple of 64 elements, so 4 elements at a multiple of 4 are 8-byte aligned
__device__ __forceinline__ void te_store4_h(const dbn_mat& m, int r, int n0, int N, const float v[4]) {
  if (n0 + 4 <= N) {
    uint2 pk;
    pk.x = pack16x2(m.dtype, v[0], v[1]); pk.y = pack16x2(m.dtype, v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(m.ptr) + (size_t)r * (size_t)m.ld + (size_t)n0) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + j < N) mat_store(m, r, n0 + j, v[j]);
  }
}
__device__ __forceinline__ void te_load4_h(const dbn_mat& m, int r, int n0, int N, float v[4]) {
  if (n0 + 4 <= N) {
    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(m.ptr) + (size_t)r * (size_t)m.ld + (size_t)n0));
    unpack16x2(m.dtype, pk.x, v[0], v[1]); unpack16x2(m.dtype, pk.y, v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (n0 + j < N) ? mat_load(m, r, n0 + j) : 0.0f;
  }
}
// Column sums over the 16 rows of a half-warp (one row per lane, 4 columns per lane): a butterfly that halves the
// columns a lane carries per step.  Afterwards the lanes with (lane & 3) == 0 hold the sum of column 2*bit3 + bit2.
__device__ __forceinline__ float te_colsum4(const float x[4], int lane) {
  float y[2];
  bool up = (lane & 8) != 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = up ? x[j] : x[j + 2], keep = up ? x[j + 2] : x[j];
    y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = (lane & 4) != 0;
  const float send = up ? y[0] : y[1], keep = up ? y[1] : y[0];
  float w = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  w += __shfl_xor_sync(0xffffffffu, w, 2);
  w += __shfl_xor_sync(0xffffffffu, w, 1);
  return w;
}
__device__ __forceinline__ void te_stat_commit(unsigned long long* acc, int n0, int N, float w, int lane) {
  const int n = n0 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  if ((lane & 3) == 0 && n < N) stat_add(acc, n, w);
}
__device__ __forceinline__ void te_stat_commit_x(int exp, unsigned long long* acc, int n0, int N, float w, int lane) {
  if (!(exp & 1)) te_stat_commit(acc, n0, N, w, lane);
}
// consume a statistic (its one owner thread): straight from L2 -- the atomics of the other CTAs landed there
__device__ __forceinline__ float te_stat_take(unsigned long long* acc, int i) {
  const unsigned long long v = __ldcg(acc + i);
  __stcg(acc + i, 0ull);
  return stat_value(v);
}

__device__ __forceinline__ void te_trace(const TeArgs& a, int step, int p, int slot) {
  if (a.trace != nullptr && step < TE_TRACE_STEPS && p < 8)
    a.trace[((size_t)blockIdx.x * TE_TRACE_STEPS + step) * TE_TRACE_SLOTS + p * 16 + slot] = clock64();
}

// ---- the activation tail of a hidden / visible phase on 4 columns of one row ---------------------------
// hid: x = act(acc + c) -> means (P0 / Pk), sample = (u < x) -> Hs (dbn/models.py:148-155, :176-183)
// vis: x = act(acc + b) -> Vt                                  (dbn/models.py:195-201)
// rows at or beyond the valid ones (short last mini-batch) store zeros, so the dW contraction over the whole
// batch tile adds nothing for them.  Column statistics: dc / db (dbn/models.py:143-144).
//
// te_act_prefetch runs BEFORE the accumulator is ready and fetches / computes everything that does not depend on
// it: the bias segment (L2), the uniforms (one Philox call; or the injected values) and, for the last visible
// phase, the v_0 segment of the db statistic.  te_act_finish is what is left behind the MMAs.
__device__ __forceinline__ void te_act_prefetch(const TeArgs& a, bool hid, int kind, int t, uint32_t epoch, int step, int grow,
                                                bool row_ok, int n0, float bias[4], float aux[4]) {
  const int N = hid ? a.H : a.V;
  te_load4_cg(hid ? a.c : a.b, n0, N, bias);
#pragma unroll
  for (int j = 0; j < 4; ++j) aux[j] = 0.0f;
  if (hid) {
    if (kind != TE_HIDK) {
      const int erow = step * a.B + grow;           // row inside the epoch slice: the Philox counter (GPU-count invariant)
      if (a.rng.injected.ptr != nullptr) {
        const float* up = reinterpret_cast<const float*>(a.rng.injected.ptr) + (size_t)erow * (size_t)a.rng.injected.ld + (size_t)t * a.H;
#pragma unroll
        for (int j = 0; j < 4; ++j) aux[j] = (row_ok && n0 + j < N) ? __ldg(up + n0 + j) : 1.0f;
      } else {
        dbn_rng g = a.rng;
        g.stream = a.rng.stream + (uint32_t)t;
        rng_uniform4(g, epoch, erow, n0, aux);
      }
    }
  } else if (t == a.k - 1 && row_ok) {
    te_load4_h(a.X, a.row0 + step * a.B + grow, n0, N, aux);                // v_0 for db = sum_rows (v_0 - v_k)
  }
}
__device__ __forceinline__ void te_act_finish(const TeArgs& a, bool hid, int kind, int t, int grow, bool in_tile, bool row_ok,
                                              int n0, const float acc[4], const float bias[4], const float aux[4], int lane) {
  const int N = hid ? a.H : a.V;
  float x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = row_ok ? apply_act<true>(a.act, acc[j] + bias[j]) : 0.0f;
  if (hid) {
    if (in_tile) te_store4_h((kind == TE_HID0) ? a.P0 : a.Pk, grow, n0, N, x);
    if (kind != TE_HIDK) {
      float s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = (aux[j] < x[j]) ? 1.0f : 0.0f;   // strict '<': dbn/models.py:155
      if (in_tile) te_store4_h(a.Hs, grow, n0, N, s);
    }
    if (kind == TE_HID0 || kind == TE_HIDK) {
      if (kind == TE_HIDK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = -x[j];
      }
      te_stat_commit_x(a.exp, a.stats, n0, N, te_colsum4(x, lane), lane);          // dc = sum_rows (h_0 - h_k)
    }
  } else {
    if (in_tile) te_store4_h(a.Vt, grow, n0, N, x);
    if (t == a.k - 1) {                                                     // db = sum_rows (v_0 - v_k)
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = row_ok ? aux[j] - x[j] : 0.0f;
      te_stat_commit_x(a.exp, a.stats + a.H, n0, N, te_colsum4(x, lane), lane);
    }
  }
}

__global__ void __launch_bounds__(TE_THREADS, 1) rbm_tc_epoch_kernel(const __grid_constant__ TeArgs a) {
  extern __shared__ uint8_t te_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(te_smem_raw) + 1023) & ~uintptr_t(1023));
  const TePlan& pl = a.pl;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TE_NBARS);
  volatile unsigned* gb_passed = reinterpret_cast<volatile unsigned*>(tmem_slot + 1);   // grid barriers this CTA has passed
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler, too
  const int lane = threadIdx.x & 31;
  const int bid = (int)blockIdx.x;
  const int np = 2 * a.k + 2;                         // phases per mini-batch
  const int steps = (a.n_rows + a.B - 1) / a.B;
  const int fmt = a.dtype == DBN_F16 ? 0 : 1;

  // tiles of this CTA (one per phase kind, or none)
  const int nt_h = pl.nt_h_rows * pl.nt_h_cols, nt_v = pl.nt_v_rows * pl.nt_v_cols * pl.ksplit, nt_u = pl.nt_u_rows * pl.nt_u_cols;
  const bool has_h = bid < nt_h, has_v = bid < nt_v, has_u = bid < nt_u;
  const int h_rb = bid / pl.nt_h_cols, h_cb = bid % pl.nt_h_cols;
  const int v_kc = bid % pl.ksplit, v_cb = (bid / pl.ksplit) % pl.nt_v_cols, v_rb = bid / (pl.ksplit * pl.nt_v_cols);
  const int u_mb = bid / pl.nt_u_cols, u_nb = bid % pl.nt_u_cols;

  const uint32_t wr_bytes = (uint32_t)(pl.hs * 128 * pl.kv_slabs);
  const uint32_t acth_bytes = (uint32_t)(pl.rb_h * 128 * pl.kv_slabs);
  const uint32_t actv_bytes = (uint32_t)(pl.rb_v * 128 * pl.kc_slabs);
  const int wc_rows = pl.kc_slabs * 64;               // hidden rows of this CTA's K chunk (boxes of <= 256 rows)
  const int wc_boxes = (wc_rows + 255) / 256;
  const int wc_box_rows = wc_rows < 256 ? wc_rows : 256;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* maps[9] = {&a.tm_x3, &a.tm_vt3, &a.tm_hs3, &a.tm_wb3, &a.tm_wb2, &a.tm_p02, &a.tm_pk2, &a.tm_x2, &a.tm_vt2};
    for (int i = 0; i < 9; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(maps[i]) : "memory");
    for (int i = 0; i < TE_NBARS; ++i) mbar_init(bar(i), i == TE_B_ACC ? TE_NISS : 1);
    *gb_passed = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (te_elect()) {
      unsigned n_gb = 0, n_acc = 0;
      auto load_v0 = [&](int s) {   // input rows of mini-batch s for this CTA's hidden tile
        mbar_expect_tx(bar(TE_B_ACT), acth_bytes);
        tma_load_3d(sbase + pl.off_act, &a.tm_x3, bar(TE_B_ACT), 0, a.row0 + s * a.B + h_rb * pl.rb_h, 0);
      };
      for (int s = 0; s < steps; ++s) {
        const int r0 = a.row0 + s * a.B;
        // (the operands on the critical path -- what the phase behind a barrier waits for: activation rows, P1, the
        //  W row slice of the next step -- are requested by the epilogue thread that watches the barrier open; this
        //  warp fetches everything that can come early)
        if (has_h && s == 0) {
          mbar_expect_tx(bar(TE_B_WR), wr_bytes);
          tma_load_3d(sbase + pl.off_wr, &a.tm_wb3, bar(TE_B_WR), 0, h_cb * pl.hs, 0);
          load_v0(0);
        }
        if (has_v) {
          mbar_expect_tx(bar(TE_B_WC), (uint32_t)(wc_boxes * wc_box_rows * 128));
          for (int i = 0; i < wc_boxes; ++i)
            tma_load_2d(sbase + pl.off_wc + (uint32_t)(i * TE_BOX_BYTES), &a.tm_wb2, bar(TE_B_WC), v_cb * pl.vs,
                        v_kc * wc_rows + i * 256);
        }
        for (int p = 0; p < np; ++p) {
          const int kind = te_phase_kind(p, np);
          const bool mine = kind == TE_UPD ? has_u : (kind == TE_VIS ? has_v : has_h);
          if (mine) ++n_acc;
          if (p == np - 1) {
            // P1 sits in the activation buffer: the next input tile follows once the update's MMAs have read it
            if (pl.p1_in_act && has_u && has_h && s + 1 < steps) {
              te_mbar_wait(a.ctl, bar(TE_B_ACC), (n_acc - 1) & 1u);
              load_v0(s + 1);
            }
            if (s + 1 == steps) break;
          }
          te_await(a.ctl, gb_passed, ++n_gb);        // every CTA has finished phase p: its outputs are visible
          if (p == np - 1) break;
          const int next = te_phase_kind(p + 1, np);
          if (p == 0 && has_u) {                     // P0 is final: its tile of the dW contraction may come now
            mbar_expect_tx(bar(TE_B_G4P), (uint32_t)TE_BOX_BYTES);
            tma_load_2d(sbase + pl.off_g4p, &a.tm_p02, bar(TE_B_G4P), u_mb * 64, 0);
          }
          if (next == TE_HIDK && has_u) {            // last visible phase done: V0 / Vk column tiles (the W column slice is dead)
            mbar_expect_tx(bar(TE_B_V0C), (uint32_t)((pl.v1c_early ? 2 : 1) * pl.n4_chunks * TE_BOX_BYTES));
            for (int ch = 0; ch < pl.n4_chunks; ++ch)
              tma_load_2d(sbase + pl.off_v0c + (uint32_t)(ch * TE_BOX_BYTES), &a.tm_x2, bar(TE_B_V0C), u_nb * pl.n4 + ch * 64, r0);
            if (pl.v1c_early)
              for (int ch = 0; ch < pl.n4_chunks; ++ch)
                tma_load_2d(sbase + pl.off_v1c + (uint32_t)(ch * TE_BOX_BYTES), &a.tm_vt2, bar(TE_B_V0C), u_nb * pl.n4 + ch * 64, 0);
          }
          if (next == TE_UPD && (!pl.p1_in_act || !has_u) && has_h && s + 1 < steps) load_v0(s + 1);
        }
      }
    }
  } else if (warp < TE_WARP_EPI) {
    // ======================================= MMA issuers =======================================
    if (te_elect()) {
      const int w = warp - 1;
      unsigned n_act = 0, n_accw = 0;
      for (int s = 0; s < steps; ++s) {
        for (int p = 0; p < np; ++p) {
          const int kind = te_phase_kind(p, np);
          const bool mine = kind == TE_UPD ? has_u : (kind == TE_VIS ? has_v : has_h);
          if (!mine) continue;
          int nk, width;
          if (kind == TE_UPD) {
            te_mbar_wait(a.ctl, bar(TE_B_G4P), (unsigned)s & 1u);
            te_mbar_wait(a.ctl, bar(TE_B_V0C), (unsigned)s & 1u);
            te_mbar_wait(a.ctl, bar(TE_B_P1), (unsigned)s & 1u);
            nk = (a.B + 15) / 16;                    // K = batch rows (zero rows beyond the valid ones)
            width = pl.n4;
          } else if (kind == TE_VIS) {
            if (p == 1) te_mbar_wait(a.ctl, bar(TE_B_WC), (unsigned)s & 1u);
            te_mbar_wait(a.ctl, bar(TE_B_ACT), n_act & 1u);
            ++n_act;
            nk = (min(wc_rows, a.H - v_kc * wc_rows) + 15) / 16;   // hidden units of this chunk (rows beyond H are zero-filled)
            width = pl.vs;
          } else {
            if (p == 0) te_mbar_wait(a.ctl, bar(TE_B_WR), (unsigned)s & 1u);
            te_mbar_wait(a.ctl, bar(TE_B_ACT), n_act & 1u);
            ++n_act;
            nk = (a.V + 15) / 16;
            width = pl.hs;
          }
          tc_fence_after();
          if (w == 0) te_trace(a, s, p, 0);
          const int upp = (nk + 3) >> 2;                               // blocks of 4 k-steps per pass
          const int nunits = kind == TE_UPD ? 2 * upp : upp;
          const int per = te_units_per_issuer(nunits);
          const int u0 = w * per, u1 = min(nunits, u0 + per);
          const uint32_t tmem_d = tmem_base + (uint32_t)(w * width);
          // descriptor bases and their increments (address field: bytes >> 4) per block of 4 k-steps / per k-step
          uint64_t dA, dB, dA1 = 0, dB1 = 0;
          uint32_t ua, ub, ia, ib, idesc, idesc1 = 0;
          if (kind == TE_UPD) {                      // both operands MN-major: 16 k rows = 2048 bytes per k-step
            idesc = make_idesc(64, width, 1, 1, 0, fmt, fmt);
            idesc1 = make_idesc(64, width, 1, 1, 1, fmt, fmt);
            dA = make_smem_desc(sbase + pl.off_g4p, TE_BOX_BYTES, 1024); dB = make_smem_desc(sbase + pl.off_v0c, TE_BOX_BYTES, 1024);
            dA1 = make_smem_desc(sbase + pl.off_p1, TE_BOX_BYTES, 1024); dB1 = make_smem_desc(sbase + pl.off_v1c, TE_BOX_BYTES, 1024);
            ua = 512; ub = 512; ia = 128; ib = 128;
          } else if (kind == TE_VIS) {               // A K-major (32 bytes per k-step inside a slab), B MN-major
            idesc = make_idesc(64, width, 0, 1, 0, fmt, fmt);
            dA = make_smem_desc(sbase + pl.off_act, 16, 1024); dB = make_smem_desc(sbase + pl.off_wc, TE_BOX_BYTES, 1024);
            ua = (uint32_t)(pl.rb_v * 128) >> 4; ub = 512; ia = 2; ib = 128;
          } else {                                   // both K-major
            idesc = make_idesc(64, width, 0, 0, 0, fmt, fmt);
            dA = make_smem_desc(sbase + pl.off_act, 16, 1024); dB = make_smem_desc(sbase + pl.off_wr, 16, 1024);
            ua = (uint32_t)(pl.rb_h * 128) >> 4; ub = (uint32_t)(pl.hs * 128) >> 4; ia = 2; ib = 2;
          }
          uint32_t accumulate = 0;
#pragma unroll 1
          for (int u = u0; u < u1; ++u) {
            const int pass = u >= upp ? 1 : 0, us = u - pass * upp;
            const int ks = min(4, nk - 4 * us);
            const uint64_t da = (pass ? dA1 : dA) + (uint64_t)(ua * (uint32_t)us), db = (pass ? dB1 : dB) + (uint64_t)(ub * (uint32_t)us);
            const uint32_t id = pass ? idesc1 : idesc;
            if (ks == 4) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                tc_mma_bf16(tmem_d, da + (uint64_t)(ia * j), db + (uint64_t)(ib * j), id, accumulate);
                accumulate = 1;
              }
            } else {
              for (int j = 0; j < ks; ++j) {
                tc_mma_bf16(tmem_d, da + (uint64_t)(ia * j), db + (uint64_t)(ib * j), id, accumulate);
                accumulate = 1;
              }
            }
          }
          if (w == 0) te_trace(a, s, p, 10);
          tc_commit(bar(TE_B_ACC));   // one arrival per issuer, when its MMAs (if any) have completed
          if (w == 0 && a.trace != nullptr) {
            te_mbar_wait(a.ctl, bar(TE_B_ACC), n_accw & 1u);
            te_trace(a, s, p, 6);
          }
          ++n_accw;
        }
      }
    }
  } else {
    // ============================ epilogue warps (+ grid barrier) ============================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int g = (warp - TE_WARP_EPI) >> 2;     // column group: chunks g, g + 4, ...
    const int tid_e = (int)threadIdx.x - 32 * TE_WARP_EPI;
    const int rloc = q * 16 + (lane & 15);       // UMMA M = 64: tile row r lives in lane 32*(r/16) + r%16 ...
    const int coff = (lane >> 4) * 4;            // ... and its 8-column chunk is split over lanes l (0-3) and l + 16 (4-7)
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t epoch = a.rng.epoch + (a.rng.epoch_ptr ? *a.rng.epoch_ptr : 0u);
    unsigned n_acc = 0, n_bar = 0, n_vis = 0;
    for (int s = 0; s < steps; ++s) {
      const int valid = min(a.B, a.n_rows - s * a.B);     // rows of this mini-batch (dbn/utils.py:14-22: short last batch)
      for (int p = 0; p < np; ++p) {
        const int kind = te_phase_kind(p, np);
        const bool mine = kind == TE_UPD ? has_u : (kind == TE_VIS ? has_v : has_h);
        if (mine && kind == TE_UPD) {
          // ---- W += scale * dW on the fp32 master, refreshed 16-bit shadow (dbn/models.py:117) ----
          // The master segments, the biases and their statistics are requested BEFORE the accumulator is ready
          // (the statistics are final since the last barrier), so their L2 latency hides behind the MMAs.
          const int m = u_mb * 64 + rloc;
          const int chunks = pl.n4 >> 3;           // <= 16: at most 4 chunks per warp
          const int nacc = te_nacc(2 * (((a.B + 15) / 16 + 3) >> 2));
          const bool row_ok = m < a.H;
          float* wrow = a.W + (size_t)m * (size_t)a.ldw;
          float wpre[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n0 = u_nb * pl.n4 + (g + 4 * i) * 8 + coff;
            if (g + 4 * i < chunks && n0 < a.V && row_ok) te_load4_cg(wrow, n0, a.V, wpre[i]);
          }
          const int hc = u_mb * 64 + tid_e, vb = u_nb * pl.n4 + tid_e;
          const bool own_c = u_nb == 0 && tid_e < 64 && hc < a.H, own_b = u_mb == 0 && tid_e < pl.n4 && vb < a.V;
          float cval = 0.f, cst = 0.f, bval = 0.f, bst = 0.f;
          if (own_c) { cval = __ldcg(a.c + hc); cst = te_stat_take(a.stats, hc); }
          if (own_b) { bval = __ldcg(a.b + vb); bst = te_stat_take(a.stats + a.H, vb); }
          if (tid_e == 0) te_trace(a, s, p, 7);
          te_mbar_wait(a.ctl, bar(TE_B_ACC), n_acc & 1u);
          ++n_acc;
          tc_fence_after();
          if (tid_e == 0) te_trace(a, s, p, 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = g + 4 * i;
            if (ch < chunks && u_nb * pl.n4 + ch * 8 < a.V) {         // warp-uniform
              const int n0 = u_nb * pl.n4 + ch * 8 + coff;
              float acc[4];
              te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, pl.n4, lane, acc);
              if (tid_e == 0 && i == 0) te_trace(a, s, p, 8);
              if (row_ok && n0 < a.V) {
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = fmaf(a.scale, acc[j], wpre[i][j]);
                te_store4_f32(wrow, n0, a.V, w);
                te_store4_h(a.Wb, m, n0, a.V, w);
              }
              if (tid_e == 0 && i == 0) te_trace(a, s, p, 9);
            }
          }
          // ---- b, c += scale * statistics, by the one owner thread of every unit (:118-119) ----
          if (own_c) a.c[hc] = cval + a.scale * cst;
          if (own_b) a.b[vb] = bval + a.scale * bst;
          tc_fence_before();
          if (tid_e == 0) te_trace(a, s, p, 2);
        } else if (mine) {
          const bool hid = kind != TE_VIS;
          const int t = hid ? p / 2 : (p - 1) / 2;
          const int rb = hid ? pl.rb_h : pl.rb_v, width = hid ? pl.hs : pl.vs;
          const int N = hid ? a.H : a.V;
          const int col0 = hid ? h_cb * pl.hs : v_cb * pl.vs;
          const int grow = (hid ? h_rb * pl.rb_h : v_rb * pl.rb_v) + rloc;      // row inside the mini-batch
          const bool lane_ok = rloc < rb && grow < a.B;
          const bool row_ok = lane_ok && grow < valid;
          const int nacc = te_nacc((((hid ? a.V : min(wc_rows, a.H - v_kc * wc_rows)) + 15) / 16 + 3) >> 2);
          const int chunks = width >> 3;             // <= 8: at most 2 chunks per warp
          float bias[2][4], aux[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int n0 = col0 + (g + 4 * i) * 8 + coff;
            if (g + 4 * i < chunks && n0 < N) te_act_prefetch(a, hid, kind, t, epoch, s, grow, row_ok, n0, bias[i], aux[i]);
          }
          if (tid_e == 0) te_trace(a, s, p, 7);
          te_mbar_wait(a.ctl, bar(TE_B_ACC), n_acc & 1u);
          ++n_acc;
          tc_fence_after();
          if (tid_e == 0) te_trace(a, s, p, 1);
          if (!hid && pl.ksplit > 1) {
            // split K: partial tile -> L2 scratch; once all ksplit partials of the tile are there, EVERY one of the
            // ksplit CTAs finishes its own rb_v / ksplit rows (sum in a fixed order: deterministic).  A half-warp
            // takes 16 rows x 4 columns, so the column statistics keep their 16-lane butterfly.
            const int tile = v_rb * pl.nt_v_cols + v_cb;
            float* part = a.scratch + ((size_t)(tile * pl.ksplit + v_kc) * 64 + rloc) * 64;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int ch = g + 4 * i;
              if (ch < chunks) {
                float v[4];
                te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, width, lane, v);
                __stcg(reinterpret_cast<float4*>(part + ch * 8 + coff), make_float4(v[0], v[1], v[2], v[3]));
              }
            }
            const int R = pl.rb_v / pl.ksplit;                         // rows this CTA finishes (a multiple of 16)
            const int quads = width >> 2;
            const int item = (warp - TE_WARP_EPI) * 2 + (lane >> 4);   // half-warp index: at most one item each
            const bool has_item = item < (R >> 4) * quads;
            const int rt = v_kc * R + (item / quads) * 16 + (lane & 15);
            const int fn0 = col0 + (item % quads) * 4;
            const int fgrow = v_rb * pl.rb_v + rt;
            const bool f_in = has_item && fgrow < a.B && fn0 < N, f_ok = f_in && fgrow < valid;
            float fb[4] = {0.f, 0.f, 0.f, 0.f}, fa[4] = {0.f, 0.f, 0.f, 0.f};
            if (has_item && fn0 < N) te_act_prefetch(a, hid, kind, t, epoch, s, fgrow, f_ok, fn0, fb, fa);
            ++n_vis;
            asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
            if (tid_e == 0) {
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&a.ctl->tile_done[tile]) : "memory");
              unsigned v = 0;
              for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&a.ctl->tile_done[tile]) : "memory");
                if (v >= (unsigned)pl.ksplit * n_vis) break;
                if (spin > (1u << 22)) te_fail(a.ctl, 4);
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
            if (__any_sync(0xffffffffu, has_item)) {                    // whole warp: the statistics butterfly shuffles with a full mask
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
              if (has_item) {
                const float* base = a.scratch + ((size_t)(tile * pl.ksplit) * 64 + rt) * 64 + (fn0 - col0);
                for (int kc = 0; kc < pl.ksplit; ++kc) {                // fixed order: deterministic sums
                  const float4 x = __ldcg(reinterpret_cast<const float4*>(base + (size_t)kc * 4096));
                  acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
                }
              }
              te_act_finish(a, hid, kind, t, fgrow, f_in, f_ok, fn0, acc, fb, fa, lane);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int ch = g + 4 * i;
              const int n0 = col0 + ch * 8 + coff;
              if (ch < chunks && col0 + ch * 8 < N) {       // warp-uniform
                float acc[4];
                te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, width, lane, acc);
                if (tid_e == 0 && i == 0) te_trace(a, s, p, 8);
                te_act_finish(a, hid, kind, t, grow, lane_ok && n0 < N, row_ok && n0 < N, n0, acc, bias[i], aux[i], lane);
                if (tid_e == 0 && i == 0) te_trace(a, s, p, 9);
              }
            }
          }
          tc_fence_before();
          if (tid_e == 0) te_trace(a, s, p, 2);
        }
        if (s + 1 == steps && p == np - 1) break;    // nothing follows the last update
        // ---- end of the phase: this CTA's outputs are complete -> grid barrier ----
        asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
        if (tid_e == 0) {
          te_trace(a, s, p, 3);
          te_grid_arrive(a.ctl, a.exp);
          te_trace(a, s, p, 4);
          ++n_bar;
          te_grid_wait(a.ctl, n_bar * gridDim.x);
          te_trace(a, s, p, 5);
          // the barrier is open: request what the next phase waits for, right here (its buffer is free -- this CTA's
          // MMAs of the phase that used it completed before its epilogue ran)
          if (p < np - 1) {
            const int next = te_phase_kind(p + 1, np);
            if (next == TE_VIS && has_v) {
              mbar_expect_tx(bar(TE_B_ACT), actv_bytes);
              tma_load_3d(sbase + pl.off_act, &a.tm_hs3, bar(TE_B_ACT), 0, v_rb * pl.rb_v, v_kc * pl.kc_slabs);
            }
            if ((next == TE_HID || next == TE_HIDK) && has_h) {
              mbar_expect_tx(bar(TE_B_ACT), acth_bytes);
              tma_load_3d(sbase + pl.off_act, &a.tm_vt3, bar(TE_B_ACT), 0, h_rb * pl.rb_h, 0);
            }
            if (next == TE_UPD && has_u) {
              mbar_expect_tx(bar(TE_B_P1), (uint32_t)((1 + (pl.v1c_early ? 0 : pl.n4_chunks)) * TE_BOX_BYTES));
              tma_load_2d(sbase + pl.off_p1, &a.tm_pk2, bar(TE_B_P1), u_mb * 64, 0);
              if (!pl.v1c_early)
                for (int ch = 0; ch < pl.n4_chunks; ++ch)
                  tma_load_2d(sbase + pl.off_v1c + (uint32_t)(ch * TE_BOX_BYTES), &a.tm_vt2, bar(TE_B_P1), u_nb * pl.n4 + ch * 64, 0);
            }
          } else if (has_h) {                        // end of the step: the updated W row slice of the next one
            mbar_expect_tx(bar(TE_B_WR), wr_bytes);
            tma_load_3d(sbase + pl.off_wr, &a.tm_wb3, bar(TE_B_WR), 0, h_cb * pl.hs, 0);
          }
          te_trace(a, s, p + 1 < np ? p + 1 : 0, 11);   // (slot 11 of the end-of-step barrier lands in the same step's phase 0: ignored)
          te_publish(gb_passed, n_bar);
        }
        // every epilogue thread passes the barrier: what the next phase prefetches (biases, statistics) is other CTAs' output
        asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
// 2-D tensor map over a 16-bit matrix stored [rows, ld]: dim0 = columns (extent `cols`: anything beyond is
// zero-filled), dim1 = rows; box = 64 columns (one 128-byte swizzle row) x box_rows.  MN-major operands: k = row.
inline int te_make_map2(CUtensorMap* map, const dbn_mat& X, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(DBN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available%s%s");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)X.ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = X.dtype == DBN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(map, dt, 2, X.ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return fail(DBN_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed with CUresult %s%s", buf);
  }
  return DBN_OK;
}

inline int& te_enabled_flag() {   // DBN_B200_TC_EPOCH=0 keeps the tiled path; dbn_set_tc_epoch() overrides (tests, A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_TC_EPOCH");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v;
}
inline bool te_enabled() { return te_enabled_flag() == 1; }

// Tile plan of a layer (V visible, H hidden, mini-batches of B rows).  Every phase gets at most TE_PLAN_CTAS
// tiles; among the shapes that fit, the one with the fewest bytes fetched on the critical path of a step wins
// (activations of the previous phase + the weight row slice).  false: the layer does not fit -- tiled path.
inline bool te_plan(int V, int H, int B, TePlan* out) {
  if (V < 64 || H < 64 || B < 16 || B > 256) return false;
  const int G = TE_PLAN_CTAS;
  const int kv = ceil_div(V, 64), kh = ceil_div(H, 64);
  long best = -1;
  TePlan bp;
  memset(&bp, 0, sizeof(bp));
  for (int rb_h : {32, 64}) {
    const int ntr_h = ceil_div(B, rb_h);
    if (rb_h == 64 && B <= 32) continue;
    int hs = 0;
    for (int c = 8; c <= 64; c += 8)
      if (ntr_h * ceil_div(H, c) <= G) { hs = c; break; }
    if (!hs) continue;
    for (int rb_v : {32, 64}) {
      if (rb_v == 64 && B <= 32) continue;
      const int ntr_v = ceil_div(B, rb_v);
      for (int S : {1, 2, 4}) {
        const int kc = ceil_div(kh, S);
        if (ceil_div(kh, kc) != S) continue;           // some CTA would have nothing to contract
        if (S > 1 && (rb_v % (16 * S)) != 0) continue; // every CTA of a split finishes whole 16-row groups of the tile
        int vs = 0;
        for (int c = 8; c <= 64; c += 8)
          if (ntr_v * ceil_div(V, c) * S <= G) { vs = c; break; }
        if (!vs) continue;
        if (S > 1 && ntr_v * ceil_div(V, vs) > TE_MAX_TILE_CTRS) continue;
        int n4 = 0;
        for (int c = 8; c <= 128; c += 8)          // <= 16 chunks: the update epilogue keeps 4 master segments per thread in registers
          if (ceil_div(H, 64) * ceil_div(V, c) <= G) { n4 = c; break; }
        if (!n4) continue;
        TePlan p;
        memset(&p, 0, sizeof(p));
        p.rb_h = rb_h; p.hs = hs; p.nt_h_rows = ntr_h; p.nt_h_cols = ceil_div(H, hs); p.kv_slabs = kv;
        p.rb_v = rb_v; p.vs = vs; p.nt_v_rows = ntr_v; p.nt_v_cols = ceil_div(V, vs); p.ksplit = S; p.kc_slabs = kc;
        p.n4 = n4; p.n4_chunks = ceil_div(n4, 64); p.nt_u_rows = ceil_div(H, 64); p.nt_u_cols = ceil_div(V, n4);
        p.grid = std::max(p.nt_h_rows * p.nt_h_cols, std::max(p.nt_v_rows * p.nt_v_cols * S, p.nt_u_rows * p.nt_u_cols));
        // shared-memory regions (1 KB aligned: 128-byte swizzle atoms)
        const uint32_t vc = (uint32_t)(p.n4_chunks * TE_BOX_BYTES);      // V0 (or Vk) column tile of the update
        uint32_t sz_wr = (uint32_t)align_up((size_t)hs * 128 * kv, 1024);
        uint32_t sz_act = (uint32_t)align_up((size_t)std::max(rb_h * 128 * kv, rb_v * 128 * kc), 1024);
        uint32_t sz_wc = (uint32_t)align_up((size_t)ceil_div(kc * 64, 256) * std::min(kc * 64, 256) * 128, 1024);
        sz_wc = std::max(sz_wc, vc);
        // Vk column tile: next to V0's in the dead W column slice (fetched a phase early), else in the dead W row slice
        p.v1c_early = sz_wc >= 2 * vc;
        if (!p.v1c_early) sz_wr = std::max(sz_wr, vc);
        // P1 tile (one box): in the W row slice if there is room, else in the activation buffer
        const uint32_t wr_used = p.v1c_early ? 0u : vc;
        if (sz_wr >= wr_used + (uint32_t)TE_BOX_BYTES) p.p1_in_act = 0;
        else { p.p1_in_act = 1; sz_act = std::max(sz_act, (uint32_t)TE_BOX_BYTES); }
        p.off_wr = 0;
        p.off_act = p.off_wr + sz_wr;
        p.off_wc = p.off_act + sz_act;
        p.off_g4p = p.off_wc + sz_wc;
        p.off_bars = p.off_g4p + (uint32_t)TE_BOX_BYTES;
        p.off_v0c = p.off_wc;
        p.off_v1c = p.v1c_early ? p.off_wc + vc : p.off_wr;
        p.off_p1 = p.p1_in_act ? p.off_act : p.off_wr + wr_used;
        p.smem_bytes = p.off_bars + 256 + 1024 /*alignment slack*/;
        if (p.smem_bytes > (uint32_t)TE_SMEM_MAX) continue;
        p.scratch_bytes = S > 1 ? (size_t)p.nt_v_rows * p.nt_v_cols * S * 64 * 64 * sizeof(float) : 0;
        // bytes a CTA fetches behind a barrier per step: H sample rows, V1 rows, P1 (+ Vk columns), W row slice
        const long cost = (long)rb_v * 128 * kc + (long)rb_h * 128 * kv + TE_BOX_BYTES + (p.v1c_early ? 0 : (long)vc) + (long)hs * 128 * kv +
                          (S > 1 ? 16384 : 0);
        if (best < 0 || cost < best) { best = cost; bp = p; }
      }
    }
  }
  if (best < 0) return false;
  if (out) *out = bp;
  return true;
}

inline int te_max_coresident() {
  static int cache[DBN_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (cache[dev] == 0) {
    int per_sm = 0, sms = 0;
    if (cudaFuncSetAttribute(rbm_tc_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_MAX) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rbm_tc_epoch_kernel, TE_THREADS, TE_SMEM_MAX) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      per_sm = 0;
    }
    cache[dev] = per_sm * sms + 1;
  }
  return cache[dev] - 1;
}

}  // namespace dbn
