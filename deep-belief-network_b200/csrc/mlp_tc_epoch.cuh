// mlp_tc_epoch.cuh -- one PERSISTENT tcgen05 kernel per fine-tune iteration (all mini-batches of one pass over the
// data): the supervised counterpart of rbm_tc_epoch.cuh.
//
// The tiled path runs a 256-row back-propagation step (dbn/models.py:439-465 -> :394-417, :471-515) as 10 launches,
// 81 us for 1.1 GFLOP.  Here ONE cooperative launch walks a PHASE PROGRAM per mini-batch and meets at a grid barrier
// between phases.  For L hidden layers the program is
//     F_0 | F_1 | .. | F_{L-1} | HEAD | B_{L-2} + U_head | B_{L-3} + U_{L-1} | .. | U_1 + U_0
//   F_l   forward contraction of layer l, epilogue +c_l, act -> a_l                                  (:406-411)
//   HEAD  narrow output head in fp32 SIMT (W_o staged in shared memory): scores, softmax / identity, loss, delta_out
//         and the delta of the last hidden layer                                                     (:594-626, :696-720, :495-499)
//   B_l   delta_l = (delta_{l+1} W_{l+1}) * f'(a_l); K split over CTAs for wide layers               (:495-501)
//   U_l   W_l = decay W_l - lr/bs delta_l^T a_{l-1}, c_l -= lr/bs sum_rows delta_l (the ones column of a_{l-1}),
//         16-bit shadow refreshed; it runs ONE phase after the delta contraction that still reads the old W_l
//         (U_head next to B_{L-2}, U_{l+1} next to B_{l-1})                                          (:508-513, :454-465)
// A phase holds up to two groups of tiles; a CTA takes tile blockIdx of each group, fetches both tiles' operands the
// moment the barrier opens and works through them in order (they share TMEM).  Every contraction tile is UMMA M = 64
// with three issuer warps and the 32-lane epilogue of the epoch kernel.  No dropout, no optional outputs (those
// run the tiled path).  Results differ from the tiled path by the summation order inside a dot product only.
#pragma once
#include "rbm_tc_epoch.cuh"

namespace dbn {

constexpr int MT_MAX_LAYERS = 4;                       // hidden layers this kernel takes
constexpr int MT_MAX_PHASES = 3 * MT_MAX_LAYERS + 1;
constexpr int MT_MAX_OUT = 16;                         // head width
enum { MT_NONE = 0, MT_FWD = 1, MT_BWD = 2, MT_UPD = 3, MT_HEAD = 4 };
enum { MT_B_LD0 = 0, MT_B_LD1, MT_B_ACC, MT_B_TFREE, MT_NBARS };

struct MtGroup {
  int kind, layer;          // FWD l: a_l;  BWD l: delta_l (from delta_{l+1}, W_{l+1});  UPD l: W_l (l == L: the head)
  int rb, width;            // FWD / BWD: tile rows x units;  UPD: 64 units x width inputs
  int nt_rows, nt_cols;
  int ksplit, kc_slabs;     // K chunking: FWD k-slabs of K;  BWD slabs per split
  int n_chunks;             // UPD: 64-wide column chunks of the B operand
  int n_tiles;
  uint32_t off_a, off_b, a_bytes, b_bytes;
  int ctr0;                 // BWD with ksplit > 1: first tile counter
  int pad;
};
struct MtPhase {
  MtGroup g[2];
};

struct MtArgs {
  CUtensorMap tm_in3[MT_MAX_LAYERS];      // input of layer l (X, a_0, ..) K-major {64, rows, slabs}: FWD A
  CUtensorMap tm_w3[MT_MAX_LAYERS];       // W_l shadow K-major: FWD B
  CUtensorMap tm_d3[MT_MAX_LAYERS];       // delta_l K-major: BWD A
  CUtensorMap tm_w2[MT_MAX_LAYERS];       // W_l shadow {columns, rows}: BWD B
  CUtensorMap tm_d2[MT_MAX_LAYERS + 1];   // delta_l {columns, rows} (l == L: 16-bit head delta): UPD A
  CUtensorMap tm_in2[MT_MAX_LAYERS + 1];  // input of layer l with its ones column {columns, rows}: UPD B
  MtPhase ph[MT_MAX_PHASES];
  int n_phases;
  int L, n_in, n_out, B, act, head, dtype;
  int dims[MT_MAX_LAYERS];
  int n_rows, row0;
  float scale, decay;                     // scale = -lr / batch_size
  dbn_mat X, dOb;
  dbn_mat a[MT_MAX_LAYERS], d[MT_MAX_LAYERS], Wb[MT_MAX_LAYERS];
  float* W[MT_MAX_LAYERS];
  int ldw[MT_MAX_LAYERS];
  float* c[MT_MAX_LAYERS];
  float* Wo;
  int ldwo;
  float* bo;
  const float* Y;
  int ldy;
  float* loss_rows;
  uint32_t off_bars, smem_bytes, wo_bytes;
  TeCtl* ctl;
  float* scratch;
  long long* trace;
  // dropout (dbn/models.py:401-411), drop != 0: the input of step s is masked into a0[s & 1] one phase ahead (every CTA a
  // share of the rows), the hidden activations in the forward epilogue; masks[j].ptr != null: injected masks (rows of
  // the iteration, != 0 keeps), else Philox keyed like the tiled path (stream + j, global row, column)
  int drop;
  float keep_p;
  dbn_rng rng;
  dbn_mat masks[MT_MAX_LAYERS + 1];
  dbn_mat a0[2];
  CUtensorMap tm_a0_3[2], tm_a0_2[2];     // a0[i] as FWD A / UPD B of layer 0
};

// debug trace (dbn_debug_set_trace): [cta][step < 8][phase < 13][8] clock64 stamps: 0 operands of the first tile landed (issuer 0),
// 1 its accumulator ready (epilogue thread 0), 2 last tile's epilogue done, 3 CTA done, 4 arrived, 5 barrier passed,
// 6 second tile's accumulator ready
constexpr int MT_TRACE_STEPS = 8, MT_TRACE_SLOTS = MT_MAX_PHASES * 8;
__device__ __forceinline__ void mt_trace(const MtArgs& a, int step, int p, int slot) {
  if (a.trace != nullptr && step < MT_TRACE_STEPS)
    a.trace[((size_t)blockIdx.x * MT_TRACE_STEPS + step) * MT_TRACE_SLOTS + p * 8 + slot] = clock64();
}

// activation derivative evaluated on the stored activation (dbn/activations.py:38, :58)
__device__ __forceinline__ float mt_prime(int act, float g) { return act == DBN_ACT_SIGMOID ? g * (1.0f - g) : (g > 0.0f ? 1.0f : 0.0f); }

// 4 consecutive elements of a 16-bit matrix that other CTAs rewrite every mini-batch: straight from L2
__device__ __forceinline__ void mt_load4_cg_h(const dbn_mat& m, int r, int n0, int N, float v[4]) {
  const uint16_t* p = reinterpret_cast<const uint16_t*>(m.ptr) + (size_t)r * (size_t)m.ld + (size_t)n0;
  if (n0 + 4 <= N) {
    const uint2 pk = __ldcg(reinterpret_cast<const uint2*>(p));
    unpack16x2(m.dtype, pk.x, v[0], v[1]); unpack16x2(m.dtype, pk.y, v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo = 0.0f, hi = 0.0f;
      if (n0 + j < N) unpack16x2(m.dtype, (uint32_t)__ldcg(p + j), lo, hi);
      v[j] = lo;
    }
  }
}

__device__ __forceinline__ void mt_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// the operands of this CTA's tiles of phase p (called by the barrier thread the moment the barrier opens)
__device__ __forceinline__ void mt_issue_loads(const MtArgs& a, int s, int p, uint32_t sbase, uint32_t bar_base) {
  const int bid = (int)blockIdx.x;
#pragma unroll 1
  for (int gi = 0; gi < 2; ++gi) {
    const MtGroup& g = a.ph[p].g[gi];
    const uint32_t bar = bar_base + 8u * (uint32_t)(MT_B_LD0 + gi);
    if (g.kind == MT_HEAD) {
      if (bid * 2 < a.B) {
        mbar_expect_tx(bar, a.wo_bytes);
        mt_bulk_load(sbase, a.Wo, a.wo_bytes, bar);
      }
      continue;
    }
    if (g.kind == MT_NONE || bid >= g.n_tiles) continue;
    const int l = g.layer;
    mbar_expect_tx(bar, g.a_bytes + g.b_bytes);
    if (g.kind == MT_FWD) {
      const int rbk = bid / g.nt_cols, cb = bid % g.nt_cols;
      if (l == 0 && a.drop) tma_load_3d(sbase + g.off_a, &a.tm_a0_3[s & 1], bar, 0, rbk * g.rb, 0);
      else tma_load_3d(sbase + g.off_a, &a.tm_in3[l], bar, 0, (l == 0 ? a.row0 + s * a.B : 0) + rbk * g.rb, 0);
      tma_load_3d(sbase + g.off_b, &a.tm_w3[l], bar, 0, cb * g.width, 0);
    } else if (g.kind == MT_BWD) {
      const int kc = bid % g.ksplit, cb = (bid / g.ksplit) % g.nt_cols, rbk = bid / (g.ksplit * g.nt_cols);
      const int krows = g.kc_slabs * 64;
      tma_load_3d(sbase + g.off_a, &a.tm_d3[l + 1], bar, 0, rbk * g.rb, kc * g.kc_slabs);
      for (int i = 0; i * 256 < krows; ++i)
        tma_load_2d(sbase + g.off_b + (uint32_t)(i * TE_BOX_BYTES), &a.tm_w2[l + 1], bar, cb * g.width, kc * krows + i * 256);
    } else {
      const int mb = bid / g.nt_cols, nb = bid % g.nt_cols;
      tma_load_2d(sbase + g.off_a, &a.tm_d2[l], bar, mb * 64, 0);
      const bool masked = l == 0 && a.drop;
      for (int ch = 0; ch < g.n_chunks; ++ch)
        tma_load_2d(sbase + g.off_b + (uint32_t)(ch * TE_BOX_BYTES), masked ? &a.tm_a0_2[s & 1] : &a.tm_in2[l], bar,
                    nb * g.width + ch * 64, (l == 0 && !masked) ? a.row0 + s * a.B : 0);
    }
  }
}

// input_data *= Binomial(1, p) (dbn/models.py:401-403) for mini-batch s, into a0[s & 1]: every CTA's epilogue threads take
// a share of the quads; rows past the end of the dataset are zero (they are operands of the dW contraction)
__device__ __forceinline__ void mt_mask_input(const MtArgs& a, int s, int tid_e, uint32_t epoch) {
  const int valid = min(a.B, a.n_rows - s * a.B);
  const int quads = (a.n_in + 3) >> 2;
  dbn_rng r = a.rng;
  r.row0 += (uint32_t)(s * a.B);
  const dbn_mat& dst = a.a0[s & 1];
  const bool inj = a.masks[0].ptr != nullptr;
  for (int q = (int)blockIdx.x * TE_EPI_THREADS + tid_e; q < a.B * quads; q += (int)gridDim.x * TE_EPI_THREADS) {
    const int row = q / quads, c = (q - row * quads) << 2, nv = min(4, a.n_in - c);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < valid) {
      float u[4] = {0.f, 0.f, 0.f, 0.f};
      mat_load4(a.X, a.row0 + s * a.B + row, c, v, nv);
      if (inj) mat_load4(a.masks[0], s * a.B + row, c, u, nv);
      else rng_uniform4(r, epoch, row, c, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (inj ? (u[j] != 0.0f) : (u[j] < a.keep_p)) ? v[j] : 0.0f;
    }
    mat_store4(dst, row, c, v, nv);
  }
}

// ---- HEAD: two rows per CTA, eight warps per row (each an eighth of K), W_o in shared memory -------------------
__device__ __forceinline__ void mt_head_phase(const MtArgs& a, int s, int valid, const float* wo_s, float (*partial)[8][32],
                                              float (*dout)[32], int warp_e, int lane) {
  const int C = a.n_out, K = a.dims[a.L - 1];
  const dbn_mat& A = a.a[a.L - 1];
  const int r_in = warp_e >> 3, part = warp_e & 7;
  const int row = (int)blockIdx.x * 2 + r_in;
  const bool active = row < a.B, row_ok = row < valid;
  const uint16_t* arow = reinterpret_cast<const uint16_t*>(A.ptr) + (size_t)row * (size_t)A.ld;
  float acc[MT_MAX_OUT];
#pragma unroll
  for (int c = 0; c < MT_MAX_OUT; ++c) acc[c] = 0.0f;
  // the target and the head bias do not depend on the scores: requested first, their DRAM / L2 latency hides behind them
  float y_pre = 0.0f, bo_pre = 0.0f;
  if (part == 0 && row_ok && lane < C) {
    y_pre = a.Y[(size_t)(a.row0 + s * a.B + row) * a.ldy + lane];
    bo_pre = __ldcg(a.bo + lane);
  }
  if (row_ok) {
    for (int k = (part * 32 + lane) * 8; k < K; k += 2048) {   // (K % 4 == 0: plan; a last chunk of 4 when K % 8 == 4)
      const bool full = k + 8 <= K;
      const uint4 t = __ldcg(reinterpret_cast<const uint4*>(arow + k));   // (the row pitch covers the whole chunk)
      float x[8];
      unpack16x2(A.dtype, t.x, x[0], x[1]); unpack16x2(A.dtype, t.y, x[2], x[3]);
      unpack16x2(A.dtype, t.z, x[4], x[5]); unpack16x2(A.dtype, t.w, x[6], x[7]);
      if (!full) x[4] = x[5] = x[6] = x[7] = 0.0f;             // the ones column and the padding behind the last unit
#pragma unroll
      for (int c = 0; c < MT_MAX_OUT; ++c) {
        if (c < C) {
          const float4 w0 = *reinterpret_cast<const float4*>(wo_s + (size_t)c * a.ldwo + k);
          const float4 w1 = full ? *reinterpret_cast<const float4*>(wo_s + (size_t)c * a.ldwo + k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          float sum = acc[c];
          sum = fmaf(x[0], w0.x, sum); sum = fmaf(x[1], w0.y, sum); sum = fmaf(x[2], w0.z, sum); sum = fmaf(x[3], w0.w, sum);
          sum = fmaf(x[4], w1.x, sum); sum = fmaf(x[5], w1.y, sum); sum = fmaf(x[6], w1.z, sum); sum = fmaf(x[7], w1.w, sum);
          acc[c] = sum;
        }
      }
    }
  }
  float z = 0.0f;   // lane c ends up holding this part's share of score c
#pragma unroll
  for (int c = 0; c < MT_MAX_OUT; ++c) {
    if (c < C) {
      float sum = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == c) z = sum;
    }
  }
  partial[r_in][part][lane] = z;
  asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
  if (part == 0) {
    const bool lv = row_ok && lane < C;
    z = 0.0f;
    if (lv) {
#pragma unroll
      for (int q = 0; q < 8; ++q) z += partial[r_in][q][lane];
      z += bo_pre;
    }
    const float y = lv ? y_pre : 0.0f;
    float o_val = z, loss = 0.0f;
    if (a.head == DBN_HEAD_SOFTMAX) {
      float mx = lv ? z : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = lv ? expf(z - mx) : 0.0f;
      float sum = e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      o_val = row_ok ? e / sum : 0.0f;
      float py = o_val * y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) py += __shfl_xor_sync(0xffffffffu, py, o);
      loss = row_ok ? -logf(py) : 0.0f;
    } else {
      const float dlt = lv ? (z - y) : 0.0f;
      loss = dlt * dlt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    }
    const float dlt = lv ? (o_val - y) : 0.0f;
    if (active && lane < C) mat_store(a.dOb, row, lane, dlt);   // 16-bit operand of the head's weight-gradient contraction (zero rows beyond the batch)
    dout[r_in][lane] = dlt;
    if (row_ok && a.loss_rows != nullptr && lane == 0) a.loss_rows[s * a.B + row] = loss;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
  if (!active) return;
  // delta of the last hidden layer: (sum_c delta_out[c] W_o[c, h]) f'(a[h]), this warp's eighth of the units
  float d[MT_MAX_OUT];
#pragma unroll
  for (int c = 0; c < MT_MAX_OUT; ++c) d[c] = (c < C) ? dout[r_in][c] : 0.0f;
  const dbn_mat& D = a.d[a.L - 1];
  for (int k = (part * 32 + lane) * 8; k < K; k += 2048) {
    const bool full = k + 8 <= K;
    float x[8], sm[8];
    if (row_ok) {
      const uint4 t = __ldcg(reinterpret_cast<const uint4*>(arow + k));
      unpack16x2(A.dtype, t.x, x[0], x[1]); unpack16x2(A.dtype, t.y, x[2], x[3]);
      unpack16x2(A.dtype, t.z, x[4], x[5]); unpack16x2(A.dtype, t.w, x[6], x[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j] = 0.0f;
#pragma unroll
    for (int c = 0; c < MT_MAX_OUT; ++c) {
      if (c < C) {
        const float4 w0 = *reinterpret_cast<const float4*>(wo_s + (size_t)c * a.ldwo + k);
        const float4 w1 = full ? *reinterpret_cast<const float4*>(wo_s + (size_t)c * a.ldwo + k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        sm[0] = fmaf(d[c], w0.x, sm[0]); sm[1] = fmaf(d[c], w0.y, sm[1]); sm[2] = fmaf(d[c], w0.z, sm[2]); sm[3] = fmaf(d[c], w0.w, sm[3]);
        sm[4] = fmaf(d[c], w1.x, sm[4]); sm[5] = fmaf(d[c], w1.y, sm[5]); sm[6] = fmaf(d[c], w1.z, sm[6]); sm[7] = fmaf(d[c], w1.w, sm[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[j] = row_ok ? sm[j] * mt_prime(a.act, x[j]) : 0.0f;
    uint4 pk;
    pk.x = pack16x2(D.dtype, sm[0], sm[1]); pk.y = pack16x2(D.dtype, sm[2], sm[3]);
    pk.z = pack16x2(D.dtype, sm[4], sm[5]); pk.w = pack16x2(D.dtype, sm[6], sm[7]);
    uint16_t* dst = reinterpret_cast<uint16_t*>(D.ptr) + (size_t)row * (size_t)D.ld + (size_t)k;
    if (full) *reinterpret_cast<uint4*>(dst) = pk;
    else *reinterpret_cast<uint2*>(dst) = make_uint2(pk.x, pk.y);   // the padding behind the last unit stays zero (K tail of the next contraction)
  }
}

__global__ void __launch_bounds__(TE_THREADS, 1) mlp_tc_epoch_kernel(const __grid_constant__ MtArgs a) {
  extern __shared__ uint8_t te_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(te_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MT_NBARS);
  float (*partial)[8][32] = reinterpret_cast<float (*)[8][32]>(tmem_slot + 4);   // HEAD: [2][8][32] partial scores
  float (*dout)[32] = reinterpret_cast<float (*)[32]>(tmem_slot + 4 + 2 * 8 * 32);   // HEAD: [2][32] delta_out
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int bid = (int)blockIdx.x;
  const int steps = (a.n_rows + a.B - 1) / a.B;
  const int fmt = a.dtype == DBN_F16 ? 0 : 1;
  const int np = a.n_phases;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < a.L; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_in3[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_w3[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_d3[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_w2[l]) : "memory");
    }
    for (int l = 0; l <= a.L; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_d2[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_in2[l]) : "memory");
    }
    if (a.drop) {
      for (int i = 0; i < 2; ++i) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_a0_3[i]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tm_a0_2[i]) : "memory");
      }
    }
    mbar_init(bar(MT_B_LD0), 1);
    mbar_init(bar(MT_B_LD1), 1);
    mbar_init(bar(MT_B_ACC), TE_NISS);
    mbar_init(bar(MT_B_TFREE), TE_EPI_WARPS);
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
    // (no producer warp: every operand is requested by the barrier thread; warp 0 only set the barriers up)
  } else if (warp < TE_WARP_EPI) {
    // ======================================= MMA issuers =======================================
    if (te_elect()) {
      const int w = warp - 1;
      unsigned n_ld[2] = {0u, 0u}, n_task = 0;
      for (int s = 0; s < steps; ++s) {
        for (int p = 0; p < np; ++p) {
#pragma unroll 1
          for (int gi = 0; gi < 2; ++gi) {
            const MtGroup& g = a.ph[p].g[gi];
            if (g.kind == MT_HEAD) {
              // SIMT phase, nothing to issue -- but the issuers must SEE this completion of the load barrier (W_o):
              // a parity wait tells adjacent phases apart only, so a thread that skipped it would take the barrier's
              // previous completion for the one it waits for next and start its MMAs on operands that have not landed
              if (bid * 2 < a.B) {
                te_mbar_wait(a.ctl, bar(MT_B_LD0 + gi), n_ld[gi] & 1u);
                ++n_ld[gi];
              }
              continue;
            }
            if (g.kind == MT_NONE || bid >= g.n_tiles) continue;
            te_mbar_wait(a.ctl, bar(MT_B_LD0 + gi), n_ld[gi] & 1u);
            ++n_ld[gi];
            if (w == 0 && gi == 0) mt_trace(a, s, p, 0);
            if (n_task > 0) te_mbar_wait(a.ctl, bar(MT_B_TFREE), (n_task - 1) & 1u);   // the previous tile's accumulators are drained
            tc_fence_after();
            const int l = g.layer;
            int nk;
            uint64_t dA, dB;
            uint32_t ua, ub, ia, ib, idesc;
            if (g.kind == MT_FWD) {           // both K-major, K = inputs of layer l
              nk = ((l == 0 ? a.n_in : a.dims[l - 1]) + 15) / 16;
              idesc = make_idesc(64, g.width, 0, 0, 0, fmt, fmt);
              dA = make_smem_desc(sbase + g.off_a, 16, 1024); dB = make_smem_desc(sbase + g.off_b, 16, 1024);
              ua = (uint32_t)(g.rb * 128) >> 4; ub = (uint32_t)(g.width * 128) >> 4; ia = 2; ib = 2;
            } else if (g.kind == MT_BWD) {    // A K-major (delta_{l+1} rows), B MN-major (W_{l+1}), K = units of layer l+1
              const int kc = bid % g.ksplit, krows = g.kc_slabs * 64;
              nk = (min(krows, a.dims[l + 1] - kc * krows) + 15) / 16;
              idesc = make_idesc(64, g.width, 0, 1, 0, fmt, fmt);
              dA = make_smem_desc(sbase + g.off_a, 16, 1024); dB = make_smem_desc(sbase + g.off_b, TE_BOX_BYTES, 1024);
              ua = (uint32_t)(g.rb * 128) >> 4; ub = 512; ia = 2; ib = 128;
            } else {                          // both MN-major, K = batch rows
              nk = (a.B + 15) / 16;
              idesc = make_idesc(64, g.width, 1, 1, 0, fmt, fmt);
              dA = make_smem_desc(sbase + g.off_a, TE_BOX_BYTES, 1024); dB = make_smem_desc(sbase + g.off_b, TE_BOX_BYTES, 1024);
              ua = 512; ub = 512; ia = 128; ib = 128;
            }
            const int nunits = (nk + 3) >> 2;
            const int per = te_units_per_issuer(nunits);
            const int u0 = w * per, u1 = min(nunits, u0 + per);
            const uint32_t tmem_d = tmem_base + (uint32_t)(w * g.width);
            uint32_t accumulate = 0;
#pragma unroll 1
            for (int u = u0; u < u1; ++u) {
              const int ks = min(4, nk - 4 * u);
              const uint64_t da = dA + (uint64_t)(ua * (uint32_t)u), db = dB + (uint64_t)(ub * (uint32_t)u);
              if (ks == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  tc_mma_bf16(tmem_d, da + (uint64_t)(ia * j), db + (uint64_t)(ib * j), idesc, accumulate);
                  accumulate = 1;
                }
              } else {
                for (int j = 0; j < ks; ++j) {
                  tc_mma_bf16(tmem_d, da + (uint64_t)(ia * j), db + (uint64_t)(ib * j), idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            tc_commit(bar(MT_B_ACC));
            ++n_task;
          }
        }
      }
    }
  } else {
    // ============================ epilogue warps (+ HEAD phase, grid barrier, operand requests) ============================
    const int q = warp & 3;
    const int wg = (warp - TE_WARP_EPI) >> 2;
    const int warp_e = warp - TE_WARP_EPI;
    const int tid_e = (int)threadIdx.x - 32 * TE_WARP_EPI;
    const int rloc = q * 16 + (lane & 15);
    const int coff = (lane >> 4) * 4;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned n_acc = 0, n_bar = 0, n_ld0 = 0;
    const uint32_t epoch = a.rng.epoch + (a.rng.epoch_ptr ? *a.rng.epoch_ptr : 0u);
    if (a.drop) {       // the first mini-batch's masked input must exist before anybody's first tile is requested
      mt_mask_input(a, 0, tid_e, epoch);
      asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
      if (tid_e == 0) {
        te_grid_arrive(a.ctl, 0);
        ++n_bar;
        te_grid_wait(a.ctl, n_bar * gridDim.x);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
    }
    if (tid_e == 0) mt_issue_loads(a, 0, 0, sbase, bar_base);
    for (int s = 0; s < steps; ++s) {
      const int valid = min(a.B, a.n_rows - s * a.B);
      for (int p = 0; p < np; ++p) {
#pragma unroll 1
        for (int gi = 0; gi < 2; ++gi) {
          const MtGroup& g = a.ph[p].g[gi];
          if (g.kind == MT_HEAD) {
            if (bid * 2 < a.B) {
              te_mbar_wait(a.ctl, bar(MT_B_LD0 + gi), n_ld0 & 1u);    // W_o has landed (HEAD is always group 0)
              ++n_ld0;
              mt_head_phase(a, s, valid, reinterpret_cast<const float*>(smem), partial, dout, warp_e, lane);
            }
            continue;
          }
          if (g.kind == MT_NONE || bid >= g.n_tiles) continue;
          if (gi == 0) ++n_ld0;                                        // (this group's load barrier completes once more)
          const int l = g.layer;
          const int chunks = g.width >> 3;
          if (g.kind == MT_UPD) {
            // ---- W_l = decay W_l + scale dW, bias from the ones column, 16-bit shadow (dbn/models.py:454-465) ----
            const int mb = bid / g.nt_cols, nb = bid % g.nt_cols;
            const bool is_head = l == a.L;
            const int M = is_head ? a.n_out : a.dims[l];
            const int Kin = l == 0 ? a.n_in : a.dims[l - 1];
            float* Wl = is_head ? a.Wo : a.W[l];
            const int ldw = is_head ? a.ldwo : a.ldw[l];
            float* bias = is_head ? a.bo : a.c[l];
            const int m = mb * 64 + rloc;
            const bool row_ok = m < M;
            float* wrow = Wl + (size_t)m * (size_t)ldw;
            const int nacc = te_nacc((((a.B + 15) / 16) + 3) >> 2);
            float wpre[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n0 = nb * g.width + (wg + 4 * i) * 8 + coff;
              if (wg + 4 * i < chunks && n0 < Kin && row_ok) te_load4_cg(wrow, n0, Kin, wpre[i]);
            }
            te_mbar_wait(a.ctl, bar(MT_B_ACC), n_acc & 1u);
            ++n_acc;
            tc_fence_after();
            if (tid_e == 0) mt_trace(a, s, p, gi == 0 ? 1 : 6);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int ch = wg + 4 * i;
              if (ch < chunks && nb * g.width + ch * 8 <= Kin) {       // warp-uniform (<=: the ones column sits at Kin)
                const int n0 = nb * g.width + ch * 8 + coff;
                float acc[4];
                te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, g.width, lane, acc);
                if (row_ok && n0 < Kin) {
                  float w4[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) w4[j] = fmaf(a.scale, acc[j], a.decay * wpre[i][j]);
                  te_store4_f32(wrow, n0, Kin, w4);
                  if (!is_head) te_store4_h(a.Wb[l], m, n0, Kin, w4);
                }
                if (row_ok && n0 <= Kin && Kin < n0 + 4)                // sum_rows delta[row, m] through the ones column: no decay on biases
                  bias[m] = __ldcg(bias + m) + a.scale * acc[Kin - n0];
              }
            }
          } else {
            const bool fwd = g.kind == MT_FWD;
            const int N = a.dims[l];
            int rbk, cb, kc = 0;
            if (fwd) { rbk = bid / g.nt_cols; cb = bid % g.nt_cols; }
            else { kc = bid % g.ksplit; cb = (bid / g.ksplit) % g.nt_cols; rbk = bid / (g.ksplit * g.nt_cols); }
            const int col0 = cb * g.width;
            const int grow = rbk * g.rb + rloc;
            const bool lane_ok = rloc < g.rb && grow < a.B;
            const bool row_ok = lane_ok && grow < valid;
            int nk;
            if (fwd) nk = ((l == 0 ? a.n_in : a.dims[l - 1]) + 15) / 16;
            else nk = (min(g.kc_slabs * 64, a.dims[l + 1] - kc * g.kc_slabs * 64) + 15) / 16;
            const int nacc = te_nacc((nk + 3) >> 2);
            // what does not depend on the accumulator: bias (forward), stored activation (backward)
            float pre[2][4];
            const bool split = !fwd && g.ksplit > 1;
            if (!split) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int n0 = col0 + (wg + 4 * i) * 8 + coff;
#pragma unroll
                for (int j = 0; j < 4; ++j) pre[i][j] = 0.0f;
                if (wg + 4 * i < chunks && n0 < N) {
                  if (fwd) te_load4_cg(a.c[l], n0, N, pre[i]);
                  else if (row_ok) mt_load4_cg_h(a.a[l], grow, n0, N, pre[i]);
                }
              }
            }
            te_mbar_wait(a.ctl, bar(MT_B_ACC), n_acc & 1u);
            ++n_acc;
            tc_fence_after();
            if (tid_e == 0) mt_trace(a, s, p, gi == 0 ? 1 : 6);
            if (split) {
              // partial tile -> L2 scratch; once all partials are there every CTA of the split finishes 16-row groups
              const int tile = rbk * g.nt_cols + cb;
              float* part = a.scratch + ((size_t)(tile * g.ksplit + kc) * 64 + rloc) * 64;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int ch = wg + 4 * i;
                if (ch < chunks) {
                  float v[4];
                  te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, g.width, lane, v);
                  __stcg(reinterpret_cast<float4*>(part + ch * 8 + coff), make_float4(v[0], v[1], v[2], v[3]));
                }
              }
              const int R = g.rb / g.ksplit;
              const int quads = g.width >> 2;
              const int item = warp_e * 2 + (lane >> 4);
              const bool has_item = item < (R >> 4) * quads;
              const int rt = kc * R + (item / quads) * 16 + (lane & 15);
              const int fn0 = col0 + (item % quads) * 4;
              const int fgrow = rbk * g.rb + rt;
              const bool f_in = has_item && fgrow < a.B && fn0 < N, f_ok = f_in && fgrow < valid;
              float gact[4] = {0.f, 0.f, 0.f, 0.f};
              if (f_ok) mt_load4_cg_h(a.a[l], fgrow, fn0, N, gact);
              tc_fence_before();
              asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
              if (tid_e == 0) {
                unsigned int* ctr = &a.ctl->tile_done[g.ctr0 + tile];
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
                unsigned v = 0;
                for (unsigned spin = 0;; ++spin) {
                  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                  if (v >= (unsigned)g.ksplit * (unsigned)(s + 1)) break;
                  if (spin > (1u << 22)) te_fail(a.ctl, 5);
                }
              }
              asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
              if (f_in) {
                const float* base = a.scratch + ((size_t)(tile * g.ksplit) * 64 + rt) * 64 + (fn0 - col0);
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int kq = 0; kq < g.ksplit; ++kq) {
                  const float4 x = __ldcg(reinterpret_cast<const float4*>(base + (size_t)kq * 4096));
                  acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
                }
                float x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = f_ok ? acc[j] * mt_prime(a.act, gact[j]) : 0.0f;
                te_store4_h(a.d[l], fgrow, fn0, N, x);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int ch = wg + 4 * i;
                const int n0 = col0 + ch * 8 + coff;
                if (ch < chunks && col0 + ch * 8 < N) {
                  float acc[4], x[4];
                  te_ld_acc4(tq + (uint32_t)(ch * 8), nacc, g.width, lane, acc);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    if (fwd) x[j] = row_ok ? apply_act<true>(a.act, acc[j] + pre[i][j]) : 0.0f;
                    else x[j] = row_ok ? acc[j] * mt_prime(a.act, pre[i][j]) : 0.0f;
                  }
                  if (fwd && a.drop && row_ok && n0 < N) {      // input_data *= Binomial(1, p): dbn/models.py:408-411
                    float u[4] = {1.f, 1.f, 1.f, 1.f};
                    const bool inj = a.masks[l + 1].ptr != nullptr;
                    if (inj) {
                      mat_load4(a.masks[l + 1], s * a.B + grow, n0, u, min(4, N - n0));
                    } else {
                      dbn_rng r = a.rng;
                      r.stream += (uint32_t)(l + 1);
                      r.row0 += (uint32_t)(s * a.B);
                      rng_uniform4(r, epoch, grow, n0, u);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = (inj ? (u[j] != 0.0f) : (u[j] < a.keep_p)) ? x[j] : 0.0f;
                  }
                  if (lane_ok && n0 < N) te_store4_h(fwd ? a.a[l] : a.d[l], grow, n0, N, x);
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(MT_B_TFREE));   // this warp has drained its share of the accumulators
        }
        if (tid_e == 0) mt_trace(a, s, p, 2);
        if (s + 1 == steps && p == np - 1) break;
        if (a.drop && p == np - 1) mt_mask_input(a, s + 1, tid_e, epoch);   // next mini-batch's input, into the other buffer
        asm volatile("bar.sync 1, %0;" ::"n"(TE_EPI_THREADS) : "memory");
        if (tid_e == 0) {
          mt_trace(a, s, p, 3);
          te_grid_arrive(a.ctl, 0);
          mt_trace(a, s, p, 4);
          ++n_bar;
          te_grid_wait(a.ctl, n_bar * gridDim.x);
          mt_trace(a, s, p, 5);
          if (p + 1 < np) mt_issue_loads(a, s, p + 1, sbase, bar_base);
          else mt_issue_loads(a, s + 1, 0, sbase, bar_base);
        }
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
inline int& mt_enabled_flag() {   // DBN_B200_TC_FINETUNE=0 keeps the tiled path; dbn_set_tc_finetune() overrides (tests, A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DBN_B200_TC_FINETUNE");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v;
}

struct MtPlan {
  MtPhase ph[MT_MAX_PHASES];
  int n_phases, grid;
  uint32_t off_bars, smem_bytes, wo_bytes;
  size_t scratch_bytes;
  // tile shapes the tensor maps need: forward (rows, units) per layer, backward (rows, slabs per split) per produced delta
  int f_rb[MT_MAX_LAYERS], f_w[MT_MAX_LAYERS], b_rb[MT_MAX_LAYERS], b_kc[MT_MAX_LAYERS];
};

// Phase program and tile plan of a network (n_in -> dims[0..L) -> n_out) for mini-batches of B rows; false: does not fit.
inline bool mt_plan(int n_in, const int32_t* dims, int L, int n_out, int B, MtPlan* out) {
  if (L < 1 || L > MT_MAX_LAYERS || n_out < 1 || n_out > MT_MAX_OUT || B < 16 || B > 256 || n_in < 32) return false;
  for (int l = 0; l < L; ++l)
    if (dims[l] < 32) return false;
  const int HL = dims[L - 1];
  if ((HL % 4) != 0) return false;                       // the head reads W_o rows as float4 (8 activations per lane, a last chunk of 4)
  const int G = TE_PLAN_CTAS;
  if ((B + 1) / 2 > G) return false;
  MtPlan P;
  memset(&P, 0, sizeof(P));
  auto in_width = [&](int l) { return l == 0 ? n_in : dims[l - 1]; };
  MtGroup fwd[MT_MAX_LAYERS], bwd[MT_MAX_LAYERS], upd[MT_MAX_LAYERS + 1];
  memset(fwd, 0, sizeof(fwd)); memset(bwd, 0, sizeof(bwd)); memset(upd, 0, sizeof(upd));
  const uint32_t slot_max = 128 * 1024;
  for (int l = 0; l < L; ++l) {                          // forward tiles
    const int ks = ceil_div(in_width(l), 64), N = dims[l];
    long best = -1;
    for (int rb : {32, 64}) {
      if (rb == 64 && B <= 32) continue;
      const int ntr = ceil_div(B, rb);
      int w = 0;
      for (int c = 8; c <= 64; c += 8)
        if (ntr * ceil_div(N, c) <= G) { w = c; break; }
      if (!w) continue;
      const long ab = (long)rb * 128 * ks, bb = (long)w * 128 * ks;
      if (ab + bb > slot_max) continue;
      if (best < 0 || ab + bb < best) {
        best = ab + bb;
        MtGroup& g = fwd[l];
        g.kind = MT_FWD; g.layer = l; g.rb = rb; g.width = w; g.nt_rows = ntr; g.nt_cols = ceil_div(N, w);
        g.ksplit = 1; g.kc_slabs = ks; g.n_tiles = g.nt_rows * g.nt_cols; g.a_bytes = (uint32_t)ab; g.b_bytes = (uint32_t)bb;
      }
    }
    if (best < 0) return false;
    P.f_rb[l] = fwd[l].rb; P.f_w[l] = fwd[l].width;
  }
  int ctr = 0;
  size_t scratch = 0;
  for (int l = 0; l + 1 < L; ++l) {                      // backward tiles: delta_l from delta_{l+1} and W_{l+1}
    const int kh = ceil_div(dims[l + 1], 64), N = dims[l];
    long best = -1;
    for (int rb : {32, 64}) {
      if (rb == 64 && B <= 32) continue;
      const int ntr = ceil_div(B, rb);
      for (int S : {1, 2, 4}) {
        const int kc = ceil_div(kh, S);
        if (ceil_div(kh, kc) != S) continue;
        if (S > 1 && (rb % (16 * S)) != 0) continue;
        int w = 0;
        for (int c = 8; c <= 64; c += 8)
          if (ntr * ceil_div(N, c) * S <= G) { w = c; break; }
        if (!w) continue;
        const long ab = (long)rb * 128 * kc, bb = (long)ceil_div(kc * 64, 256) * std::min(kc * 64, 256) * 128;
        if (ab + bb > slot_max) continue;
        const long cost = ab + bb + (S > 1 ? 98304 : 0);   // the exchange of a split costs about as much as fetching 96 KB
        if (best < 0 || cost < best) {
          best = cost;
          MtGroup& g = bwd[l];
          g.kind = MT_BWD; g.layer = l; g.rb = rb; g.width = w; g.nt_rows = ntr; g.nt_cols = ceil_div(N, w);
          g.ksplit = S; g.kc_slabs = kc; g.n_tiles = g.nt_rows * g.nt_cols * S; g.a_bytes = (uint32_t)ab; g.b_bytes = (uint32_t)bb;
        }
      }
    }
    if (best < 0) return false;
    if (bwd[l].ksplit > 1) {
      bwd[l].ctr0 = ctr;
      ctr += bwd[l].nt_rows * bwd[l].nt_cols;
      scratch += (size_t)bwd[l].n_tiles * 64 * 64 * sizeof(float);
    }
    P.b_rb[l] = bwd[l].rb; P.b_kc[l] = bwd[l].kc_slabs;
  }
  if (ctr > TE_MAX_TILE_CTRS) return false;
  for (int l = 0; l <= L; ++l) {                         // update tiles (l == L: the head); the ones column of the input carries the bias
    const int M = l == L ? n_out : dims[l], Kin = l == L ? HL : in_width(l);
    int n4 = 0;
    for (int c = 8; c <= 128; c += 8)
      if (ceil_div(M, 64) * ceil_div(Kin + 1, c) <= G) { n4 = c; break; }
    if (!n4) return false;
    MtGroup& g = upd[l];
    g.kind = MT_UPD; g.layer = l; g.rb = 64; g.width = n4; g.nt_rows = ceil_div(M, 64); g.nt_cols = ceil_div(Kin + 1, n4);
    g.ksplit = 1; g.n_chunks = ceil_div(n4, 64); g.n_tiles = g.nt_rows * g.nt_cols;
    g.a_bytes = (uint32_t)TE_BOX_BYTES; g.b_bytes = (uint32_t)(g.n_chunks * TE_BOX_BYTES);
  }
  // the program:  F_0 .. F_{L-1} | HEAD | (B_l + U_{l+2}) for l = L-2 .. 0 | U_1 + U_0
  int np = 0;
  for (int l = 0; l < L; ++l) P.ph[np++].g[0] = fwd[l];
  P.ph[np].g[0].kind = MT_HEAD;
  P.ph[np].g[0].n_tiles = (B + 1) / 2;
  ++np;
  static const bool serial = getenv("DBN_B200_MT_SERIAL") != nullptr && getenv("DBN_B200_MT_SERIAL")[0] == '1';   // debug: one group per phase
  for (int l = L - 2; l >= 0; --l) {
    P.ph[np].g[0] = bwd[l];
    if (serial) ++np;
    P.ph[np].g[serial ? 0 : 1] = upd[l + 2];
    ++np;
  }
  P.ph[np].g[0] = upd[1];
  if (serial) ++np;
  P.ph[np].g[serial ? 0 : 1] = upd[0];
  ++np;
  P.n_phases = np;
  const int ldwo_bytes = n_out * HL * 4;
  P.wo_bytes = (uint32_t)ldwo_bytes;
  if ((P.wo_bytes % 16) != 0) return false;
  uint32_t end = (uint32_t)align_up(P.wo_bytes, 1024);
  P.grid = (B + 1) / 2;
  for (int p = 0; p < np; ++p) {
    uint32_t off = 0;
    for (int gi = 0; gi < 2; ++gi) {
      MtGroup& g = P.ph[p].g[gi];
      if (g.kind == MT_NONE || g.kind == MT_HEAD) continue;
      g.off_a = off;
      off = (uint32_t)align_up(off + g.a_bytes, 1024);
      g.off_b = off;
      off = (uint32_t)align_up(off + g.b_bytes, 1024);
      P.grid = std::max(P.grid, g.n_tiles);
    }
    end = std::max(end, off);
  }
  P.off_bars = end;
  P.smem_bytes = end + 4096 + 1024;
  if (P.smem_bytes > (uint32_t)TE_SMEM_MAX || P.grid > G) return false;
  P.scratch_bytes = scratch;
  if (out) *out = P;
  return true;
}

inline int mt_max_coresident() {
  static int cache[DBN_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (cache[dev] == 0) {
    int per_sm = 0, sms = 0;
    if (cudaFuncSetAttribute(mlp_tc_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_MAX) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mlp_tc_epoch_kernel, TE_THREADS, TE_SMEM_MAX) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      per_sm = 0;
    }
    cache[dev] = per_sm * sms + 1;
  }
  return cache[dev] - 1;
}

}  // namespace dbn
