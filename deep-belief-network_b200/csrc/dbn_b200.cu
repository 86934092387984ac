// dbn_b200.cu -- C-ABI entry points (include/dbn_b200.h) of the B200-native CD-k / backprop path.
// Host-side glue only: argument checks, workspace carving and kernel launches.  All arithmetic
// lives in simt_gemm.cuh (strict FP32), umma_gemm.cuh (tcgen05 BF16) and misc_kernels.cuh.
#include "dbn_common.cuh"
#include "epilogue.cuh"
#include "simt_gemm.cuh"
#include "umma_gemm.cuh"
#include "misc_kernels.cuh"
#include "rbm_persistent.cuh"
#include "rbm_tc_epoch.cuh"
#include "mlp_tc_epoch.cuh"
#include "dp_kernels.cuh"
#include <memory>

using namespace dbn;

namespace {

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// the two tensor-core modes differ only in the 16-bit element type of every operand buffer
inline bool is_tc(int precision) { return precision == DBN_PREC_BF16 || precision == DBN_PREC_F16; }
inline int tc_dtype(int precision) { return precision == DBN_PREC_F16 ? DBN_F16 : DBN_BF16; }

inline dbn_mat make_mat(void* p, int ld, int dtype) {
  dbn_mat m;
  m.ptr = p; m.ld = ld; m.dtype = dtype;
  return m;
}
inline dbn_mat null_mat() { return make_mat(nullptr, 0, DBN_F32); }
inline size_t elt_size(int dtype) { return dtype == DBN_F32 ? 4 : 2; }
inline dbn_mat offset_cols(const dbn_mat& m, int cols) {
  dbn_mat r = m;
  if (r.ptr) r.ptr = static_cast<char*>(r.ptr) + (size_t)cols * elt_size(r.dtype);
  return r;
}
inline dbn_operands make_ops(const dbn_mat& A, int tA, int row0A, const dbn_mat& B, int tB, int row0B, int K) {
  dbn_operands o;
  o.A = A; o.B = B; o.transA = tA; o.transB = tB; o.K = K; o.row0_A = row0A; o.row0_B = row0B; o._pad = 0;
  return o;
}
inline dbn_act_epilogue blank_act() {
  dbn_act_epilogue e;
  memset(&e, 0, sizeof(e));
  e.rng.keep_p = 1.0f;
  return e;
}
inline dbn_update_epilogue blank_upd() {
  dbn_update_epilogue e;
  memset(&e, 0, sizeof(e));
  e.decay = 1.0f;
  return e;
}

int g_device_count = -1;
int device_count_cached() {
  if (g_device_count < 0) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
      int major = 0;
      if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    g_device_count = ok;
  }
  return g_device_count;
}
#define DBN_NEED_DEVICE()                                                                           \
  do {                                                                                              \
    if (device_count_cached() <= 0)                                                                 \
      return ::dbn::fail(DBN_ERR_NO_DEVICE, "%s: no sm_100 CUDA device (this library has no CPU fallback)%s", __func__, ""); \
  } while (0)

// bf16 row pitch: room for the ones column, multiple of 64 elements so that every row starts on a
// 128-byte line (TMA needs 16 B; line-aligned rows halve the L2 requests of a 128-byte box row).
inline int bf16_pitch(int cols) { return (int)align_up((size_t)cols + 1, 64); }

struct RbmWorkspace {
  dbn_mat P0, Hs, Vt, Pk;
  unsigned long long* stats;   // [H + V] fixed-point bias statistics dc | db (dbn_colsum)
  TeCtl* ctl;                  // tensor-core modes: control block + split-K scratch of the persistent epoch kernel
  float* scratch;
};
size_t rbm_ws_layout(int precision, int Bmax, int V, int H, char* base, RbmWorkspace* ws) {
  const int dt = is_tc(precision) ? tc_dtype(precision) : DBN_F32;
  const int ldH = is_tc(precision) ? bf16_pitch(H) : H;
  const int ldV = is_tc(precision) ? bf16_pitch(V) : V;
  size_t off = 0;
  auto carve = [&](int ld) {
    dbn_mat m = make_mat(base ? base + off : nullptr, ld, dt);
    off = align_up(off + (size_t)Bmax * ld * elt_size(dt), 1024);
    return m;
  };
  RbmWorkspace w;
  w.P0 = carve(ldH); w.Hs = carve(ldH); w.Vt = carve(ldV); w.Pk = carve(ldH);
  w.stats = reinterpret_cast<unsigned long long*>(base ? base + off : nullptr);
  off = align_up(off + (size_t)(H + V) * sizeof(unsigned long long), 1024);
  w.ctl = nullptr;
  w.scratch = nullptr;
  if (is_tc(precision)) {
    w.ctl = reinterpret_cast<TeCtl*>(base ? base + off : nullptr);
    off = align_up(off + sizeof(TeCtl), 1024);
    TePlan pl;
    if (te_plan(V, H, Bmax, &pl) && pl.scratch_bytes > 0) {
      w.scratch = reinterpret_cast<float*>(base ? base + off : nullptr);
      off = align_up(off + pl.scratch_bytes, 1024);
    }
  }
  if (ws) *ws = w;
  return off;
}

// Bias statistics by column sums in the Gibbs-chain epilogues instead of the ones row / column of dW: on request,
// or -- for an in-place update -- when the extra row / column would cost the dW contraction a whole tile row /
// column (H or V a multiple of 64: 4096 -> 17 x 17 instead of 16 x 16 CTA-pair tiles, +12.9 % MMA work).
inline int stats_override() {   // DBN_B200_BIAS_STATS=0/1 forces the choice (experiments); unset = automatic
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("DBN_B200_BIAS_STATS");
    v = (e == nullptr || e[0] == 0) ? -1 : (e[0] != '0');
  }
  return v;
}
inline bool rbm_use_stats(const dbn_rbm_step_args* a) {
  if (!is_tc(a->precision)) return false;
  if (a->flags & DBN_STEP_BIAS_STATS) return true;
  if (a->flags & DBN_STEP_SKIP_UPDATE) return false;   // deltas come later from dbn_rbm_delta_rows (ones row / column)
  if (stats_override() >= 0) return stats_override() == 1;
  return (a->H % 64) == 0 || (a->V % 64) == 0;
}

struct MlpWorkspace {
  dbn_mat a0;                         // masked input (dropout only)
  dbn_mat act[DBN_MAX_LAYERS];        // a_1..a_L
  dbn_mat delta[DBN_MAX_LAYERS];      // delta_1..delta_L
  dbn_mat Z, dO;                      // head scores/outputs, head delta (fp32)
  dbn_mat dOb;                        // bf16 copy of the head delta (BF16 mode: operand of the head-gradient contraction)
  TeCtl* ctl;                         // tensor-core modes: control block + split-K scratch of the persistent fine-tune kernel
  float* scratch;
  dbn_mat a0b;                        // tensor-core modes: second masked-input buffer (the persistent kernel masks one step ahead)
};
size_t mlp_ws_layout(int precision, int Bmax, int n_in, const int32_t* dims, int L, int n_out, char* base,
                     MlpWorkspace* ws) {
  const int dt = is_tc(precision) ? tc_dtype(precision) : DBN_F32;
  size_t off = 0;
  auto carve = [&](int cols, int dtype, bool ones) {
    const int ld = (dtype != DBN_F32) ? bf16_pitch(cols) : cols;
    (void)ones;
    dbn_mat m = make_mat(base ? base + off : nullptr, ld, dtype);
    off = align_up(off + (size_t)Bmax * ld * elt_size(dtype), 1024);
    return m;
  };
  MlpWorkspace w;
  w.a0 = carve(n_in, dt, true);
  for (int l = 0; l < L; ++l) w.act[l] = carve(dims[l], dt, true);
  for (int l = 0; l < L; ++l) w.delta[l] = carve(dims[l], dt, false);
  w.Z = carve(n_out, DBN_F32, false);
  w.dO = carve(n_out, DBN_F32, false);
  w.dOb = is_tc(precision) ? carve(n_out, dt, false) : null_mat();
  w.ctl = nullptr;
  w.scratch = nullptr;
  if (is_tc(precision)) {
    w.ctl = reinterpret_cast<TeCtl*>(base ? base + off : nullptr);
    off = align_up(off + sizeof(TeCtl), 1024);
    MtPlan pl;
    if (mt_plan(n_in, dims, L, n_out, Bmax, &pl) && pl.scratch_bytes > 0) {
      w.scratch = reinterpret_cast<float*>(base ? base + off : nullptr);
      off = align_up(off + pl.scratch_bytes, 1024);
    }
    w.a0b = carve(n_in, dt, true);
  } else {
    w.a0b = null_mat();
  }
  if (ws) *ws = w;
  return off;
}

// Fork / join between the two streams of a fine-tune step.  Events are created once (per process, timing disabled)
// and reused round-robin; a record followed by a wait is legal inside a stream capture, where it becomes a graph edge.
cudaEvent_t next_event() {
  static cudaEvent_t ring[64];
  static std::atomic<unsigned> n{0};
  static std::atomic<uint64_t> made{0};
  const unsigned i = n.fetch_add(1) % 64;
  if (!((made.load() >> i) & 1)) {
    cudaEventCreateWithFlags(&ring[i], cudaEventDisableTiming);
    made.fetch_or(uint64_t(1) << i);
  }
  return ring[i];
}
// `to` waits for everything issued so far on `from`
int stream_edge(cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = next_event();
  DBN_CUDA_OK(cudaEventRecord(e, from));
  DBN_CUDA_OK(cudaStreamWaitEvent(to, e, 0));
  return DBN_OK;
}

int copy2d(cudaStream_t st, const dbn_mat& dst, const dbn_mat& src, int rows, int cols) {
  if (dst.dtype == src.dtype) {
    DBN_CUDA_OK(cudaMemcpy2DAsync(dst.ptr, (size_t)dst.ld * elt_size(dst.dtype), src.ptr,
                                  (size_t)src.ld * elt_size(src.dtype), (size_t)cols * elt_size(src.dtype), rows,
                                  cudaMemcpyDeviceToDevice, st));
    return DBN_OK;
  }
  dbn_rng none;
  memset(&none, 0, sizeof(none));
  gather_rows_kernel<<<min(rows, 1184), 256, 0, st>>>(src, nullptr, rows, cols, dst, -1, none);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

}  // namespace

extern "C" {

int dbn_abi_version(void) { return DBN_B200_ABI_VERSION; }
const char* dbn_last_error(void) { return last_error_ref().c_str(); }
int dbn_device_count(void) { return device_count_cached(); }
uint64_t dbn_launch_count(void) { return launch_counter().load(); }
uint64_t dbn_tail_split_count(void) { return ts_counter().load(); }
int dbn_set_tail_split(int enabled) {
  const int prev = ts_enabled_flag();
  ts_enabled_flag() = enabled ? 1 : 0;
  return prev;
}
int dbn_set_sm_reserve(int n_sms) {
  DBN_REQUIRE(n_sms >= 0 && n_sms <= 64, "reserve out of range");
  umma_sm_reserve() = n_sms;
  return DBN_OK;
}
int dbn_debug_set_trace(void* device_buf) {
  umma_trace_ptr() = static_cast<long long*>(device_buf);
  return DBN_OK;
}
size_t dbn_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(dbn_mat);
    case 1: return sizeof(dbn_rng);
    case 2: return sizeof(dbn_operands);
    case 3: return sizeof(dbn_act_epilogue);
    case 4: return sizeof(dbn_update_epilogue);
    case 5: return sizeof(dbn_rbm_step_args);
    case 6: return sizeof(dbn_mlp_step_args);
    case 7: return sizeof(dbn_colsum);
    case 8: return sizeof(dbn_dp_exchange);
    default: return 0;
  }
}

int dbn_philox_uniforms_host(uint64_t seed, uint32_t stream, uint32_t epoch, uint32_t row0, int rows, int cols,
                             float* out_host) {
  DBN_REQUIRE(out_host && rows >= 0 && cols >= 0, "bad argument");
  for (int r = 0; r < rows; ++r)
    for (int q = 0; q * 4 < cols; ++q) {
      uint32_t o[4];
      philox4x32_10(row0 + (uint32_t)r, (uint32_t)q, epoch, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
      for (int j = 0; j < 4 && q * 4 + j < cols; ++j) out_host[(size_t)r * cols + q * 4 + j] = u01_from_bits(o[j]);
    }
  return DBN_OK;
}

// ------------------------------------------------------------------------------------------------
int dbn_gemm_act(void* stream, int precision, int M, int N, const dbn_operands* ops, int n_ops,
                 const dbn_act_epilogue* epi) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(ops && epi && (n_ops == 1 || n_ops == 2), "bad operands");
  DBN_REQUIRE(M > 0 && N > 0, "empty output");
  if (is_tc(precision) && umma_gemm_supported(M, N, ops, n_ops))
    return umma_gemm_act(as_stream(stream), M, N, ops, n_ops, *epi);
  return simt_gemm_act(as_stream(stream), M, N, ops, n_ops, *epi);
}

int dbn_gemm_update(void* stream, int precision, int M, int N, int augM, int augN, const dbn_operands* ops,
                    int n_ops, const dbn_update_epilogue* epi) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(ops && epi && (n_ops == 1 || n_ops == 2), "bad operands");
  DBN_REQUIRE(M > 0 && N > 0, "empty output");
  if (is_tc(precision) && umma_gemm_supported(M + augM, N + augN, ops, n_ops))
    return umma_gemm_update(as_stream(stream), M, N, augM, augN, ops, n_ops, *epi);
  return simt_gemm_update(as_stream(stream), M, N, augM, augN, ops, n_ops, *epi);
}

// ------------------------------------------------------------------------------------------------
int dbn_rbm_hidden_means(void* stream, int precision, const dbn_mat* X, int row0, int B, int V, int H,
                         const dbn_mat* W, const float* c, int act, const dbn_mat* out) {
  DBN_REQUIRE(X && W && out && out->ptr, "null argument");
  dbn_operands o = make_ops(*X, 0, row0, *W, 0, 0, V);
  dbn_act_epilogue e = blank_act();
  e.bias = c; e.act = act; e.out = *out;
  return dbn_gemm_act(stream, precision, B, H, &o, 1, &e);
}

int dbn_rbm_visible_means(void* stream, int precision, const dbn_mat* Hm, int B, int V, int H, const dbn_mat* W,
                          const float* b, int act, const dbn_mat* out) {
  DBN_REQUIRE(Hm && W && out && out->ptr, "null argument");
  dbn_operands o = make_ops(*Hm, 0, 0, *W, 1, 0, H);
  dbn_act_epilogue e = blank_act();
  e.bias = b; e.act = act; e.out = *out;
  return dbn_gemm_act(stream, precision, B, V, &o, 1, &e);
}

int dbn_rbm_reconstruction_sq(void* stream, int precision, const dbn_mat* Hm, int B, int V, int H, const dbn_mat* W,
                              const float* b, int act, const dbn_mat* X, int row0, unsigned long long* acc) {
  DBN_REQUIRE(Hm && W && X && X->ptr && acc, "null argument");
  dbn_operands o = make_ops(*Hm, 0, 0, *W, 1, 0, H);
  dbn_act_epilogue e = blank_act();
  e.bias = b; e.act = act;
  e.colsum.acc = acc; e.colsum.sign = 1.0f; e.colsum.square = 1;
  e.colsum.ref = *X; e.colsum.ref_row0 = row0;
  return dbn_gemm_act(stream, precision, B, V, &o, 1, &e);
}

size_t dbn_rbm_workspace_bytes(int precision, int Bmax, int V, int H) {
  return rbm_ws_layout(precision, Bmax, V, H, nullptr, nullptr);
}

int dbn_rbm_workspace_init(void* stream, int precision, int Bmax, int V, int H, void* workspace) {
  DBN_NEED_DEVICE();
  RbmWorkspace ws;
  size_t bytes = rbm_ws_layout(precision, Bmax, V, H, static_cast<char*>(workspace), &ws);
  DBN_CUDA_OK(cudaMemsetAsync(workspace, 0, bytes, as_stream(stream)));
  if (is_tc(precision)) {
    DBN_TRY(dbn_fill_ones_column(stream, &ws.P0, Bmax, H));
    DBN_TRY(dbn_fill_ones_column(stream, &ws.Pk, Bmax, H));
    DBN_TRY(dbn_fill_ones_column(stream, &ws.Vt, Bmax, V));
  }
  return DBN_OK;
}

// The workspace layout (and its ones columns) is fixed by the row count it was initialised for.
static inline int rbm_ws_rows(const dbn_rbm_step_args* a) { return a->ws_rows > 0 ? a->ws_rows : a->B; }

// t = 0 hidden phase on hidden units [h0, h1): P0 (kept for dW == h_0 of dbn/models.py:140) and the
// first sample (dbn/models.py:135 -> :148-155).
static int rbm_first_hidden_impl(void* stream, const dbn_rbm_step_args* a, int h0, int h1) {
  const bool tc = is_tc(a->precision);
  RbmWorkspace ws;
  rbm_ws_layout(a->precision, rbm_ws_rows(a), a->V, a->H, static_cast<char*>(a->workspace), &ws);
  const dbn_mat Wop = tc ? a->Wb : make_mat(a->W, a->ldw, DBN_F32);
  dbn_operands o = make_ops(a->X, 0, a->row0, Wop, 0, h0, a->V);
  dbn_act_epilogue e = blank_act();
  e.bias = a->c + h0; e.act = a->act;
  e.rng = a->rng;
  e.rng.mode = DBN_RNG_SAMPLE;
  e.rng.col0 = a->rng.col0 + (uint32_t)h0;
  if (a->rng.injected.ptr) e.rng.injected = offset_cols(a->rng.injected, h0);
  e.out = offset_cols(ws.P0, h0);
  e.out2 = offset_cols(a->P0, h0);
  e.sample = offset_cols(ws.Hs, h0);
  if (rbm_use_stats(a)) {   // dc += sum_rows h_0 (dbn/models.py:144)
    e.colsum.acc = ws.stats + h0;
    e.colsum.sign = 1.0f;
  }
  return dbn_gemm_act(stream, a->precision, a->B, h1 - h0, &o, 1, &e);
}

int dbn_rbm_first_hidden(void* stream, const dbn_rbm_step_args* a, int h0, int h1) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(a != nullptr, "null args");
  DBN_REQUIRE(h0 >= 0 && h0 < h1 && h1 <= a->H && (h0 % 64) == 0, "bad unit range (h0 must be a multiple of 64)");
  if (a->k == 0) return DBN_OK;
  return rbm_first_hidden_impl(stream, a, h0, h1);
}

int dbn_rbm_cd_step(void* stream, const dbn_rbm_step_args* a) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(a != nullptr, "null args");
  DBN_REQUIRE(a->B > 0 && a->V > 0 && a->H > 0 && a->k >= 0, "bad shape");
  DBN_REQUIRE(a->act == DBN_ACT_SIGMOID || a->act == DBN_ACT_RELU, "Invalid activation function.");
  DBN_REQUIRE(a->ws_rows == 0 || a->ws_rows >= a->B, "ws_rows smaller than the mini-batch");
  DBN_REQUIRE(a->workspace_bytes >= dbn_rbm_workspace_bytes(a->precision, rbm_ws_rows(a), a->V, a->H), "workspace too small");
  const bool tc = is_tc(a->precision);
  DBN_REQUIRE(!tc || (a->Wb.ptr && a->X.dtype == tc_dtype(a->precision) && a->Wb.dtype == a->X.dtype),
              "a tensor-core mode needs X and the shadow of W in its 16-bit type (the two operands of an MMA share one format)");
  if (a->k == 0) return DBN_OK;  // v_k == v_0: all deltas are zero (dbn/models.py:134 loop is empty)
  RbmWorkspace ws;
  rbm_ws_layout(a->precision, rbm_ws_rows(a), a->V, a->H, static_cast<char*>(a->workspace), &ws);
  const dbn_mat Wop = tc ? a->Wb : make_mat(a->W, a->ldw, DBN_F32);
  const int B = a->B, V = a->V, H = a->H;
  const bool stats = rbm_use_stats(a);

  for (int t = 0; t < a->k; ++t) {
    // h_t ~ P(h | v_t): dbn/models.py:135 -> :148-155
    if (t == 0) {
      if (!(a->flags & DBN_STEP_SKIP_FIRST_HIDDEN)) DBN_TRY(rbm_first_hidden_impl(stream, a, 0, H));
    } else {
      dbn_operands o = make_ops(ws.Vt, 0, 0, Wop, 0, 0, V);
      dbn_act_epilogue e = blank_act();
      e.bias = a->c; e.act = a->act;
      e.rng = a->rng;
      e.rng.mode = DBN_RNG_SAMPLE;
      e.rng.stream = a->rng.stream + (uint32_t)t;
      if (a->rng.injected.ptr) e.rng.injected = offset_cols(a->rng.injected, t * H);
      e.out = ws.Pk;  // only the t = 0 means are kept (== h_0 of :140)
      e.sample = ws.Hs;
      DBN_TRY(dbn_gemm_act(stream, a->precision, B, H, &o, 1, &e));
    }
    if (t == 0 && a->H0s.ptr) DBN_TRY(copy2d(as_stream(stream), a->H0s, ws.Hs, B, H));
    // v_{t+1} = mean visible: dbn/models.py:136 -> :185-201
    dbn_operands o2 = make_ops(ws.Hs, 0, 0, Wop, 1, 0, H);
    dbn_act_epilogue e2 = blank_act();
    e2.bias = a->b; e2.act = a->act; e2.out = ws.Vt;
    if (t == a->k - 1) {
      e2.out2 = a->Vk;
      if (stats) {   // db = sum_rows (v_0 - v_k) (dbn/models.py:143): v_0 re-read from the (L2-hot) mini-batch
        e2.colsum.acc = ws.stats + H;
        e2.colsum.sign = 1.0f;
        e2.colsum.ref = a->X;
        e2.colsum.ref_row0 = a->row0;
      }
    }
    DBN_TRY(dbn_gemm_act(stream, a->precision, B, V, &o2, 1, &e2));
  }
  {  // h_k = P(h | v_k): dbn/models.py:141
    dbn_operands o = make_ops(ws.Vt, 0, 0, Wop, 0, 0, V);
    dbn_act_epilogue e = blank_act();
    e.bias = a->c; e.act = a->act; e.out = ws.Pk; e.out2 = a->Pk;
    if (stats) {   // dc -= sum_rows h_k
      e.colsum.acc = ws.stats;
      e.colsum.sign = -1.0f;
    }
    DBN_TRY(dbn_gemm_act(stream, a->precision, B, H, &o, 1, &e));
  }
  if (a->flags & DBN_STEP_SKIP_UPDATE) return DBN_OK;
  {  // dW = h_0^T v_0 - h_k^T v_k, db, dc (dbn/models.py:142-144) + update (:117-119)
    dbn_operands o[2];
    o[0] = make_ops(ws.P0, 1, 0, a->X, 1, a->row0, B);
    o[1] = make_ops(ws.Pk, 1, 0, ws.Vt, 1, 0, B);
    dbn_update_epilogue u = blank_upd();
    u.W = a->W; u.ldw = a->ldw; u.Wb = tc ? a->Wb : null_mat();
    u.vecM = a->c; u.vecN = a->b; u.decay = 1.0f; u.scale = a->lr_over_bs; u.raw = a->raw_delta;
    if (!stats) {
      DBN_TRY(dbn_gemm_update(stream, a->precision, H, V, 1, 1, o, 2, &u));
    } else {
      // the statistics were taken in the chain's epilogues: H x V accumulator, no ones row / column; the update
      // epilogue's owner threads fold them into b and c (and clear them)
      u.statM = ws.stats; u.statN = ws.stats + H;
      DBN_TRY(dbn_gemm_update(stream, a->precision, H, V, 0, 0, o, 2, &u));
      if (a->raw_delta.ptr) {
        stats_to_raw_kernel<<<ceil_div(H + V, 256), 256, 0, as_stream(stream)>>>(ws.stats, H, V, a->raw_delta);
        DBN_LAUNCH_CHECK();
      }
    }
  }
  return DBN_OK;
}

int dbn_rbm_persistent_plan(int precision, int V, int H, int batch_size) {
  if (precision != DBN_PREC_FP32) return 0;
  int Hc = 0;
  return rp_plan(V, H, batch_size, &Hc);
}

int dbn_rbm_fit_epoch_persistent(void* stream, const dbn_rbm_step_args* t, const int32_t* idx, int n_rows,
                                 int batch_size) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(t && idx && n_rows >= 0 && batch_size > 0, "bad argument");
  DBN_REQUIRE(t->precision == DBN_PREC_FP32 && t->X.dtype == DBN_F32, "the persistent kernel is the strict-FP32 path");
  DBN_REQUIRE(t->act == DBN_ACT_SIGMOID || t->act == DBN_ACT_RELU, "Invalid activation function.");
  DBN_REQUIRE(!t->rng.injected.ptr || t->rng.injected.dtype == DBN_F32, "injected uniforms must be fp32");
  if (n_rows == 0 || t->k == 0) return DBN_OK;
  RbmPersistArgs a;
  memset(&a, 0, sizeof(a));
  a.X = static_cast<const float*>(t->X.ptr); a.ldx = t->X.ld;
  a.idx = idx; a.n_rows = n_rows; a.batch_size = batch_size;
  a.V = t->V; a.H = t->H; a.k = t->k; a.act = t->act;
  a.lr_over_bs = t->lr_over_bs;
  a.W = t->W; a.ldw = t->ldw; a.b = t->b; a.c = t->c;
  a.rng = t->rng;
  return rbm_persistent_launch(as_stream(stream), a, t->V, t->H, batch_size);
}

// Whole epoch in ONE cooperative launch of rbm_tc_epoch_kernel when the layer fits its tile plan (tensor-core modes,
// mini-batches of 16..256 rows, plain in-place training); *done = 0: not applicable, the caller runs the tiled path.
static int rbm_fit_epoch_tc(cudaStream_t st, const dbn_rbm_step_args* t, int n_rows, int batch_size, int* done) {
  *done = 0;
  const int B = t->ws_rows > 0 ? t->ws_rows : std::min(batch_size, n_rows);
  if (!te_enabled() || !is_tc(t->precision) || t->k < 1 || n_rows <= 0) return DBN_OK;
  if (B < std::min(batch_size, n_rows) || (n_rows > B && B != batch_size)) return DBN_OK;   // the step stride is the tile height
  if (t->flags & (DBN_STEP_SKIP_UPDATE | DBN_STEP_SKIP_FIRST_HIDDEN)) return DBN_OK;
  if (t->P0.ptr || t->Vk.ptr || t->Pk.ptr || t->H0s.ptr || t->raw_delta.ptr) return DBN_OK;
  if (t->act != DBN_ACT_SIGMOID && t->act != DBN_ACT_RELU) return DBN_OK;
  if (t->rng.injected.ptr && t->rng.injected.dtype != DBN_F32) return DBN_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return DBN_OK;   // a cooperative launch is not captured: graphs keep the tiled path
  }
  TeArgs a;
  memset(&a, 0, sizeof(a));
  if (!te_plan(t->V, t->H, B, &a.pl)) return DBN_OK;
  if (a.pl.grid > te_max_coresident()) return DBN_OK;
  if (t->workspace_bytes < dbn_rbm_workspace_bytes(t->precision, B, t->V, t->H)) return DBN_OK;
  if (!umma_operand_ok(t->X, t->V) || !umma_operand_ok(t->Wb, t->V) || t->X.dtype != t->Wb.dtype) return DBN_OK;
  RbmWorkspace ws;
  rbm_ws_layout(t->precision, B, t->V, t->H, static_cast<char*>(t->workspace), &ws);
  if (a.pl.scratch_bytes > 0 && ws.scratch == nullptr) return DBN_OK;
  a.V = t->V; a.H = t->H; a.B = B; a.k = t->k; a.act = t->act; a.dtype = t->X.dtype;
  a.n_rows = n_rows; a.row0 = t->row0; a.scale = t->lr_over_bs;
  a.X = t->X; a.P0 = ws.P0; a.Hs = ws.Hs; a.Vt = ws.Vt; a.Pk = ws.Pk; a.Wb = t->Wb;
  a.W = t->W; a.ldw = t->ldw; a.b = t->b; a.c = t->c; a.stats = ws.stats;
  a.rng = t->rng;
  a.ctl = ws.ctl; a.scratch = ws.scratch;
  a.trace = umma_trace_ptr();
  { const char* e = getenv("DBN_B200_TE_EXP"); a.exp = e ? atoi(e) : 0; }
  const TePlan& pl = a.pl;
  const uint64_t xrows = (uint64_t)t->row0 + (uint64_t)n_rows;
  // K-major operands: one copy brings every 64-wide k-slab of a tile's rows
  DBN_TRY(make_operand_map(&a.tm_x3, t->X, xrows, pl.kv_slabs, pl.rb_h, pl.kv_slabs));
  DBN_TRY(make_operand_map(&a.tm_vt3, ws.Vt, B, pl.kv_slabs, pl.rb_h, pl.kv_slabs));
  DBN_TRY(make_operand_map(&a.tm_hs3, ws.Hs, B, ceil_div(t->H, 64), pl.rb_v, pl.kc_slabs));
  DBN_TRY(make_operand_map(&a.tm_wb3, t->Wb, t->H, pl.kv_slabs, pl.hs, pl.kv_slabs));
  // MN-major operands: 64 columns x up to 256 k rows per copy
  DBN_TRY(te_make_map2(&a.tm_wb2, t->Wb, t->Wb.ld, t->H, std::min(pl.kc_slabs * 64, 256)));
  DBN_TRY(te_make_map2(&a.tm_p02, ws.P0, ws.P0.ld, B, 256));
  DBN_TRY(te_make_map2(&a.tm_pk2, ws.Pk, ws.Pk.ld, B, 256));
  DBN_TRY(te_make_map2(&a.tm_x2, t->X, t->X.ld, xrows, 256));
  DBN_TRY(te_make_map2(&a.tm_vt2, ws.Vt, ws.Vt.ld, B, 256));
  static std::atomic<uint64_t> attr_set{0};
  if (first_use_on_device(attr_set))
    DBN_CUDA_OK(cudaFuncSetAttribute(rbm_tc_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_MAX));
  DBN_CUDA_OK(cudaMemsetAsync(ws.ctl, 0, sizeof(TeCtl), st));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(TE_THREADS);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident or the launch fails: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, rbm_tc_epoch_kernel, a) != cudaSuccess) {
    (void)cudaGetLastError();   // e.g. the grid cannot be co-resident right now (another context on the GPU): tiled path
    return DBN_OK;
  }
  DBN_LAUNCH_CHECK();
  *done = 1;
  return DBN_OK;
}

int dbn_rbm_tc_epoch_plan(int precision, int V, int H, int batch_size) {
  TePlan pl;
  if (!is_tc(precision) || !te_enabled() || !te_plan(V, H, batch_size, &pl)) return 0;
  return pl.grid;
}
int dbn_set_tc_epoch(int enabled) {
  const int prev = te_enabled_flag();
  te_enabled_flag() = enabled ? 1 : 0;
  return prev;
}
int dbn_rbm_tc_epoch_plan_info(int V, int H, int batch_size, int32_t* out, int n) {
  DBN_REQUIRE(out && n >= 0, "bad argument");
  TePlan pl;
  if (!te_plan(V, H, batch_size, &pl)) return 0;
  const int32_t v[] = {pl.grid, pl.rb_h, pl.hs, pl.nt_h_rows * pl.nt_h_cols, pl.rb_v, pl.vs, pl.ksplit, pl.kc_slabs,
                       pl.nt_v_rows * pl.nt_v_cols * pl.ksplit, pl.n4, pl.nt_u_rows * pl.nt_u_cols, (int32_t)pl.smem_bytes,
                       pl.v1c_early, pl.p1_in_act, (int32_t)pl.scratch_bytes, pl.kv_slabs};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
  return m;
}

int dbn_rbm_fit_epoch(void* stream, const dbn_rbm_step_args* tmpl, int n_rows, int batch_size) {
  DBN_REQUIRE(tmpl && n_rows >= 0 && batch_size > 0, "bad argument");
  {
    DBN_NEED_DEVICE();
    int done = 0;
    DBN_TRY(rbm_fit_epoch_tc(as_stream(stream), tmpl, n_rows, batch_size, &done));
    if (done) return DBN_OK;
  }
  for (int s = 0; s < n_rows; s += batch_size) {  // batch_generator: dbn/utils.py:14-22 (short last batch)
    dbn_rbm_step_args a = *tmpl;
    a.B = std::min(batch_size, n_rows - s);
    if (a.ws_rows == 0) a.ws_rows = std::min(batch_size, n_rows);   // one layout for every batch of the epoch
    a.row0 = tmpl->row0 + s;
    a.rng.row0 = tmpl->rng.row0 + (uint32_t)s;
    if (a.rng.injected.ptr)
      a.rng.injected.ptr = static_cast<char*>(a.rng.injected.ptr) + (size_t)s * a.rng.injected.ld * sizeof(float);
    DBN_TRY(dbn_rbm_cd_step(stream, &a));
  }
  return DBN_OK;
}

int dbn_rbm_delta_rows(void* stream, const dbn_rbm_step_args* a, int h0, int h1) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(a && a->raw_delta.ptr, "needs args with a raw_delta buffer");
  DBN_REQUIRE(h0 >= 0 && h0 < h1 && h1 <= a->H && (h0 % 64) == 0, "bad row range (h0 must be a multiple of 64)");
  if (a->k == 0) return DBN_OK;
  RbmWorkspace ws;
  rbm_ws_layout(a->precision, rbm_ws_rows(a), a->V, a->H, static_cast<char*>(a->workspace), &ws);
  dbn_operands o[2];
  o[0] = make_ops(offset_cols(ws.P0, h0), 1, 0, a->X, 1, a->row0, a->B);
  o[1] = make_ops(offset_cols(ws.Pk, h0), 1, 0, ws.Vt, 1, 0, a->B);
  dbn_update_epilogue u = blank_upd();
  u.raw = a->raw_delta;
  u.raw.ptr = static_cast<char*>(u.raw.ptr) + (size_t)h0 * u.raw.ld * sizeof(float);
  if (!(is_tc(a->precision) && (a->flags & DBN_STEP_BIAS_STATS)))
    return dbn_gemm_update(stream, a->precision, h1 - h0, a->V, h1 == a->H ? 1 : 0, 1, o, 2, &u);
  // the chain took the bias statistics in its epilogues: plain H x V rows, then dc (column V of these rows) and,
  // with the last block, db (row H) out of the accumulators
  DBN_TRY(dbn_gemm_update(stream, a->precision, h1 - h0, a->V, 0, 0, o, 2, &u));
  stats_rows_to_raw_kernel<<<ceil_div((h1 - h0) + a->V, 256), 256, 0, as_stream(stream)>>>(ws.stats, a->H, a->V, h0, h1,
                                                                                         a->raw_delta);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_rbm_apply_delta(void* stream, int V, int H, int h0, int h1, int with_db, float scale, const float* raw, int ld,
                        int raw_row0, float* W, int ldw, float* b, float* c, const dbn_mat* Wb) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(raw && W && b && c, "null argument");
  DBN_REQUIRE(h0 >= 0 && h0 <= h1 && h1 <= H && (h0 < h1 || with_db), "bad row range");
  DBN_REQUIRE(raw_row0 <= h0 || h0 == h1, "raw_row0 beyond the first row");
  const int rows = (h1 - h0) + (with_db ? 1 : 0);
  const long total = (long)rows * ((V + 4) / 4);
  const int blocks = (int)std::min<long>((total + 255) / 256, 148L * 8);
  apply_delta_kernel<<<blocks, 256, 0, as_stream(stream)>>>(V, H, h0, h1 - h0, with_db, scale, raw, ld, raw_row0, W, ldw, b,
                                                           c, Wb ? *Wb : null_mat());
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

unsigned long long* dbn_rbm_workspace_stats(int precision, int Bmax, int V, int H, void* workspace) {
  RbmWorkspace ws;
  rbm_ws_layout(precision, Bmax, V, H, static_cast<char*>(workspace), &ws);
  return ws.stats;
}

// ---- data-parallel exchange over peer memory ---------------------------------------------------------
int dbn_peer_alloc(size_t bytes, void** ptr_out, void* handle_out) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(bytes > 0 && ptr_out && handle_out, "bad argument");
  void* p = nullptr;
  DBN_CUDA_OK(cudaMalloc(&p, bytes));
  DBN_CUDA_OK(cudaMemset(p, 0, bytes));
  DBN_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == DBN_PEER_HANDLE_BYTES, "handle size");
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {   // single-process use (one rank, emulation) works without a handle
    (void)cudaGetLastError();
    memset(&h, 0, sizeof(h));
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return DBN_OK;
}
int dbn_peer_free(void* ptr) {
  if (ptr) DBN_CUDA_OK(cudaFree(ptr));
  return DBN_OK;
}
int dbn_peer_open(const void* handle, void** ptr_out) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(handle && ptr_out, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  DBN_CUDA_OK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return DBN_OK;
}
int dbn_peer_close(void* ptr) {
  if (ptr) DBN_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return DBN_OK;
}

static int dp_check(const dbn_dp_exchange* x) {
  DBN_REQUIRE(x != nullptr, "null exchange");
  DBN_REQUIRE(x->n_peers >= 1 && x->n_peers <= DBN_MAX_PEERS && x->rank >= 0 && x->rank < x->n_peers, "bad rank / peer count");
  DBN_REQUIRE(x->H > 0 && x->V > 0 && x->rows_per_peer * x->n_peers == x->H, "H must be n_peers * rows_per_peer");
  DBN_REQUIRE(x->ld_recv >= x->V && (x->ld_recv % 4) == 0, "ld_recv must be a multiple of 4 and >= V");
  for (int j = 0; j < x->n_peers; ++j)
    DBN_REQUIRE(x->recv[j] && x->stats[j] && x->flags[j] && x->Wb[j].ptr &&
                (x->Wb[j].dtype == DBN_BF16 || x->Wb[j].dtype == DBN_F16) && x->Wb[j].dtype == x->Wb[0].dtype, "bad peer pointer / shadow type");
  return DBN_OK;
}

int dbn_rbm_delta_scatter(void* stream, const dbn_rbm_step_args* a, const dbn_dp_exchange* x) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(a != nullptr, "null args");
  DBN_TRY(dp_check(x));
  DBN_REQUIRE(is_tc(a->precision) && x->H == a->H && x->V == a->V, "exchange does not match the layer (tensor-core modes only)");
  if (a->k == 0) return DBN_OK;
  RbmWorkspace ws;
  rbm_ws_layout(a->precision, rbm_ws_rows(a), a->V, a->H, static_cast<char*>(a->workspace), &ws);
  dbn_operands o[2];
  o[0] = make_ops(ws.P0, 1, 0, a->X, 1, a->row0, a->B);
  o[1] = make_ops(ws.Pk, 1, 0, ws.Vt, 1, 0, a->B);
  dbn_update_epilogue u = blank_upd();
  for (int j = 0; j < x->n_peers; ++j) u.raw_peers[j] = x->recv[j];
  u.rows_per_peer = x->rows_per_peer;
  u.raw_slot = x->rank;
  u.raw.ld = x->ld_recv; u.raw.dtype = DBN_F32;
  return dbn_gemm_update(stream, a->precision, a->H, a->V, 0, 0, o, 2, &u);
}

int dbn_dp_signal(void* stream, const dbn_dp_exchange* x, unsigned long long* local_stats, unsigned long long value) {
  DBN_NEED_DEVICE();
  DBN_TRY(dp_check(x));
  DBN_REQUIRE(local_stats != nullptr, "null statistics");
  dp_signal_kernel<<<x->n_peers, 512, 0, as_stream(stream)>>>(*x, local_stats, value);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_dp_apply(void* stream, const dbn_dp_exchange* x, float scale, float* W, int ldw, float* b, float* c,
                 unsigned long long* local_stats, unsigned int* counter, unsigned long long value) {
  DBN_NEED_DEVICE();
  DBN_TRY(dp_check(x));
  DBN_REQUIRE(W && b && c && local_stats && counter && ldw >= x->V, "bad argument");
  // A small persistent grid on purpose: with every warp of a huge grid resident at once the kernel runs as two bulk
  // phases (all loads of the slots, then all peer stores); a few iterations per thread let the peer stores of
  // iteration i ride under the slot reads of iteration i+1 (dp_apply_blocks() = DBN_B200_DP_APPLY_BLOCKS or 2 per SM).
  const long total = (long)x->rows_per_peer * ((x->V + 7) / 8);
  static int per_sm = -1;
  if (per_sm < 0) {
    const char* e = getenv("DBN_B200_DP_APPLY_BLOCKS");
    per_sm = e ? std::max(1, atoi(e)) : 2;
  }
  const int blocks = (int)std::max<long>(1, std::min<long>((total + 255) / 256, 148L * per_sm));
  dp_apply_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*x, scale, W, ldw, b, c, local_stats, counter, value);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_dp_wait(void* stream, const dbn_dp_exchange* x, int which, unsigned long long value) {
  DBN_NEED_DEVICE();
  DBN_TRY(dp_check(x));
  DBN_REQUIRE(which == 0 || which == 1, "bad flag index");
  dp_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(*x, which, value);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_gather_rows_peer(void* stream, void* const* srcs, int n_srcs, int rows_per_src, int ld, const int32_t* idx,
                         int n_rows, const dbn_mat* dst) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(srcs && idx && dst && dst->ptr && n_srcs >= 1 && n_srcs <= DBN_MAX_PEERS && rows_per_src > 0, "bad argument");
  DBN_REQUIRE(dst->dtype != DBN_F32 && dst->ld == ld && (ld % 8) == 0, "16-bit rows of the same pitch (multiple of 8) on both sides");
  if (n_rows <= 0) return DBN_OK;
  PeerSrcs ps;
  memset(&ps, 0, sizeof(ps));
  for (int j = 0; j < n_srcs; ++j) {
    DBN_REQUIRE(srcs[j] != nullptr, "null source");
    ps.p[j] = srcs[j];
  }
  const int blocks = std::min(ceil_div(n_rows, 8), 148 * 8);
  gather_rows_peer_kernel<<<blocks, 256, 0, as_stream(stream)>>>(ps, rows_per_src, ld, idx, n_rows, *dst);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

// ------------------------------------------------------------------------------------------------
size_t dbn_mlp_workspace_bytes(int precision, int Bmax, int n_in, const int32_t* dims, int n_layers, int n_out) {
  return mlp_ws_layout(precision, Bmax, n_in, dims, n_layers, n_out, nullptr, nullptr);
}

int dbn_mlp_workspace_init(void* stream, int precision, int Bmax, int n_in, const int32_t* dims, int n_layers,
                           int n_out, void* workspace) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(n_layers >= 1 && n_layers <= DBN_MAX_LAYERS, "bad layer count");
  MlpWorkspace ws;
  size_t bytes = mlp_ws_layout(precision, Bmax, n_in, dims, n_layers, n_out, static_cast<char*>(workspace), &ws);
  DBN_CUDA_OK(cudaMemsetAsync(workspace, 0, bytes, as_stream(stream)));
  if (is_tc(precision)) {
    DBN_TRY(dbn_fill_ones_column(stream, &ws.a0, Bmax, n_in));
    DBN_TRY(dbn_fill_ones_column(stream, &ws.a0b, Bmax, n_in));
    for (int l = 0; l < n_layers; ++l) DBN_TRY(dbn_fill_ones_column(stream, &ws.act[l], Bmax, dims[l]));
  }
  return DBN_OK;
}

int dbn_mlp_train_step(void* stream, const dbn_mlp_step_args* a) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(a != nullptr, "null args");
  const int L = a->n_layers, B = a->B;
  DBN_REQUIRE(L >= 1 && L <= DBN_MAX_LAYERS && B > 0 && a->n_in > 0 && a->n_out > 0, "bad shape");
  DBN_REQUIRE(a->act == DBN_ACT_SIGMOID || a->act == DBN_ACT_RELU, "Invalid activation function.");
  const int ws_rows = a->ws_rows > 0 ? a->ws_rows : B;   // the layout (and its ones columns) is fixed by this
  DBN_REQUIRE(ws_rows >= B, "ws_rows smaller than the mini-batch");
  DBN_REQUIRE(a->workspace_bytes >= dbn_mlp_workspace_bytes(a->precision, ws_rows, a->n_in, a->dims, L, a->n_out),
              "workspace too small");
  const bool tc = is_tc(a->precision);
  const bool dropout = a->keep_p < 1.0f;
  const bool raw = a->raw_grads[0].ptr != nullptr;
  cudaStream_t st = as_stream(stream);
  MlpWorkspace ws;
  mlp_ws_layout(a->precision, ws_rows, a->n_in, a->dims, L, a->n_out, static_cast<char*>(a->workspace), &ws);
  auto Wop = [&](int l) { return tc ? a->Wb[l] : make_mat(a->W[l], a->ldw[l], DBN_F32); };
  const int prime = a->act == DBN_ACT_SIGMOID ? DBN_MUL_PRIME_SIGMOID : DBN_MUL_PRIME_RELU;

  // ---- forward (dbn/models.py:394-417) ----
  dbn_mat in0 = a->X;
  int in0_row0 = a->row0;
  if (dropout) {  // input_data *= Binomial(1, p): dbn/models.py:401-403
    dbn_rng r = a->rng;
    r.mode = DBN_RNG_MASK; r.keep_p = a->keep_p; r.stream = a->rng.stream + 0u; r.injected = a->masks[0];
    dbn_mat src = a->X;
    src.ptr = static_cast<char*>(src.ptr) + (size_t)a->row0 * src.ld * elt_size(src.dtype);
    gather_rows_kernel<<<min(B, 1184), 256, 0, st>>>(src, nullptr, B, a->n_in, ws.a0, -1, r);
    DBN_LAUNCH_CHECK();
    in0 = ws.a0;
    in0_row0 = 0;
  }
  for (int l = 0; l < L; ++l) {
    const int K = l == 0 ? a->n_in : a->dims[l - 1];
    dbn_operands o = make_ops(l == 0 ? in0 : ws.act[l - 1], 0, l == 0 ? in0_row0 : 0, Wop(l), 0, 0, K);
    dbn_act_epilogue e = blank_act();
    e.bias = a->c[l]; e.act = a->act; e.out = ws.act[l];
    if (dropout) {  // dbn/models.py:408-411
      e.rng = a->rng;
      e.rng.mode = DBN_RNG_MASK; e.rng.keep_p = a->keep_p; e.rng.stream = a->rng.stream + (uint32_t)(l + 1);
      e.rng.injected = a->masks[l + 1];
    }
    DBN_TRY(dbn_gemm_act(stream, a->precision, B, a->dims[l], &o, 1, &e));
  }
  const int HL = a->dims[L - 1];
  const dbn_mat WoM = make_mat(a->Wo, a->ldwo, DBN_F32);
  bool head_did_delta = false;
  if (a->n_out <= 32) {  // fused head: scores + softmax/identity + delta + loss in one launch
    // ... and, in the same launch, the delta of the last hidden layer (dbn/models.py:495-499)
    DBN_CUDA_OK(launch_pdl(head_fused_kernel, dim3(ceil_div(B, 2)), dim3(256), 0, st, ws.act[L - 1], HL,
                           (const float*)a->Wo, a->ldwo, (const float*)a->bo, B, a->n_out, a->head,
                           a->Y + (size_t)a->row0 * a->ldy, a->ldy, static_cast<float*>(ws.Z.ptr), ws.Z.ld,
                           static_cast<float*>(ws.dO.ptr), ws.dO.ld, ws.dOb, a->loss_rows, ws.delta[L - 1], prime));
    DBN_LAUNCH_CHECK();
    head_did_delta = true;
  } else {
    // head scores (dbn/models.py:601 / :703), strict FP32
    dbn_operands o = make_ops(ws.act[L - 1], 0, 0, WoM, 0, 0, HL);
    dbn_act_epilogue e = blank_act();
    e.bias = a->bo; e.act = DBN_ACT_IDENTITY; e.out = ws.Z;
    DBN_TRY(simt_gemm_act(st, B, a->n_out, &o, 1, e));
    DBN_TRY(dbn_head_delta(stream, a->head, static_cast<float*>(ws.Z.ptr), ws.Z.ld, a->Y + (size_t)a->row0 * a->ldy,
                           a->ldy, B, a->n_out, static_cast<float*>(ws.dO.ptr), ws.dO.ld, a->loss_rows));
  }
  if (a->out.ptr) DBN_TRY(copy2d(st, a->out, ws.Z, B, a->n_out));

  // ---- backward deltas (dbn/models.py:489-501) and gradients + decayed SGD update (:508-513, :454-465) ----
  // delta_l needs W_{l+1} as it was in the forward pass; the update of W_{l+1} needs delta_{l+1} and a_l.  So
  //     delta_{L-1};  then for l = L-1 .. 1:   { delta_{l-1} (reads W_l) }  ||  { update of the layer above };  update W_0
  // i.e. the update of layer l+1 may run NEXT TO the delta contraction of layer l-1... precisely: update(head) after
  // delta_{L-1}; update(W_l) after delta_{l-1}.  With an auxiliary stream the updates go there (a second branch of the
  // captured graph): the critical path is forward + deltas + the last update instead of forward + deltas + all updates.
  cudaStream_t aux = a->aux_stream ? as_stream(a->aux_stream) : st;
  void* auxp = a->aux_stream ? a->aux_stream : stream;
  auto update_head = [&]() -> int {
    if (tc && a->n_out <= 32) {
      // narrow head in a tensor-core mode: the same tcgen05 contraction as the hidden layers (M = n_out rows of a
      // 128-row tile), operands = 16-bit copy of the head delta and the last activations with their ones column
      dbn_operands o = make_ops(ws.dOb, 1, 0, ws.act[L - 1], 1, 0, B);
      dbn_update_epilogue u = blank_upd();
      u.W = a->Wo; u.ldw = a->ldwo; u.vecM = a->bo; u.decay = a->decay; u.scale = -a->lr_over_bs;
      if (raw) u.raw = a->raw_grads[L];
      return dbn_gemm_update(auxp, a->precision, a->n_out, HL, 0, 1, &o, 1, &u);
    }
    if (a->n_out <= 32 && (size_t)B * a->n_out * sizeof(float) <= 48 * 1024) {   // narrow head, FP32 mode
      float* rawp = raw ? static_cast<float*>(a->raw_grads[L].ptr) : nullptr;
      DBN_CUDA_OK(launch_pdl(head_grad_kernel, dim3(ceil_div(HL + 1, 128)), dim3(128), (size_t)B * a->n_out * sizeof(float), aux,
                             ws.act[L - 1], HL, (const float*)ws.dO.ptr, ws.dO.ld, B, a->n_out, a->Wo, a->ldwo, a->bo, a->decay,
                             -a->lr_over_bs, rawp, raw ? a->raw_grads[L].ld : 0));
      DBN_LAUNCH_CHECK();
      return DBN_OK;
    }
    dbn_operands o = make_ops(ws.dO, 1, 0, ws.act[L - 1], 1, 0, B);
    dbn_update_epilogue u = blank_upd();
    u.W = a->Wo; u.ldw = a->ldwo; u.vecM = a->bo; u.decay = a->decay; u.scale = -a->lr_over_bs;
    if (raw) u.raw = a->raw_grads[L];
    return simt_gemm_update(aux, a->n_out, HL, 0, 1, &o, 1, u);
  };
  auto update_layer = [&](int l, void* on) -> int {
    const int K_in = l == 0 ? a->n_in : a->dims[l - 1];
    dbn_operands o = make_ops(ws.delta[l], 1, 0, l == 0 ? in0 : ws.act[l - 1], 1, l == 0 ? in0_row0 : 0, B);
    dbn_update_epilogue u = blank_upd();
    u.W = a->W[l]; u.ldw = a->ldw[l]; u.Wb = tc ? a->Wb[l] : null_mat();
    u.vecM = a->c[l]; u.decay = a->decay; u.scale = -a->lr_over_bs;
    if (raw) u.raw = a->raw_grads[l];
    return dbn_gemm_update(on, a->precision, a->dims[l], K_in, 0, 1, &o, 1, &u);
  };
  const bool forked = aux != st;
  if (!head_did_delta) {  // delta_{L-1} = (delta_out W_o) * f'(a_{L-1})   (wide heads; narrow ones did it in the head kernel)
    dbn_operands o = make_ops(ws.dO, 0, 0, WoM, 1, 0, a->n_out);
    dbn_act_epilogue e = blank_act();
    e.act = DBN_ACT_IDENTITY; e.mul = prime; e.aux = ws.act[L - 1]; e.out = ws.delta[L - 1];
    DBN_TRY(simt_gemm_act(st, B, HL, &o, 1, e));
  }
  if (forked) DBN_TRY(stream_edge(st, aux));   // the head update reads delta_out (done) and rewrites W_o (delta_{L-1} has read it)
  DBN_TRY(update_head());
  for (int l = L - 2; l >= 0; --l) {
    dbn_operands o = make_ops(ws.delta[l + 1], 0, 0, Wop(l + 1), 1, 0, a->dims[l + 1]);
    dbn_act_epilogue e = blank_act();
    e.act = DBN_ACT_IDENTITY; e.mul = prime; e.aux = ws.act[l]; e.out = ws.delta[l];
    DBN_TRY(dbn_gemm_act(stream, a->precision, B, a->dims[l], &o, 1, &e));
    if (forked) DBN_TRY(stream_edge(st, aux));  // W_{l+1} has been read by delta_l: its update may start on the side
    DBN_TRY(update_layer(l + 1, auxp));
  }
  DBN_TRY(update_layer(0, stream));             // the last update stays on the main stream (nothing left to overlap)
  if (forked) DBN_TRY(stream_edge(aux, st));    // join: the next step's forward pass reads every updated layer
  return DBN_OK;
}

// Whole iteration (every mini-batch of one pass) in ONE cooperative launch of mlp_tc_epoch_kernel when the network fits its
// phase plan (tensor-core modes, <= 4 hidden layers, head <= 16 wide, mini-batches of 16..256 rows, no optional outputs;
// dropout by Philox or injected masks included); *done = 0: not applicable, the caller runs the tiled path.
static int mlp_fit_epoch_tc(cudaStream_t st, const dbn_mlp_step_args* t, int n_rows, int batch_size, int* done) {
  *done = 0;
  const int L = t->n_layers;
  const int B = t->ws_rows > 0 ? t->ws_rows : std::min(batch_size, n_rows);
  if (mt_enabled_flag() != 1 || !is_tc(t->precision) || n_rows <= 0 || L < 1 || L > MT_MAX_LAYERS) return DBN_OK;
  if (B < std::min(batch_size, n_rows) || (n_rows > B && B != batch_size)) return DBN_OK;
  if (t->out.ptr || t->raw_grads[0].ptr) return DBN_OK;
  if (t->act != DBN_ACT_SIGMOID && t->act != DBN_ACT_RELU) return DBN_OK;
  if (t->head != DBN_HEAD_SOFTMAX && t->head != DBN_HEAD_LINEAR) return DBN_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return DBN_OK;
  }
  MtPlan pl;
  if (!mt_plan(t->n_in, t->dims, L, t->n_out, B, &pl)) return DBN_OK;
  if (pl.grid > mt_max_coresident()) return DBN_OK;
  if (t->workspace_bytes < dbn_mlp_workspace_bytes(t->precision, B, t->n_in, t->dims, L, t->n_out)) return DBN_OK;
  if (!t->Y || !t->Wo || !t->bo || t->ldwo != t->dims[L - 1] || (reinterpret_cast<uintptr_t>(t->Wo) & 15) != 0) return DBN_OK;
  if (!umma_operand_ok(t->X, t->n_in)) return DBN_OK;
  for (int l = 0; l < L; ++l)
    if (!umma_operand_ok(t->Wb[l], l == 0 ? t->n_in : t->dims[l - 1]) || t->Wb[l].dtype != t->X.dtype || !t->W[l] || !t->c[l]) return DBN_OK;
  MlpWorkspace ws;
  mlp_ws_layout(t->precision, B, t->n_in, t->dims, L, t->n_out, static_cast<char*>(t->workspace), &ws);
  if (pl.scratch_bytes > 0 && ws.scratch == nullptr) return DBN_OK;
  std::unique_ptr<MtArgs> args(new MtArgs);   // (7 KB; the launch copies it.  Not a static: two host threads may fit two models)
  MtArgs& a = *args;
  memset(&a, 0, sizeof(a));
  memcpy(a.ph, pl.ph, sizeof(a.ph));
  a.n_phases = pl.n_phases;
  a.L = L; a.n_in = t->n_in; a.n_out = t->n_out; a.B = B; a.act = t->act; a.head = t->head; a.dtype = t->X.dtype;
  a.n_rows = n_rows; a.row0 = t->row0; a.scale = -t->lr_over_bs; a.decay = t->decay;
  a.X = t->X; a.dOb = ws.dOb;
  for (int l = 0; l < L; ++l) {
    a.dims[l] = t->dims[l];
    a.a[l] = ws.act[l]; a.d[l] = ws.delta[l]; a.Wb[l] = t->Wb[l];
    a.W[l] = t->W[l]; a.ldw[l] = t->ldw[l]; a.c[l] = t->c[l];
  }
  a.Wo = t->Wo; a.ldwo = t->ldwo; a.bo = t->bo; a.Y = t->Y; a.ldy = t->ldy; a.loss_rows = t->loss_rows;
  a.off_bars = pl.off_bars; a.smem_bytes = pl.smem_bytes; a.wo_bytes = pl.wo_bytes;
  a.ctl = ws.ctl; a.scratch = ws.scratch; a.trace = umma_trace_ptr();
  const uint64_t xrows = (uint64_t)t->row0 + (uint64_t)n_rows;
  for (int l = 0; l < L; ++l) {
    const dbn_mat in = l == 0 ? t->X : ws.act[l - 1];
    const uint64_t in_rows = l == 0 ? xrows : (uint64_t)B;
    const int Kin = l == 0 ? t->n_in : t->dims[l - 1];
    const int ks = ceil_div(Kin, 64);
    DBN_TRY(make_operand_map(&a.tm_in3[l], in, in_rows, ks, pl.f_rb[l], ks));
    DBN_TRY(make_operand_map(&a.tm_w3[l], t->Wb[l], t->dims[l], ks, pl.f_w[l], ks));
    DBN_TRY(te_make_map2(&a.tm_in2[l], in, in.ld, in_rows, 256));
    DBN_TRY(te_make_map2(&a.tm_d2[l], ws.delta[l], ws.delta[l].ld, B, 256));
    if (l >= 1) {   // delta_l and W_l feed the contraction that produces delta_{l-1}
      DBN_TRY(make_operand_map(&a.tm_d3[l], ws.delta[l], B, ceil_div(t->dims[l], 64), pl.b_rb[l - 1], pl.b_kc[l - 1]));
      DBN_TRY(te_make_map2(&a.tm_w2[l], t->Wb[l], t->Wb[l].ld, t->dims[l], std::min(pl.b_kc[l - 1] * 64, 256)));
    }
  }
  if (t->keep_p < 1.0f) {   // dropout: layer 0 reads the masked copy of its mini-batch (double-buffered, written one phase ahead)
    a.drop = 1;
    a.keep_p = t->keep_p;
    a.rng = t->rng;
    for (int j = 0; j <= L; ++j) a.masks[j] = t->masks[j];
    a.a0[0] = ws.a0;
    a.a0[1] = ws.a0b;
    const int ks0 = ceil_div(t->n_in, 64);
    for (int i = 0; i < 2; ++i) {
      DBN_TRY(make_operand_map(&a.tm_a0_3[i], a.a0[i], B, ks0, pl.f_rb[0], ks0));
      DBN_TRY(te_make_map2(&a.tm_a0_2[i], a.a0[i], a.a0[i].ld, B, 256));
    }
  }
  DBN_TRY(te_make_map2(&a.tm_in2[L], ws.act[L - 1], ws.act[L - 1].ld, B, 256));
  DBN_TRY(te_make_map2(&a.tm_d2[L], ws.dOb, ws.dOb.ld, B, 256));
  static std::atomic<uint64_t> attr_set{0};
  if (first_use_on_device(attr_set))
    DBN_CUDA_OK(cudaFuncSetAttribute(mlp_tc_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_MAX));
  DBN_CUDA_OK(cudaMemsetAsync(ws.ctl, 0, sizeof(TeCtl), st));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(TE_THREADS);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, mlp_tc_epoch_kernel, a) != cudaSuccess) {
    (void)cudaGetLastError();   // e.g. the grid cannot be co-resident right now: tiled path
    return DBN_OK;
  }
  DBN_LAUNCH_CHECK();
  *done = 1;
  return DBN_OK;
}

int dbn_mlp_tc_epoch_plan(int precision, int n_in, const int32_t* dims, int n_layers, int n_out, int batch_size) {
  MtPlan pl;
  if (!dims || !is_tc(precision) || mt_enabled_flag() != 1 || !mt_plan(n_in, dims, n_layers, n_out, batch_size, &pl)) return 0;
  return pl.grid;
}
int dbn_set_tc_finetune(int enabled) {
  const int prev = mt_enabled_flag();
  mt_enabled_flag() = enabled ? 1 : 0;
  return prev;
}

int dbn_mlp_fit_epoch(void* stream, const dbn_mlp_step_args* tmpl, int n_rows, int batch_size) {
  DBN_REQUIRE(tmpl && n_rows >= 0 && batch_size > 0, "bad argument");
  {
    DBN_NEED_DEVICE();
    int done = 0;
    DBN_TRY(mlp_fit_epoch_tc(as_stream(stream), tmpl, n_rows, batch_size, &done));
    if (done) return DBN_OK;
  }
  for (int s = 0; s < n_rows; s += batch_size) {
    dbn_mlp_step_args a = *tmpl;
    a.B = std::min(batch_size, n_rows - s);
    if (a.ws_rows == 0) a.ws_rows = std::min(batch_size, n_rows);
    a.row0 = tmpl->row0 + s;
    a.rng.row0 = tmpl->rng.row0 + (uint32_t)s;
    if (a.loss_rows) a.loss_rows += s;
    for (int j = 0; j <= a.n_layers; ++j)
      if (a.masks[j].ptr)
        a.masks[j].ptr = static_cast<char*>(a.masks[j].ptr) + (size_t)s * a.masks[j].ld * sizeof(float);
    DBN_TRY(dbn_mlp_train_step(stream, &a));
  }
  return DBN_OK;
}

// ------------------------------------------------------------------------------------------------
int dbn_gather_rows(void* stream, const dbn_mat* src, const int32_t* idx, int n_rows, int n_cols, const dbn_mat* dst,
                    int ones_col, const dbn_rng* input_dropout) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(src && dst && src->ptr && dst->ptr, "null argument");
  if (n_rows <= 0) return DBN_OK;
  dbn_rng r;
  memset(&r, 0, sizeof(r));
  if (input_dropout) r = *input_dropout;
  gather_rows_kernel<<<std::min(n_rows, 148 * 16), 256, 0, as_stream(stream)>>>(*src, idx, n_rows, n_cols, *dst,
                                                                              ones_col, r);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_scale_cast(void* stream, const float* src, int lds, int rows, int cols, float scale, const dbn_mat* dst) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(src && dst && dst->ptr, "null argument");
  const long total = (long)rows * ((cols + 3) / 4);
  if (total <= 0) return DBN_OK;
  const int blocks = (int)std::min<long>((total + 255) / 256, 148L * 8);
  scale_cast_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, lds, rows, cols, scale, *dst);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_fill_ones_column(void* stream, const dbn_mat* m, int rows, int col) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(m && m->ptr && col < m->ld, "bad matrix");
  fill_ones_column_kernel<<<ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(*m, rows, col);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_sum_sq_diff(void* stream, const dbn_mat* A, const dbn_mat* B, int row0_B, int rows, int cols, float inv_denom,
                    float* out) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(A && B && out, "null argument");
  const long total = (long)rows * ((cols + 3) / 4);
  if (total <= 0) return DBN_OK;
  const int blocks = (int)std::min<long>((total + 255) / 256, 148L * 4);
  sum_sq_diff_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*A, *B, row0_B, rows, cols, inv_denom, out);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

int dbn_head_delta(void* stream, int head, float* Z, int ldz, const float* Y, int ldy, int B, int C, float* delta,
                   int ldd, float* loss_rows) {
  DBN_NEED_DEVICE();
  DBN_REQUIRE(Z && B > 0 && C > 0, "bad argument");
  head_delta_kernel<<<ceil_div(B * 32, 128), 128, 0, as_stream(stream)>>>(head, Z, ldz, Y, ldy, B, C, delta, ldd,
                                                                        loss_rows);
  DBN_LAUNCH_CHECK();
  return DBN_OK;
}

}  // extern "C"
