/*
 * dbn_b200.h -- C-ABI of the B200-native CD-k / backprop hot path.
 *
 * Drop-in boundary for the training hot path of albertbup/deep-belief-network (reference 1.0.5).
 * The reference is pure Python; its "backend seam" is the set of methods a backend module
 * overrides (see dbn/tensorflow/models.py:78-95, :199-236 and the abstract hooks at
 * dbn/models.py:364-382).  Every entry point below names the reference function(s) it replaces
 * (paths relative to the reference root).  The Python estimators in
 * deep-belief-network_b200/models.py are the only callers; INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - Every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - Matrices are row-major; `ld` is the leading dimension in ELEMENTS.
 *   - All calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate
 *     device memory and never synchronise, unless stated.  The caller owns every buffer.
 *   - Return value: 0 = ok, non-zero = error (DBN_ERR_*); dbn_last_error() returns the message of
 *     the calling thread's last failure.  Nothing throws across the boundary.
 *   - There is no CPU fallback: on a machine without an sm_100 device every compute call returns
 *     DBN_ERR_NO_DEVICE.
 */
#ifndef DBN_B200_H
#define DBN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBN_B200_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------------- */
enum {
  DBN_OK = 0,
  DBN_ERR_INVALID_ARG = 1,
  DBN_ERR_CUDA = 2,
  DBN_ERR_NO_DEVICE = 3,
  DBN_ERR_UNSUPPORTED = 4
};

/* ---- enums ------------------------------------------------------------------------------ */
/* activation_function strings of the reference: 'sigmoid' / 'relu' (dbn/models.py:57-69;
 * dbn/activations.py:23-38, :43-58).  IDENTITY is the linear regression head (dbn/models.py:696). */
enum { DBN_ACT_IDENTITY = 0, DBN_ACT_SIGMOID = 1, DBN_ACT_RELU = 2 };

/* element types of a dbn_mat (the 16-bit type of a tensor-core operand follows the precision mode) */
enum { DBN_F32 = 0, DBN_BF16 = 1, DBN_F16 = 2 };

/* arithmetic of the contraction: strict FP32 FFMA (1e-5 parity mode) or tcgen05 with 16-bit operands and FP32
 * accumulation in TMEM (2e-3 parity mode).  The two tensor-core modes differ only in the element type of EVERY
 * operand buffer (weight shadow, staged inputs, workspace): bf16 has fp32's range (unbounded inputs, relu outputs,
 * back-propagated deltas), fp16 has 11 significant bits instead of 8 -- the better choice where every operand is a
 * probability, a binary state or a weight, i.e. sigmoid RBM pre-training: weight quantisation is what bounds the
 * parity of the 16-bit modes against float64 weights (profiles/r02_bf16_error_report.txt).  The hardware takes ONE
 * format per MMA (mixing A = bf16 with B = fp16 traps as an illegal instruction on sm_100a), hence a mode, not a
 * per-matrix choice.  Conversions to fp16 saturate at +-65504. */
enum { DBN_PREC_FP32 = 0, DBN_PREC_BF16 = 1, DBN_PREC_F16 = 2 };

/* multiplicative term of the activation epilogue: the activation derivative evaluated on the
 * (post-dropout) activation, dbn/activations.py:38 `a*(1-a)` and :58 `(a>0)`. */
enum { DBN_MUL_NONE = 0, DBN_MUL_PRIME_SIGMOID = 1, DBN_MUL_PRIME_RELU = 2 };

/* random use in the activation epilogue */
enum {
  DBN_RNG_NONE = 0,
  DBN_RNG_SAMPLE = 1, /* s = (u < p), strict '<'   -- dbn/models.py:155                 */
  DBN_RNG_MASK = 2    /* a *= Bernoulli(keep_p), no 1/p rescale -- dbn/models.py:401-411 */
};

/* output head: softmax + cross-entropy (dbn/models.py:594-626) or linear + squared error
 * (dbn/models.py:696-720) */
enum { DBN_HEAD_SOFTMAX = 0, DBN_HEAD_LINEAR = 1 };

#define DBN_MAX_PEERS 8 /* GPUs of one NVSwitch box */

/* ---- plain structs ------------------------------------------------------------------------ */
typedef struct dbn_mat {
  void* ptr;     /* device pointer (may be NULL = absent) */
  int32_t ld;    /* leading dimension, elements           */
  int32_t dtype; /* DBN_F32 / DBN_BF16                    */
} dbn_mat;

/* Counter-based randomness.  Uniforms are Philox4x32-10 with key = seed and counter =
 * (row0 + m, (col0 + n) / 4, *epoch_ptr + epoch, stream): a draw depends only on the GLOBAL row index of
 * the sample inside the epoch, the unit index, the epoch and the purpose -- not on the batch
 * split or the number of GPUs.  u = (x >> 8) * 2^-24, so every uniform is exact in fp32.
 * If `injected.ptr` is non-NULL the values are read from it instead (fp32, [M, ld]): uniforms for
 * DBN_RNG_SAMPLE, {0,1} keep-masks for DBN_RNG_MASK.  This is how the parity tests feed the
 * reference's draws (np.random.random_sample / np.random.binomial) to the kernels. */
typedef struct dbn_rng {
  int32_t mode;              /* DBN_RNG_*                         */
  float keep_p;              /* DBN_RNG_MASK: p = 1 - dropout_p   */
  uint64_t seed;
  uint32_t stream;           /* purpose id (Gibbs step / layer)   */
  uint32_t epoch;            /* added to *epoch_ptr               */
  const uint32_t* epoch_ptr; /* device counter or NULL            */
  uint32_t row0;             /* global row index of row m = 0     */
  uint32_t col0;             /* global unit index of column n = 0 (multiple of 4): column blocks of one layer */
  dbn_mat injected;
} dbn_rng;

/* One operand pair of a contraction  acc[m,n] += sign * sum_k A(m,k) * B(n,k).
 * transX = 0: X is stored [rows = M or N, cols = K] (K contiguous);
 * transX = 1: X is stored [rows = K, cols = M or N] (M/N contiguous).
 * row0_X offsets the stored ROW index (slices of an epoch-sized array without pointer math that
 * would break TMA alignment). */
typedef struct dbn_operands {
  dbn_mat A;
  dbn_mat B;
  int32_t transA, transB;
  int32_t K;
  int32_t row0_A, row0_B;
  int32_t _pad;
} dbn_operands;

/* Column statistics of an activation epilogue: acc[n] += round(2^32 * sign * x[m,n]) summed over the rows m of
 * the launch, where x = a[m,n] (the epilogue's output value) or, with `ref` set, x = ref[ref_row0 + m, n] - a[m,n].
 * This is how the bias statistics of a CD step are taken where the values already sit in registers --
 * dc = sum_rows(h_0 - h_k), db = sum_rows(v_0 - v_k) (dbn/models.py:143-144 summed by :115-116) -- instead of
 * through an extra ones row / column of the dW contraction.  The accumulators are 64-bit FIXED POINT (2^-32):
 * integer addition is associative, so the result does not depend on the order in which tiles, warps or GPUs
 * contribute (deterministic, and partial sums of several GPUs add exactly). */
#define DBN_STAT_ONE 4294967296.0 /* 2^32 */
typedef struct dbn_colsum {
  unsigned long long* acc; /* [N] fixed-point accumulators, or NULL = off */
  float sign;              /* +1 / -1                                     */
  int32_t ref_row0;
  dbn_mat ref;             /* optional [rows, ld] matrix (16-bit / fp32)  */
  int32_t square;          /* != 0: accumulate x^2 instead of sign * x -- with `ref` the squared reconstruction distance
                              sum_rows (v - v')^2 of dbn/models.py:212-220, taken in the epilogue that produces v' */
  int32_t _pad;
} dbn_colsum;

/* Activation-type epilogue:
 *     x = acc + bias[n];  a = act(x);  a *= prime(aux[m,n]);  a *= mask;   out = a
 *     sample = (u < a)
 * Replaces the element-wise tails of _compute_hidden_units_matrix (dbn/models.py:176-183),
 * _compute_visible_units_matrix (:195-201), _sample_hidden_units (:148-155), the dropout lines
 * of _compute_activations (:401-411) and the delta line of _backpropagation (:498-499). */
typedef struct dbn_act_epilogue {
  const float* bias; /* [N] or NULL */
  int32_t act;       /* DBN_ACT_*   */
  int32_t mul;       /* DBN_MUL_*   */
  dbn_mat aux;       /* activation a_l for DBN_MUL_PRIME_* */
  dbn_rng rng;
  dbn_mat out;       /* value, may be NULL          */
  dbn_mat out2;      /* second copy (other dtype), may be NULL */
  dbn_mat sample;    /* {0,1} sample, may be NULL   */
  dbn_colsum colsum; /* column statistics of `out`, off when colsum.acc == NULL */
} dbn_act_epilogue;

/* Update-type epilogue over an (M+augM) x (N+augN) accumulator whose optional extra row / column
 * hold the contraction with a vector of ones (bias statistics):
 *     W[m,n]  = decay * W[m,n] + scale * acc[m,n]          (m < M, n < N)
 *     vecN[n] = vecN[n]        + scale * acc[M,n]          (augM: visible-bias statistics)
 *     vecM[m] = vecM[m]        + scale * acc[m,N]          (augN: hidden-bias statistics)
 * or, if `raw.ptr` != NULL, acc is written unmodified to raw [(M+augM), ld] and nothing is updated
 * (parity tests, multi-GPU all-reduce).
 * Replaces dbn/models.py:114-119 (RBM accumulate + update) and :445-465 (fine-tune accumulate,
 * L2 decay and SGD update). */
typedef struct dbn_update_epilogue {
  float* W;      /* fp32 master [M, ldw]  */
  int32_t ldw;
  int32_t _pad;
  dbn_mat Wb;    /* bf16 shadow of W (tensor-core operand), may be NULL */
  float* vecM;   /* [M] or NULL */
  float* vecN;   /* [N] or NULL */
  float decay;
  float scale;
  dbn_mat raw;   /* fp32, may be NULL */
  /* Bias statistics taken by dbn_colsum instead of the ones row / column (augM = augN = 0 then): the thread
   * that owns row m / column n applies vecM[m] += scale * statM[m] * 2^-32 and clears statM[m] (likewise N). */
  unsigned long long* statM; /* [M] or NULL */
  unsigned long long* statN; /* [N] or NULL */
  /* Scatter of the raw accumulator over peer GPUs (data-parallel reduce-scatter fused into the epilogue):
   * if raw_peers[0] != NULL, accumulator row m goes to the GPU that owns it, o = m / rows_per_peer, at
   *   (float*)raw_peers[o] + (raw_slot * rows_per_peer + m - o * rows_per_peer) * raw.ld + n
   * i.e. into slot `raw_slot` (= the sender's rank) of the owner's receive buffer [n_peers][rows_per_peer][raw.ld].
   * raw_peers[] are device pointers valid on THIS device (peer mappings, or local memory for the owner itself). */
  void* raw_peers[DBN_MAX_PEERS];
  int32_t rows_per_peer;
  int32_t raw_slot;
} dbn_update_epilogue;

/* ---- library / device ------------------------------------------------------------------- */
int dbn_abi_version(void);
const char* dbn_last_error(void);
/* Number of sm_100 devices visible; 0 on a CPU-only box (no CUDA call fails). */
int dbn_device_count(void);
/* Counts kernels launched by this library since load (bench.py's `gpu_launches`). */
uint64_t dbn_launch_count(void);

/* CTA-pair contraction (4096-wide layers): when the last wave of 256 x 256 tiles fills at most half of the SM pairs
 * (256 tiles on 74 pairs: 3 waves and one 46 % full), each of its tiles is contracted by two pairs, half of K each, and
 * the halves are summed before the fused epilogue (csrc/umma_gemm.cuh, ts_work).  dbn_set_tail_split(0) contracts every
 * tile whole (also DBN_B200_TAIL_SPLIT=0); returns the previous setting.  dbn_tail_split_count: launches that split. */
int dbn_set_tail_split(int enabled);
uint64_t dbn_tail_split_count(void);
/* Keep n_sms SMs free of the persistent contraction kernels (0 = use all; default).  Data-parallel
 * callers set this so the collective's kernels can run next to the contraction they overlap with. */
int dbn_set_sm_reserve(int n_sms);
/* Debug: when device_buf != NULL every single-CTA tcgen05 contraction writes 16 int64 phase timestamps
 * per CTA to device_buf[blockIdx.x * 16 ...]: [0] entry (globaltimer ns); [1..7] clock64 at entry /
 * setup done / first operands / mainloop issued / accumulator ready / epilogue done / exit; [8] ns when
 * the programmatic-dependency wait returned; [9] ns at exit.  NULL switches it off (default). */
int dbn_debug_set_trace(void* device_buf);
/* sizeof() of the structs above as compiled (0 dbn_mat, 1 dbn_rng, 2 dbn_operands, 3 dbn_act_epilogue,
 * 4 dbn_update_epilogue, 5 dbn_rbm_step_args, 6 dbn_mlp_step_args, 7 dbn_colsum, 8 dbn_dp_exchange): lets a
 * binding check its layout. */
size_t dbn_struct_size(int which);

/* The uniforms the kernels draw (see dbn_rng) computed on the HOST into out_host[rows*cols]: lets a
 * caller or a test reproduce the device's Philox stream without a GPU. */
int dbn_philox_uniforms_host(uint64_t seed, uint32_t stream, uint32_t epoch, uint32_t row0, int rows,
                             int cols, float* out_host);

/* ---- fused contractions (building blocks of every path-level call) ----------------------- */
/* acc = ops[0] - ops[1] (n_ops = 1 or 2), M x N output, activation epilogue.
 * precision = DBN_PREC_FP32: operands fp32, SIMT FFMA kernel.
 * precision = DBN_PREC_BF16: operands bf16, TMA + tcgen05.mma kernel.  Operand contract: base 128-byte
 *   aligned, ld a multiple of 64 elements and >= the extent rounded up to 64; the columns between an
 *   operand's K extent and the next multiple of 64 are read, so they must be finite in both operands of a
 *   pair and zero in at least one (the engine zero-initialises every staged buffer and weight shadow;
 *   only the ones column lives there).  Anything else runs on the FP32 SIMT kernel. */
int dbn_gemm_act(void* stream, int precision, int M, int N, const dbn_operands* ops, int n_ops,
                 const dbn_act_epilogue* epi);

/* Same contraction with the update epilogue.  augM / augN (0/1) append the ones row / column:
 * for DBN_PREC_FP32 they are virtual; for DBN_PREC_BF16 the caller's operand buffers must hold
 * 1.0 in column M of A-storage / column N of B-storage (see dbn_fill_ones_column). */
int dbn_gemm_update(void* stream, int precision, int M, int N, int augM, int augN,
                    const dbn_operands* ops, int n_ops, const dbn_update_epilogue* epi);

/* ---- RBM: BinaryRBM hot path ---------------------------------------------------------------- */
/* _compute_reconstruction_error (dbn/models.py:212-220) for a block of rows without materialising the reconstruction:
 * v' = act(Hm W + b) and acc[j] += sum_rows (X[row0 + m, j] - v'[m, j])^2 (fixed point, see dbn_colsum) in the epilogue of
 * the contraction.  Hm: hidden means of the same rows.  The caller sums acc[0..V) and divides by 2^32 and by the number of rows. */
int dbn_rbm_reconstruction_sq(void* stream, int precision, const dbn_mat* Hm, int B, int V, int H, const dbn_mat* W,
                              const float* b, int act, const dbn_mat* X, int row0, unsigned long long* acc);

/* _compute_hidden_units_matrix (dbn/models.py:176-183): out[B,H] = act(X W^T + c). */
int dbn_rbm_hidden_means(void* stream, int precision, const dbn_mat* X, int row0, int B, int V, int H,
                         const dbn_mat* W, const float* c, int act, const dbn_mat* out);
/* _compute_visible_units_matrix (dbn/models.py:195-201): out[B,V] = act(Hm W + b). */
int dbn_rbm_visible_means(void* stream, int precision, const dbn_mat* Hm, int B, int V, int H,
                          const dbn_mat* W, const float* b, int act, const dbn_mat* out);

/* dbn_rbm_step_args.flags */
#define DBN_STEP_SKIP_UPDATE 1 /* run the Gibbs chain only; deltas are produced later by dbn_rbm_delta_rows */
#define DBN_STEP_SKIP_FIRST_HIDDEN 2 /* the t = 0 hidden phase was already run by dbn_rbm_first_hidden */
#define DBN_STEP_BIAS_STATS 4 /* BF16 mode: take dc / db as column statistics in the Gibbs-chain epilogues (left in the
                                 workspace, see dbn_rbm_workspace_stats) instead of the ones row / column of dW.  Without
                                 the flag the library does so by itself when the extra row / column would cost a tile
                                 (H or V a multiple of 64) and the step updates in place. */

/* Workspace of one CD-k step for a batch of up to Bmax rows (bytes). */
size_t dbn_rbm_workspace_bytes(int precision, int Bmax, int V, int H);
/* Zero the workspace and (BF16 mode) write the ones columns the bias statistics ride on.  Call once
 * after allocating the workspace, before the first dbn_rbm_cd_step. */
int dbn_rbm_workspace_init(void* stream, int precision, int Bmax, int V, int H, void* workspace);

typedef struct dbn_rbm_step_args {
  int32_t precision, act;
  int32_t B, V, H, k;      /* B rows in this mini-batch                              */
  float lr_over_bs;        /* learning_rate / configured batch_size (models.py:117-119) */
  int32_t row0;            /* first row of the mini-batch inside X                   */
  dbn_mat X;               /* [rows, ldx] fp32 (FP32) / bf16 with ones column V (BF16) */
  float* W;                /* fp32 master (H, ldw) */
  int32_t ldw;
  int32_t flags;           /* DBN_STEP_* */
  float* b;                /* (V,) */
  float* c;                /* (H,) */
  dbn_mat Wb;              /* bf16 shadow of W (BF16 mode), else NULL               */
  dbn_rng rng;             /* mode ignored; .injected = U[B, k*H] or NULL; stream = Gibbs step is added */
  void* workspace;         /* >= dbn_rbm_workspace_bytes(precision, ws_rows, V, H) */
  size_t workspace_bytes;
  int32_t ws_rows;         /* rows (Bmax) the workspace was sized and initialised for; 0 = B.  The layout of the
                              workspace depends on it, so a short last batch must pass the SAME value. */
  int32_t _pad2;
  /* optional outputs (parity tests): */
  dbn_mat P0;              /* fp32 [B,H]  hidden means at t=0  (h_0, models.py:140)   */
  dbn_mat Vk;              /* fp32 [B,V]                                            */
  dbn_mat Pk;              /* fp32 [B,H]  (h_k, models.py:141)                       */
  dbn_mat H0s;             /* fp32 [B,H]  first hidden sample (models.py:155)        */
  dbn_mat raw_delta;       /* fp32 [(H+1), ld>=V+1]: dW | dc column ; db row. If set, W/b/c are NOT updated */
} dbn_rbm_step_args;

/* One mini-batch of BinaryRBM._stochastic_gradient_descent (dbn/models.py:108-119) ==
 * batched _contrastive_divergence (:124-146) + parameter update. */
int dbn_rbm_cd_step(void* stream, const dbn_rbm_step_args* args);

/* One epoch's mini-batch loop (dbn/models.py:108-119 with dbn/utils.py:14-22): rows
 * [tmpl->row0, tmpl->row0 + n_rows) of tmpl->X in steps of batch_size, last batch short.  The
 * template's optional outputs must be NULL.  Pure launches: capture it in a CUDA graph. */
int dbn_rbm_fit_epoch(void* stream, const dbn_rbm_step_args* tmpl, int n_rows, int batch_size);

/* Tensor-core modes: dbn_rbm_fit_epoch runs the whole epoch as ONE cooperative launch of the persistent tcgen05
 * kernel (csrc/rbm_tc_epoch.cuh: <= 148 co-resident CTAs, a grid barrier between the phases of a mini-batch instead
 * of a launch boundary, the CTA's weight slices resident in shared memory for the step, P0 / samples / V1 / Pk kept
 * in L2) when the layer fits its tile plan -- mini-batches of 16..256 rows, plain in-place training (no optional
 * outputs, no DBN_STEP_SKIP_* flags), stream not being captured -- and as the loop of dbn_rbm_cd_step otherwise.
 * Replaces the whole mini-batch loop of dbn/models.py:105-119.  Returns the number of CTAs the plan uses, 0 when the
 * shape does not fit (or DBN_B200_TC_EPOCH=0): callers skip CUDA-graph capture of the epoch when it is non-zero. */
int dbn_rbm_tc_epoch_plan(int precision, int V, int H, int batch_size);
/* Switch the persistent epoch kernel on / off at run time (default: on unless DBN_B200_TC_EPOCH=0); returns the
 * previous setting.  Parity tests run the same epoch through both paths. */
int dbn_set_tc_epoch(int enabled);
/* The tile plan as numbers (tests, profiles/ notes): out[0..] = CTAs, hidden tile rows, hidden tile units, hidden tiles,
 * visible tile rows, visible tile units, K split, k-slabs per chunk, visible tiles, dW tile width, dW tiles,
 * shared-memory bytes, Vk-columns-fetched-early, P1-in-activation-buffer, split-K scratch bytes, k-slabs of V.
 * Returns the number of values the plan has (0: the shape does not fit). */
int dbn_rbm_tc_epoch_plan_info(int V, int H, int batch_size, int32_t* out, int n);

/* On-chip persistent variant of dbn_rbm_fit_epoch for layers whose weights fit in the shared memory of
 * one thread-block cluster (strict FP32 only): ONE launch trains the whole epoch with W resident in
 * distributed shared memory; hidden units are sharded over the cluster's CTAs.  tmpl->X is the
 * UN-permuted fp32 dataset, idx[n_rows] the epoch order (device), tmpl->rng.injected (optional) the
 * uniforms in epoch order.  dbn_rbm_persistent_plan returns the cluster size it would use, 0 if the
 * shape does not fit (then use dbn_rbm_fit_epoch). */
int dbn_rbm_persistent_plan(int precision, int V, int H, int batch_size);
int dbn_rbm_fit_epoch_persistent(void* stream, const dbn_rbm_step_args* tmpl, const int32_t* idx, int n_rows,
                                 int batch_size);

/* The t = 0 hidden phase (means P0 + first sample) of the step described by args, restricted to hidden
 * units [h0, h1) (h0 % 64 == 0): it needs only rows [h0, h1) of W, so a data-parallel caller can start
 * it as soon as that row block of W has been updated, while later blocks are still being all-reduced.
 * Follow with dbn_rbm_cd_step(flags |= DBN_STEP_SKIP_FIRST_HIDDEN). */
int dbn_rbm_first_hidden(void* stream, const dbn_rbm_step_args* args, int h0, int h1);

/* Rows [h0, h1) of the raw delta (dW | dc) of the chain left in the workspace by a
 * dbn_rbm_cd_step(flags = DBN_STEP_SKIP_UPDATE) with the same args; h1 == H also writes the db row H.
 * h0 % 64 == 0.  Lets the caller pipeline delta production with the all-reduce of earlier rows. */
int dbn_rbm_delta_rows(void* stream, const dbn_rbm_step_args* args, int h0, int h1);

/* Parameter update from a summed raw delta buffer (layout [(H+1), ld] = dW | dc ; db):
 *   rows [h0, h1):  W += s*dW, c += s*dc, bf16 shadow refreshed        (h0 == h1: none)
 *   with_db != 0 :  b += s*db from row H
 * `raw` points at stored row `raw_row0` of that layout, so a reduce-scatter shard (rows [h0, h1) only)
 * is passed with raw_row0 = h0 and a full buffer with raw_row0 = 0. */
int dbn_rbm_apply_delta(void* stream, int V, int H, int h0, int h1, int with_db, float scale, const float* raw,
                        int ld, int raw_row0, float* W, int ldw, float* b, float* c, const dbn_mat* Wb);

/* The fixed-point bias statistics of the step (DBN_STEP_BIAS_STATS): H + V accumulators inside the workspace,
 * dc[0..H) then db[0..V), value = sum * 2^32 as a two's-complement 64-bit integer.  Zero after
 * dbn_rbm_workspace_init; consumed (and cleared) by the in-place update, by the raw-delta output and by
 * dbn_dp_signal. */
unsigned long long* dbn_rbm_workspace_stats(int precision, int Bmax, int V, int H, void* workspace);

/* ---- data-parallel exchange over NVLink peer memory (BF16 mode) ------------------------------------------
 * One process per GPU.  Every rank allocates one peer-visible arena (dbn_peer_alloc), ships its 64-byte handle to
 * the other ranks (any transport: the engine uses torch.distributed), and maps theirs (dbn_peer_open).  A step is
 *   1. dbn_rbm_cd_step(flags = SKIP_UPDATE | BIAS_STATS)      Gibbs chain on the rank's rows of the mini-batch
 *   2. dbn_rbm_delta_scatter                                    dW = P0^T V0 - Pk^T Vk; the epilogue stores every
 *                                                               accumulator tile straight into the receive buffer of the
 *                                                               GPU that owns those rows of W (reduce-scatter fused
 *                                                               into the contraction: the wire time sits under the
 *                                                               mainloop of the following tiles)
 *   3. dbn_dp_signal                                            statistics -> every peer; "my deltas are out" flags
 *   4. dbn_dp_apply                                             owner: ordered sum of the n_peers slots, update of its
 *                                                               rows of the fp32 master, bf16 rows written into EVERY
 *                                                               rank's shadow (all-gather fused into the update);
 *                                                               b, c from the exact integer sum of the statistics
 *   5. dbn_dp_wait (before the next step's first contraction)  every owner's rows have arrived
 * No NCCL on the data path; summation order is fixed (rank 0..n-1), so all replicas hold bit-identical parameters
 * and a run is reproducible.  Replaces the accumulate + update of dbn/models.py:114-119 across GPUs. */
#define DBN_PEER_HANDLE_BYTES 64
/* cudaMalloc'ed, zero-filled device memory that other processes of this box can map; handle_out[64]. */
int dbn_peer_alloc(size_t bytes, void** ptr_out, void* handle_out);
int dbn_peer_free(void* ptr);
/* Map another process's arena from its handle (enables peer access); close when done. */
int dbn_peer_open(const void* handle, void** ptr_out);
int dbn_peer_close(void* ptr);

typedef struct dbn_dp_exchange {
  int32_t n_peers, rank;
  int32_t V, H;
  int32_t rows_per_peer;           /* H / n_peers rows of W owned by each rank                          */
  int32_t ld_recv;                 /* row pitch of the receive buffers in floats (multiple of 4, >= V)  */
  void* recv[DBN_MAX_PEERS];       /* rank j's receive buffer [n_peers][rows_per_peer][ld_recv] fp32    */
  void* stats[DBN_MAX_PEERS];      /* rank j's statistics slots [n_peers][H + V] uint64                 */
  void* flags[DBN_MAX_PEERS];      /* rank j's flag words [2][DBN_MAX_PEERS] uint64: deltas out / weights in */
  dbn_mat Wb[DBN_MAX_PEERS];       /* rank j's bf16 shadow of W                                         */
} dbn_dp_exchange;                 /* every pointer as mapped on THIS device; entry `rank` is local memory */

/* Step 2: rows [0, H) of dW of the chain left in the workspace by dbn_rbm_cd_step, scattered by owner. */
int dbn_rbm_delta_scatter(void* stream, const dbn_rbm_step_args* args, const dbn_dp_exchange* x);
int dbn_dp_signal(void* stream, const dbn_dp_exchange* x, unsigned long long* local_stats, unsigned long long value);
/* `local_stats`: the accumulators handed to dbn_dp_signal (cleared here for the next step).
 * `counter`: one zero-initialised uint32 of local device memory (last-block detection; left at zero). */
int dbn_dp_apply(void* stream, const dbn_dp_exchange* x, float scale, float* W, int ldw, float* b, float* c,
                 unsigned long long* local_stats, unsigned int* counter, unsigned long long value);
/* which = 0: deltas of every peer are out; 1: every owner's weight rows are in. */
int dbn_dp_wait(void* stream, const dbn_dp_exchange* x, int which, unsigned long long value);

/* Epoch shuffle over NVLink: dst[i, :] = row (idx[i] % rows_per_src) of srcs[idx[i] / rows_per_src], bf16 rows of
 * `ld` elements copied whole (values, ones column and padding).  Each rank stages only ITS slice of the dataset
 * (rows [r * rows_per_src, ...)) and the ranks read each other's slices through peer mappings, instead of every
 * rank uploading the whole dataset (dbn/models.py:106-107 + dbn/utils.py:12-13 across GPUs). */
int dbn_gather_rows_peer(void* stream, void* const* srcs, int n_srcs, int rows_per_src, int ld, const int32_t* idx,
                         int n_rows, const dbn_mat* dst);

/* ---- supervised fine-tune: _backpropagation + SGD update ------------------------------------ */
#define DBN_MAX_LAYERS 16

typedef struct dbn_mlp_step_args {
  int32_t precision, act, head;     /* DBN_PREC_*, DBN_ACT_*, DBN_HEAD_*             */
  int32_t n_layers;                 /* hidden layers L                               */
  int32_t B, n_in, n_out;           /* rows in this mini-batch, input width, head width */
  int32_t dims[DBN_MAX_LAYERS];     /* hidden widths H_1..H_L                        */
  float* W[DBN_MAX_LAYERS];         /* fp32 masters (H_l, ldw[l])                    */
  int32_t ldw[DBN_MAX_LAYERS];
  float* c[DBN_MAX_LAYERS];         /* hidden biases                                 */
  dbn_mat Wb[DBN_MAX_LAYERS];       /* bf16 shadows (BF16 mode)                      */
  float* Wo; int32_t ldwo, _pad0;   /* head (n_out, ldwo) fp32                       */
  float* bo;
  dbn_mat X;  int32_t row0, _pad1;  /* inputs [rows, ld] (fp32 / bf16+ones column)   */
  const float* Y; int32_t ldy, _pad2; /* targets in network format [rows, ldy] fp32 (one-hot / regression), row0 applies */
  float lr_over_bs;                 /* learning_rate / batch_size  (models.py:458-465) */
  float decay;                      /* 1 - lr*l2/N                 (models.py:456-457) */
  float keep_p;                     /* 1 - dropout_p; 1 = no dropout                 */
  int32_t _pad3;
  dbn_rng rng;                      /* dropout: .injected unused here; see masks[]   */
  dbn_mat masks[DBN_MAX_LAYERS + 1];/* injected keep-masks fp32 (input, H_1..H_L) or NULL ptr -> Philox */
  void* workspace; size_t workspace_bytes;
  int32_t ws_rows, _pad4;           /* rows (Bmax) the workspace was sized / initialised for; 0 = B */
  float* loss_rows;                 /* optional [B]: -log p[label] or sum_t (pred-y)^2 */
  dbn_mat out;                      /* optional fp32 [B, n_out] head outputs          */
  dbn_mat raw_grads[DBN_MAX_LAYERS + 1]; /* optional fp32 [(out_l), in_l + 1] raw gradient | bias column; if raw_grads[0].ptr set nothing is updated */
  void* aux_stream;                 /* optional second cudaStream_t: the weight-gradient / update contraction of layer l+1
                                       then runs NEXT TO the delta contraction of layer l instead of after it (they share
                                       delta_{l+1} and nothing else; fork / join by events, also inside a stream capture, where
                                       it becomes two branches of the graph).  NULL = one stream, strictly serial. */
} dbn_mlp_step_args;

size_t dbn_mlp_workspace_bytes(int precision, int Bmax, int n_in, const int32_t* dims, int n_layers,
                               int n_out);
int dbn_mlp_workspace_init(void* stream, int precision, int Bmax, int n_in, const int32_t* dims,
                           int n_layers, int n_out, void* workspace);

/* One mini-batch of NumPyAbstractSupervisedDBN._stochastic_gradient_descent
 * (dbn/models.py:439-465) == batched _compute_activations (:394-417) + _backpropagation
 * (:471-515) + output head (:594-626 / :696-720) + decayed SGD update. */
int dbn_mlp_train_step(void* stream, const dbn_mlp_step_args* args);

/* One iteration's mini-batch loop (dbn/models.py:439-465). loss_rows (if set) is [n_rows]. */
int dbn_mlp_fit_epoch(void* stream, const dbn_mlp_step_args* tmpl, int n_rows, int batch_size);

/* Tensor-core modes: dbn_mlp_fit_epoch runs the whole iteration as ONE cooperative launch of the persistent fine-tune
 * kernel (csrc/mlp_tc_epoch.cuh: a phase program F_0..F_{L-1} | HEAD | B_l + U_{l+2} .. | U_1 + U_0 per mini-batch, grid
 * barriers between phases) when the network fits its plan -- <= 4 hidden layers, head <= 16 wide, mini-batches of 16..256
 * rows, no optional outputs, stream not being captured (dropout by Philox or injected masks is handled inside the kernel)
 * -- and as the loop of dbn_mlp_train_step otherwise.
 * Replaces the mini-batch loop of dbn/models.py:439-465.  Returns the number of CTAs of the plan, 0 when it does not fit
 * (or DBN_B200_TC_FINETUNE=0): callers then capture the iteration in a CUDA graph as before. */
int dbn_mlp_tc_epoch_plan(int precision, int n_in, const int32_t* dims, int n_layers, int n_out, int batch_size);
/* Switch it on / off at run time; returns the previous setting (parity tests run the same iteration through both paths). */
int dbn_set_tc_finetune(int enabled);

/* ---- small helpers the host loops need -------------------------------------------------------- */
/* dst[i, :] = src[idx[i], :] with optional dtype conversion and dropout on the input
 * (permutation + batch_generator of the reference: dbn/models.py:106-107, dbn/utils.py:11-13).
 * If ones_col >= 0, dst[i, ones_col] = 1 (bias-statistics column for the tensor-core path). */
int dbn_gather_rows(void* stream, const dbn_mat* src, const int32_t* idx, int n_rows, int n_cols,
                    const dbn_mat* dst, int ones_col, const dbn_rng* input_dropout);
/* dst = scale * src (fp32 -> fp32 / bf16) -- the /p and *p rescale of dbn/models.py:533-535, :546-548
 * and the refresh of bf16 shadows. */
int dbn_scale_cast(void* stream, const float* src, int lds, int rows, int cols, float scale,
                   const dbn_mat* dst);
/* column `col` of a [rows, ld] matrix := 1.0 */
int dbn_fill_ones_column(void* stream, const dbn_mat* m, int rows, int col);
/* out[0] += sum_rows sum_cols (A - B)^2 / denom   (reconstruction error, dbn/models.py:212-220) */
int dbn_sum_sq_diff(void* stream, const dbn_mat* A, const dbn_mat* B, int row0_B, int rows, int cols,
                    float inv_denom, float* out);
/* softmax / identity of head scores in place + delta + loss (dbn/models.py:594-626, :696-720).
 * Y == NULL or delta == NULL: outputs only (predict, dbn/models.py:607-615 / :705-711). */
int dbn_head_delta(void* stream, int head, float* Z, int ldz, const float* Y, int ldy, int B, int C,
                   float* delta, int ldd, float* loss_rows);

#ifdef __cplusplus
}
#endif
#endif /* DBN_B200_H */
