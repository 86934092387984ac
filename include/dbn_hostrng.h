/* dbn_hostrng.h -- host-side helper of the dbn_b200 backend (libdbn_hostrng.so, plain C, no CUDA).
 *
 * Not part of the compute boundary (that is include/dbn_b200.h): it only makes the parameter initialisation the
 * reference performs with np.random.randn (dbn/models.py:55-69) cheaper on the host while producing the same
 * numbers.  Binding: deep-belief-network_b200/_hostrng.py. */
#ifndef DBN_HOSTRNG_H
#define DBN_HOSTRNG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* out[0..n) = what n consecutive calls of NumPy's legacy Gaussian (RandomState.standard_normal / randn: MT19937
 * words -> 53-bit uniforms -> polar method with one cached value) return from the generator state
 * (key[624], *pos, *has_gauss, *gauss) = np.random.get_state()[1:5]; the state is advanced exactly as NumPy would
 * advance it, so it can be handed back with np.random.set_state().  n_threads (1..16) parallelises only the
 * per-pair sqrt(-2 log r2 / r2).  Returns 0, or -1 if scratch memory could not be allocated (state untouched). */
int dbn_host_legacy_randn(uint32_t* key, int* pos, int* has_gauss, double* gauss, double* out, size_t n, int n_threads);

/* Same draws, same final state, but out32[i] = (float)(g_i / divisor): `np.random.randn(h, v) / sqrt(v)`
 * (dbn/models.py:55-57) rounded once to the fp32 the device stores, written straight into caller memory (pinned, for
 * the upload) without a float64 array of the layer's size.  Returns 0; -1 if nothing was consumed. */
int dbn_host_legacy_randn_f32(uint32_t* key, int* pos, int* has_gauss, double* gauss, float* out32, size_t n, double divisor,
                              int n_threads);

/* Rows of a host dataset in the device's operand layout (csrc/host_stage.c), gathered in the order `rows` gives (the
 * reference's `data = _data[idx]`, dbn/models.py:106-107; rows == NULL: 0..m-1) and rounded exactly as the device
 * rounds them (bf16: nearest even; f16: nearest even, saturating at +-65504), so that a tensor-core fit can upload
 * 16-bit operands instead of fp32 rows:
 *   out[i, 0..cols) = cast(X[rows[i], 0..cols)),  out[i, cols] = 1 if ones,  remaining columns up to ldo = 0.
 * X: fp32, pitch ldx elements; out: pitch ldo elements of `kind`; every rows[i] must be a valid row of X.
 * n_threads 1..16.  Returns 0, or -1 on a bad argument. */
#define DBN_HOST_F32 0
#define DBN_HOST_BF16 1
#define DBN_HOST_F16 2
#define DBN_HOST_SCALAR 256   /* or-ed into kind: use the portable scalar conversions (tests) */
int dbn_host_stage_rows(const float* X, size_t ldx, const int64_t* rows, size_t m, size_t cols, void* out, size_t ldo,
                        int kind, int ones, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
