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

#ifdef __cplusplus
}
#endif
#endif
