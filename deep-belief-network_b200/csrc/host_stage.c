/* host_stage.c -- rows of a host dataset, gathered in epoch order and written in the device's operand layout.
 *
 * The tensor-core modes contract 16-bit operands: the device path uploads fp32 rows and rounds them there
 * (csrc/misc_kernels.cuh gather_rows_kernel -> dbn_common.cuh mat_store: round-to-nearest-even, fp16 saturating at
 * +-65504).  Doing the SAME rounding on the host while the rows are being gathered for the upload (the reference's
 * `data = _data[idx]`, dbn/models.py:106-107, dbn/utils.py:12-13) halves the bytes that cross PCIe -- the bound of an
 * end-to-end fit of a large dataset -- and gives bit-identical operands.  Plain C + pthreads, part of
 * libdbn_hostrng.so (host-side helper, not the compute boundary).  tests/test_host_logic.py checks every kind
 * against torch's CPU casts, ties / subnormals / overflow included.
 *
 * Row i of the output (pitch `ldo` elements) = [ cast(X[rows[i], 0..cols)) | 1 if ones | 0 ... 0 ].
 */
#include "dbn_hostrng.h"

#include <pthread.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define DBN_X86 1
#else
#define DBN_X86 0
#endif

static inline uint32_t f32_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

/* __float2bfloat16_rn: round to nearest even on the upper 16 bits, NaN -> 0x7FFF */
static inline uint16_t bf16_rn(float f) {
  uint32_t u = f32_bits(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFFu;
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

/* __float2half_rn(fminf(fmaxf(v, -65504), 65504)): the clamp sends NaN to -65504 (fmaxf returns its other operand);
 * after it every value is within half's finite range */
static inline uint16_t f16_sat_rn(float f) {
  uint32_t u, sign, o;
  float c = f > -65504.0f ? f : -65504.0f;        /* NaN compares false -> -65504 */
  c = c < 65504.0f ? c : 65504.0f;
  u = f32_bits(c);
  sign = (u >> 16) & 0x8000u;
  u &= 0x7FFFFFFFu;
  if (u < (113u << 23)) {                        /* below 2^-14: subnormal half; the fp32 add does the rounding */
    const uint32_t magic = 126u << 23;           /* 0.5f: its ulp is 2^-24, the subnormal half's ulp */
    float m, t;
    memcpy(&m, &magic, 4);
    memcpy(&t, &u, 4);
    t += m;
    o = f32_bits(t) - magic;
  } else {
    const uint32_t odd = (u >> 13) & 1u;
    u += ((uint32_t)(15 - 127) << 23) + 0xFFFu + odd;
    o = u >> 13;
  }
  return (uint16_t)(o | sign);
}

static void row_bf16(const float* src, uint16_t* dst, size_t n) {
  size_t j;
  for (j = 0; j < n; j++) dst[j] = bf16_rn(src[j]);
}

#if DBN_X86
__attribute__((target("avx2"))) static void row_bf16_avx2(const float* src, uint16_t* dst, size_t n) {
  size_t j = 0;
  const __m256i abs_mask = _mm256_set1_epi32(0x7FFFFFFF), inf = _mm256_set1_epi32(0x7F800000);
  const __m256i bias = _mm256_set1_epi32(0x7FFF), one = _mm256_set1_epi32(1), nan16 = _mm256_set1_epi32(0x7FFF);
  for (; j + 16 <= n; j += 16) {
    __m256i a = _mm256_loadu_si256((const __m256i*)(src + j)), b = _mm256_loadu_si256((const __m256i*)(src + j + 8));
    __m256i na = _mm256_cmpgt_epi32(_mm256_and_si256(a, abs_mask), inf), nb = _mm256_cmpgt_epi32(_mm256_and_si256(b, abs_mask), inf);
    __m256i ra = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(a, bias), _mm256_and_si256(_mm256_srli_epi32(a, 16), one)), 16);
    __m256i rb = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(b, bias), _mm256_and_si256(_mm256_srli_epi32(b, 16), one)), 16);
    __m256i p;
    ra = _mm256_blendv_epi8(ra, nan16, na);
    rb = _mm256_blendv_epi8(rb, nan16, nb);
    p = _mm256_packus_epi32(ra, rb);                    /* lanes interleaved per 128-bit half */
    p = _mm256_permute4x64_epi64(p, 0xD8);
    _mm256_storeu_si256((__m256i*)(dst + j), p);
  }
  for (; j < n; j++) dst[j] = bf16_rn(src[j]);
}

__attribute__((target("avx2,f16c"))) static void row_f16_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t j = 0;
  const __m256 lo = _mm256_set1_ps(-65504.0f), hi = _mm256_set1_ps(65504.0f);
  for (; j + 8 <= n; j += 8) {
    __m256 v = _mm256_loadu_ps(src + j);
    v = _mm256_max_ps(v, lo);                           /* NaN in the first operand -> second operand, like fmaxf */
    v = _mm256_min_ps(v, hi);
    _mm_storeu_si128((__m128i*)(dst + j), _mm256_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  for (; j < n; j++) dst[j] = f16_sat_rn(src[j]);
}
#endif

static void row_f16(const float* src, uint16_t* dst, size_t n) {
  size_t j;
  for (j = 0; j < n; j++) dst[j] = f16_sat_rn(src[j]);
}

typedef void (*row_fn)(const float*, uint16_t*, size_t);

static row_fn pick_row_fn(int kind, int scalar) {
#if DBN_X86
  __builtin_cpu_init();
  if (scalar) return kind == DBN_HOST_BF16 ? row_bf16 : row_f16;
  if (kind == DBN_HOST_BF16 && __builtin_cpu_supports("avx2")) return row_bf16_avx2;
  if (kind == DBN_HOST_F16 && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("f16c")) return row_f16_f16c;
#endif
  return kind == DBN_HOST_BF16 ? row_bf16 : row_f16;
}

typedef struct {
  const float* X;
  size_t ldx;
  const int64_t* rows;
  size_t i0, i1, cols, ldo;
  char* out;
  int kind, ones;
  row_fn fn;
  pthread_t th;
  int started;
} stage_job;

static void* stage_worker(void* arg) {
  stage_job* j = (stage_job*)arg;
  const size_t esz = j->kind == DBN_HOST_F32 ? 4 : 2;
  const size_t tail0 = j->cols + (j->ones ? 1 : 0);
  size_t i;
  for (i = j->i0; i < j->i1; i++) {
    const float* src = j->X + (size_t)(j->rows ? j->rows[i] : (int64_t)i) * j->ldx;
    char* dst = j->out + i * j->ldo * esz;
    if (j->kind == DBN_HOST_F32) {
      memcpy(dst, src, j->cols * 4);
      if (j->ones) ((float*)dst)[j->cols] = 1.0f;
    } else {
      j->fn(src, (uint16_t*)dst, j->cols);
      if (j->ones) ((uint16_t*)dst)[j->cols] = j->kind == DBN_HOST_BF16 ? 0x3F80u : 0x3C00u;
    }
    if (tail0 < j->ldo) memset(dst + tail0 * esz, 0, (j->ldo - tail0) * esz);
  }
  return 0;
}

int dbn_host_stage_rows(const float* X, size_t ldx, const int64_t* rows, size_t m, size_t cols, void* out, size_t ldo,
                        int kind, int ones, int n_threads) {
  stage_job jobs[16];
  int nt, t;
  size_t per;
  const int scalar = (kind & DBN_HOST_SCALAR) != 0;     /* tests: the portable conversions instead of AVX2 / F16C */
  kind &= ~DBN_HOST_SCALAR;
  if (!X || !out || cols > ldx || cols + (ones ? 1 : 0) > ldo) return -1;
  if (kind != DBN_HOST_F32 && kind != DBN_HOST_BF16 && kind != DBN_HOST_F16) return -1;
  if (m == 0) return 0;
  nt = n_threads < 1 ? 1 : (n_threads > 16 ? 16 : n_threads);
  if (m * cols < ((size_t)1 << 16)) nt = 1;
  if ((size_t)nt > m) nt = (int)m;
  per = (m + (size_t)nt - 1) / (size_t)nt;
  for (t = 0; t < nt; t++) {
    stage_job* j = &jobs[t];
    j->X = X; j->ldx = ldx; j->rows = rows; j->cols = cols; j->ldo = ldo; j->out = (char*)out; j->kind = kind; j->ones = ones;
    j->fn = kind == DBN_HOST_F32 ? 0 : pick_row_fn(kind, scalar);
    j->i0 = (size_t)t * per;
    j->i1 = j->i0 + per < m ? j->i0 + per : m;
    if (j->i0 > m) j->i0 = m;
    j->started = 0;
  }
  for (t = 1; t < nt; t++) jobs[t].started = pthread_create(&jobs[t].th, 0, stage_worker, &jobs[t]) == 0;
  stage_worker(&jobs[0]);
  for (t = 1; t < nt; t++) {
    if (jobs[t].started) pthread_join(jobs[t].th, 0);
    else stage_worker(&jobs[t]);                        /* thread could not be created: do its share here */
  }
  return 0;
}
