/* host_rng.c -- NumPy's legacy Gaussian stream, bit for bit, at several times the speed (host code, plain C).
 *
 * The reference initialises every RBM with np.random.randn (dbn/models.py:55-69), and this backend keeps that
 * contract (same seed -> same initial weights).  At 784-[500,500,2000] those 1.6 M draws cost NumPy ~26 ms --
 * more than the three training epochs take on the GPU -- so they bound the end-to-end fit.  This file restates
 * the published algorithm of NumPy's legacy generator (numpy/random/src/legacy/legacy-distributions.c,
 * numpy/random/src/mt19937/mt19937.c; NumPy is a dependency of the reference, not part of it):
 *
 *   mt19937 word        y = key[pos++] (regenerate 624 words when pos == 624), tempered
 *   uniform double      a = word >> 5, b = word >> 6;  u = (a * 2^26 + b) / 2^53
 *   gauss (polar)       do { x1 = 2u-1; x2 = 2u'-1; r2 = x1^2 + x2^2 } while (r2 >= 1 || r2 == 0)
 *                       f = sqrt(-2 log(r2) / r2);  return f*x2 now, f*x1 at the next call (has_gauss / gauss)
 *
 * and splits it in two: the rejection pass is sequential in the stream but needs no transcendental (it is done a
 * 624-word block at a time as plain array code); the log / sqrt / divide of the accepted pairs is independent per
 * pair and runs on a few threads.  Every output is
 * produced by the same double-precision operations in the same order as NumPy's (same libm `log`; compile without
 * -ffast-math and with -ffp-contract=off), and the generator state handed back is exactly the state NumPy would have.
 * tests/test_host_logic.py checks both against np.random for sizes and states of every parity.
 */
#include "dbn_hostrng.h"

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define MT_N 624
#define MT_M 397
#define MATRIX_A 0x9908b0dfUL
#define UPPER_MASK 0x80000000UL
#define LOWER_MASK 0x7fffffffUL

/* The array loops below are plain integer / double code with no cross-iteration dependency inside a vector's reach;
 * function multi-versioning lets the compiler emit an AVX2 copy and pick it at load time on CPUs that have it (the
 * library is built on one machine and runs on another).  Integer and IEEE double operations give identical results
 * in either copy (no FMA contraction: -ffp-contract=off). */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define DBN_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define DBN_CLONES
#endif

DBN_CLONES
static void mt_regenerate(uint32_t* mt) {
  int i;
  uint32_t y;
  for (i = 0; i < MT_N - MT_M; i++) {
    y = (mt[i] & UPPER_MASK) | (mt[i + 1] & LOWER_MASK);
    mt[i] = mt[i + MT_M] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
  }
  for (; i < MT_N - 1; i++) {
    y = (mt[i] & UPPER_MASK) | (mt[i + 1] & LOWER_MASK);
    mt[i] = mt[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
  }
  y = (mt[MT_N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK);
  mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
}

static inline uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680UL;
  y ^= (y << 15) & 0xefc60000UL;
  y ^= (y >> 18);
  return y;
}

typedef struct {
  double* out;       /* pairs: out[2i] holds x2, out[2i+1] holds x1 on entry; f*x2, f*x1 on exit */
  const double* r2;
  size_t begin, end; /* pair range */
} finish_job;

static void* finish_pairs(void* arg) {
  finish_job* j = (finish_job*)arg;
  size_t i;
  for (i = j->begin; i < j->end; i++) {
    const double r2 = j->r2[i];
    const double f = sqrt(-2.0 * log(r2) / r2);
    j->out[2 * i] = f * j->out[2 * i];
    j->out[2 * i + 1] = f * j->out[2 * i + 1];
  }
  return 0;
}

/* The rejection pass as a resumable stream of accepted pairs.  Every attempt of the polar loop consumes exactly
 * four words whether it is accepted or not, so the attempts of a whole 624-word block are evaluated as straight
 * (vectorisable) array code -- temper, two uniforms, x1, x2, r2 -- and only the accept / compact pass is scalar.
 * Up to three words are carried from one block into the next. */
typedef struct {
  uint32_t* key;
  int p;            /* next unread word of the current key block (MT_N = block exhausted) */
  int carry;        /* words of the previous block waiting at the front of w[] */
  int block_p, block_carry;   /* p and carry when the current attempt list was built */
  int n_att, t;     /* attempts in the list, next attempt to look at */
  uint32_t w[MT_N + 4];
  double x1[MT_N / 4 + 1], x2[MT_N / 4 + 1], r2[MT_N / 4 + 1];
} pair_stream;

/* temper `n` state words */
DBN_CLONES
static void temper_block(const uint32_t* key, uint32_t* w, int n) {
  int i;
  for (i = 0; i < n; i++) {
    uint32_t y = key[i];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    w[i] = y;
  }
}

/* the polar loop's attempts t = 0 .. n_att-1, each from four consecutive words */
DBN_CLONES
static void eval_attempts(const uint32_t* w, int n_att, double* x1o, double* x2o, double* r2o) {
  int t;
  for (t = 0; t < n_att; t++) {
    const int32_t a0 = (int32_t)(w[4 * t] >> 5), b0 = (int32_t)(w[4 * t + 1] >> 6);
    const int32_t a1 = (int32_t)(w[4 * t + 2] >> 5), b1 = (int32_t)(w[4 * t + 3] >> 6);
    const double u1 = (a0 * 67108864.0 + b0) / 9007199254740992.0;
    const double u2 = (a1 * 67108864.0 + b1) / 9007199254740992.0;
    const double x1 = 2.0 * u1 - 1.0, x2 = 2.0 * u2 - 1.0;
    x1o[t] = x1;
    x2o[t] = x2;
    r2o[t] = x1 * x1 + x2 * x2;
  }
}

static void stream_refill(pair_stream* s) {
  int avail, total, i, t;
  if (s->n_att > 0) {         /* the previous list is used up: keep its unconsumed tail words */
    const int total_prev = s->block_carry + (MT_N - s->block_p);
    s->carry = total_prev - 4 * s->n_att;
    for (i = 0; i < s->carry; i++) s->w[i] = s->w[4 * s->n_att + i];
    s->p = MT_N;
  }
  if (s->p >= MT_N) {
    mt_regenerate(s->key);
    s->p = 0;
  }
  s->block_p = s->p;
  s->block_carry = s->carry;
  avail = MT_N - s->p;
  temper_block(s->key + s->p, s->w + s->carry, avail);
  total = s->carry + avail;
  s->n_att = total / 4;
  s->t = 0;
  eval_attempts(s->w, s->n_att, s->x1, s->x2, s->r2);
  (void)t; (void)i;
  if (s->n_att == 0) {        /* fewer than four words in hand (only possible with a nearly exhausted block) */
    s->carry = total;
    s->p = MT_N;
    s->block_p = MT_N;        /* so that the next refill keeps exactly these words */
    s->block_carry = total;
  }
}

/* Accepts `count` pairs: pair i goes to dst[2i] (x2: the value legacy_gauss returns first) and dst[2i+1] (x1),
 * its r2 to r2_out[i]. */
static void stream_take(pair_stream* s, size_t count, double* dst, double* r2_out) {
  size_t got = 0;
  while (got < count) {
    if (s->t >= s->n_att) {
      stream_refill(s);
      continue;
    }
    for (; s->t < s->n_att && got < count; s->t++) {
      const double rr = s->r2[s->t];
      if (rr >= 1.0 || rr == 0.0) continue;
      r2_out[got] = rr;
      dst[2 * got] = s->x2[s->t];
      dst[2 * got + 1] = s->x1[s->t];
      got++;
    }
  }
}

/* generator position after the last attempt looked at */
static int stream_pos(const pair_stream* s) {
  if (s->n_att == 0) return s->p;
  return s->block_p + 4 * s->t - s->block_carry;
}

/* The finishing pass of one chunk on up to 16 helper threads, started without waiting: the caller goes on with the
 * (sequential) rejection pass of the next chunk and joins later. */
typedef struct {
  pthread_t th[16];
  finish_job jobs[16];
  int started[16];
  int nt;
} finish_batch;

static void finish_start(finish_batch* fb, double* pairs, const double* r2, size_t count, int nt) {
  int t;
  fb->nt = 0;
  if (count == 0) return;
  if (count < 16384 || nt <= 1) {
    finish_job j = {pairs, r2, 0, count};
    finish_pairs(&j);
    return;
  }
  if (nt > 16) nt = 16;
  fb->nt = nt;
  for (t = 0; t < nt; t++) {
    fb->jobs[t].out = pairs;
    fb->jobs[t].r2 = r2;
    fb->jobs[t].begin = count * (size_t)t / (size_t)nt;
    fb->jobs[t].end = count * (size_t)(t + 1) / (size_t)nt;
    fb->started[t] = pthread_create(&fb->th[t], 0, finish_pairs, &fb->jobs[t]) == 0;
    if (!fb->started[t]) finish_pairs(&fb->jobs[t]);   /* a share whose thread did not start */
  }
}

static void finish_join(finish_batch* fb) {
  int t;
  for (t = 0; t < fb->nt; t++)
    if (fb->started[t]) pthread_join(fb->th[t], 0);
  fb->nt = 0;
}

/* Fills out[0..n) with what n consecutive legacy_gauss() calls would return, starting from the generator state
 * (key[624], *pos, *has_gauss, *gauss), and leaves that state exactly as those calls would.  Returns 0, or -1 if
 * scratch memory could not be allocated (nothing is consumed then).  Works in chunks of 64 Ki pairs so that the
 * scratch and the freshly written outputs are still in cache when the second pass reads them. */
#define CHUNK_PAIRS 262144
int dbn_host_legacy_randn(uint32_t* key, int* pos, int* has_gauss, double* gauss, double* out, size_t n, int n_threads) {
  size_t done = 0, full_pairs, taken = 0;
  double* r2[2];
  pair_stream* s;
  finish_batch fb[2];
  int odd, cur = 0;
  if (n == 0) return 0;
  r2[0] = (double*)malloc(2 * sizeof(double) * CHUNK_PAIRS + sizeof(pair_stream));
  if (!r2[0]) return -1;
  r2[1] = r2[0] + CHUNK_PAIRS;
  s = (pair_stream*)(r2[1] + CHUNK_PAIRS);
  s->key = key;
  s->p = *pos;
  s->carry = 0;
  s->n_att = 0;
  s->t = 0;
  s->block_p = *pos;
  s->block_carry = 0;
  fb[0].nt = fb[1].nt = 0;
  if (*has_gauss) {          /* the cached half of an earlier pair comes first */
    out[done++] = *gauss;
    *has_gauss = 0;
    *gauss = 0.0;
  }
  full_pairs = (n - done) / 2;           /* pairs whose both halves are returned */
  odd = (int)((n - done) & 1);           /* one more pair: first half returned, second half stays cached */
  while (taken < full_pairs) {
    const size_t c = full_pairs - taken < CHUNK_PAIRS ? full_pairs - taken : CHUNK_PAIRS;
    double* dst = out + done + 2 * taken;
    finish_join(&fb[cur]);               /* the r2 buffer of this slot is free again */
    stream_take(s, c, dst, r2[cur]);     /* sequential: rejection pass of chunk i ... */
    finish_start(&fb[cur], dst, r2[cur], c, n_threads);   /* ... while chunk i-1 is still being finished */
    taken += c;
    cur ^= 1;
  }
  finish_join(&fb[0]);
  finish_join(&fb[1]);
  if (odd) {
    double last[2], rr, f;
    stream_take(s, 1, last, &rr);
    f = sqrt(-2.0 * log(rr) / rr);
    out[n - 1] = f * last[0];
    *gauss = f * last[1];
    *has_gauss = 1;
  }
  *pos = stream_pos(s);
  free(r2[0]);
  return 0;
}
