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
 * pair and runs on a few threads.  For half a million draws and more the rejection pass goes to the threads as well
 * (an attempt always consumes four words, so attempt t sits at words [4t, 4t+4) of the stream) and only the MT19937
 * recurrence stays sequential: 16.7 M draws of a 4096 x 4096 layer in ~1/5 of the time.  Every output is
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

/* ---- large draws: the rejection pass on threads, too ---------------------------------------------------------------
 * For millions of draws (a 4096 x 4096 layer: 16.7 M) the sequential rejection pass above is the bound (~9 ns per
 * attempt).  Only the MT19937 recurrence itself is sequential; an attempt always consumes exactly four words, so attempt
 * t of the stream uses words [4t, 4t + 4) whatever was accepted before it.  The calling thread therefore only produces
 * RAW state words, a chunk of CHUNK_ATT attempts at a time; worker threads temper / evaluate / compact their share of
 * the chunk's attempts, meet at a barrier, and finish (log, sqrt) their accepted pairs straight into `out` at the offset
 * given by the counts of the threads before them -- while the caller already produces the next chunk's words.  Values and
 * final state are those of the sequential algorithm: same operations per attempt and per pair, same order of outputs. */
#include <string.h>

#define PAR_MIN_N ((size_t)1 << 17)   /* the binding passes n_threads = 1 below its own (measured) threshold */
#define CHUNK_ATT ((size_t)1 << 18)
#define PAR_MAX_T 16

/* copy `nwords` raw words of the generator into dst, regenerating blocks as they are used up */
static void raw_fill(uint32_t* key, int* p, uint32_t* dst, size_t nwords) {
  while (nwords > 0) {
    size_t take;
    if (*p >= MT_N) {
      mt_regenerate(key);
      *p = 0;
    }
    take = (size_t)(MT_N - *p);
    if (take > nwords) take = nwords;
    memcpy(dst, key + *p, take * sizeof(uint32_t));
    dst += take;
    *p += (int)take;
    nwords -= take;
  }
}

typedef struct par_ctx par_ctx;
typedef struct {
  par_ctx* cx;
  int tid;
  double* px[2];         /* per chunk slot: accepted pairs, px[2i] = x2, px[2i+1] = x1 */
  double* pr2[2];
  uint32_t* patt[2];     /* attempt index of accepted pair i */
  size_t n_acc[2];
  pthread_t th;
} par_job;
struct par_ctx {
  int nt;
  par_job jobs[PAR_MAX_T];
  pthread_barrier_t start, mid, done;   /* start / done: workers + caller; mid: workers */
  volatile int quit;
  /* the chunk being worked on */
  volatile int slot;
  const uint32_t* raw;   /* 4 * n_att raw words */
  size_t n_att;
  size_t base;           /* pairs accepted by earlier chunks */
  /* the call */
  double* out;           /* where pair 0 goes */
  float* out32;          /* instead of `out`, when set: (float)(value / divisor) -- initial weights in the device's form */
  double divisor;
  size_t want, full_pairs;
  double last[2];        /* the odd pair: first half, second half */
};

static void* par_worker(void* arg) {
  par_job* j = (par_job*)arg;
  par_ctx* cx = j->cx;
  uint32_t w[1024];
  double x1[256], x2[256], r2[256];
  for (;;) {
    size_t a, att0, att1, off = 0, emit, i, n_acc = 0;
    int t, k, sl;
    pthread_barrier_wait(&cx->start);
    if (cx->quit) return 0;
    sl = cx->slot;
    att0 = cx->n_att * (size_t)j->tid / (size_t)cx->nt;
    att1 = cx->n_att * (size_t)(j->tid + 1) / (size_t)cx->nt;
    for (a = att0; a < att1; a += 256) {
      const int na = (int)(att1 - a < 256 ? att1 - a : 256);
      temper_block(cx->raw + 4 * a, w, 4 * na);
      eval_attempts(w, na, x1, x2, r2);
      for (t = 0; t < na; t++) {
        const double rr = r2[t];
        if (rr >= 1.0 || rr == 0.0) continue;
        j->pr2[sl][n_acc] = rr;
        j->px[sl][2 * n_acc] = x2[t];
        j->px[sl][2 * n_acc + 1] = x1[t];
        j->patt[sl][n_acc] = (uint32_t)(a + (size_t)t);
        n_acc++;
      }
    }
    j->n_acc[sl] = n_acc;
    pthread_barrier_wait(&cx->mid);
    for (k = 0; k < j->tid; k++) off += cx->jobs[k].n_acc[sl];
    off += cx->base;                                 /* index of this thread's first pair in the whole call */
    emit = off >= cx->want ? 0 : (cx->want - off < n_acc ? cx->want - off : n_acc);
    for (i = 0; i < emit; i++) {
      const double rr = j->pr2[sl][i];
      const double f = sqrt(-2.0 * log(rr) / rr);
      const size_t g = off + i;
      if (g < cx->full_pairs) {
        if (cx->out32) {
          cx->out32[2 * g] = (float)(f * j->px[sl][2 * i] / cx->divisor);
          cx->out32[2 * g + 1] = (float)(f * j->px[sl][2 * i + 1] / cx->divisor);
        } else {
          cx->out[2 * g] = f * j->px[sl][2 * i];
          cx->out[2 * g + 1] = f * j->px[sl][2 * i + 1];
        }
      } else {                                       /* the odd pair: its second half stays cached in the generator */
        cx->last[0] = f * j->px[sl][2 * i];
        cx->last[1] = f * j->px[sl][2 * i + 1];
      }
    }
    pthread_barrier_wait(&cx->done);
  }
}

/* n >= PAR_MIN_N draws on nt >= 2 threads.  Returns 0 on success; -1 if nothing was consumed (no memory / threads). */
static int randn_parallel(uint32_t* key, int* pos, int* has_gauss, double* gauss, double* out, size_t n, int nt,
                          float* out32, double divisor) {
  size_t done = 0, full_pairs, want, base = 0;
  int odd, cur = 0, t, p = *pos, started = 0, in_flight = 0;
  uint32_t key_work[MT_N], snap_key[2][MT_N];
  int snap_p[2] = {0, 0};
  uint32_t* raw[2];
  par_ctx* cx;
  char* scratch;
  size_t per_thread_att, job_bytes;
  if (nt > PAR_MAX_T) nt = PAR_MAX_T;
  per_thread_att = CHUNK_ATT / (size_t)nt + 2;
  job_bytes = per_thread_att * (3 * sizeof(double) + sizeof(uint32_t)) + 64;
  raw[0] = (uint32_t*)malloc(2 * (4 * CHUNK_ATT * sizeof(uint32_t)) + sizeof(par_ctx) + 2 * (size_t)nt * job_bytes);
  if (!raw[0]) return -1;
  raw[1] = raw[0] + 4 * CHUNK_ATT;
  cx = (par_ctx*)(raw[1] + 4 * CHUNK_ATT);
  scratch = (char*)(cx + 1);
  memset(cx, 0, sizeof(*cx));
  cx->nt = nt;
  for (t = 0; t < nt; t++) {
    int sl;
    cx->jobs[t].cx = cx;
    cx->jobs[t].tid = t;
    for (sl = 0; sl < 2; sl++) {
      char* b = scratch + ((size_t)sl * (size_t)nt + (size_t)t) * job_bytes;
      cx->jobs[t].pr2[sl] = (double*)b;
      cx->jobs[t].px[sl] = (double*)b + per_thread_att;
      cx->jobs[t].patt[sl] = (uint32_t*)((double*)b + 3 * per_thread_att);
    }
  }
  pthread_barrier_init(&cx->start, 0, (unsigned)nt + 1);
  pthread_barrier_init(&cx->mid, 0, (unsigned)nt);
  pthread_barrier_init(&cx->done, 0, (unsigned)nt + 1);
  for (t = 0; t < nt; t++) {
    if (pthread_create(&cx->jobs[t].th, 0, par_worker, &cx->jobs[t]) != 0) break;
    started++;
  }
  if (started < nt) {
    /* nothing has been consumed yet: release the workers that did start (pthread_barrier_wait is a cancellation point
     * nowhere guaranteed, so they are told to quit through a barrier of their own size) and take the sequential path */
    pthread_barrier_t only;
    (void)only;
    cx->quit = 1;
    for (t = started; t < nt; t++) {      /* stand-ins for the missing workers so that `start` (nt + 1 arrivals) opens */
      if (pthread_create(&cx->jobs[t].th, 0, par_worker, &cx->jobs[t]) != 0) {
        /* still no thread: the waiting workers cannot be released safely -- leave them parked (they hold no lock) */
        return -1;
      }
    }
    pthread_barrier_wait(&cx->start);
    for (t = 0; t < nt; t++) pthread_join(cx->jobs[t].th, 0);
    free(raw[0]);
    return -1;
  }
  memcpy(key_work, key, sizeof(key_work));
  if (*has_gauss) {
    if (out32) out32[done++] = (float)(*gauss / divisor);
    else out[done++] = *gauss;
  }
  full_pairs = (n - done) / 2;
  odd = (int)((n - done) & 1);
  want = full_pairs + (size_t)odd;
  cx->out = out32 ? 0 : out + done;
  cx->out32 = out32 ? out32 + done : 0;
  cx->divisor = divisor;
  cx->want = want;
  cx->full_pairs = full_pairs;
  for (;;) {
    /* produce the next chunk's raw words (sized to what may still be missing, with a margin for the rejections) while the
     * workers are busy with the previous chunk */
    const int prev = cur ^ 1;
    size_t natt = (size_t)((double)(want - base) * 1.2732395447351628 * 1.01) + 4096;   /* 4 / pi attempts per pair */
    if (natt > CHUNK_ATT) natt = CHUNK_ATT;
    memcpy(snap_key[cur], key_work, sizeof(key_work));
    snap_p[cur] = p;
    raw_fill(key_work, &p, raw[cur], 4 * natt);
    if (in_flight) {
      size_t acc = 0, prev_base = cx->base;
      pthread_barrier_wait(&cx->done);
      in_flight = 0;
      for (t = 0; t < nt; t++) acc += cx->jobs[t].n_acc[prev];
      if (prev_base + acc >= want) {
        /* finished inside the previous chunk: the state is the snapshot taken before it, advanced by the words it used */
        size_t need = want - prev_base, seen = 0, used_att = 0, left;
        for (t = 0; t < nt; t++) {
          if (seen + cx->jobs[t].n_acc[prev] >= need) {
            used_att = (size_t)cx->jobs[t].patt[prev][need - seen - 1] + 1;
            break;
          }
          seen += cx->jobs[t].n_acc[prev];
        }
        memcpy(key, snap_key[prev], sizeof(key_work));
        p = snap_p[prev];
        left = 4 * used_att;
        while (left > 0) {       /* advance without storing */
          size_t take;
          if (p >= MT_N) { mt_regenerate(key); p = 0; }
          take = (size_t)(MT_N - p);
          if (take > left) take = left;
          p += (int)take;
          left -= take;
        }
        if (odd) {
          if (out32) out32[n - 1] = (float)(cx->last[0] / divisor);
          else out[n - 1] = cx->last[0];
          *gauss = cx->last[1];
          *has_gauss = 1;
        } else {
          *has_gauss = 0;
          *gauss = 0.0;
        }
        *pos = p;
        cx->quit = 1;
        pthread_barrier_wait(&cx->start);
        for (t = 0; t < nt; t++) pthread_join(cx->jobs[t].th, 0);
        pthread_barrier_destroy(&cx->start);
        pthread_barrier_destroy(&cx->mid);
        pthread_barrier_destroy(&cx->done);
        free(raw[0]);
        return 0;
      }
      base = prev_base + acc;
    }
    cx->slot = cur;
    cx->raw = raw[cur];
    cx->n_att = natt;
    cx->base = base;
    pthread_barrier_wait(&cx->start);
    in_flight = 1;
    cur ^= 1;
  }
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
  if (n >= PAR_MIN_N && n_threads >= 2) {
    if (randn_parallel(key, pos, has_gauss, gauss, out, n, n_threads, 0, 1.0) == 0) return 0;
  }
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

/* out32[i] = (float)(g_i / divisor) for the same n draws g_i (and the same final state) as dbn_host_legacy_randn: the
 * initial weights `np.random.randn(h, v) / sqrt(v)` (dbn/models.py:55-57) in the form the device takes them, written
 * straight into (pinned) memory -- no float64 array of the layer's size, no divide pass, no narrowing pass.  Each value
 * goes through the same two roundings as NumPy's `(randn / s).astype(float32)`. */
int dbn_host_legacy_randn_f32(uint32_t* key, int* pos, int* has_gauss, double* gauss, float* out32, size_t n, double divisor,
                              int n_threads) {
  size_t done = 0;
  double* tmp;
  if (n == 0) return 0;
  if (!out32 || divisor == 0.0) return -1;
  if (n >= PAR_MIN_N && n_threads >= 2 &&
      randn_parallel(key, pos, has_gauss, gauss, 0, n, n_threads, out32, divisor) == 0)
    return 0;
  tmp = (double*)malloc((n < 65536 ? n : 65536) * sizeof(double));
  if (!tmp) return -1;
  while (done < n) {       /* small draws (or no threads): the sequential path, a cache-sized piece at a time */
    const size_t m = n - done < 65536 ? n - done : 65536;
    size_t i;
    if (dbn_host_legacy_randn(key, pos, has_gauss, gauss, tmp, m, 1) != 0) {
      free(tmp);
      return done == 0 ? -1 : -2;       /* -2: the state has advanced */
    }
    for (i = 0; i < m; i++) out32[done + i] = (float)(tmp[i] / divisor);
    done += m;
  }
  free(tmp);
  return 0;
}
