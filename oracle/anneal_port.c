/* CPU restatement of the multi-replica simulated annealer -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference calls the third-party Haskell/C package `ising-glass-annealer`
 * (pinned =0.4.1.2 in conda-annealing.yml:8, >=0.3 in setup.py:34) at
 *   annealing_sign_problem/common.py:242-248          sa.anneal(h, seed, number_sweeps,
 *                                                     repetitions, only_best)
 *   experiments/full_hilbert_space.py:84-90, 212-218  same, only_best=False -> (xs, es)
 * It is NOT vendored and cannot be built here (no ghc), so trajectories, beta schedule
 * and RNG are PARITY UNPINNED.  What the reference does pin, and this file honours:
 *   - energy convention E(s) = sum_ij J_ij s_i s_j + sum_i h_i s_i over the full
 *     symmetric matrix incl. the diagonal (full_hilbert_space.py:143-145, common.py:757-760);
 *   - bit layout: word i/64, bit i%64, 1 <=> s_i = +1 (cbits/build_matrix.c:72-73);
 *   - R independent chains, S sweeps each, best configuration per chain returned.
 *
 * Published algorithm restated (single-spin-flip Metropolis simulated annealing as in
 * the package's README / Kirkpatrick et al.): for every sweep t (inverse temperature
 * beta_t), visit the spins in index order; flipping spin i changes the energy by
 *   dE = -s_i * (4 * sum_{j != i} J_ij s_j + 2 h_i);
 * accept when dE <= 0, else with probability exp(-beta_t dE).  After each sweep the chain
 * remembers its configuration if its energy is the lowest seen so far.
 *
 * Determinism contract shared with the CUDA kernel (DESIGN.md "SA chain definition"):
 * every floating-point operation below is a single correctly-rounded IEEE-754 add / mul /
 * fma (binary64 for energies, binary32 for the acceptance probability), so a GPU that
 * performs the same sequence reproduces every accept decision bit for bit.  Compile with
 * -ffp-contract=off.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Philox4x32-10 (Salmon et al., SC'11), counter-based --------------------------- */
static inline void philox4x32_10(uint32_t const ctr[4], uint32_t const key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t const p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t const p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t const n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t const n1 = (uint32_t)p1;
    uint32_t const n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t const n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(uint32_t const *ctr, uint32_t const *key, uint32_t *out) {
  philox4x32_10(ctr, key, out);
}

/* Random word for (replica r, sweep t, position p): element p%4 of the Philox block with
 * counter (p/4, t, r, 0) and key (seed_lo, seed_hi).  t = 0xFFFFFFFF draws the initial
 * configuration. */
static inline uint32_t draw(uint64_t seed, uint32_t r, uint32_t t, uint64_t p) {
  uint32_t const ctr[4] = {(uint32_t)(p >> 2), t, r, (uint32_t)(p >> 34)};
  uint32_t const key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return out[p & 3];
}

/* exp(-x) for 0 < x < 23 in IEEE binary32 from add / mul / fma only (no libm call on the
 * decision path), relative error < 2e-7 -- far below anything an acceptance rate can resolve:
 * t = x log2(e) = k + g with k the nearest integer (magic-number rounding, no conversion) and
 * |g| <= 1/2;  exp(-x) = 2^-k exp(-g ln 2), degree-7 Taylor, 2^-k applied to the exponent bits. */
static inline float exp_neg_f32(float x) {
  float const t = x * 1.44269502f;
  float const r = t + 12582912.0f; /* 1.5 * 2^23: r's low mantissa bits hold round(t) */
  uint32_t rbits;
  memcpy(&rbits, &r, sizeof rbits);
  int32_t const k = (int32_t)(rbits - 0x4B400000u);
  float const kf = r - 12582912.0f;
  float const g = t - kf; /* exact, |g| <= 1/2 */
  float const w = g * -0.693147182f;
  float p = 1.0f / 5040.0f;
  p = fmaf(p, w, 1.0f / 720.0f);
  p = fmaf(p, w, 1.0f / 120.0f);
  p = fmaf(p, w, 1.0f / 24.0f);
  p = fmaf(p, w, 1.0f / 6.0f);
  p = fmaf(p, w, 0.5f);
  p = fmaf(p, w, 1.0f);
  p = fmaf(p, w, 1.0f);
  uint32_t pbits;
  memcpy(&pbits, &p, sizeof pbits);
  pbits -= (uint32_t)k << 23; /* p in [0.70, 1.42], k <= 34: stays a normal number */
  memcpy(&p, &pbits, sizeof p);
  return p;
}

/* Metropolis test for an uphill move with 0 < x = beta dE < 23 and the 32-bit variate rnd:
 * u = (rnd + 1/2) 2^-32 (binary32) < exp(-x). */
static inline int accept_uphill(double x, uint32_t rnd) {
  float const u = fmaf((float)rnd, 2.32830644e-10f, 1.16415322e-10f);
  return u < exp_neg_f32((float)x);
}

double oracle_exp_neg(double x) { return (double)exp_neg_f32((float)x); }

#define ORACLE_REJECT_ABOVE 23.0 /* exp(-23) < 2^-33 = smallest uniform variate */

/* One replica: S sequential sweeps over positions 0..n-1.
 *   spin  : working configuration, one byte per position (0/1)
 *   best  : packed best-so-far configuration (out)
 * Returns the fixed-point running energy offset of the best configuration relative to the
 * start (units of 1/escale). */
static int64_t anneal_one(uint64_t n, int64_t const *indptr, int32_t const *cols,
                          double const *vals, double const *field, uint32_t S,
                          double const *betas, uint64_t seed, uint32_t r, double escale,
                          uint8_t *spin, uint64_t *best, int64_t *final_rel) {
  uint64_t const words = (n + 63) / 64;
  int64_t rel = 0, best_rel = 0;
  memset(best, 0, words * sizeof(uint64_t));
  for (uint64_t i = 0; i < n; ++i)
    if (spin[i]) best[i >> 6] |= (uint64_t)1 << (i & 63);
  for (uint32_t t = 0; t < S; ++t) {
    double const beta = betas[t];
    for (uint64_t i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
        uint64_t const j = (uint64_t)cols[k];
        if (j == i) continue;
        acc = acc + (spin[j] ? vals[k] : -vals[k]);
      }
      double const g = 4.0 * acc + 2.0 * (field ? field[i] : 0.0);
      double const dE = spin[i] ? -g : g;
      int accept;
      if (dE <= 0.0) {
        accept = 1;
      } else {
        double const x = beta * dE;
        if (x >= ORACLE_REJECT_ABOVE) {
          accept = 0;
        } else {
          accept = accept_uphill(x, draw(seed, r, t, i));
        }
      }
      if (accept) {
        spin[i] ^= 1;
        rel += llrint(dE * escale);
      }
    }
    if (rel < best_rel) {
      best_rel = rel;
      memset(best, 0, words * sizeof(uint64_t));
      for (uint64_t i = 0; i < n; ++i)
        if (spin[i]) best[i >> 6] |= (uint64_t)1 << (i & 63);
    }
  }
  *final_rel = rel;
  return best_rel;
}

/* R replicas.  x0 == NULL: replica r starts from bit (draw(seed, r, 0xFFFFFFFF, p) & 1);
 * otherwise every replica starts from the packed configuration x0.
 * out_best: [R][ceil(n/64)] packed; out_best_rel / out_final_rel: [R] fixed-point offsets.
 * Replicas are independent, so they are dealt to `threads` POSIX threads. */
typedef struct {
  uint64_t n;
  int64_t const *indptr;
  int32_t const *cols;
  double const *vals;
  double const *field;
  uint32_t R, S;
  double const *betas;
  uint64_t seed;
  uint64_t const *x0;
  double escale;
  uint64_t *out_best;
  int64_t *out_best_rel;
  int64_t *out_final_rel;
  uint32_t *next; /* shared ticket counter */
} anneal_job;

static void *anneal_worker(void *arg) {
  anneal_job const *job = (anneal_job const *)arg;
  uint64_t const n = job->n, words = (n + 63) / 64;
  uint8_t *spin = (uint8_t *)malloc(n ? n : 1);
  for (;;) {
    uint32_t const r = __atomic_fetch_add(job->next, 1u, __ATOMIC_RELAXED);
    if (r >= job->R) break;
    for (uint64_t i = 0; i < n; ++i)
      spin[i] = job->x0 ? (uint8_t)((job->x0[i >> 6] >> (i & 63)) & 1)
                        : (uint8_t)(draw(job->seed, r, 0xFFFFFFFFu, i) & 1);
    job->out_best_rel[r] =
        anneal_one(n, job->indptr, job->cols, job->vals, job->field, job->S, job->betas,
                   job->seed, r, job->escale, spin, job->out_best + (uint64_t)r * words,
                   &job->out_final_rel[r]);
  }
  free(spin);
  return NULL;
}

void oracle_anneal(uint64_t n, int64_t const *indptr, int32_t const *cols, double const *vals,
                   double const *field, uint32_t R, uint32_t S, double const *betas,
                   uint64_t seed, uint64_t const *x0, double escale, uint64_t *out_best,
                   int64_t *out_best_rel, int64_t *out_final_rel, uint32_t threads) {
  uint32_t next = 0;
  anneal_job job = {n, indptr, cols, vals, field, R, S, betas, seed, x0, escale,
                    out_best, out_best_rel, out_final_rel, &next};
  if (threads < 1) threads = 1;
  if (threads > R) threads = R ? R : 1;
  pthread_t *tid = (pthread_t *)malloc(threads * sizeof(pthread_t));
  for (uint32_t k = 1; k < threads; ++k) pthread_create(&tid[k], NULL, anneal_worker, &job);
  anneal_worker(&job);
  for (uint32_t k = 1; k < threads; ++k) pthread_join(tid[k], NULL);
  free(tid);
}
