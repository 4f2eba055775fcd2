/* CPU restatement of the reference's Ising-extraction kernel -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates, does not copy, /root/reference/cbits/build_matrix.c:
 *   - key comparison            build_matrix.c:7-20   (lexicographic from words[0] upward)
 *   - build_matrix              build_matrix.c:22-53  (per candidate: search the sorted key
 *                               set; hit -> COO triplet, miss -> external-field term)
 *   - extract_signs             build_matrix.c:67-76  (bit i%64 of word i/64 set iff psi>0)
 * and the canonical CSR the reference's live path ends with
 *   - sum duplicates, sort columns   annealing_sign_problem/common.py:193-196.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * Pinned against oracle/_ref (the reference's own C file compiled where it lies) by
 * tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t w[8];
} key512;

/* Three-way order of two 512-bit keys: the first differing word, counted from word 0,
 * decides (build_matrix.c:7-20). */
static inline int key512_order(key512 const *x, key512 const *y) {
  int k = 0;
  while (k < 8 && x->w[k] == y->w[k]) ++k;
  if (k == 8) return 0;
  return x->w[k] < y->w[k] ? -1 : 1;
}

/* Position of `needle` in the ascending, duplicate-free array `hay[0..n)`, or -1. */
static int64_t find512(key512 const *hay, uint64_t n, key512 const *needle) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t const mid = lo + ((hi - lo) >> 1);
    if (key512_order(&hay[mid], needle) < 0)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < n && key512_order(&hay[lo], needle) == 0) return (int64_t)lo;
  return -1;
}

static int64_t find64(uint64_t const *hay, uint64_t n, uint64_t needle) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t const mid = lo + ((hi - lo) >> 1);
    if (hay[mid] < needle)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < n && hay[lo] == needle) ? (int64_t)lo : -1;
}

/* build_matrix.c:22-53 with the reference's argument list and arithmetic association
 * t = ((counts*coeff)*|psi_i|)*|psi'_j|  (line 39-40), field term signed in psi'_j (line 49). */
uint64_t oracle_build_matrix(uint64_t n, key512 const *spins, int64_t const *counts,
                             double const *psi, key512 const *other_spins,
                             double const *other_coeffs, int64_t const *other_counts,
                             double const *other_psi, uint32_t *row_indices,
                             uint32_t *col_indices, double *elements, double *field) {
  uint64_t emitted = 0, cursor = 0;
  for (uint64_t r = 0; r < n; ++r) field[r] = 0.0;
  for (uint64_t r = 0; r < n; ++r) {
    double const a_r = fabs(psi[r]);
    for (int64_t k = 0; k < other_counts[r]; ++k, ++cursor) {
      int64_t const c = find512(spins, n, &other_spins[cursor]);
      double const w = (double)counts[r] * other_coeffs[cursor] * a_r;
      if (c >= 0) {
        row_indices[emitted] = (uint32_t)r;
        col_indices[emitted] = (uint32_t)c;
        elements[emitted] = w * fabs(other_psi[cursor]);
        ++emitted;
      } else {
        field[r] += w * other_psi[cursor];
      }
    }
  }
  return emitted;
}

/* Same contract on plain 64-bit keys (every BASELINE config has <= 64 spins; the
 * reference's Python glue only ever fills words[0]: common.py:58-68, :100). */
uint64_t oracle_build_matrix_u64(uint64_t n, uint64_t const *spins, int64_t const *counts,
                                 double const *psi, uint64_t const *other_spins,
                                 double const *other_coeffs, int64_t const *other_counts,
                                 double const *other_psi, uint32_t *row_indices,
                                 uint32_t *col_indices, double *elements, double *field) {
  uint64_t emitted = 0, cursor = 0;
  for (uint64_t r = 0; r < n; ++r) field[r] = 0.0;
  for (uint64_t r = 0; r < n; ++r) {
    double const a_r = fabs(psi[r]);
    for (int64_t k = 0; k < other_counts[r]; ++k, ++cursor) {
      int64_t const c = find64(spins, n, other_spins[cursor]);
      double const w = (double)counts[r] * other_coeffs[cursor] * a_r;
      if (c >= 0) {
        row_indices[emitted] = (uint32_t)r;
        col_indices[emitted] = (uint32_t)c;
        elements[emitted] = w * fabs(other_psi[cursor]);
        ++emitted;
      } else {
        field[r] += w * other_psi[cursor];
      }
    }
  }
  return emitted;
}

/* build_matrix.c:67-76 */
void oracle_extract_signs(uint64_t n, double const *psi, uint64_t *signs) {
  uint64_t const words = (n + 63) / 64;
  for (uint64_t k = 0; k < words; ++k) signs[k] = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (psi[i] > 0) signs[i >> 6] |= (uint64_t)1 << (i & 63);
}

/* Canonical CSR of row-monotone COO triplets (what build_matrix emits): inside each row a
 * STABLE sort by column, duplicates summed in generation order (SURVEY.md 8c "parity
 * definition"; matches scipy's csr + sort_indices at common.py:193-195 on the index side).
 * indptr has n+1 entries; returns the merged nnz. out_cols/out_vals need capacity nnz_in. */
uint64_t oracle_coo_to_canonical_csr(uint64_t n, uint64_t nnz_in, uint32_t const *rows,
                                     uint32_t const *cols, double const *vals,
                                     int64_t *indptr, int32_t *out_cols, double *out_vals) {
  uint64_t out = 0, k = 0;
  uint64_t cap = 64;
  uint64_t *order = (uint64_t *)malloc(cap * sizeof(uint64_t));
  for (uint64_t r = 0; r < n; ++r) {
    indptr[r] = (int64_t)out;
    uint64_t const begin = k;
    while (k < nnz_in && rows[k] == r) ++k;
    uint64_t const len = k - begin;
    if (len > cap) {
      cap = 2 * len;
      order = (uint64_t *)realloc(order, cap * sizeof(uint64_t));
    }
    /* stable insertion sort of positions by column */
    for (uint64_t a = 0; a < len; ++a) {
      uint64_t const p = begin + a;
      uint64_t b = a;
      while (b > 0 && cols[order[b - 1]] > cols[p]) {
        order[b] = order[b - 1];
        --b;
      }
      order[b] = p;
    }
    for (uint64_t a = 0; a < len; ++a) {
      uint64_t const p = order[a];
      if (a > 0 && cols[p] == cols[order[a - 1]]) {
        out_vals[out - 1] += vals[p];
      } else {
        out_cols[out] = (int32_t)cols[p];
        out_vals[out] = vals[p];
        ++out;
      }
    }
  }
  indptr[n] = (int64_t)out;
  free(order);
  return out;
}

/* E(s) = sum_ij J_ij s_i s_j + sum_i h_i s_i over the full matrix incl. diagonal
 * (energy convention proved by the KAT at experiments/full_hilbert_space.py:143-145 and
 * common.py:757-760). bits: LSB-first, 1 <=> s = +1. */
double oracle_energy(uint64_t n, int64_t const *indptr, int32_t const *cols,
                     double const *vals, double const *field, uint64_t const *bits) {
  double e = 0.0;
  for (uint64_t i = 0; i < n; ++i) {
    double const si = ((bits[i >> 6] >> (i & 63)) & 1) ? 1.0 : -1.0;
    double acc = 0.0;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      uint64_t const j = (uint64_t)cols[k];
      double const sj = ((bits[j >> 6] >> (j & 63)) & 1) ? 1.0 : -1.0;
      acc += vals[k] * sj;
    }
    e += si * (acc + (field ? field[i] : 0.0));
  }
  return e;
}
