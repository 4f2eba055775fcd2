/* CPU restatement of the annealing plan's colouring and relabelling -- TEST INFRASTRUCTURE ONLY.
 *
 * The SA chain (DESIGN.md "SA chain definition"; the reference's annealer is third-party Haskell, absent:
 * PARITY UNPINNED) visits the positions of a RELABELLED model in order.  The relabelling is part of the
 * definition: greedy colouring of the coupling graph in descending priority, priority = (hash(index), index)
 * with the 32-bit mixer below, every vertex taking the smallest colour none of its already coloured
 * neighbours has; positions = stable counting sort by colour (ascending index inside a class), every class
 * padded to a multiple of 4 positions; the relabelled CSR keeps the stored entry order of each row and drops
 * the diagonal.  The product builds the same thing on the device by Jones-Plassmann rounds
 * (csrc/anneal.cu: colour_round_kernel -- a vertex colours itself once no uncoloured neighbour outranks it,
 * which IS greedy colouring in priority order); this file is the sequential statement of it, so that
 * oracle/anneal_port.c can be run on the ORIGINAL model without anything exported by the product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

typedef struct {
  uint32_t hash, index;
} prio_t;

static int prio_desc(void const *pa, void const *pb) { /* highest (hash, index) first */
  prio_t const *a = (prio_t const *)pa, *b = (prio_t const *)pb;
  if (a->hash != b->hash) return a->hash > b->hash ? -1 : 1;
  return a->index > b->index ? -1 : (a->index < b->index ? 1 : 0);
}

/* colour[n] out; returns the number of colours */
uint32_t oracle_colour(uint64_t n, int64_t const *indptr, int32_t const *cols, int32_t *colour) {
  prio_t *order = (prio_t *)malloc((n ? n : 1) * sizeof(prio_t));
  for (uint64_t i = 0; i < n; ++i) {
    order[i].hash = hash_u32((uint32_t)i);
    order[i].index = (uint32_t)i;
    colour[i] = -1;
  }
  qsort(order, n, sizeof(prio_t), prio_desc);
  uint32_t classes = 0;
  uint64_t used_cap = 64;
  uint8_t *used = (uint8_t *)calloc(used_cap, 1);
  for (uint64_t q = 0; q < n; ++q) {
    uint32_t const i = order[q].index;
    uint64_t const deg = (uint64_t)(indptr[i + 1] - indptr[i]);
    if (deg + 2 > used_cap) {
      used_cap = 2 * (deg + 2);
      used = (uint8_t *)realloc(used, used_cap);
    }
    memset(used, 0, deg + 2);
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      uint32_t const j = (uint32_t)cols[k];
      if (j == i) continue;
      if (colour[j] >= 0 && (uint64_t)colour[j] <= deg) used[colour[j]] = 1;
    }
    int32_t c = 0;
    while (used[c]) ++c;
    colour[i] = c;
    if ((uint32_t)c + 1 > classes) classes = (uint32_t)c + 1;
  }
  free(used);
  free(order);
  return classes;
}

/* class_ptr[classes + 1], order[n_padded] (position -> original spin or -1), position[n]; returns n_padded */
uint64_t oracle_positions(uint64_t n, uint32_t classes, int32_t const *colour, int64_t *class_ptr, int32_t *order, int32_t *position) {
  int64_t *size = (int64_t *)calloc(classes ? classes : 1, sizeof(int64_t));
  for (uint64_t i = 0; i < n; ++i) ++size[colour[i]];
  class_ptr[0] = 0;
  for (uint32_t c = 0; c < classes; ++c) class_ptr[c + 1] = class_ptr[c] + (size[c] + 3) / 4 * 4;
  uint64_t const n_padded = (uint64_t)class_ptr[classes];
  if (order) {
    for (uint64_t p = 0; p < n_padded; ++p) order[p] = -1;
    int64_t *cursor = (int64_t *)malloc((classes ? classes : 1) * sizeof(int64_t));
    for (uint32_t c = 0; c < classes; ++c) cursor[c] = class_ptr[c];
    for (uint64_t i = 0; i < n; ++i) {
      int64_t const p = cursor[colour[i]]++;
      order[p] = (int32_t)i;
      position[i] = (int32_t)p;
    }
    free(cursor);
  }
  free(size);
  return n_padded;
}

/* relabelled CSR without the diagonal: out_indptr[n_padded + 1], out_cols / out_vals (capacity nnz of the input), out_field[n_padded];
 * returns its number of entries */
uint64_t oracle_relabel(uint64_t n, uint64_t n_padded, int64_t const *indptr, int32_t const *cols, double const *vals, double const *field,
                        int32_t const *order, int32_t const *position, int64_t *out_indptr, int32_t *out_cols, double *out_vals,
                        double *out_field) {
  (void)n;
  uint64_t out = 0;
  for (uint64_t p = 0; p < n_padded; ++p) {
    out_indptr[p] = (int64_t)out;
    int32_t const i = order[p];
    out_field[p] = (i >= 0 && field) ? field[i] : 0.0;
    if (i < 0) continue;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      if (cols[k] == i) continue;
      out_cols[out] = position[cols[k]];
      out_vals[out] = vals[k];
      ++out;
    }
  }
  out_indptr[n_padded] = (int64_t)out;
  return out;
}
