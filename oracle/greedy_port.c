/* CPU restatement of the greedy solver -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference calls `ising_glass_annealer.greedy_solve` (third-party Haskell, pinned
 * =0.4.1.2 in conda-annealing.yml:8, NOT vendored, no ghc here) at
 * annealing_sign_problem/common.py:249-250.  PARITY UNPINNED.  The algorithm is restated from the
 * Python the reference preserves as a comment at common.py:298-438:
 *   - couplings visited in descending |J| (common.py:313-320: argsort(|data|)[::-1], s1 < s2);
 *   - an edge between two different clusters merges them, one of them flipped when the edge is
 *     frustrated (common.py:359-372); two free spins start a cluster with the edge satisfied
 *     (common.py:398-404);
 *   - then sweeps "flip every spin with positive local energy until nothing changes"
 *     (common.py:417-433).
 * Deviations shared with the CUDA implementation (csrc/greedy.cu): ties in |J| are
 * broken by ascending (i, j); a single free spin joins a cluster through the joining edge alone
 * (the reference sums all its couplings to the cluster, common.py:374-396); every cluster is
 * normalised so that its smallest index is +1; the sweeps visit the spins in index order.
 * With a strict edge order step 1 is Kruskal's maximum spanning forest with edge-satisfying signs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double w;   /* |J| */
  uint64_t e; /* (i << 32) | j, i < j */
  int neg;    /* J < 0 */
} edge_t;

static int edge_cmp(void const *pa, void const *pb) {
  edge_t const *a = (edge_t const *)pa, *b = (edge_t const *)pb;
  if (a->w != b->w) return a->w > b->w ? -1 : 1;
  return a->e < b->e ? -1 : (a->e > b->e ? 1 : 0);
}

/* union-find with the sign of every element relative to its root */
static uint32_t find(uint32_t *parent, int8_t *rel, uint32_t x, int *sign) {
  int s = 1;
  uint32_t r = x;
  while (parent[r] != r) {
    s *= rel[r];
    r = parent[r];
  }
  /* path compression, keeping signs relative to the root */
  int acc = s;
  while (parent[x] != r && x != r) {
    uint32_t const next = parent[x];
    int const step = rel[x];
    parent[x] = r;
    rel[x] = (int8_t)acc;
    acc *= step;
    x = next;
  }
  *sign = s;
  return r;
}

/* spins out: +1 / -1 per position; returns the number of descent sweeps (the last one flips nothing) */
uint32_t oracle_greedy(uint64_t n, int64_t const *indptr, int32_t const *cols, double const *vals,
                       double const *field, int8_t *spin) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i)
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
      if ((uint64_t)cols[k] > i && vals[k] != 0.0) ++m;
  edge_t *edges = (edge_t *)malloc((m ? m : 1) * sizeof(edge_t));
  m = 0;
  for (uint64_t i = 0; i < n; ++i)
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
      if ((uint64_t)cols[k] > i && vals[k] != 0.0) {
        edges[m].w = fabs(vals[k]);
        edges[m].e = (i << 32) | (uint64_t)cols[k];
        edges[m].neg = vals[k] < 0.0;
        ++m;
      }
  qsort(edges, m, sizeof(edge_t), edge_cmp);
  uint32_t *parent = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  int8_t *rel = (int8_t *)malloc(n ? n : 1);
  for (uint64_t i = 0; i < n; ++i) {
    parent[i] = (uint32_t)i;
    rel[i] = 1;
  }
  for (uint64_t q = 0; q < m; ++q) {
    uint32_t const i = (uint32_t)(edges[q].e >> 32), j = (uint32_t)edges[q].e;
    int si, sj;
    uint32_t const ri = find(parent, rel, i, &si), rj = find(parent, rel, j, &sj);
    if (ri == rj) continue; /* all earlier couplings were stronger: leave the cluster alone */
    int const t = edges[q].neg ? 1 : -1; /* the edge is satisfied: s_i s_j = -sign(J) */
    parent[ri] = rj;
    rel[ri] = (int8_t)(t * si * sj);
  }
  /* signs relative to the roots, then "the smallest index of a cluster is +1" */
  uint32_t *root = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *smallest = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  for (uint64_t i = 0; i < n; ++i) smallest[i] = UINT32_MAX;
  for (uint64_t i = 0; i < n; ++i) {
    int s;
    root[i] = find(parent, rel, (uint32_t)i, &s);
    spin[i] = (int8_t)s;
    if (smallest[root[i]] == UINT32_MAX) smallest[root[i]] = (uint32_t)i; /* ascending i: the first is the smallest */
  }
  for (uint64_t i = 0; i < n; ++i) root[i] = (uint32_t)spin[smallest[root[i]]] == 1u ? 0u : 1u; /* 1 = flip */
  for (uint64_t i = 0; i < n; ++i)
    if (root[i]) spin[i] = (int8_t)-spin[i];
  free(root);
  free(smallest);
  free(parent);
  free(rel);
  free(edges);
  /* local descent */
  uint32_t sweeps = 0;
  for (;;) {
    uint64_t flips = 0;
    for (uint64_t i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
        if ((uint64_t)cols[k] == i) continue;
        acc = acc + (spin[cols[k]] > 0 ? vals[k] : -vals[k]);
      }
      double const g = 4.0 * acc + 2.0 * (field ? field[i] : 0.0);
      double const dE = spin[i] > 0 ? -g : g;
      if (dE < 0.0) {
        spin[i] = (int8_t)-spin[i];
        ++flips;
      }
    }
    ++sweeps;
    if (!flips) break;
  }
  return sweeps;
}
