// asp_operator: compiles the reference's YAML-style bond list (two-site 4x4 terms,
// physical_systems/*.yaml) into delta-sorted XOR moves + diagonal table + symmetry-group
// bit permutations (Benes networks), and mirrors them on the device.
#include <algorithm>
#include <climits>
#include <cmath>
#include <functional>
#include <numeric>

#include "operator.cuh"

namespace asp {

static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

cudaError_t keep_pool_memory() {
  static bool done[64] = {};
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
  cudaMemPool_t pool;
  e = cudaDeviceGetDefaultMemPool(&pool, device);
  if (e != cudaSuccess) return e;
  uint64_t threshold = UINT64_MAX;
  e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  if (e == cudaSuccess) done[device] = true;
  return e;
}

// ---- Benes network construction (looping algorithm) ---------------------------------
// Stages: k = 0..5 swap distance 32 >> k (input side), k = 6..10 distance 2 << (k-6)
// (output side).  out bit q = in bit src[q].
static void benes_route(const std::vector<int> &src, int base, int d, int level, BitPerm &net) {
  // src: for local output position q in [0, 2d) the local input position it takes from.
  const int n = 2 * d;
  if (d == 0) return;
  std::vector<int> dst(n);
  for (int q = 0; q < n; ++q) dst[src[q]] = q;
  if (n == 2) {  // single switch, realised on the input side (stage 5, distance 1)
    if (src[0] == 1) net.mask[5] |= 1ull << base;
    return;
  }
  // side[p]: sub-network (0 = low half, 1 = high half) the element entering at p uses.
  std::vector<int> side(n, -1);
  for (int start = 0; start < n; ++start) {
    if (side[start] != -1) continue;
    int p = start, s = 0;
    while (side[p] == -1) {
      side[p] = s;
      // the element sharing p's input switch must use the other sub-network
      int partner_in = p ^ d;
      side[partner_in] = s ^ 1;
      // the element sharing partner_in's OUTPUT switch must differ from partner_in
      int q = dst[partner_in];
      int q_partner = q ^ d;
      p = src[q_partner];
      // p must use side s again
    }
  }
  std::vector<int> sub_src[2] = {std::vector<int>(d), std::vector<int>(d)};
  for (int p = 0; p < n; ++p) {
    const int lo = p & (d - 1);  // index of the switch
    if (p < d && side[p] == 1) net.mask[level] |= 1ull << (base + lo);  // input switch crossed
    const int q = dst[p];
    const int qlo = q & (d - 1);
    if (q < d && side[p] == 1) net.mask[10 - level] |= 1ull << (base + qlo);  // output switch crossed
    sub_src[side[p]][qlo] = lo;
  }
  benes_route(sub_src[0], base, d / 2, level + 1, net);
  benes_route(sub_src[1], base + d, d / 2, level + 1, net);
}

static inline uint64_t delta_swap(uint64_t x, uint64_t m, int d) {
  const uint64_t t = ((x >> d) ^ x) & m;
  return x ^ t ^ (t << d);
}

static uint64_t apply_perm_host(const BitPerm &net, uint64_t x) {
  for (int k = 0; k < 6; ++k) x = delta_swap(x, net.mask[k], 32 >> k);
  for (int k = 6; k < 11; ++k) x = delta_swap(x, net.mask[k], 2 << (k - 6));
  return x;
}

static bool make_bit_perm(const uint32_t *perm, uint32_t number_spins, BitPerm &net) {
  std::vector<int> src(64);
  std::iota(src.begin(), src.end(), 0);
  for (uint32_t k = 0; k < number_spins; ++k) src[k] = static_cast<int>(perm[k]);
  std::vector<int> seen(64, 0);
  for (int v : src) {
    if (v < 0 || v >= 64 || seen[v]) return false;
    seen[v] = 1;
  }
  std::memset(&net, 0, sizeof(net));
  benes_route(src, 0, 32, 0, net);
  for (int q = 0; q < 64; ++q)
    if (apply_perm_host(net, 1ull << src[q]) != (1ull << q)) return false;
  return true;
}

// Closed form of the diagonal (operator.cuh: DiagGroup).  With t[bl][bh] the entry of a bond for
// bits (bl, bh) at sites lo < hi:  t = t00 + (t10 - t00) bl + (t01 - t00) bh + (t11 - t10 - t01 + t00) bl bh.
static void build_diag_groups(asp_operator *op) {
  op->diag_groups.clear();
  op->diag_scale = -1;
  int scale = 0;
  for (const DiagBond &db : op->diag)
    for (double v : db.d) {
      if (!std::isfinite(v)) return;
      while (scale <= 24 && std::ldexp(v, scale) != std::nearbyint(std::ldexp(v, scale))) ++scale;
      if (scale > 24) return;
    }
  // |entries| * 2^scale * 4 (the mixed coefficient) must fit int32 and every sum of <= 2^12 bonds must be exact
  if (op->diag.size() > 4096) return;
  for (const DiagBond &db : op->diag)
    for (double v : db.d)
      if (std::fabs(std::ldexp(v, scale)) >= static_cast<double>(1 << 26)) return;
  int64_t c0 = 0;
  std::vector<int64_t> site_weight(64, 0);
  struct Pair {
    uint32_t shift;
    int64_t weight;
    uint64_t sites;
  };
  std::vector<Pair> pairs;
  for (const DiagBond &db : op->diag) {
    const uint32_t lo = std::min(db.i, db.j), hi = std::max(db.i, db.j);
    auto t = [&](int bl, int bh) {  // entry for bit(lo) = bl, bit(hi) = bh; db.d index is 2 * bit(i) + bit(j)
      const int bi = db.i == lo ? bl : bh, bj = db.i == lo ? bh : bl;
      return static_cast<int64_t>(std::ldexp(db.d[2 * bi + bj], scale));
    };
    c0 += t(0, 0);
    site_weight[lo] += t(1, 0) - t(0, 0);
    site_weight[hi] += t(0, 1) - t(0, 0);
    const int64_t mixed = t(1, 1) - t(1, 0) - t(0, 1) + t(0, 0);
    if (mixed == 0) continue;
    bool placed = false;
    for (Pair &p : pairs)
      if (p.shift == hi - lo && p.weight == mixed && !(p.sites & (1ull << lo))) {  // a site pair listed twice opens a new group
        p.sites |= 1ull << lo;
        placed = true;
        break;
      }
    if (!placed) pairs.push_back({hi - lo, mixed, 1ull << lo});
  }
  std::vector<Pair> singles;
  for (uint32_t site = 0; site < 64; ++site) {
    if (site_weight[site] == 0) continue;
    bool placed = false;
    for (Pair &p : singles)
      if (p.weight == site_weight[site]) {
        p.sites |= 1ull << site;
        placed = true;
        break;
      }
    if (!placed) singles.push_back({0u, site_weight[site], 1ull << site});
  }
  for (const auto *list : {&singles, &pairs})
    for (const Pair &p : *list) {
      if (p.weight > INT32_MAX || p.weight < INT32_MIN) return;
      op->diag_groups.push_back({p.sites, static_cast<int32_t>(p.weight), p.shift});
    }
  if (op->diag_groups.size() > 512) {  // no gain over the bond loop
    op->diag_groups.clear();
    return;
  }
  op->diag_c0 = c0;
  op->diag_scale = scale;
}

}  // namespace asp

using asp::DiagBond;
using asp::Move;

extern "C" {

int asp_version(void) { return 100; }
const char *asp_last_error(void) { return asp::g_error; }
uint64_t asp_kernel_launch_count(void) { return asp::g_launches.load(); }

int asp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int asp_operator_create(asp_operator **out, uint32_t number_spins, int32_t hamming_weight,
                        int32_t spin_inversion, uint32_t num_terms, double const *matrices,
                        uint32_t const *term_offsets, uint32_t const *sites, uint32_t num_perms,
                        uint32_t const *perms, double const *characters) {
  ASP_REQUIRE(out != nullptr, "out is NULL");
  ASP_REQUIRE(number_spins >= 1 && number_spins <= 64, "number_spins must be in [1, 64]");
  ASP_REQUIRE(spin_inversion >= -1 && spin_inversion <= 1, "spin_inversion must be -1, 0 or +1");
  ASP_REQUIRE(num_terms == 0 || (matrices && term_offsets && sites), "NULL term arrays");
  auto *op = new asp_operator();
  op->number_spins = number_spins;
  op->hamming_weight = hamming_weight;
  op->spin_inversion = spin_inversion;
  op->state_mask = number_spins == 64 ? ~0ull : ((1ull << number_spins) - 1);

  struct Keyed {
    Move m;
    __int128 delta;
    uint32_t seq;
  };
  std::vector<Keyed> keyed;
  for (uint32_t t = 0; t < num_terms; ++t) {
    const double *mat = matrices + 16 * t;
    for (uint32_t b = term_offsets[t]; b < term_offsets[t + 1]; ++b) {
      const uint32_t i = sites[2 * b], j = sites[2 * b + 1];
      if (i >= number_spins || j >= number_spins || i == j) {
        delete op;
        asp::set_error("bond %u: bad sites (%u, %u)", b, i, j);
        return ASP_ERR_ARG;
      }
      DiagBond db;
      db.i = i;
      db.j = j;
      for (int a = 0; a < 4; ++a) db.d[a] = mat[4 * a + a];
      op->diag.push_back(db);
      const uint64_t bi = 1ull << i, bj = 1ull << j;
      for (int a = 0; a < 4; ++a) {
        for (int ap = 0; ap < 4; ++ap) {
          const double c = mat[4 * a + ap];
          if (a == ap || c == 0.0) continue;
          Keyed k;
          k.m.mask = bi | bj;
          k.m.need = ((a >> 1) ? bi : 0) | ((a & 1) ? bj : 0);
          const uint64_t target = ((ap >> 1) ? bi : 0) | ((ap & 1) ? bj : 0);
          k.m.flip = k.m.need ^ target;
          k.m.coef = c;
          __int128 delta = 0;
          for (uint64_t bit : {bi, bj})
            if (k.m.flip & bit) delta += (k.m.need & bit) ? -static_cast<__int128>(bit) : static_cast<__int128>(bit);
          k.delta = delta;
          k.seq = static_cast<uint32_t>(keyed.size());
          keyed.push_back(k);
        }
      }
    }
  }
  std::stable_sort(keyed.begin(), keyed.end(), [](const Keyed &x, const Keyed &y) { return x.delta < y.delta; });
  // identical moves (same bond listed twice) are merged, coefficients summed in listed order
  op->n_down = 0;
  for (const auto &k : keyed) {
    bool merged = false;
    for (auto &m : op->moves)
      if (m.mask == k.m.mask && m.need == k.m.need && m.flip == k.m.flip) {
        m.coef += k.m.coef;
        merged = true;
        break;
      }
    if (merged) continue;
    if (k.delta < 0) ++op->n_down;
    op->moves.push_back(k.m);
  }
  // Two different moves with the same flip that can both apply to one word would emit the
  // same candidate twice: then rows are not duplicate-free and the canonicalising path must
  // be used.  (The two directions of an exchange share a flip but never apply together.)
  op->distinct_flips = true;
  for (size_t a = 0; a < op->moves.size() && op->distinct_flips; ++a)
    for (size_t b = a + 1; b < op->moves.size(); ++b) {
      if (op->moves[b].flip != op->moves[a].flip) continue;
      const uint64_t shared = op->moves[a].mask & op->moves[b].mask;
      if (((op->moves[a].need ^ op->moves[b].need) & shared) == 0) {
        op->distinct_flips = false;
        break;
      }
    }
  asp::build_diag_groups(op);
  for (uint32_t g = 0; g < num_perms; ++g) {
    asp::BitPerm net;
    if (!asp::make_bit_perm(perms + static_cast<size_t>(g) * number_spins, number_spins, net)) {
      delete op;
      asp::set_error("symmetry %u is not a permutation of 0..%u", g, number_spins - 1);
      return ASP_ERR_ARG;
    }
    op->perms.push_back(net);
    op->characters.push_back(characters ? characters[g] : 1.0);
    // out bit k = in bit perm[k]: the element sends bit perm[k] to bit k (bits beyond number_spins stay)
    const size_t base = op->perm_dst.size();
    op->perm_dst.resize(base + 64);
    for (uint32_t k = 0; k < 64; ++k) op->perm_dst[base + k] = static_cast<uint8_t>(k);
    for (uint32_t k = 0; k < number_spins; ++k) op->perm_dst[base + perms[static_cast<size_t>(g) * number_spins + k]] = static_cast<uint8_t>(k);
  }

  if (asp_device_count() > 0) {
    if (cudaGetDevice(&op->device) != cudaSuccess) op->device = -1;
    auto upload = [&](auto *&dst, const auto &vec) -> int {
      using T = typename std::remove_reference<decltype(vec[0])>::type;
      const size_t bytes = std::max<size_t>(vec.size(), 1) * sizeof(T);
      ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&dst), bytes));
      if (!vec.empty()) ASP_CUDA_CHECK(cudaMemcpy(dst, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice));
      return ASP_OK;
    };
    int rc = upload(op->d_moves, op->moves);
    if (rc == ASP_OK) rc = upload(op->d_diag, op->diag);
    if (rc == ASP_OK) rc = upload(op->d_diag_groups, op->diag_groups);
    if (rc == ASP_OK) rc = upload(op->d_perms, op->perms);
    if (rc == ASP_OK) rc = upload(op->d_characters, op->characters);
    if (rc == ASP_OK) rc = upload(op->d_perm_dst, op->perm_dst);
    if (rc != ASP_OK) {
      asp_operator_destroy(op);
      return rc;
    }
  }
  *out = op;
  return ASP_OK;
}

void asp_operator_destroy(asp_operator *op) {
  if (!op) return;
  if (op->d_moves) cudaFree(op->d_moves);
  if (op->d_diag) cudaFree(op->d_diag);
  if (op->d_diag_groups) cudaFree(op->d_diag_groups);
  if (op->d_perms) cudaFree(op->d_perms);
  if (op->d_characters) cudaFree(op->d_characters);
  if (op->d_perm_dst) cudaFree(op->d_perm_dst);
  delete op;
}

uint32_t asp_operator_max_candidates(asp_operator const *op) { return op ? op->max_candidates() : 0; }
int asp_operator_is_sorted_emitter(asp_operator const *op) { return op && op->sorted_emitter() ? 1 : 0; }

}  // extern "C"
