// Neighbour generation on the device with the contract of lattice_symmetries'
// Operator.batched_apply (as called at annealing_sign_problem/common.py:96), including the
// symmetrised case: image -> orbit representative min_g g(s'), coefficient scaled by
// chi(g) * norm(rep)/norm(s), norm(x)^2 = sum_g chi(g)[g x == x] / |G|.
// Plus the canonicalisation kernel (stable per-row column sort + duplicate merge) that
// turns generation-order rows into the canonical CSR of common.py:193-195.
#include <algorithm>

#include "operator.cuh"

namespace asp {

struct SymmetryView {
  const BitPerm *perms;
  const double *characters;
  int num_perms;        // non-identity elements
  int spin_inversion;   // 0, +1, -1
  uint64_t state_mask;
  double group_order;   // (num_perms + 1) * (inversion ? 2 : 1)
};

__device__ __forceinline__ uint64_t delta_swap(uint64_t x, uint64_t m, int d) {
  const uint64_t t = ((x >> d) ^ x) & m;
  return x ^ t ^ (t << d);
}

__device__ __forceinline__ uint64_t permute(const BitPerm &p, uint64_t x) {
  // the masks are the same for every lane (one group element at a time), so skipping the empty
  // stages of a network is a uniform branch
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const uint64_t m = p.mask[k];
    if (m) x = delta_swap(x, m, 32 >> k);
  }
#pragma unroll
  for (int k = 6; k < 11; ++k) {
    const uint64_t m = p.mask[k];
    if (m) x = delta_swap(x, m, 2 << (k - 6));
  }
  return x;
}

// -> representative, character of the element reaching it, norm
__device__ __forceinline__ void state_info(const SymmetryView &g, const BitPerm *s_perms, uint64_t c, uint64_t &rep, double &chi_rep, double &norm) {
  rep = c;
  chi_rep = 1.0;
  double stab = 0.0;
  for (int e = -1; e < g.num_perms; ++e) {
    const uint64_t y0 = e < 0 ? c : permute(s_perms[e], c);
    const double chi0 = e < 0 ? 1.0 : g.characters[e];
    if (y0 == c) stab += chi0;
    if (y0 < rep) {
      rep = y0;
      chi_rep = chi0;
    }
    if (g.spin_inversion) {
      const uint64_t y1 = ~y0 & g.state_mask;
      const double chi1 = chi0 * static_cast<double>(g.spin_inversion);
      if (y1 == c) stab += chi1;
      if (y1 < rep) {
        rep = y1;
        chi_rep = chi1;
      }
    }
  }
  norm = sqrt(fmax(stab, 0.0) / g.group_order);
}

struct ApplyArgs {
  uint64_t num_rows;
  const uint64_t *spins;
  const Move *moves;
  int n_moves, n_down;
  const DiagBond *diag;
  int n_diag;
  const DiagGroup *groups;  // closed form of the diagonal (operator.cuh), n_groups < 0: not available
  int n_groups, diag_scale;
  long long diag_c0;
  SymmetryView sym;
  bool symmetrised;
  int64_t *counts;         // pass 1 out / unused
  const int64_t *offsets;  // pass 2 in
  uint64_t *other_spins;
  double *other_coeffs;
};

constexpr int kApplyThreads = 128;

template <bool kFill>
__global__ void __launch_bounds__(kApplyThreads) apply_kernel(const ApplyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Move *s_moves = reinterpret_cast<Move *>(smem_raw);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(s_moves + a.n_moves);
  BitPerm *s_perms = reinterpret_cast<BitPerm *>(s_diag + a.n_diag);
  for (int k = threadIdx.x; k < a.n_moves; k += blockDim.x) s_moves[k] = a.moves[k];
  for (int k = threadIdx.x; k < a.n_diag; k += blockDim.x) s_diag[k] = a.diag[k];
  for (int k = threadIdx.x; k < a.sym.num_perms; k += blockDim.x) s_perms[k] = a.sym.perms[k];
  __syncthreads();
  const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= a.num_rows) return;
  const uint64_t s = a.spins[r];
  double norm_s = 1.0;
  if (a.symmetrised) {
    uint64_t rep;
    double chi;
    state_info(a.sym, s_perms, s, rep, chi, norm_s);
  }
  int64_t out = kFill ? a.offsets[r] : 0;
  int64_t cnt = 0;
  auto emit = [&](uint64_t c, double coef) {
    if (a.symmetrised) {
      uint64_t rep;
      double chi, norm;
      state_info(a.sym, s_perms, c, rep, chi, norm);
      if (norm == 0.0) return;  // outside the symmetry sector
      c = rep;
      coef = coef * ((chi * norm) / norm_s);
    }
    if (kFill) {
      a.other_spins[out + cnt] = c;
      a.other_coeffs[out + cnt] = coef;
    }
    ++cnt;
  };
  for (int m = 0; m < a.n_down; ++m) {
    const Move mv = s_moves[m];
    if ((s & mv.mask) == mv.need) emit(s ^ mv.flip, mv.coef);
  }
  {
    double d = 0.0;
    if (kFill)
      for (int k = 0; k < a.n_diag; ++k) {
        const DiagBond db = s_diag[k];
        d += db.d[((s >> db.i) & 1) * 2 + ((s >> db.j) & 1)];
      }
    emit(s, d);
  }
  for (int m = a.n_down; m < a.n_moves; ++m) {
    const Move mv = s_moves[m];
    if ((s & mv.mask) == mv.need) emit(s ^ mv.flip, mv.coef);
  }
  if (!kFill) a.counts[r] = cnt;
}

// ---- symmetrised operators whose characters are all +1 (every shipped system: sectors 0, spin
// inversion +1): norm(x)^2 = |stabiliser| / |G| never vanishes, so the number of candidates of a row
// is known without visiting a single orbit, and only the fill pass pays for the group.
__global__ void __launch_bounds__(256) apply_count_plain_kernel(uint64_t num_rows, const uint64_t *__restrict__ spins, const Move *__restrict__ moves,
                                                                int n_moves, int64_t *__restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *s_need = s_mask + n_moves;
  for (int k = threadIdx.x; k < n_moves; k += blockDim.x) {
    s_mask[k] = moves[k].mask;
    s_need[k] = moves[k].need;
  }
  __syncthreads();
  const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= num_rows) return;
  const uint64_t s = spins[r];
  int cnt = 1;  // the diagonal
  for (int m = 0; m < n_moves; ++m) cnt += (s & s_mask[m]) == s_need[m];
  counts[r] = cnt;
}

// orbit representative, multiplicity of the stabiliser (characters all +1)
__device__ __forceinline__ void orbit_info(const SymmetryView &g, const BitPerm *s_perms, uint64_t c, uint64_t &rep, uint32_t &stab) {
  rep = c;
  stab = 1;  // the identity
  if (g.spin_inversion) {
    const uint64_t y1 = ~c & g.state_mask;
    rep = min(rep, y1);  // y1 != c always
  }
  for (int e = 0; e < g.num_perms; ++e) {
    const uint64_t y0 = permute(s_perms[e], c);
    stab += y0 == c;
    rep = min(rep, y0);
    if (g.spin_inversion) {
      const uint64_t y1 = ~y0 & g.state_mask;
      stab += y1 == c;
      rep = min(rep, y1);
    }
  }
}

// Fill pass: every lane finds ITS next applicable move by itself (cheap, divergent) and then all
// lanes walk the group together on their own candidate (expensive, converged): the orbit loop runs
// with ~90 % of the lanes busy instead of the ~25 % of a move-by-move loop.
__global__ void __launch_bounds__(kApplyThreads) apply_fill_positive_kernel(const ApplyArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Move *s_moves = reinterpret_cast<Move *>(smem_raw);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(s_moves + a.n_moves);
  BitPerm *s_perms = reinterpret_cast<BitPerm *>(s_diag + a.n_diag);
  for (int k = threadIdx.x; k < a.n_moves; k += blockDim.x) s_moves[k] = a.moves[k];
  for (int k = threadIdx.x; k < a.n_diag; k += blockDim.x) s_diag[k] = a.diag[k];
  for (int k = threadIdx.x; k < a.sym.num_perms; k += blockDim.x) s_perms[k] = a.sym.perms[k];
  __syncthreads();
  const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool live = r < a.num_rows;
  const uint64_t s = live ? a.spins[r] : 0ull;
  uint64_t rep;
  uint32_t stab_s = 1;
  orbit_info(a.sym, s_perms, s, rep, stab_s);  // converged: dead lanes walk the orbit of 0
  int64_t out = live ? a.offsets[r] : 0;
  int m = 0;
  bool diag_done = !live;
  if (!live) m = a.n_moves;
  for (;;) {
    uint64_t c = 0;
    double coef = 0.0;
    bool have = false;
    while (!have && (m < a.n_moves || !diag_done)) {
      if (!diag_done && m >= a.n_down) {  // the diagonal sits between the negative and the positive deltas
        double d = 0.0;
        for (int k = 0; k < a.n_diag; ++k) {
          const DiagBond db = s_diag[k];
          d += db.d[((s >> db.i) & 1) * 2 + ((s >> db.j) & 1)];
        }
        c = s;
        coef = d;
        diag_done = true;
        have = true;
      } else {
        const Move mv = s_moves[m++];
        if ((s & mv.mask) == mv.need) {
          c = s ^ mv.flip;
          coef = mv.coef;
          have = true;
        }
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;
    uint32_t stab_c = 1;
    orbit_info(a.sym, s_perms, c, rep, stab_c);
    if (have) {
      // coef * (chi * norm(c)) / norm(s) with chi = 1 and norm = sqrt(stabiliser / |G|): the operation
      // sequence of apply_kernel<true>, so both kernels agree bit for bit
      const double norm_c = sqrt(fmax(static_cast<double>(stab_c), 0.0) / a.sym.group_order);
      const double norm_s = sqrt(fmax(static_cast<double>(stab_s), 0.0) / a.sym.group_order);
      a.other_spins[out] = rep;
      a.other_coeffs[out] = coef * ((1.0 * norm_c) / norm_s);
      ++out;
    }
  }
}

// Fill pass, one WARP per row: the lanes share the group (element e = lane + 32 k).  g(s ^ flip) = g(s) ^ g(flip) and
// g(flip) is two bits looked up in the element's site table, so the Benes networks run ONCE per row (for g(s)) and a
// candidate costs two byte loads, two shifts and two XORs per group element instead of eleven delta swaps; the orbit
// minimum and the stabiliser count are warp reductions.  Same representatives, same coefficient arithmetic as
// apply_fill_positive_kernel (bitwise equal outputs).
constexpr int kOrbitK = 10;  // group elements per lane: up to 320 non-identity elements
constexpr int kOrbitThreads = 256;

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
  const uint32_t hi = static_cast<uint32_t>(v >> 32), lo = static_cast<uint32_t>(v);
  const uint32_t hi_min = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t lo_min = __reduce_min_sync(0xffffffffu, hi == hi_min ? lo : 0xFFFFFFFFu);
  return (static_cast<uint64_t>(hi_min) << 32) | lo_min;
}

template <bool kTopInHi>  // more than 32 spins: the top bit of a state is in its high word
__global__ void __launch_bounds__(kOrbitThreads, 3) apply_fill_orbit_kernel(const ApplyArgs a, const uint8_t *__restrict__ perm_dst, int number_spins) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Move *s_moves = reinterpret_cast<Move *>(smem_raw);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(s_moves + a.n_moves);
  BitPerm *s_perms = reinterpret_cast<BitPerm *>(s_diag + a.n_diag);
  uint2 *s_image = reinterpret_cast<uint2 *>(s_perms + a.sym.num_perms);  // [bit][element]: g(1 << bit) as {low word, high word}
  const int stride = (a.sym.num_perms + 31) & ~31;
  const int order = static_cast<int>(a.sym.group_order);
  // sqrt(stabiliser / |G|) for every possible stabiliser size, and the two flipped bits of every move: a candidate then
  // costs one table read instead of a double-precision square root and two find-first-set chains on all 32 lanes
  double *s_norm = reinterpret_cast<double *>(s_image + static_cast<size_t>(number_spins + 1) * stride);  // [order + 1]
  uchar2 *s_bits = reinterpret_cast<uchar2 *>(s_norm + order + 1);                                    // [n_moves]
  for (int k = threadIdx.x; k <= order; k += blockDim.x) s_norm[k] = sqrt(fmax(static_cast<double>(k), 0.0) / a.sym.group_order);
  for (int k = threadIdx.x; k < a.n_moves; k += blockDim.x) {
    const uint64_t flip = a.moves[k].flip, rest = flip & (flip - 1);
    s_bits[k] = make_uchar2(static_cast<unsigned char>(__ffsll(static_cast<long long>(flip)) - 1),
                            rest ? static_cast<unsigned char>(__ffsll(static_cast<long long>(rest)) - 1) : static_cast<unsigned char>(255));
  }
  for (int k = threadIdx.x; k < a.n_moves; k += blockDim.x) s_moves[k] = a.moves[k];
  for (int k = threadIdx.x; k < a.n_diag; k += blockDim.x) s_diag[k] = a.diag[k];
  for (int k = threadIdx.x; k < a.sym.num_perms; k += blockDim.x) s_perms[k] = a.sym.perms[k];
  // image of every single bit under every group element, laid out [bit][element]: the lanes of a warp (consecutive
  // elements) read consecutive 8-byte words
  for (int k = threadIdx.x; k < (number_spins + 1) * stride; k += blockDim.x) {  // row number_spins: all zero ("no bit")
    const int bit = k / stride, e = k % stride;
    const uint64_t image = (bit < number_spins && e < a.sym.num_perms) ? 1ull << perm_dst[e * 64 + bit] : 0ull;
    s_image[k] = make_uint2(static_cast<uint32_t>(image), static_cast<uint32_t>(image >> 32));
  }
  const uint32_t zero_row = static_cast<uint32_t>(number_spins);

  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const int rounds = (a.sym.num_perms + 31) / 32;  // group elements per lane actually present (<= kOrbitK)
  const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (kOrbitThreads / 32);
  const bool inversion = a.sym.spin_inversion != 0;
  const uint64_t top = 1ull << (number_spins - 1);
  // under spin inversion x and ~x are one state: the member of the pair without the top bit is the smaller one
  auto fold = [&](uint64_t x) { return (inversion && (x & top)) ? ~x & a.sym.state_mask : x; };
  for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * (kOrbitThreads / 32) + (threadIdx.x >> 5); r < a.num_rows; r += warps) {
    const uint64_t s = a.spins[r];
    uint64_t gs[kOrbitK];
#pragma unroll
    for (int k = 0; k < kOrbitK; ++k) {
      const int e = static_cast<int>(lane) + 32 * k;
      // a lane without an element in the last round carries 2^51: above every state, unchanged by the (all-zero) images
      // and by the fold, so it never is the minimum and never equals the candidate -- no lane tests in the loops below
      gs[k] = 1ull << 51;
      if (k < rounds && e < a.sym.num_perms) {
        const BitPerm &p = s_perms[e];
        uint64_t x = s;
#pragma unroll
        for (int q = 0; q < 6; ++q) x = delta_swap(x, p.mask[q], 32 >> q);
#pragma unroll
        for (int q = 6; q < 11; ++q) x = delta_swap(x, p.mask[q], 2 << (q - 6));
        gs[k] = x;
      }
    }
    // orbit of the candidate s ^ flip (flip = bits b0, b1; 255 = none): representative and stabiliser size.
    // min(y, ~y) = fold(y), and [y == c] + [~y == c] = [fold(y) == fold(c)] (y and ~y differ in the top bit).
    // A state (< 2^52) is carried as the bit pattern of the double 2^52 + y: minimum and equality are then ONE
    // instruction each (DMNMX, DSETP) instead of two-word integer compare-and-select chains, and positive doubles order
    // like their bit patterns, so the warp minimum can stay on the integer halves.
    const uint32_t top_shift = 31u - ((number_spins - 1) & 31);
    const uint32_t mask_lo = static_cast<uint32_t>(a.sym.state_mask), mask_hi = static_cast<uint32_t>(a.sym.state_mask >> 32);
    const uint32_t inv_all = inversion ? 0xFFFFFFFFu : 0u;
    constexpr uint32_t kBiasHi = 0x43300000u;  // high word of 2^52
    auto orbit = [&](uint64_t c, uint32_t b0, uint32_t b1, uint64_t &rep, uint32_t &stab) {
      const uint64_t c_folded = fold(c);
      const double c_d = __hiloint2double(static_cast<int>(static_cast<uint32_t>(c_folded >> 32) | kBiasHi), static_cast<int>(static_cast<uint32_t>(c_folded)));
      double best = __hiloint2double(0x43400000, 0);  // 2^53: above every state
      uint32_t fixed = 0;
      if (lane == 0) {  // the identity
        best = c_d;
        fixed = 1;
      }
      const uint2 *const image0 = s_image + (b0 == 255u ? zero_row : b0) * stride + lane, *const image1 = s_image + (b1 == 255u ? zero_row : b1) * stride + lane;
      auto element = [&](int k) {
        const uint2 f0 = image0[32 * k], f1 = image1[32 * k];
        uint32_t lo = static_cast<uint32_t>(gs[k]) ^ f0.x ^ f1.x, hi = static_cast<uint32_t>(gs[k] >> 32) ^ f0.y ^ f1.y;
        // fold: under inversion a state with the top bit set is replaced by its complement
        const uint32_t m = static_cast<uint32_t>(static_cast<int32_t>((kTopInHi ? hi : lo) << top_shift) >> 31) & inv_all;
        lo ^= m & mask_lo;
        hi ^= m & mask_hi;
        const double y = __hiloint2double(static_cast<int>(hi | kBiasHi), static_cast<int>(lo));
        fixed += y == c_d;
        best = y < best ? y : best;  // (no NaNs here: a plain compare-and-select, fmin would add its NaN fix-up)
      };
#pragma unroll
      for (int k = 0; k < kOrbitK; ++k) {
        if (k >= rounds) break;  // (uniform) no group elements beyond
        element(k);
      }
      const uint32_t bhi = static_cast<uint32_t>(__double2hiint(best)), blo = static_cast<uint32_t>(__double2loint(best));
      const uint32_t hi_min = __reduce_min_sync(0xffffffffu, bhi);
      const uint32_t lo_min = __reduce_min_sync(0xffffffffu, bhi == hi_min ? blo : 0xFFFFFFFFu);
      rep = (static_cast<uint64_t>(hi_min & ~kBiasHi & 0x000FFFFFu) << 32) | lo_min;
      stab = __reduce_add_sync(0xffffffffu, fixed);
    };
    uint64_t rep_s;
    uint32_t stab_s;
    orbit(s, 255u, 255u, rep_s, stab_s);
    const double norm_s = s_norm[stab_s];
    int64_t out = a.offsets[r];
    auto emit = [&](uint64_t c, double coef, uint32_t b0, uint32_t b1) {
      uint64_t rep;
      uint32_t stab_c;
      orbit(c, b0, b1, rep, stab_c);
      if (lane == 0) {
        // the same sqrt(stab / |G|) the general kernels evaluate; equal stabilisers (the usual case) give x / x = 1 exactly
        const double ratio = stab_c == stab_s ? 1.0 : (1.0 * s_norm[stab_c]) / norm_s;
        a.other_spins[out] = rep;
        a.other_coeffs[out] = coef * ratio;
      }
      ++out;
    };
    // which moves apply: lane l tests moves l, l + 32, ... and a ballot makes the answer a warp-uniform bit mask,
    // so the candidate loops below visit applicable moves only (a row has ~37 of 144)
    auto applicable = [&](int base) -> uint32_t {
      const int m = base + static_cast<int>(lane);
      bool on = false;
      if (m < a.n_moves) {
        const Move mv = s_moves[m];
        on = (s & mv.mask) == mv.need;
      }
      return __ballot_sync(0xffffffffu, on);
    };
    auto emit_moves = [&](int first, int last) {  // moves [first, last) in order
      for (int base = first & ~31; base < last; base += 32) {
        uint32_t bits = applicable(base);
        if (base < first) bits &= ~0u << (first - base);
        if (base + 32 > last) bits &= (last - base) >= 32 ? ~0u : ((1u << (last - base)) - 1u);
        while (bits) {
          const int m = base + __ffs(static_cast<int>(bits)) - 1;
          bits &= bits - 1;
          const Move mv = s_moves[m];
          const uchar2 b = s_bits[m];
          emit(s ^ mv.flip, mv.coef, b.x, b.y);
        }
      }
    };
    emit_moves(0, a.n_down);
    {
      double d = 0.0;
      if (a.n_groups >= 0) {
        // every diagonal entry a small dyadic rational: the sum in exact integer arithmetic (a few popcounts) is bit for
        // bit the bond-by-bond f64 sum of the general kernels (the same closed form extract_csr_kernel uses)
        long long acc = a.diag_c0;
        for (int g = 0; g < a.n_groups; ++g) {
          const DiagGroup grp = a.groups[g];
          acc += static_cast<long long>(grp.weight) * __popcll(s & (s >> grp.shift) & grp.sites);
        }
        d = scalbn(static_cast<double>(acc), -a.diag_scale);
      } else {
        for (int k = 0; k < a.n_diag; ++k) {
          const DiagBond db = s_diag[k];
          d += db.d[((s >> db.i) & 1) * 2 + ((s >> db.j) & 1)];
        }
      }
      emit(s, d, 255u, 255u);
    }
    emit_moves(a.n_down, a.n_moves);
  }
}

// ---- canonicalisation: one warp per row, bitonic sort of (col, seq) keys in smem --------
constexpr int kCanonThreads = 128;
constexpr int kCanonMaxRow = 1024;

__global__ void __launch_bounds__(kCanonThreads) canonicalize_kernel(uint64_t num_rows, int row_capacity /* power of two */, const int64_t *__restrict__ row_offsets,
                                                                        const uint32_t *__restrict__ cols, const double *__restrict__ vals,
                                                                        uint32_t *__restrict__ tmp_cols, double *__restrict__ tmp_vals,
                                                                        int64_t *__restrict__ merged_counts, int *__restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char canon_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int kWarps = kCanonThreads / 32;
  uint64_t *key = reinterpret_cast<uint64_t *>(canon_smem) + static_cast<size_t>(w) * row_capacity;
  double *val = reinterpret_cast<double *>(canon_smem) + static_cast<size_t>(kWarps) * row_capacity + static_cast<size_t>(w) * row_capacity;
  const uint64_t r = static_cast<uint64_t>(blockIdx.x) * kWarps + w;
  if (r >= num_rows) return;
  const int64_t begin = row_offsets[r];
  const int len = static_cast<int>(row_offsets[r + 1] - begin);
  if (len > row_capacity) {
    if (lane == 0) {
      atomicExch(overflow, 1);
      merged_counts[r] = 0;
    }
    return;
  }
  int P = 1;
  while (P < len) P <<= 1;
  for (int e = lane; e < P; e += 32) {
    key[e] = e < len ? (static_cast<uint64_t>(cols[begin + e]) << 32 | static_cast<uint32_t>(e)) : ~0ull;
    val[e] = e < len ? vals[begin + e] : 0.0;
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool asc = (i & k) == 0;
          const uint64_t a = key[i], b = key[ixj];
          if ((a > b) == asc) {
            key[i] = b;
            key[ixj] = a;
            const double t = val[i];
            val[i] = val[ixj];
            val[ixj] = t;
          }
        }
      }
      __syncwarp();
    }
  }
  // heads of equal-column runs sum their run in generation (seq) order
  int written = 0;
  for (int base = 0; base < len; base += 32) {
    const int e = base + lane;
    bool head = false;
    uint32_t col = 0;
    double sum = 0.0;
    if (e < len) {
      col = static_cast<uint32_t>(key[e] >> 32);
      head = e == 0 || static_cast<uint32_t>(key[e - 1] >> 32) != col;
      if (head) {
        sum = val[e];
        for (int f = e + 1; f < len && static_cast<uint32_t>(key[f] >> 32) == col; ++f) sum += val[f];
      }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int pos = written + __popc(ballot & ((1u << lane) - 1));
      tmp_cols[begin + pos] = col;
      tmp_vals[begin + pos] = sum;
    }
    written += __popc(ballot);
  }
  if (lane == 0) merged_counts[r] = written;
}

__global__ void __launch_bounds__(256) compact_rows_kernel(uint64_t num_rows, const int64_t *__restrict__ row_offsets, const int64_t *__restrict__ indptr,
                                                           const uint32_t *__restrict__ tmp_cols, const double *__restrict__ tmp_vals,
                                                           int32_t *__restrict__ indices, double *__restrict__ data) {
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= num_rows) return;
  const int64_t src = row_offsets[warp], dst = indptr[warp];
  const int64_t len = indptr[warp + 1] - dst;
  for (int64_t e = lane; e < len; e += 32) {
    indices[dst + e] = static_cast<int32_t>(tmp_cols[src + e]);
    data[dst + e] = tmp_vals[src + e];
  }
}

}  // namespace asp

using namespace asp;

static int g_apply_mode = 0;  // test hook: 1 = always the general kernels, 2 = the lane-per-row fast kernel instead of the warp-per-row one

extern "C" {

void asp_debug_set_apply_mode(int mode) { g_apply_mode = mode; }

int asp_operator_apply_dev(asp_operator const *op, uint64_t num_rows, uint64_t const *d_spins, uint64_t *d_other_spins,
                           double *d_other_coeffs, int64_t *d_counts, uint64_t capacity, uint64_t *h_total, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  ASP_REQUIRE(op != nullptr && op->d_moves != nullptr, "operator is NULL or has no device mirror");
  ASP_REQUIRE(h_total != nullptr, "h_total is NULL");
  *h_total = 0;
  if (num_rows == 0) return ASP_OK;
  ASP_REQUIRE(d_spins && d_counts, "NULL buffer");
  ApplyArgs a{};
  a.num_rows = num_rows;
  a.spins = d_spins;
  a.moves = op->d_moves;
  a.n_moves = static_cast<int>(op->moves.size());
  a.n_down = static_cast<int>(op->n_down);
  a.diag = op->d_diag;
  a.n_diag = static_cast<int>(op->diag.size());
  a.groups = op->d_diag_groups;
  a.n_groups = (op->diag_scale >= 0 && op->d_diag_groups) ? static_cast<int>(op->diag_groups.size()) : -1;
  a.diag_scale = op->diag_scale;
  a.diag_c0 = op->diag_c0;
  a.symmetrised = op->symmetrised();
  a.sym.perms = op->d_perms;
  a.sym.characters = op->d_characters;
  a.sym.num_perms = static_cast<int>(op->perms.size());
  a.sym.spin_inversion = op->spin_inversion;
  a.sym.state_mask = op->state_mask;
  a.sym.group_order = static_cast<double>((op->perms.size() + 1) * (op->spin_inversion ? 2 : 1));
  a.counts = d_counts;
  const size_t smem = op->moves.size() * sizeof(Move) + op->diag.size() * sizeof(DiagBond) + op->perms.size() * sizeof(BitPerm) + 16;
  ASP_REQUIRE(smem <= 200 * 1024, "operator too large for shared memory");
  const unsigned blocks = static_cast<unsigned>((num_rows + kApplyThreads - 1) / kApplyThreads);
  ASP_CUDA_CHECK(cudaFuncSetAttribute(apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ASP_CUDA_CHECK(cudaFuncSetAttribute(apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // all characters +1 (and no odd spin-inversion sector): the fast pair of kernels
  bool positive = a.symmetrised && op->spin_inversion >= 0 && g_apply_mode != 1;
  for (double chi : op->characters) positive = positive && chi == 1.0;
  if (positive) {
    const size_t csmem = op->moves.size() * 16 + 16;
    ASP_CUDA_CHECK(cudaFuncSetAttribute(apply_count_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(csmem)));
    apply_count_plain_kernel<<<static_cast<unsigned>((num_rows + 255) / 256), 256, csmem, s>>>(num_rows, d_spins, op->d_moves, a.n_moves, d_counts);
  } else {
    apply_kernel<false><<<blocks, kApplyThreads, smem, s>>>(a);
  }
  ASP_LAUNCH_CHECK();
  int64_t *d_offsets = nullptr;
  void *d_tmp = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_offsets), (num_rows + 1) * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(&d_tmp, scan_tmp_bytes(num_rows), s));
  int rc = scan_exclusive_i64(d_counts, d_offsets, num_rows, d_tmp, s);
  if (rc != ASP_OK) return rc;
  int64_t total = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&total, d_offsets + num_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  *h_total = static_cast<uint64_t>(total);
  if (static_cast<uint64_t>(total) > capacity || (total > 0 && (!d_other_spins || !d_other_coeffs))) {
    cudaFreeAsync(d_offsets, s);
    cudaFreeAsync(d_tmp, s);
    set_error("candidate capacity %llu < total %lld", static_cast<unsigned long long>(capacity), static_cast<long long>(total));
    return ASP_ERR_WORKSPACE;
  }
  a.offsets = d_offsets;
  a.other_spins = d_other_spins;
  a.other_coeffs = d_other_coeffs;
  // moves that flip one or two bits (every two-site term) and a group of at most 32 * kOrbitK elements: one warp per row
  bool warp_per_row = positive && g_apply_mode != 2 && op->d_perm_dst != nullptr && op->perms.size() <= 32u * kOrbitK &&
                      op->number_spins <= 51;  // the orbit kernel carries a state as the double 2^52 + state, 2^51 marks "no element"
  for (const Move &mv : op->moves) warp_per_row = warp_per_row && __builtin_popcountll(mv.flip) <= 2;
  if (warp_per_row) {
    const size_t osmem = smem + ((op->perms.size() + 31) / 32 * 32) * static_cast<size_t>(op->number_spins + 1) * sizeof(uint2) +
                         (static_cast<size_t>(a.sym.group_order) + 1) * sizeof(double) + ((op->moves.size() * 2 + 15) & ~size_t(15));
    ASP_REQUIRE(osmem <= 200 * 1024, "operator too large for shared memory");
    void (*const orbit_kernel)(const ApplyArgs, const uint8_t *, int) = op->number_spins > 32 ? apply_fill_orbit_kernel<true> : apply_fill_orbit_kernel<false>;
    ASP_CUDA_CHECK(cudaFuncSetAttribute(orbit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(osmem)));
    int per_sm = 0;
    ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, orbit_kernel, kOrbitThreads, osmem));
    const uint64_t want = (num_rows + kOrbitThreads / 32 - 1) / (kOrbitThreads / 32);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(want, static_cast<uint64_t>(kNumSMs) * std::max(per_sm, 1)));
    orbit_kernel<<<grid, kOrbitThreads, osmem, s>>>(a, op->d_perm_dst, static_cast<int>(op->number_spins));
  } else if (positive) {
    ASP_CUDA_CHECK(cudaFuncSetAttribute(apply_fill_positive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    apply_fill_positive_kernel<<<blocks, kApplyThreads, smem, s>>>(a);
  } else {
    apply_kernel<true><<<blocks, kApplyThreads, smem, s>>>(a);
  }
  ASP_LAUNCH_CHECK();
  ASP_CUDA_CHECK(cudaFreeAsync(d_offsets, s));
  ASP_CUDA_CHECK(cudaFreeAsync(d_tmp, s));
  return ASP_OK;
}

int asp_csr_canonicalize(uint64_t num_rows, uint32_t max_row_len, int64_t const *d_row_offsets, uint32_t const *d_cols,
                         double const *d_vals, uint64_t nnz_in, int64_t *d_indptr, int32_t *d_indices, double *d_data,
                         uint64_t *h_nnz, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  ASP_REQUIRE(h_nnz != nullptr && d_indptr != nullptr, "NULL output");
  *h_nnz = 0;
  if (num_rows == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), s));
    return ASP_OK;
  }
  uint32_t *tmp_cols = nullptr;
  double *tmp_vals = nullptr;
  int64_t *merged = nullptr;
  int *overflow = nullptr;
  void *scan_tmp = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&tmp_cols), (nnz_in + 1) * sizeof(uint32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&tmp_vals), (nnz_in + 1) * sizeof(double), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&merged), num_rows * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&overflow), sizeof(int), s));
  ASP_CUDA_CHECK(cudaMallocAsync(&scan_tmp, scan_tmp_bytes(num_rows), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(overflow, 0, sizeof(int), s));
  int row_capacity = 32;
  while (row_capacity < static_cast<int>(max_row_len ? max_row_len : kCanonMaxRow)) row_capacity <<= 1;
  ASP_REQUIRE(row_capacity <= kCanonMaxRow, "max_row_len above 1024 is not supported");
  constexpr int kWarps = kCanonThreads / 32;
  const size_t canon_smem = static_cast<size_t>(kWarps) * row_capacity * 16;
  ASP_CUDA_CHECK(cudaFuncSetAttribute(canonicalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(canon_smem)));
  canonicalize_kernel<<<static_cast<unsigned>((num_rows + kWarps - 1) / kWarps), kCanonThreads, canon_smem, s>>>(
      num_rows, row_capacity, d_row_offsets, d_cols, d_vals, tmp_cols, tmp_vals, merged, overflow);
  ASP_LAUNCH_CHECK();
  int rc = scan_exclusive_i64(merged, d_indptr, num_rows, scan_tmp, s);
  if (rc != ASP_OK) return rc;
  int64_t total = 0;
  int h_overflow = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&total, d_indptr + num_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaMemcpyAsync(&h_overflow, overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  *h_nnz = static_cast<uint64_t>(total);
  if (!h_overflow && total > 0 && d_indices && d_data) {
    compact_rows_kernel<<<static_cast<unsigned>((num_rows * 32 + 255) / 256), 256, 0, s>>>(num_rows, d_row_offsets, d_indptr, tmp_cols, tmp_vals, d_indices, d_data);
    ASP_LAUNCH_CHECK();
  }
  ASP_CUDA_CHECK(cudaFreeAsync(tmp_cols, s));
  ASP_CUDA_CHECK(cudaFreeAsync(tmp_vals, s));
  ASP_CUDA_CHECK(cudaFreeAsync(merged, s));
  ASP_CUDA_CHECK(cudaFreeAsync(overflow, s));
  ASP_CUDA_CHECK(cudaFreeAsync(scan_tmp, s));
  if (h_overflow) {
    set_error("a row has more than %d raw entries", row_capacity);
    return ASP_ERR_UNSUPPORTED;
  }
  return ASP_OK;
}

}  // extern "C"
