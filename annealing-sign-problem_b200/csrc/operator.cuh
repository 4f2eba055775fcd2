// Hamiltonian bond list compiled to XOR "moves" (replaces what lattice_symmetries'
// Operator.batched_apply computes for the reference at annealing_sign_problem/common.py:96).
#pragma once
#include <vector>

#include "common.cuh"

namespace asp {

// One off-diagonal matrix element family: applicable to basis word s when
// (s & mask) == need; the image is s ^ flip with coefficient coef.
// candidate - s == delta (signed), independent of the other bits of s, so a move list
// sorted by delta emits the candidates of ANY row in ascending key order.
struct Move {
  uint64_t mask;
  uint64_t need;
  uint64_t flip;
  double coef;
};

// Diagonal contribution of one bond: d[2*bit(s,i) + bit(s,j)].
struct DiagBond {
  uint32_t i, j;
  double d[4];
};

// Diagonal in closed form, available when every diagonal entry is a dyadic rational small enough
// that sums over all bonds are exact in any order (then it equals the bond-by-bond f64 sum bit for
// bit):  d(s) = 2^-scale * (c0 + sum_g weight_g * popcount(s & (s >> shift_g) & sites_g)).
// shift 0 groups carry the single-site part, shift k > 0 groups the bonds (i, i + k).
struct DiagGroup {
  uint64_t sites;
  int32_t weight;
  uint32_t shift;
};

// A permutation of <= 64 bits as a Benes network of delta swaps:
// x = delta_swap(x, mask[k], shift[k]) for k = 0..stages-1.
struct BitPerm {
  uint64_t mask[11];
};

}  // namespace asp

struct asp_operator {
  uint32_t number_spins = 0;
  int32_t hamming_weight = -1;
  int32_t spin_inversion = 0;
  uint64_t state_mask = 0;
  std::vector<asp::Move> moves;  // sorted by delta ascending; [0, n_down) have delta < 0
  uint32_t n_down = 0;
  std::vector<asp::DiagBond> diag;  // in (term, bond) order
  std::vector<asp::DiagGroup> diag_groups;  // closed form of the diagonal (valid when diag_scale >= 0)
  int64_t diag_c0 = 0;
  int32_t diag_scale = -1;
  bool distinct_flips = true;
  // symmetry group (non-identity permutations), real characters
  std::vector<asp::BitPerm> perms;
  std::vector<double> characters;
  std::vector<uint8_t> perm_dst;  // [perms.size()][64]: group element g sends bit i to bit perm_dst[g * 64 + i]
  // device mirrors (device current at creation)
  asp::Move *d_moves = nullptr;
  asp::DiagBond *d_diag = nullptr;
  asp::DiagGroup *d_diag_groups = nullptr;
  asp::BitPerm *d_perms = nullptr;
  double *d_characters = nullptr;
  uint8_t *d_perm_dst = nullptr;
  int device = -1;

  bool symmetrised() const { return spin_inversion != 0 || !perms.empty(); }
  bool sorted_emitter() const { return !symmetrised() && distinct_flips; }
  uint32_t max_candidates() const { return static_cast<uint32_t>(moves.size()) + 1; }
};
