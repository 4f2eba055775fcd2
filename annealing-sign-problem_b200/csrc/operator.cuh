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
  bool distinct_flips = true;
  // symmetry group (non-identity permutations), real characters
  std::vector<asp::BitPerm> perms;
  std::vector<double> characters;
  // device mirrors (device current at creation)
  asp::Move *d_moves = nullptr;
  asp::DiagBond *d_diag = nullptr;
  asp::BitPerm *d_perms = nullptr;
  double *d_characters = nullptr;
  int device = -1;

  bool symmetrised() const { return spin_inversion != 0 || !perms.empty(); }
  bool sorted_emitter() const { return !symmetrised() && distinct_flips; }
  uint32_t max_candidates() const { return static_cast<uint32_t>(moves.size()) + 1; }
};
