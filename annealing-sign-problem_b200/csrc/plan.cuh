// The annealing plan (coloured + relabelled Ising model on the device), shared by the SA kernels
// (anneal.cu) and the greedy solver (greedy.cu).
#pragma once
#include <vector>

#include "common.cuh"

struct asp_sa_plan {
  uint64_t n = 0;         // original spins
  uint64_t n_padded = 0;  // positions (classes padded to multiples of 4)
  uint64_t nnz = 0;       // entries of the relabelled CSR (diagonal removed)
  uint32_t num_classes = 0;
  // originals (borrowed device pointers; must outlive the plan)
  const int64_t *d_indptr0 = nullptr;
  const int32_t *d_indices0 = nullptr;
  const double *d_data0 = nullptr;
  const double *d_field0 = nullptr;
  // relabelled model (owned)
  int32_t *d_order = nullptr;     // [n_padded] position -> original spin or -1
  int32_t *d_position = nullptr;  // [n] original spin -> position
  int64_t *d_indptr = nullptr;    // [n_padded + 1]
  int32_t *d_indices = nullptr;   // [nnz] positions
  double *d_data = nullptr;       // [nnz]
  double *d_field = nullptr;      // [n_padded]
  int64_t *d_class_ptr = nullptr; // [num_classes + 1]
  int4 *d_bounds = nullptr;       // [n_padded / 4] row boundaries of the sweep kernel's 4-position tasks (relative to the first entry)
  bool has_field = false;         // any h_p != 0
  uint64_t long_tasks = 0;        // 4-position tasks with more than 32 CSR entries (one staged chunk)
  std::vector<int64_t> class_ptr;
  double diag_sum = 0.0;          // sum_i J_ii of the original model (constant part of the energy)
  double max_de = 0.0;            // max_p (4 sum_j |J_pj| + 2 |h_p|): bound on any single-flip energy change
};

