// Greedy solver on sm_100a -- replaces ising_glass_annealer.greedy_solve as the reference calls it
// (annealing_sign_problem/common.py:249-250; the 32/36-spin production runs use it through
// `--no-annealing`, Makefile:111,125,139).  The library is third-party Haskell and absent; the
// algorithm is restated from the Python the reference preserves at common.py:298-438:
//
//   1. strongest couplings first: edges (i < j, J_ij != 0) in descending |J_ij| (ties: ascending
//      (i, j)); an edge that joins two different clusters merges them and fixes their relative sign
//      so that the edge is satisfied (s_i s_j J_ij < 0) -- common.py:346-372, :398-404;
//   2. local descent: sweep over the spins, flip every spin whose flip lowers the energy, until a
//      sweep changes nothing -- common.py:417-433.
//
// Two documented deviations make the result a function of the model alone and
// computable in parallel: a single spin joins a cluster through the joining edge, like the
// cluster-cluster branch (the reference sums all couplings to the cluster, common.py:374-396), and
// the descent visits the spins in the plan's position order (the reference: dict insertion order).
//
// Step 1 is then the maximum spanning forest under a strict total order of the edges, which does
// not depend on how it is built: the device runs Boruvka rounds (every cluster picks its best
// outgoing edge; picks only form 2-cycles, broken towards the smaller root; pointer jumping with
// sign products), the oracle runs Kruskal (oracle/greedy_port.c) -- same forest, same signs once
// each cluster is normalised to "its smallest position is +1".  Step 2 reuses the plan's colouring:
// spins of one colour class do not interact, so flipping a class in parallel IS the sequential sweep.
#include <climits>

#include "plan.cuh"

namespace asp {

constexpr unsigned long long kNoEdge = ~0ull;

__device__ __forceinline__ unsigned long long pack_edge(uint32_t a, uint32_t b) {
  return a < b ? (static_cast<unsigned long long>(a) << 32) | b : (static_cast<unsigned long long>(b) << 32) | a;
}

// best outgoing edge of every cluster, part 1: largest |J| (non-negative doubles order like their bits)
__global__ void __launch_bounds__(256) greedy_best_weight_kernel(uint32_t np, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                 const double *__restrict__ data, const int32_t *__restrict__ comp,
                                                                 unsigned long long *__restrict__ best_w) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  const int32_t c = comp[v];
  unsigned long long w = 0;
  for (int64_t k = indptr[v]; k < indptr[v + 1]; ++k) {
    if (comp[indices[k]] == c) continue;
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(fabs(data[k])));
    w = max(w, bits);
  }
  if (w != 0) atomicMax(&best_w[c], w);
}

// part 2: among the edges of that weight the smallest (min, max) pair
__global__ void __launch_bounds__(256) greedy_best_edge_kernel(uint32_t np, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                               const double *__restrict__ data, const int32_t *__restrict__ comp,
                                                               const unsigned long long *__restrict__ best_w, unsigned long long *__restrict__ best_e) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  const int32_t c = comp[v];
  const unsigned long long target = best_w[c];
  if (target == 0) return;
  unsigned long long e = kNoEdge;
  for (int64_t k = indptr[v]; k < indptr[v + 1]; ++k) {
    const uint32_t u = static_cast<uint32_t>(indices[k]);
    if (comp[u] == c) continue;
    if (static_cast<unsigned long long>(__double_as_longlong(fabs(data[k]))) == target) e = min(e, pack_edge(v, u));
  }
  if (e != kNoEdge) atomicMin(&best_e[c], e);
}

// every root with an outgoing edge hooks onto the cluster at the other end; a mutual pick keeps the
// smaller root.  link[c] = (parent root) | sign bit 63 when the relative sign of the two roots is -1.
__global__ void __launch_bounds__(256) greedy_hook_kernel(uint32_t np, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                          const double *__restrict__ data, const int32_t *__restrict__ comp,
                                                          const signed char *__restrict__ sigma, const unsigned long long *__restrict__ best_e,
                                                          unsigned long long *__restrict__ link, unsigned int *__restrict__ hooks) {
  const uint32_t c = blockIdx.x * 256u + threadIdx.x;
  if (c >= np) return;
  unsigned long long out = c;  // stays a root
  if (comp[c] == static_cast<int32_t>(c) && best_e[c] != kNoEdge) {
    const unsigned long long e = best_e[c];
    const uint32_t a = static_cast<uint32_t>(e >> 32), b = static_cast<uint32_t>(e);
    const uint32_t i = comp[a] == static_cast<int32_t>(c) ? a : b, j = comp[a] == static_cast<int32_t>(c) ? b : a;
    const uint32_t d = static_cast<uint32_t>(comp[j]);
    if (!(best_e[d] == e && c < d)) {  // not the surviving end of a mutual pick
      double coupling = 0.0;
      for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
        if (static_cast<uint32_t>(indices[k]) == j) {
          coupling = data[k];
          break;
        }
      // the edge is satisfied: s_i s_j = -sign(J); with s = sigma * (root sign): root_c = rel * root_d
      const int rel = (coupling > 0.0 ? -1 : 1) * sigma[i] * sigma[j];
      out = static_cast<unsigned long long>(d) | (rel < 0 ? 1ull << 63 : 0ull);
      atomicAdd(hooks, 1u);
    }
  }
  link[c] = out;
}

// one pointer-jumping step over the hooked roots (ping-pong: reads `in`, writes `out`)
__global__ void __launch_bounds__(256) greedy_jump_kernel(uint32_t np, const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
                                                          unsigned int *__restrict__ changed) {
  const uint32_t c = blockIdx.x * 256u + threadIdx.x;
  if (c >= np) return;
  const unsigned long long mine = in[c];
  const uint32_t p = static_cast<uint32_t>(mine);
  const unsigned long long up = in[p];
  if (static_cast<uint32_t>(up) != p) {
    out[c] = static_cast<unsigned long long>(static_cast<uint32_t>(up)) | ((mine ^ up) & (1ull << 63));
    *changed = 1u;
  } else {
    out[c] = mine;
  }
}

__global__ void __launch_bounds__(256) greedy_relabel_kernel(uint32_t np, const unsigned long long *__restrict__ link, int32_t *__restrict__ comp,
                                                             signed char *__restrict__ sigma) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  const unsigned long long l = link[comp[v]];
  comp[v] = static_cast<int32_t>(static_cast<uint32_t>(l));
  if (l >> 63) sigma[v] = static_cast<signed char>(-sigma[v]);
}

__global__ void __launch_bounds__(256) greedy_init_kernel(uint32_t np, int32_t *__restrict__ comp, signed char *__restrict__ sigma, int32_t *__restrict__ smallest) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  comp[v] = static_cast<int32_t>(v);
  sigma[v] = 1;
  smallest[v] = INT_MAX;
}

__global__ void __launch_bounds__(256) greedy_smallest_kernel(uint32_t np, const int32_t *__restrict__ comp, int32_t *__restrict__ smallest) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  atomicMin(&smallest[comp[v]], static_cast<int32_t>(v));
}

// normalise: the smallest position of every cluster is +1
__global__ void __launch_bounds__(256) greedy_normalise_kernel(uint32_t np, const int32_t *__restrict__ comp, const int32_t *__restrict__ smallest,
                                                               const signed char *__restrict__ sigma, signed char *__restrict__ spin) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;
  if (v >= np) return;
  spin[v] = static_cast<signed char>(sigma[v] * sigma[smallest[comp[v]]]);
}

// local descent on one colour class: flip when dE = -s (4 sum_j J s_j + 2 h) < 0 (row summed in stored order)
__global__ void __launch_bounds__(256) greedy_descent_kernel(uint64_t begin, uint64_t end, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                             const double *__restrict__ data, const double *__restrict__ field, signed char *spin,
                                                             unsigned long long *__restrict__ flips) {
  const uint64_t p = begin + static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= end) return;
  double acc = 0.0;
  for (int64_t k = indptr[p]; k < indptr[p + 1]; ++k) {
    const double v = data[k];
    acc = __dadd_rn(acc, spin[indices[k]] > 0 ? v : -v);
  }
  const double g = __dadd_rn(__dmul_rn(4.0, acc), __dmul_rn(2.0, field[p]));
  const double dE = spin[p] > 0 ? -g : g;
  if (dE < 0.0) {
    spin[p] = static_cast<signed char>(-spin[p]);
    atomicAdd(flips, 1ull);
  }
}

__global__ void __launch_bounds__(256) greedy_pack_kernel(uint64_t n, const int32_t *__restrict__ position, const signed char *__restrict__ spin,
                                                          uint64_t *__restrict__ bits) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  const bool up = i < n && spin[position[i]] > 0;
  const uint32_t word = __ballot_sync(0xffffffffu, up);
  if ((threadIdx.x & 31) == 0 && i < n) reinterpret_cast<uint32_t *>(bits)[i >> 5] = word;
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_greedy_solve(asp_sa_plan *plan, uint64_t *d_bits, double *d_energy, uint32_t *h_rounds, uint32_t *h_sweeps, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_REQUIRE(plan != nullptr && d_bits != nullptr, "NULL argument");
  ASP_CUDA_CHECK(keep_pool_memory());
  const uint64_t n = plan->n;
  const uint32_t np = static_cast<uint32_t>(plan->n_padded);
  const uint64_t words = (n + 63) / 64;
  if (h_rounds) *h_rounds = 0;
  if (h_sweeps) *h_sweeps = 0;
  if (n == 0) return ASP_OK;
  const unsigned blocks = (np + 255) / 256;

  int32_t *comp = nullptr, *smallest = nullptr;
  signed char *sigma = nullptr, *spin = nullptr;
  unsigned long long *best_w = nullptr, *best_e = nullptr, *link_a = nullptr, *link_b = nullptr, *counters = nullptr;
  auto alloc = [&](auto *&ptr, size_t count) { return cudaMallocAsync(reinterpret_cast<void **>(&ptr), std::max<size_t>(count, 1) * sizeof(*ptr), s); };
  ASP_CUDA_CHECK(alloc(comp, np));
  ASP_CUDA_CHECK(alloc(smallest, np));
  ASP_CUDA_CHECK(alloc(sigma, np));
  ASP_CUDA_CHECK(alloc(spin, np));
  ASP_CUDA_CHECK(alloc(best_w, np));
  ASP_CUDA_CHECK(alloc(best_e, np));
  ASP_CUDA_CHECK(alloc(link_a, np));
  ASP_CUDA_CHECK(alloc(link_b, np));
  ASP_CUDA_CHECK(alloc(counters, 2));
  auto release = [&]() {
    for (void *ptr : {static_cast<void *>(comp), static_cast<void *>(smallest), static_cast<void *>(sigma), static_cast<void *>(spin),
                      static_cast<void *>(best_w), static_cast<void *>(best_e), static_cast<void *>(link_a), static_cast<void *>(link_b),
                      static_cast<void *>(counters)})
      cudaFreeAsync(ptr, s);
  };
  unsigned int *hooks = reinterpret_cast<unsigned int *>(counters);
  unsigned int *changed = hooks + 1;
  unsigned long long *flips = counters + 1;

  greedy_init_kernel<<<blocks, 256, 0, s>>>(np, comp, sigma, smallest);
  ASP_LAUNCH_CHECK();
  // ---- step 1: maximum spanning forest by Boruvka rounds ------------------------------------
  uint32_t rounds = 0;
  for (;; ++rounds) {
    ASP_CUDA_CHECK(cudaMemsetAsync(best_w, 0, static_cast<size_t>(np) * sizeof(unsigned long long), s));
    ASP_CUDA_CHECK(cudaMemsetAsync(best_e, 0xFF, static_cast<size_t>(np) * sizeof(unsigned long long), s));
    ASP_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), s));
    greedy_best_weight_kernel<<<blocks, 256, 0, s>>>(np, plan->d_indptr, plan->d_indices, plan->d_data, comp, best_w);
    ASP_LAUNCH_CHECK();
    greedy_best_edge_kernel<<<blocks, 256, 0, s>>>(np, plan->d_indptr, plan->d_indices, plan->d_data, comp, best_w, best_e);
    ASP_LAUNCH_CHECK();
    greedy_hook_kernel<<<blocks, 256, 0, s>>>(np, plan->d_indptr, plan->d_indices, plan->d_data, comp, sigma, best_e, link_a, hooks);
    ASP_LAUNCH_CHECK();
    unsigned int h_hooks = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&h_hooks, hooks, sizeof(h_hooks), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (h_hooks == 0) break;
    for (;;) {  // pointer jumping until every hooked root points at a surviving root
      ASP_CUDA_CHECK(cudaMemsetAsync(changed, 0, sizeof(unsigned int), s));
      greedy_jump_kernel<<<blocks, 256, 0, s>>>(np, link_a, link_b, changed);
      ASP_LAUNCH_CHECK();
      std::swap(link_a, link_b);
      unsigned int h_changed = 0;
      ASP_CUDA_CHECK(cudaMemcpyAsync(&h_changed, changed, sizeof(h_changed), cudaMemcpyDeviceToHost, s));
      ASP_CUDA_CHECK(cudaStreamSynchronize(s));
      if (!h_changed) break;
    }
    greedy_relabel_kernel<<<blocks, 256, 0, s>>>(np, link_a, comp, sigma);
    ASP_LAUNCH_CHECK();
    if (rounds > 64) {
      release();
      set_error("greedy: the cluster merge did not converge");
      return ASP_ERR_CUDA;
    }
  }
  greedy_smallest_kernel<<<blocks, 256, 0, s>>>(np, comp, smallest);
  ASP_LAUNCH_CHECK();
  greedy_normalise_kernel<<<blocks, 256, 0, s>>>(np, comp, smallest, sigma, spin);
  ASP_LAUNCH_CHECK();
  // ---- step 2: local descent, colour class by colour class, until a sweep flips nothing -----
  uint32_t sweeps = 0;
  for (;;) {
    ASP_CUDA_CHECK(cudaMemsetAsync(flips, 0, sizeof(unsigned long long), s));
    for (uint32_t c = 0; c < plan->num_classes; ++c) {
      const uint64_t begin = static_cast<uint64_t>(plan->class_ptr[c]), end = static_cast<uint64_t>(plan->class_ptr[c + 1]);
      if (end == begin) continue;
      greedy_descent_kernel<<<static_cast<unsigned>((end - begin + 255) / 256), 256, 0, s>>>(begin, end, plan->d_indptr, plan->d_indices, plan->d_data,
                                                                                             plan->d_field, spin, flips);
      ASP_LAUNCH_CHECK();
    }
    ++sweeps;
    unsigned long long h_flips = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&h_flips, flips, sizeof(h_flips), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (h_flips == 0) break;
  }
  ASP_CUDA_CHECK(cudaMemsetAsync(d_bits, 0, words * sizeof(uint64_t), s));
  greedy_pack_kernel<<<static_cast<unsigned>((words * 64 + 255) / 256), 256, 0, s>>>(n, plan->d_position, spin, d_bits);
  ASP_LAUNCH_CHECK();
  release();
  if (h_rounds) *h_rounds = rounds;
  if (h_sweeps) *h_sweeps = sweeps;
  if (d_energy) return asp_energy(n, plan->d_indptr0, plan->d_indices0, plan->d_data0, plan->d_field0, 1, d_bits, d_energy, stream);
  return ASP_OK;
}

}  // extern "C"
