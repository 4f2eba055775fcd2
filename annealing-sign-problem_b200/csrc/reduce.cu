// Per-replica reductions over packed sign vectors:
//   asp_energy            replaces sa.Hamiltonian.energy (experiments/full_hilbert_space.py:144;
//                         convention E = sum_ij J_ij s_i s_j + sum_i h_i s_i, common.py:757-760)
//   asp_accuracy_overlap  replaces compute_accuracy_and_overlap (common.py:211-229)
//   asp_csr_symmetrize    replaces 0.5*(M + M.T) (common.py:194) for structurally symmetric J
// All reductions are two-stage with a fixed tree, so results do not depend on scheduling.
#include <algorithm>

#include "common.cuh"

namespace asp {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *smem /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__device__ __forceinline__ double sign_of(const uint64_t *bits, uint64_t i) {
  return ((__ldg(&bits[i >> 6]) >> (i & 63)) & 1) ? 1.0 : -1.0;
}

// grid = (row tiles, replicas); lane per row, rows summed in column order.
__global__ void __launch_bounds__(kRedThreads) energy_partial_kernel(
    uint64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
    const double *__restrict__ data, const double *__restrict__ field, const uint64_t *__restrict__ bits,
    uint64_t words, double *__restrict__ partial /*[replicas][tiles]*/) {
  __shared__ double smem[32];
  const uint64_t *b = bits + static_cast<uint64_t>(blockIdx.y) * words;
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  double e = 0.0;
  if (i < n) {
    double acc = 0.0;
    const int64_t end = indptr[i + 1];
    for (int64_t k = indptr[i]; k < end; ++k) acc += __ldg(&data[k]) * sign_of(b, static_cast<uint64_t>(__ldg(&indices[k])));
    if (field) acc += field[i];
    e = sign_of(b, i) * acc;
  }
  const double total = block_sum(e, smem);
  if (threadIdx.x == 0) partial[static_cast<uint64_t>(blockIdx.y) * gridDim.x + blockIdx.x] = total;
}

// one CTA per replica: fixed-order sum of its tile partials
__global__ void __launch_bounds__(kRedThreads) sum_partials_kernel(const double *__restrict__ partial, uint64_t tiles, double *__restrict__ out) {
  __shared__ double smem[32];
  const double *p = partial + static_cast<uint64_t>(blockIdx.x) * tiles;
  double acc = 0.0;
  for (uint64_t t = threadIdx.x; t < tiles; t += kRedThreads) acc += p[t];
  const double total = block_sum(acc, smem);
  if (threadIdx.x == 0) out[blockIdx.x] = total;
}

// partial: [replicas][tiles][3] = {matches, sum w s_p s_e, sum w}
__global__ void __launch_bounds__(kRedThreads) overlap_partial_kernel(
    uint64_t n, const uint64_t *__restrict__ predicted, const uint64_t *__restrict__ exact,
    const double *__restrict__ weights, uint64_t words, double *__restrict__ partial) {
  __shared__ double smem[32];
  const uint64_t *p = predicted + static_cast<uint64_t>(blockIdx.y) * words;
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  double match = 0.0, dot = 0.0, wsum = 0.0;
  if (i < n) {
    const double sp = sign_of(p, i), se = sign_of(exact, i);
    const double w = weights ? weights[i] : 1.0;
    match = sp == se ? 1.0 : 0.0;
    dot = sp * se * w;
    wsum = w;
  }
  const double m = block_sum(match, smem);
  const double d = block_sum(dot, smem);
  const double w = block_sum(wsum, smem);
  if (threadIdx.x == 0) {
    double *o = partial + (static_cast<uint64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 3;
    o[0] = m;
    o[1] = d;
    o[2] = w;
  }
}

__global__ void __launch_bounds__(kRedThreads) overlap_final_kernel(const double *__restrict__ partial, uint64_t tiles, uint64_t n,
                                                                    double *__restrict__ accuracy, double *__restrict__ overlap) {
  __shared__ double smem[32];
  const double *p = partial + static_cast<uint64_t>(blockIdx.x) * tiles * 3;
  double m = 0.0, d = 0.0, w = 0.0;
  for (uint64_t t = threadIdx.x; t < tiles; t += kRedThreads) {
    m += p[3 * t];
    d += p[3 * t + 1];
    w += p[3 * t + 2];
  }
  m = block_sum(m, smem);
  d = block_sum(d, smem);
  w = block_sum(w, smem);
  if (threadIdx.x == 0) {
    const double frac = m / static_cast<double>(n);
    accuracy[blockIdx.x] = fmax(frac, 1.0 - frac);
    overlap[blockIdx.x] = fabs(d / w);
  }
}

// Position of column `col` inside row `row` (columns ascending), or -1.
__device__ __forceinline__ int64_t row_find(const int64_t *indptr, const int32_t *indices, int64_t row, int32_t col) {
  int64_t lo = indptr[row], hi = indptr[row + 1];
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&indices[mid]) < col)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < indptr[row + 1] && indices[lo] == col) ? lo : -1;
}

template <bool kApply>
__global__ void __launch_bounds__(kRedThreads) symmetrize_kernel(uint64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                 double *__restrict__ data, unsigned long long *__restrict__ missing) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  if (i >= n) return;
  for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
    const int32_t j = indices[k];
    if (static_cast<uint64_t>(j) == i) continue;
    if (kApply) {
      if (static_cast<uint64_t>(j) < i) continue;  // the (i < j) owner writes both triangles
      const int64_t t = row_find(indptr, indices, j, static_cast<int32_t>(i));
      const double avg = 0.5 * (data[k] + data[t]);
      data[k] = avg;
      data[t] = avg;
    } else {
      if (row_find(indptr, indices, j, static_cast<int32_t>(i)) < 0) atomicAdd(missing, 1ull);
    }
  }
}


// ---- cluster sparsification helpers -----------------------------------------------
// strongest off-diagonal coupling of every row (common.py:525-541)
__global__ void __launch_bounds__(kRedThreads) strongest_offdiag_kernel(uint64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                        const double *__restrict__ data, double *__restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  if (i >= n) return;
  double best = 0.0;
  for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
    if (static_cast<uint64_t>(indices[k]) != i) best = fmax(best, fabs(data[k]));
  out[i] = best;
}

// max |data| (non-negative doubles order like their bit patterns)
__global__ void __launch_bounds__(kRedThreads) max_abs_kernel(uint64_t m, const double *__restrict__ data, unsigned long long *__restrict__ out) {
  unsigned long long best = 0;
  for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x; k < m; k += static_cast<uint64_t>(gridDim.x) * kRedThreads)
    best = max(best, static_cast<unsigned long long>(__double_as_longlong(fabs(data[k]))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

// an entry survives the global cutoff (common.py:621-631) when it is non-zero and either not below
// reltol * max|J| or between two frozen spins
__device__ __forceinline__ bool survives_cutoff(double v, double cutoff, bool both_frozen) { return v != 0.0 && (both_frozen || !(fabs(v) < cutoff)); }

// connected components of the surviving couplings: minimum-label propagation + pointer jumping;
// labels only ever decrease and stay inside the component, so unsynchronised updates are benign.
// Converges to label = smallest vertex of the component.
__global__ void __launch_bounds__(kRedThreads) cc_propagate_kernel(uint64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                   const double *__restrict__ data, const unsigned long long *__restrict__ max_bits,
                                                                   double reltol, const unsigned char *__restrict__ frozen, int32_t *labels,
                                                                   unsigned int *__restrict__ changed) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  if (i >= n) return;
  const double cutoff = reltol * __longlong_as_double(static_cast<long long>(*max_bits));
  const int32_t mine = labels[i];
  int32_t best = mine;
  const bool fi = frozen && frozen[i];
  for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
    const int32_t j = indices[k];
    if (survives_cutoff(data[k], cutoff, fi && frozen[j])) best = min(best, labels[j]);
  }
  if (best < mine) {
    atomicMin(&labels[i], best);
    atomicMin(&labels[mine], best);  // hook the old representative as well
    *changed = 1u;
  }
}

__global__ void __launch_bounds__(kRedThreads) cc_jump_kernel(uint64_t n, int32_t *labels) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  if (i >= n) return;
  int32_t l = labels[i];
  for (;;) {
    const int32_t up = labels[l];
    if (up == l) break;
    l = up;
  }
  labels[i] = l;
}

__global__ void __launch_bounds__(kRedThreads) iota_kernel(uint64_t n, int32_t *labels) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kRedThreads + threadIdx.x;
  if (i < n) labels[i] = static_cast<int32_t>(i);
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_energy(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices, double const *d_data,
               double const *d_field, uint32_t num_replicas, uint64_t const *d_bits, double *d_energy, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  if (num_replicas == 0) return ASP_OK;
  ASP_REQUIRE(d_energy && d_bits, "NULL buffer");
  if (n == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_energy, 0, num_replicas * sizeof(double), s));
    return ASP_OK;
  }
  ASP_REQUIRE(num_replicas <= 65535, "at most 65535 replicas per call");
  const uint64_t tiles = (n + kRedThreads - 1) / kRedThreads;
  double *partial = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&partial), tiles * num_replicas * sizeof(double), s));
  energy_partial_kernel<<<dim3(static_cast<unsigned>(tiles), num_replicas), kRedThreads, 0, s>>>(
      n, d_indptr, d_indices, d_data, d_field, d_bits, (n + 63) / 64, partial);
  ASP_LAUNCH_CHECK();
  sum_partials_kernel<<<num_replicas, kRedThreads, 0, s>>>(partial, tiles, d_energy);
  ASP_LAUNCH_CHECK();
  ASP_CUDA_CHECK(cudaFreeAsync(partial, s));
  return ASP_OK;
}

int asp_accuracy_overlap(uint64_t n, uint32_t num_replicas, uint64_t const *d_predicted, uint64_t const *d_exact,
                         double const *d_weights, double *d_accuracy, double *d_overlap, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  if (num_replicas == 0) return ASP_OK;
  ASP_REQUIRE(n > 0, "n must be positive");
  ASP_REQUIRE(d_predicted && d_exact && d_accuracy && d_overlap, "NULL buffer");
  ASP_REQUIRE(num_replicas <= 65535, "at most 65535 replicas per call");
  const uint64_t tiles = (n + kRedThreads - 1) / kRedThreads;
  double *partial = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&partial), tiles * num_replicas * 3 * sizeof(double), s));
  overlap_partial_kernel<<<dim3(static_cast<unsigned>(tiles), num_replicas), kRedThreads, 0, s>>>(
      n, d_predicted, d_exact, d_weights, (n + 63) / 64, partial);
  ASP_LAUNCH_CHECK();
  overlap_final_kernel<<<num_replicas, kRedThreads, 0, s>>>(partial, tiles, n, d_accuracy, d_overlap);
  ASP_LAUNCH_CHECK();
  ASP_CUDA_CHECK(cudaFreeAsync(partial, s));
  return ASP_OK;
}

int asp_csr_symmetrize(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices, double *d_data,
                       uint64_t *h_asymmetric, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  ASP_REQUIRE(h_asymmetric != nullptr, "h_asymmetric is NULL");
  *h_asymmetric = 0;
  if (n == 0) return ASP_OK;
  unsigned long long *missing = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&missing), sizeof(unsigned long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(missing, 0, sizeof(unsigned long long), s));
  const unsigned blocks = static_cast<unsigned>((n + kRedThreads - 1) / kRedThreads);
  symmetrize_kernel<false><<<blocks, kRedThreads, 0, s>>>(n, d_indptr, d_indices, d_data, missing);
  ASP_LAUNCH_CHECK();
  unsigned long long h = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&h, missing, sizeof(h), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  ASP_CUDA_CHECK(cudaFreeAsync(missing, s));
  *h_asymmetric = h;
  if (h != 0) return ASP_OK;
  symmetrize_kernel<true><<<blocks, kRedThreads, 0, s>>>(n, d_indptr, d_indices, d_data, nullptr);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

int asp_csr_strongest_offdiag(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices, double const *d_data,
                              double *d_out, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  if (n == 0) return ASP_OK;
  ASP_REQUIRE(d_indptr && d_out, "NULL buffer");
  strongest_offdiag_kernel<<<static_cast<unsigned>((n + kRedThreads - 1) / kRedThreads), kRedThreads, 0, s>>>(n, d_indptr, d_indices, d_data, d_out);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

int asp_cutoff_components(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices, double const *d_data, uint64_t nnz,
                          double reltol, unsigned char const *d_frozen, int32_t *d_labels, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  if (n == 0) return ASP_OK;
  ASP_REQUIRE(d_indptr && d_labels, "NULL buffer");
  ASP_REQUIRE(n < (1ull << 31), "labels are int32");
  unsigned long long *scratch = nullptr;  // [0] max |J| bits, [1] changed flag
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&scratch), 2 * sizeof(unsigned long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 2 * sizeof(unsigned long long), s));
  const unsigned blocks = static_cast<unsigned>((n + kRedThreads - 1) / kRedThreads);
  if (nnz) {
    max_abs_kernel<<<static_cast<unsigned>(std::min<uint64_t>((nnz + kRedThreads - 1) / kRedThreads, 4 * 148)), kRedThreads, 0, s>>>(nnz, d_data, scratch);
    ASP_LAUNCH_CHECK();
  }
  iota_kernel<<<blocks, kRedThreads, 0, s>>>(n, d_labels);
  ASP_LAUNCH_CHECK();
  unsigned int *changed = reinterpret_cast<unsigned int *>(scratch + 1);
  for (int round = 0;; ++round) {
    ASP_CUDA_CHECK(cudaMemsetAsync(changed, 0, sizeof(unsigned int), s));
    cc_propagate_kernel<<<blocks, kRedThreads, 0, s>>>(n, d_indptr, d_indices, d_data, scratch, reltol, d_frozen, d_labels, changed);
    ASP_LAUNCH_CHECK();
    cc_jump_kernel<<<blocks, kRedThreads, 0, s>>>(n, d_labels);
    ASP_LAUNCH_CHECK();
    unsigned int h = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&h, changed, sizeof(h), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (!h) break;
    if (round > 100000) {
      cudaFreeAsync(scratch, s);
      asp::set_error("connected components did not converge");
      return ASP_ERR_CUDA;
    }
  }
  ASP_CUDA_CHECK(cudaFreeAsync(scratch, s));
  return ASP_OK;
}


}  // extern "C"
