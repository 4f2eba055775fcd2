// Sampling front-end of the extraction: the two data-parallel pieces that feed n
// and psi into the hot path.
//
//   asp_batched_index   lattice_symmetries' basis.batched_index / ls.batched_index as the reference
//                       calls it (annealing_sign_problem/common.py:283, :817;
//                       experiments/sampled_connected_components.py:720, :730): position of every
//                       needle in the sorted basis (3.15e7 representatives for kagome_36), -1 = absent.
//   asp_sample_indices  monte_carlo_sampling (common.py:269-278) = legacy np.random.choice(n, m,
//                       replace=True, p): cdf = cumsum(p); cdf /= cdf[-1]; index = searchsorted(cdf,
//                       u, side="right") for uniform draws u the CALLER takes from numpy's global
//                       stream (so a seeded run draws the same numbers as the reference).
//
// Both are searches in sorted arrays: a 2^k-entry first-level table is not worth its build here
// (one-shot calls, needles << basis), so the search is a branch-free halving with the top of the
// tree served from L2.  The cumulative sum is a three-kernel scan (block scan, scan of the block
// totals, fix-up + normalisation); its rounding differs from numpy's sequential cumsum by a few
// ulp, which moves a sample only when u lies within ~1e-16 of a bin edge.
#include "common.cuh"

namespace asp {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ double weight_of(double psi, double power, int mode) {
  const double a = fabs(psi);
  return mode == 2 ? a * a : mode == 1 ? a : pow(a, power);  // numpy squares for exponent 2 as well
}

__device__ __forceinline__ double block_inclusive_scan_f64(double v, double *total, double *smem /*[33]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    const double w = lane < nw ? smem[lane] : 0.0;
    double wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < nw) smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  const double out = smem[warp] + incl;
  *total = smem[32];
  __syncthreads();
  return out;
}

// cdf[i] = inclusive sum of the weights inside the tile; tile_total[b] = sum of tile b
__global__ void __launch_bounds__(kScanThreads) cdf_tile_kernel(const double *__restrict__ psi, uint64_t n, double power, int mode,
                                                                double *__restrict__ cdf, double *__restrict__ tile_total) {
  __shared__ double smem[33];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
  double w[kScanItems], run = 0.0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    w[k] = base + k < n ? weight_of(psi[base + k], power, mode) : 0.0;
    run += w[k];
    w[k] = run;
  }
  double total;
  const double before = block_inclusive_scan_f64(run, &total, smem) - run;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) cdf[base + k] = before + w[k];
  if (threadIdx.x == 0) tile_total[blockIdx.x] = total;
}

// exclusive scan of the tile totals in place (one CTA, any number of tiles); total[tiles] = grand total
// and total[tiles + 1] = the value the fix-up gives the LAST cdf entry (same expression, so cdf[-1] == 1)
__global__ void __launch_bounds__(1024) cdf_totals_kernel(double *tile_total, uint64_t tiles, const double *cdf, uint64_t n) {
  __shared__ double smem[33];
  __shared__ double carry;
  if (threadIdx.x == 0) carry = 0.0;
  __syncthreads();
  for (uint64_t start = 0; start < tiles; start += 1024) {
    const uint64_t i = start + threadIdx.x;
    const double v = i < tiles ? tile_total[i] : 0.0;
    double total;
    const double incl = block_inclusive_scan_f64(v, &total, smem);
    const double c = carry;
    if (i < tiles) tile_total[i] = c + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_total[tiles] = carry;
    tile_total[tiles + 1] = tile_total[tiles - 1] + cdf[n - 1];
  }
}

__global__ void __launch_bounds__(kScanThreads) cdf_fix_kernel(double *__restrict__ cdf, uint64_t n, const double *__restrict__ tile_total,
                                                               uint64_t tiles) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kScanThreads + threadIdx.x;
  if (i >= n) return;
  cdf[i] = (tile_total[i / kScanTile] + cdf[i]) / tile_total[tiles + 1];  // cdf /= cdf[-1] (np.random.choice); the last entry is exactly 1
}

// index[j] = number of cdf entries <= u[j]  (searchsorted side = "right"), clipped to n-1
__global__ void __launch_bounds__(256) sample_search_kernel(const double *__restrict__ cdf, uint64_t n, const double *__restrict__ u, uint64_t m,
                                                            int64_t *__restrict__ index) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (j >= m) return;
  const double x = u[j];
  uint64_t lo = 0, len = n;
  while (len > 0) {
    const uint64_t half = len >> 1;
    const bool right = __ldg(&cdf[lo + half]) <= x;
    lo = right ? lo + half + 1 : lo;
    len = right ? len - half - 1 : half;
  }
  index[j] = static_cast<int64_t>(lo < n ? lo : n - 1);
}

__global__ void __launch_bounds__(256) batched_index_kernel(const uint64_t *__restrict__ sorted, uint64_t n, const uint64_t *__restrict__ needles,
                                                            uint64_t m, int64_t *__restrict__ index, unsigned long long *missing) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (j >= m) return;
  const uint64_t key = needles[j];
  uint64_t lo = 0, len = n;
  while (len > 0) {  // lower bound
    const uint64_t half = len >> 1;
    const bool right = __ldg(&sorted[lo + half]) < key;
    lo = right ? lo + half + 1 : lo;
    len = right ? len - half - 1 : half;
  }
  const bool found = lo < n && __ldg(&sorted[lo]) == key;
  index[j] = found ? static_cast<int64_t>(lo) : -1;
  if (!found) atomicAdd(missing, 1ull);
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_batched_index(uint64_t n, uint64_t const *d_sorted, uint64_t m, uint64_t const *d_needles, int64_t *d_index,
                      uint64_t *h_missing, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_REQUIRE(m == 0 || (d_needles && d_index), "asp_batched_index: NULL buffer");
  ASP_REQUIRE(n == 0 || d_sorted, "asp_batched_index: NULL basis");
  ASP_CUDA_CHECK(keep_pool_memory());
  unsigned long long *d_missing = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(&d_missing, sizeof(unsigned long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(d_missing, 0, sizeof(unsigned long long), s));
  if (m > 0) {
    batched_index_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(d_sorted, n, d_needles, m, d_index, d_missing);
    ASP_LAUNCH_CHECK();
  }
  unsigned long long missing = 0;
  if (h_missing) {
    ASP_CUDA_CHECK(cudaMemcpyAsync(&missing, d_missing, sizeof(missing), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    *h_missing = missing;
  }
  ASP_CUDA_CHECK(cudaFreeAsync(d_missing, s));
  return ASP_OK;
}

int asp_sample_indices(uint64_t n, double const *d_psi, double power, uint64_t m, double const *d_uniform, int64_t *d_index,
                       double *d_cdf, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_REQUIRE(n > 0 && d_psi != nullptr, "asp_sample_indices: empty distribution");
  ASP_REQUIRE(m == 0 || (d_uniform && d_index), "asp_sample_indices: NULL buffer");
  ASP_CUDA_CHECK(keep_pool_memory());
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  double *cdf = d_cdf, *totals = nullptr;
  if (!cdf) ASP_CUDA_CHECK(cudaMallocAsync(&cdf, n * sizeof(double), s));
  ASP_CUDA_CHECK(cudaMallocAsync(&totals, (tiles + 2) * sizeof(double), s));
  const int mode = power == 2.0 ? 2 : power == 1.0 ? 1 : 0;
  cdf_tile_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(d_psi, n, power, mode, cdf, totals);
  ASP_LAUNCH_CHECK();
  cdf_totals_kernel<<<1, 1024, 0, s>>>(totals, tiles, cdf, n);
  ASP_LAUNCH_CHECK();
  cdf_fix_kernel<<<static_cast<unsigned>((n + kScanThreads - 1) / kScanThreads), kScanThreads, 0, s>>>(cdf, n, totals, tiles);
  ASP_LAUNCH_CHECK();
  if (m > 0) {
    sample_search_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(cdf, n, d_uniform, m, d_index);
    ASP_LAUNCH_CHECK();
  }
  ASP_CUDA_CHECK(cudaFreeAsync(totals, s));
  if (!d_cdf) ASP_CUDA_CHECK(cudaFreeAsync(cdf, s));
  return ASP_OK;
}

}  // extern "C"
