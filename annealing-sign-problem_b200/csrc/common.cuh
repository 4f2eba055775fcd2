// Shared helpers for the sm_100a kernels of the extraction + annealing hot path.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/asp_b200.h"

namespace asp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// -- error channel ------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define ASP_CUDA_CHECK(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::asp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                       cudaGetErrorString(_e));                                      \
      return ASP_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define ASP_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    ::asp::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    ASP_CUDA_CHECK(cudaGetLastError());                                              \
  } while (0)

#define ASP_REQUIRE(cond, msg)                                                       \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ::asp::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                        \
      return ASP_ERR_ARG;                                                            \
    }                                                                                \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// -- device helpers -----------------------------------------------------------------
__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t *p) { return __ldg(p); }
__device__ __forceinline__ double ld_nc_f64(const double *p) { return __ldg(p); }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan across a CTA (blockDim.x multiple of 32, <= 1024); smem: 33 x int64.
// Every thread of the CTA must call it.
__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v, int64_t *total, int64_t *smem /*[33]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    int64_t w = lane < nw ? smem[lane] : 0;
    int64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < nw) smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  const int64_t base = smem[warp];
  *total = smem[32];
  __syncthreads();
  return base + incl - v;
}

// The stream-ordered allocations (cudaMallocAsync) of a call are returned to the default pool at
// its end; keep them cached there instead of handing them back to the driver at every
// synchronisation.  Call once at the top of every entry point that allocates.
cudaError_t keep_pool_memory();

// Internal scan entry (extract.cu) reused by other translation units.
int scan_exclusive_i64(const int64_t *d_in, int64_t *d_out, uint64_t m, void *d_tmp, cudaStream_t s);
size_t scan_tmp_bytes(uint64_t m);

}  // namespace asp
