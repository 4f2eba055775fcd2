// Replica-parallel simulated annealing on sm_100a.
//
// Replaces ising_glass_annealer.anneal as the reference calls it
// (annealing_sign_problem/common.py:242-248; experiments/full_hilbert_space.py:212-218).
//
// Chain definition (DESIGN.md "SA chain definition", mirrored by oracle/anneal_port.c):
//   sweep t visits positions p = 0..n_padded-1 of the RELABELLED model in order;
//   dE = -s_p (4 sum_{j != p} J_pj s_j + 2 h_p);  accept iff dE <= 0 or
//   (beta_t dE < 23 and u < exp_neg(beta_t dE)),  u = (philox(p, t, replica) + 1/2) 2^-32.
//   After each sweep a replica snapshots its configuration when its running energy
//   (fixed point, exact integer sums) is the lowest so far.
//
// Parallelisation: the relabelling puts every colour class of the coupling graph into a
// contiguous, 4-aligned range of positions.  Spins of one class do not interact, so
// updating a whole class concurrently IS the sequential sweep.  Layout: one lane = one
// replica, one warp-task = 4 consecutive positions of a class for 32 replicas; the spins of
// 32 replicas at one position are ONE 32-bit word, so a CSR row is read once per 32
// replicas and the local field is summed in f64 in stored row order (bitwise the oracle's).
// A team of CTAs owns one group of 32 replicas and synchronises between classes.
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

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
  std::vector<int64_t> class_ptr;
};

namespace asp {

// ---- colouring (Jones-Plassmann with hashed priorities; = greedy in priority order) ----
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool prio_less(uint32_t a, uint32_t b) {  // (hash, index) order
  const uint32_t ha = hash_u32(a), hb = hash_u32(b);
  return ha < hb || (ha == hb && a < b);
}

// One round: reads the colours of the previous round only (ping-pong), so the result does
// not depend on thread timing.
__global__ void __launch_bounds__(256) colour_round_kernel(uint32_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                           const int32_t *__restrict__ colour_in, int32_t *__restrict__ colour_out,
                                                           unsigned long long *__restrict__ remaining) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const int32_t mine = colour_in[i];
  if (mine >= 0) {
    colour_out[i] = mine;
    return;
  }
  const int64_t b = indptr[i], e = indptr[i + 1];
  for (int64_t k = b; k < e; ++k) {
    const uint32_t j = static_cast<uint32_t>(indices[k]);
    if (j == i) continue;
    if (colour_in[j] < 0 && prio_less(i, j)) {  // an uncoloured neighbour outranks us: wait
      colour_out[i] = -1;
      atomicAdd(remaining, 1ull);
      return;
    }
  }
  for (int32_t base = 0;; base += 64) {
    uint64_t used = 0;
    for (int64_t k = b; k < e; ++k) {
      const uint32_t j = static_cast<uint32_t>(indices[k]);
      if (j == i) continue;
      const int32_t c = colour_in[j];
      if (c >= base && c < base + 64) used |= 1ull << (c - base);
    }
    if (~used) {
      colour_out[i] = base + __ffsll(static_cast<long long>(~used)) - 1;
      return;
    }
  }
}

// ---- relabelled CSR -----------------------------------------------------------------
__global__ void __launch_bounds__(256) relabel_count_kernel(uint64_t n_padded, const int32_t *__restrict__ order, const int64_t *__restrict__ indptr0,
                                                            const int32_t *__restrict__ indices0, int64_t *__restrict__ row_len) {
  const uint64_t p = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n_padded) return;
  const int32_t i = order[p];
  int64_t len = 0;
  if (i >= 0)
    for (int64_t k = indptr0[i]; k < indptr0[i + 1]; ++k) len += indices0[k] != i;
  row_len[p] = len;
}

__global__ void __launch_bounds__(256) relabel_fill_kernel(uint64_t n_padded, const int32_t *__restrict__ order, const int32_t *__restrict__ position,
                                                           const int64_t *__restrict__ indptr0, const int32_t *__restrict__ indices0,
                                                           const double *__restrict__ data0, const double *__restrict__ field0,
                                                           const int64_t *__restrict__ indptr, int32_t *__restrict__ indices,
                                                           double *__restrict__ data, double *__restrict__ field) {
  const uint64_t p = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= n_padded) return;
  const int32_t i = order[p];
  field[p] = (i >= 0 && field0) ? field0[i] : 0.0;
  if (i < 0) return;
  int64_t out = indptr[p];
  for (int64_t k = indptr0[i]; k < indptr0[i + 1]; ++k) {
    const int32_t j = indices0[k];
    if (j == i) continue;
    indices[out] = position[j];
    data[out] = data0[k];
    ++out;
  }
}

// ---- RNG + deterministic exp (same operation sequence as oracle/anneal_port.c) --------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ double exp_neg(double x) {
  const double t = __dmul_rn(x, 1.4426950408889634);
  const double kf = floor(t);
  const double z = __dmul_rn(__dadd_rn(__dadd_rn(t, -kf), -0.5), 0.6931471805599453);
  const double w = -z;
  double p = 1.0 / 6227020800.0;
  p = __fma_rn(p, w, 1.0 / 479001600.0);
  p = __fma_rn(p, w, 1.0 / 39916800.0);
  p = __fma_rn(p, w, 1.0 / 3628800.0);
  p = __fma_rn(p, w, 1.0 / 362880.0);
  p = __fma_rn(p, w, 1.0 / 40320.0);
  p = __fma_rn(p, w, 1.0 / 5040.0);
  p = __fma_rn(p, w, 1.0 / 720.0);
  p = __fma_rn(p, w, 1.0 / 120.0);
  p = __fma_rn(p, w, 1.0 / 24.0);
  p = __fma_rn(p, w, 1.0 / 6.0);
  p = __fma_rn(p, w, 0.5);
  p = __fma_rn(p, w, 1.0);
  p = __fma_rn(p, w, 1.0);
  p = __dmul_rn(p, 0.7071067811865476);
  const double scale = __longlong_as_double((1023ll - static_cast<long long>(kf)) << 52);
  return __dmul_rn(p, scale);
}

constexpr double kRejectAbove = 23.0;  // exp(-23) < 2^-33, the smallest uniform variate

// ---- initial configuration ----------------------------------------------------------
// words[g][p]: bit `lane` = spin of replica 32 g + lane at position p.
__global__ void __launch_bounds__(256) sa_init_kernel(uint64_t n_padded, uint32_t groups, uint32_t replica_offset, uint64_t seed, const int32_t *__restrict__ order,
                                                      const uint64_t *__restrict__ x0, uint32_t *__restrict__ words) {
  const uint64_t tasks = n_padded / 4;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= tasks * groups) return;
  const uint32_t g = static_cast<uint32_t>(warp / tasks);
  const uint64_t p0 = (warp % tasks) * 4;
  uint32_t w[4];
  if (x0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int32_t i = order[p0 + j];
      w[j] = (i >= 0 && ((x0[i >> 6] >> (i & 63)) & 1)) ? 0xFFFFFFFFu : 0u;
    }
  } else {
    const uint32_t r = replica_offset + g * 32 + lane;
    const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(p0 >> 2), 0xFFFFFFFFu, r, static_cast<uint32_t>(p0 >> 34)),
                                    make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    w[0] = __ballot_sync(0xffffffffu, rnd.x & 1);
    w[1] = __ballot_sync(0xffffffffu, rnd.y & 1);
    w[2] = __ballot_sync(0xffffffffu, rnd.z & 1);
    w[3] = __ballot_sync(0xffffffffu, rnd.w & 1);
  }
  if (lane == 0) *reinterpret_cast<uint4 *>(words + static_cast<uint64_t>(g) * n_padded + p0) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- the sweep kernel -----------------------------------------------------------------
struct SaArgs {
  uint64_t n_padded;
  const int64_t *indptr;
  const int32_t *indices;
  const double *data;
  const double *field;
  const int64_t *class_ptr;
  uint32_t num_classes;
  uint32_t groups;      // replica groups of 32
  uint32_t replica_offset;  // global index of replica 0 of this launch
  uint32_t team_size;   // CTAs per group (>= 1)
  uint32_t num_teams;
  uint32_t num_sweeps;
  const double *betas;
  uint64_t seed;
  double escale;
  uint32_t *words;       // [groups][n_padded] current
  uint32_t *best_words;  // [groups][n_padded]
  long long *rel;        // [groups*32] running fixed-point energy (zeroed by the host)
  long long *best_rel;   // [groups*32] out
  unsigned long long *barriers;  // [num_teams] zeroed by the host
};

constexpr int kSaThreads = 512;
constexpr int kSaWarps = kSaThreads / 32;

struct TeamBarrier {
  unsigned long long *counter;
  unsigned long long target;
  uint32_t team_size;
  __device__ __forceinline__ void sync() {
    if (team_size == 1) {
      __syncthreads();
      return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      target += team_size;
      __threadfence();
      atomicAdd(counter, 1ull);
      while (*reinterpret_cast<volatile unsigned long long *>(counter) < target) {
      }
      __threadfence();
    }
    __syncthreads();
  }
};

struct __align__(16) StagedEntry {
  double val;
  uint32_t word;
  uint32_t pad;
};

__global__ void __launch_bounds__(kSaThreads, 1) sa_sweep_kernel(const SaArgs a) {
  __shared__ StagedEntry s_stage[kSaWarps][32];
  const uint32_t lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
  const uint32_t team = blockIdx.x / a.team_size;
  if (team >= a.num_teams) return;
  const uint32_t member = blockIdx.x % a.team_size;
  const uint32_t team_warps = a.team_size * kSaWarps;
  const uint32_t my_warp = member * kSaWarps + warp_in_cta;
  TeamBarrier bar{a.barriers + team, 0ull, a.team_size};
  StagedEntry *stage = s_stage[warp_in_cta];
  const uint2 key = make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));

  for (uint32_t g = team; g < a.groups; g += a.num_teams) {
    uint32_t *words = a.words + static_cast<uint64_t>(g) * a.n_padded;
    uint32_t *best = a.best_words + static_cast<uint64_t>(g) * a.n_padded;
    const uint32_t replica = g * 32 + lane;                  // local slot
    const uint32_t stream_id = a.replica_offset + replica;   // global replica: RNG stream
    long long best_rel = 0;  // every warp of the team tracks the same value
    for (uint32_t t = 0; t < a.num_sweeps; ++t) {
      const double beta = a.betas[t];
      long long rel_delta = 0;
      for (uint32_t c = 0; c < a.num_classes; ++c) {
        const uint64_t q_begin = static_cast<uint64_t>(a.class_ptr[c]) >> 2, q_end = static_cast<uint64_t>(a.class_ptr[c + 1]) >> 2;
        for (uint64_t q = q_begin + my_warp; q < q_end; q += team_warps) {
          const uint64_t p0 = q * 4;
          const uint4 cur = __ldcg(reinterpret_cast<const uint4 *>(words + p0));
          const uint32_t cur_w[4] = {cur.x, cur.y, cur.z, cur.w};
          double dE[4];
          bool need_rng = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t e_begin = a.indptr[p0 + j], e_end = a.indptr[p0 + j + 1];
            double acc = 0.0;
            for (int64_t chunk = e_begin; chunk < e_end; chunk += 32) {
              const int64_t e = chunk + lane;
              StagedEntry se;
              se.val = 0.0;
              se.word = 0;
              se.pad = 0;
              if (e < e_end) {
                se.val = __ldg(&a.data[e]);
                se.word = __ldcg(&words[__ldg(&a.indices[e])]);
              }
              __syncwarp();
              stage[lane] = se;
              __syncwarp();
              const int cnt = static_cast<int>(min(static_cast<int64_t>(32), e_end - chunk));
              for (int k = 0; k < cnt; ++k) {
                const StagedEntry x = stage[k];
                acc = __dadd_rn(acc, ((x.word >> lane) & 1) ? x.val : -x.val);
              }
            }
            const double gsum = __dadd_rn(__dmul_rn(4.0, acc), __dmul_rn(2.0, a.field[p0 + j]));
            dE[j] = ((cur_w[j] >> lane) & 1) ? -gsum : gsum;
            need_rng |= dE[j] > 0.0 && __dmul_rn(beta, dE[j]) < kRejectAbove;
          }
          uint32_t rnd[4] = {0, 0, 0, 0};
          if (__any_sync(0xffffffffu, need_rng)) {
            const uint4 r4 = philox4x32_10(make_uint4(static_cast<uint32_t>(q), t, stream_id, static_cast<uint32_t>(p0 >> 34)), key);
            rnd[0] = r4.x;
            rnd[1] = r4.y;
            rnd[2] = r4.z;
            rnd[3] = r4.w;
          }
          uint32_t new_w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bool accept;
            if (dE[j] <= 0.0) {
              accept = true;
            } else {
              const double x = __dmul_rn(beta, dE[j]);
              if (x >= kRejectAbove) {
                accept = false;
              } else {
                const double u = __dmul_rn(__dadd_rn(static_cast<double>(rnd[j]), 0.5), 2.3283064365386963e-10);
                accept = u < exp_neg(x);
              }
            }
            if (accept) rel_delta += __double2ll_rn(__dmul_rn(dE[j], a.escale));
            new_w[j] = cur_w[j] ^ __ballot_sync(0xffffffffu, accept);
          }
          if (lane == 0) __stcg(reinterpret_cast<uint4 *>(words + p0), make_uint4(new_w[0], new_w[1], new_w[2], new_w[3]));
        }
        bar.sync();
      }
      // end of sweep: publish running energies, snapshot improved replicas
      if (rel_delta != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&a.rel[replica]), static_cast<unsigned long long>(rel_delta));
      bar.sync();
      const long long rel_now = __ldcg(&a.rel[replica]);
      const bool improved = rel_now < best_rel;
      if (improved) best_rel = rel_now;
      const uint32_t mask = __ballot_sync(0xffffffffu, improved);
      if (mask) {  // uniform across the team: every warp reads the same rel[]
        const uint64_t stride = static_cast<uint64_t>(team_warps) * 32;
        for (uint64_t p = static_cast<uint64_t>(my_warp) * 32 + lane; p < a.n_padded; p += stride)
          best[p] = (best[p] & ~mask) | (__ldcg(&words[p]) & mask);
        bar.sync();
      }
    }
    if (member == 0 && warp_in_cta == 0) a.best_rel[replica] = best_rel;
  }
}

// ---- outputs: relabelled [groups][n_padded] words -> original-order packed [R][words64] --
__global__ void __launch_bounds__(256) sa_unpermute_kernel(uint64_t n, uint64_t n_padded, uint32_t num_replicas, const int32_t *__restrict__ position,
                                                           const uint32_t *__restrict__ best_words, uint64_t *__restrict__ out) {
  const uint64_t words64 = (n + 63) / 64;
  const uint64_t idx = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= words64 * num_replicas) return;
  const uint32_t r = static_cast<uint32_t>(idx / words64);
  const uint64_t w = idx % words64;
  const uint32_t *src = best_words + static_cast<uint64_t>(r >> 5) * n_padded;
  uint64_t bits = 0;
  for (int b = 0; b < 64; ++b) {
    const uint64_t i = w * 64 + b;
    if (i >= n) break;
    bits |= static_cast<uint64_t>((src[position[i]] >> (r & 31)) & 1) << b;
  }
  out[idx] = bits;
}

}  // namespace asp

using namespace asp;

extern "C" {

void asp_sa_plan_destroy(asp_sa_plan *plan) {
  if (!plan) return;
  for (void *p : {static_cast<void *>(plan->d_order), static_cast<void *>(plan->d_position), static_cast<void *>(plan->d_indptr),
                  static_cast<void *>(plan->d_indices), static_cast<void *>(plan->d_data), static_cast<void *>(plan->d_field),
                  static_cast<void *>(plan->d_class_ptr)})
    if (p) cudaFree(p);
  delete plan;
}

int asp_sa_plan_create(asp_sa_plan **out, uint64_t n, int64_t const *d_indptr, int32_t const *d_indices,
                       double const *d_data, double const *d_field, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_REQUIRE(out != nullptr, "out is NULL");
  ASP_REQUIRE(n > 0 && n < (1ull << 31), "n must be in [1, 2^31)");
  ASP_REQUIRE(d_indptr && d_indices && d_data, "NULL CSR arrays");
  auto *plan = new asp_sa_plan();
  plan->n = n;
  plan->d_indptr0 = d_indptr;
  plan->d_indices0 = d_indices;
  plan->d_data0 = d_data;
  plan->d_field0 = d_field;
  struct Guard {
    asp_sa_plan *p;
    ~Guard() {
      if (p) asp_sa_plan_destroy(p);
    }
  } guard{plan};

  // 1. colour the coupling graph on the device
  int32_t *d_colour[2] = {nullptr, nullptr};
  unsigned long long *d_remaining = nullptr;
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_colour[0]), n * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_colour[1]), n * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_remaining), sizeof(unsigned long long)));
  ASP_CUDA_CHECK(cudaMemsetAsync(d_colour[0], 0xFF, n * sizeof(int32_t), s));
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  int cur = 0;
  for (int round = 0;; ++round) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_remaining, 0, sizeof(unsigned long long), s));
    colour_round_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), d_indptr, d_indices, d_colour[cur], d_colour[cur ^ 1], d_remaining);
    ASP_LAUNCH_CHECK();
    cur ^= 1;
    unsigned long long remaining = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&remaining, d_remaining, sizeof(remaining), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (remaining == 0) break;
    if (round > 100000) {
      cudaFree(d_colour[0]);
      cudaFree(d_colour[1]);
      cudaFree(d_remaining);
      set_error("colouring did not converge");
      return ASP_ERR_CUDA;
    }
  }
  // 2. stable counting sort by colour on the host (O(n) bookkeeping): positions
  std::vector<int32_t> colour(n);
  ASP_CUDA_CHECK(cudaMemcpy(colour.data(), d_colour[cur], n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  cudaFree(d_colour[0]);
  cudaFree(d_colour[1]);
  cudaFree(d_remaining);
  int32_t max_colour = 0;
  for (int32_t c : colour) max_colour = std::max(max_colour, c);
  const uint32_t classes = static_cast<uint32_t>(max_colour) + 1;
  std::vector<int64_t> size(classes, 0);
  for (int32_t c : colour) ++size[c];
  plan->class_ptr.assign(classes + 1, 0);
  for (uint32_t c = 0; c < classes; ++c) plan->class_ptr[c + 1] = plan->class_ptr[c] + (size[c] + 3) / 4 * 4;
  plan->num_classes = classes;
  plan->n_padded = static_cast<uint64_t>(plan->class_ptr[classes]);
  std::vector<int32_t> order(plan->n_padded, -1), position(n);
  {
    std::vector<int64_t> cursor(plan->class_ptr.begin(), plan->class_ptr.end() - 1);
    for (uint64_t i = 0; i < n; ++i) {
      const int64_t p = cursor[colour[i]]++;
      order[p] = static_cast<int32_t>(i);
      position[i] = static_cast<int32_t>(p);
    }
  }
  const uint64_t np = plan->n_padded;
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_order), np * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_position), n * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_class_ptr), (classes + 1) * sizeof(int64_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_indptr), (np + 1) * sizeof(int64_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_field), np * sizeof(double)));
  ASP_CUDA_CHECK(cudaMemcpyAsync(plan->d_order, order.data(), np * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  ASP_CUDA_CHECK(cudaMemcpyAsync(plan->d_position, position.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  ASP_CUDA_CHECK(cudaMemcpyAsync(plan->d_class_ptr, plan->class_ptr.data(), (classes + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  // 3. relabelled CSR without the diagonal (the diagonal never enters dE)
  int64_t *d_len = nullptr;
  void *d_tmp = nullptr;
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_len), np * sizeof(int64_t)));
  ASP_CUDA_CHECK(cudaMalloc(&d_tmp, scan_tmp_bytes(np)));
  const unsigned pblocks = static_cast<unsigned>((np + 255) / 256);
  relabel_count_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_order, d_indptr, d_indices, d_len);
  ASP_LAUNCH_CHECK();
  int rc = scan_exclusive_i64(d_len, plan->d_indptr, np, d_tmp, s);
  if (rc != ASP_OK) return rc;
  int64_t nnz = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&nnz, plan->d_indptr + np, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(d_len);
  cudaFree(d_tmp);
  plan->nnz = static_cast<uint64_t>(nnz);
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_indices), std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&plan->d_data), std::max<int64_t>(nnz, 1) * sizeof(double)));
  relabel_fill_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_order, plan->d_position, d_indptr, d_indices, d_data, d_field,
                                              plan->d_indptr, plan->d_indices, plan->d_data, plan->d_field);
  ASP_LAUNCH_CHECK();
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  guard.p = nullptr;
  *out = plan;
  return ASP_OK;
}

int asp_sa_plan_info(asp_sa_plan const *plan, uint64_t *n_padded, uint32_t *num_classes, uint64_t *nnz) {
  ASP_REQUIRE(plan != nullptr, "plan is NULL");
  if (n_padded) *n_padded = plan->n_padded;
  if (num_classes) *num_classes = plan->num_classes;
  if (nnz) *nnz = plan->nnz;
  return ASP_OK;
}

int asp_sa_plan_export(asp_sa_plan const *plan, int32_t *h_order, int64_t *h_class_ptr, int64_t *h_indptr,
                       int32_t *h_indices, double *h_data, double *h_field) {
  ASP_REQUIRE(plan != nullptr, "plan is NULL");
  const uint64_t np = plan->n_padded;
  if (h_order) ASP_CUDA_CHECK(cudaMemcpy(h_order, plan->d_order, np * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (h_class_ptr) std::copy(plan->class_ptr.begin(), plan->class_ptr.end(), h_class_ptr);
  if (h_indptr) ASP_CUDA_CHECK(cudaMemcpy(h_indptr, plan->d_indptr, (np + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (h_indices && plan->nnz) ASP_CUDA_CHECK(cudaMemcpy(h_indices, plan->d_indices, plan->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (h_data && plan->nnz) ASP_CUDA_CHECK(cudaMemcpy(h_data, plan->d_data, plan->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  if (h_field) ASP_CUDA_CHECK(cudaMemcpy(h_field, plan->d_field, np * sizeof(double), cudaMemcpyDeviceToHost));
  return ASP_OK;
}

int asp_sa_anneal(asp_sa_plan *plan, uint32_t num_replicas, uint32_t replica_offset, uint32_t num_sweeps, double const *h_betas,
                  uint64_t seed, uint64_t const *d_x0, double energy_scale, uint64_t *d_best_bits, double *d_best_energy,
                  void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_REQUIRE(plan != nullptr, "plan is NULL");
  ASP_REQUIRE(num_replicas > 0, "need at least one replica");
  ASP_REQUIRE(num_sweeps == 0 || h_betas != nullptr, "h_betas is NULL");
  ASP_REQUIRE(d_best_bits != nullptr, "d_best_bits is NULL");
  ASP_REQUIRE(energy_scale > 0.0, "energy_scale must be positive");
  ASP_REQUIRE(replica_offset % 32 == 0, "replica_offset must be a multiple of 32");
  const uint32_t groups = (num_replicas + 31) / 32;
  const uint64_t np = plan->n_padded;

  int device = 0, sms = 0, per_sm = 0, coop = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&device));
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  ASP_REQUIRE(coop != 0, "device does not support cooperative launches");
  ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sa_sweep_kernel, kSaThreads, 0));
  ASP_REQUIRE(per_sm >= 1, "sweep kernel does not fit on an SM");
  const uint32_t max_ctas = static_cast<uint32_t>(sms * per_sm);

  SaArgs a{};
  a.n_padded = np;
  a.indptr = plan->d_indptr;
  a.indices = plan->d_indices;
  a.data = plan->d_data;
  a.field = plan->d_field;
  a.class_ptr = plan->d_class_ptr;
  a.num_classes = plan->num_classes;
  a.groups = groups;
  if (groups >= max_ctas) {
    a.team_size = 1;
    a.num_teams = max_ctas;
  } else {
    a.team_size = max_ctas / groups;
    a.num_teams = groups;
  }
  // a team larger than the work of the biggest class only adds barrier latency
  {
    int64_t biggest = 4;
    for (uint32_t c = 0; c < plan->num_classes; ++c) biggest = std::max(biggest, plan->class_ptr[c + 1] - plan->class_ptr[c]);
    const uint32_t useful = static_cast<uint32_t>((biggest / 4 + kSaWarps - 1) / kSaWarps);
    a.team_size = std::max(1u, std::min(a.team_size, useful));
  }
  a.num_sweeps = num_sweeps;
  a.replica_offset = replica_offset;
  a.seed = seed;
  a.escale = energy_scale;

  double *d_betas = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_betas), std::max<size_t>(num_sweeps, 1) * sizeof(double), s));
  if (num_sweeps) ASP_CUDA_CHECK(cudaMemcpyAsync(d_betas, h_betas, num_sweeps * sizeof(double), cudaMemcpyHostToDevice, s));
  a.betas = d_betas;
  const size_t word_bytes = static_cast<size_t>(groups) * np * sizeof(uint32_t);
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.words), word_bytes, s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.best_words), word_bytes, s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.rel), groups * 32 * sizeof(long long), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.best_rel), groups * 32 * sizeof(long long), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.barriers), a.num_teams * sizeof(unsigned long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(a.rel, 0, groups * 32 * sizeof(long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(a.best_rel, 0, groups * 32 * sizeof(long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(a.barriers, 0, a.num_teams * sizeof(unsigned long long), s));

  {
    const uint64_t warps = (np / 4) * groups;
    const uint64_t threads = warps * 32;
    sa_init_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(np, groups, replica_offset, seed, plan->d_order, d_x0, a.words);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(a.best_words, a.words, word_bytes, cudaMemcpyDeviceToDevice, s));
  }
  if (num_sweeps > 0) {
    void *params[] = {&a};
    const unsigned grid = a.num_teams * a.team_size;
    ASP_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(sa_sweep_kernel), dim3(grid), dim3(kSaThreads), params, 0, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  {
    const uint64_t total = ((plan->n + 63) / 64) * num_replicas;
    sa_unpermute_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(plan->n, np, num_replicas, plan->d_position, a.best_words, d_best_bits);
    ASP_LAUNCH_CHECK();
  }
  int rc = ASP_OK;
  if (d_best_energy)
    rc = asp_energy(plan->n, plan->d_indptr0, plan->d_indices0, plan->d_data0, plan->d_field0, num_replicas, d_best_bits, d_best_energy, s);
  ASP_CUDA_CHECK(cudaFreeAsync(d_betas, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.words, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.best_words, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.rel, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.best_rel, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.barriers, s));
  return rc;
}

}  // extern "C"
