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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "plan.cuh"

namespace cg = cooperative_groups;

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

// ---- positions: stable sort of the spins by colour, on the device ------------------------
// One stable split per bit of the colour (flags -> exclusive scan -> scatter), lowest bit first; then the spins of a
// colour are a run of the permutation in ascending index order, exactly the oracle's counting sort.
__global__ void __launch_bounds__(256) colour_max_kernel(uint32_t n, const int32_t *__restrict__ colour, unsigned int *__restrict__ out) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  unsigned int c = i < n ? static_cast<unsigned int>(colour[i]) : 0u;
  c = __reduce_max_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) atomicMax(out, c);
}

__global__ void __launch_bounds__(256) split_flag_kernel(uint32_t n, const int32_t *__restrict__ perm, const int32_t *__restrict__ colour, int bit,
                                                         int64_t *__restrict__ flag) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= n) return;
  const int32_t i = perm ? perm[k] : static_cast<int32_t>(k);
  flag[k] = ((colour[i] >> bit) & 1) ? 0 : 1;
}

__global__ void __launch_bounds__(256) split_scatter_kernel(uint32_t n, const int32_t *__restrict__ perm, const int32_t *__restrict__ colour, int bit,
                                                            const int64_t *__restrict__ zeros_before, int32_t *__restrict__ perm_out) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= n) return;
  const int32_t i = perm ? perm[k] : static_cast<int32_t>(k);
  const int64_t z = zeros_before[k], total_zeros = zeros_before[n];
  const int64_t dst = ((colour[i] >> bit) & 1) ? total_zeros + (static_cast<int64_t>(k) - z) : z;
  perm_out[dst] = i;
}

// start[c] = first slot of colour c in the sorted permutation (every colour 0..max is used: a spin takes the
// smallest colour its neighbours leave free); start[classes] = n
__global__ void __launch_bounds__(256) class_start_kernel(uint32_t n, uint32_t classes, const int32_t *__restrict__ perm, const int32_t *__restrict__ colour,
                                                          int64_t *__restrict__ start) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= n) return;
  const int32_t c = colour[perm ? perm[k] : static_cast<int32_t>(k)];
  if (k == 0 || colour[perm ? perm[k - 1] : static_cast<int32_t>(k - 1)] != c) start[c] = k;
  if (k == n - 1) start[classes] = n;
}

__global__ void __launch_bounds__(256) place_kernel(uint32_t n, const int32_t *__restrict__ perm, const int32_t *__restrict__ colour,
                                                    const int64_t *__restrict__ start, const int64_t *__restrict__ class_ptr, int32_t *__restrict__ order,
                                                    int32_t *__restrict__ position) {
  const uint32_t k = blockIdx.x * 256u + threadIdx.x;
  if (k >= n) return;
  const int32_t i = perm ? perm[k] : static_cast<int32_t>(k);
  const int32_t c = colour[i];
  const int64_t p = class_ptr[c] + (static_cast<int64_t>(k) - start[c]);
  order[p] = i;
  position[i] = static_cast<int32_t>(p);
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

// the same function with the ten round keys precomputed on the host (kernel arguments: constant-bank operands)
struct PhiloxKeys {
  uint32_t x[10], y[10];
};
__device__ __forceinline__ uint4 philox4x32_10_scheduled(uint4 c, const PhiloxKeys &k) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x[round], lo1, hi0 ^ c.w ^ k.y[round], lo0);
  }
  return c;
}

// exp(-x), 0 < x < 23, binary32, the operation sequence of oracle/anneal_port.c:exp_neg_f32
// (explicit _rn intrinsics: nothing is contracted or reassociated).
__device__ __forceinline__ float exp_neg_f32(float x) {
  const float t = __fmul_rn(x, 1.44269502f);
  const float r = __fadd_rn(t, 12582912.0f);
  const int k = __float_as_int(r) - 0x4B400000;
  const float g = __fadd_rn(t, -__fadd_rn(r, -12582912.0f));
  const float w = __fmul_rn(g, -0.693147182f);
  float p = 1.0f / 5040.0f;
  p = __fmaf_rn(p, w, 1.0f / 720.0f);
  p = __fmaf_rn(p, w, 1.0f / 120.0f);
  p = __fmaf_rn(p, w, 1.0f / 24.0f);
  p = __fmaf_rn(p, w, 1.0f / 6.0f);
  p = __fmaf_rn(p, w, 0.5f);
  p = __fmaf_rn(p, w, 1.0f);
  p = __fmaf_rn(p, w, 1.0f);
  return __int_as_float(__float_as_int(p) - (k << 23));
}

// Metropolis test for an uphill move, 0 < x = beta dE < 23 (oracle: accept_uphill).
__device__ __forceinline__ bool accept_uphill(double x, uint32_t rnd) {
  const float u = __fmaf_rn(__uint2float_rn(rnd), 2.32830644e-10f, 1.16415322e-10f);
  return u < exp_neg_f32(__double2float_rn(x));
}

// llrint(v) for |v| < 2^51 without a conversion instruction: adding 1.5 * 2^52 leaves the
// integer, rounded to nearest-even, in the low mantissa bits (asp_sa_anneal checks the bound).
__device__ __forceinline__ long long round_to_ll(double v) {
  return __double_as_longlong(__dadd_rn(v, 6755399441055744.0)) - 0x4338000000000000ll;
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
  const int4 *bounds;    // [n_padded / 4] row boundaries of each warp-task, relative to its first entry
  PhiloxKeys round_keys; // key schedule of `seed`
  unsigned int *tickets; // [num_teams][num_classes] chunk tickets, zeroed by the host
  uint32_t stage_slots;  // staged entries per warp: kStageSlots, or kStageSlotsWide when a task of the model spans more than 32 entries
};

constexpr int kSaThreads = 256;
constexpr int kSaWarps = kSaThreads / 32;
#ifndef ASP_SA_CTAS_PER_SM
#define ASP_SA_CTAS_PER_SM 3
#endif
#ifndef ASP_SA_CHUNK
#define ASP_SA_CHUNK 8
#endif
constexpr uint32_t kSaChunk = ASP_SA_CHUNK;  // consecutive tasks per ticket
constexpr int kSaCtasPerSm = ASP_SA_CTAS_PER_SM;  // 24 warps/SM: hides the indptr -> entries -> spin-word load chain

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

// +val when bit `lane` of `word` is set, -val otherwise (sign-bit XOR: exact); up = 31 - lane
__device__ __forceinline__ double signed_by_bit(double val, uint32_t word, uint32_t up) {
  const uint32_t flip = ~(word << up) & 0x80000000u;
  return __hiloint2double(__double2hiint(val) ^ static_cast<int>(flip), __double2loint(val));
}

// A warp-task = 4 consecutive positions x 32 replicas.  The entries of the 4 rows are ONE
// contiguous span of the CSR: lanes load (value, column) coalesced, gather the 32-replica
// spin word of their column, and stage the pairs in shared memory; then every lane (= one
// replica) walks the staged span and sums each row's local field in stored order.
struct TaskRows {
  int32_t b[5];  // row boundaries relative to the first entry of the span
};

// sums[j] = sum over row j of (+/-) val in stored order, for this lane's replica.
// first-chunk (value, word) of the span arrive in (pv, wv); later chunks are loaded here.
__device__ __forceinline__ void accumulate_rows(const TaskRows &rows, int64_t e_begin, double pv, uint32_t wv, const int32_t *__restrict__ indices,
                                                const double *__restrict__ data, const uint32_t *words, StagedEntry *stage, uint32_t lane,
                                                double acc[4]) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
  const int32_t span = rows.b[4];
  for (int32_t cb = 0; cb < span; cb += 32) {
    if (cb != 0) {
      pv = 0.0;
      wv = 0u;
      if (cb + static_cast<int32_t>(lane) < span) {
        const int64_t e = e_begin + cb + lane;
        pv = __ldg(&data[e]);
        wv = __ldcg(&words[__ldg(&indices[e])]);
      }
    }
    __syncwarp();
    StagedEntry se;
    se.val = pv;
    se.word = wv;
    se.pad = 0;
    stage[lane] = se;
    __syncwarp();
    const uint32_t up = 31u - lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int32_t lo = max(rows.b[j], cb) - cb, hi = min(rows.b[j + 1], cb + 32) - cb;
#pragma unroll 1  // rows hold ~4 entries: a tight loop (two entries per trip, the second one predicated) beats an unrolled one
      for (int32_t k = lo; k < hi; k += 2) {
        const StagedEntry x0 = stage[k], x1 = stage[k + 1];  // stage has a 33rd slot
        acc[j] = __dadd_rn(acc[j], signed_by_bit(x0.val, x0.word, up));
        if (k + 1 < hi) acc[j] = __dadd_rn(acc[j], signed_by_bit(x1.val, x1.word, up));
      }
    }
  }
}

// Row boundaries of a warp-task relative to its first entry (b[0] = 0 is implied): one 16-byte word,
// read by all lanes at once, replaces five 64-bit row pointers and their shuffles.
__global__ void __launch_bounds__(256) task_bounds_kernel(uint64_t tasks, const int64_t *__restrict__ indptr, int4 *__restrict__ bounds) {
  const uint64_t q = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (q >= tasks) return;
  const int64_t b0 = indptr[q * 4];
  bounds[q] = make_int4(static_cast<int>(indptr[q * 4 + 1] - b0), static_cast<int>(indptr[q * 4 + 2] - b0), static_cast<int>(indptr[q * 4 + 3] - b0),
                        static_cast<int>(indptr[q * 4 + 4] - b0));
}

// how many tasks span more than one staged chunk (32 entries)
__global__ void __launch_bounds__(256) long_tasks_kernel(uint64_t tasks, const int4 *__restrict__ bounds, unsigned long long *__restrict__ out) {
  const uint64_t q = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  const unsigned int n = __popc(__ballot_sync(0xffffffffu, q < tasks && bounds[q].w > 32));
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, static_cast<unsigned long long>(n));
}

__global__ void __launch_bounds__(256) any_field_kernel(uint64_t n_padded, const double *__restrict__ field, uint32_t *__restrict__ out) {
  const uint64_t p = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (p < n_padded && field[p] != 0.0) *out = 1u;
}

// sum of one staged row [lo, hi) for this lane's replica, stored order, two entries per trip (one 16-byte shared
// load per entry -- x, y: value; z: spin word of the column); pointers instead of indices: 12 instructions per trip
__device__ __forceinline__ double staged_entry(const uint4 e, uint32_t up) {
  return __hiloint2double(static_cast<int>(e.y ^ (~(e.z << up) & 0x80000000u)), static_cast<int>(e.x));
}

__device__ __forceinline__ double staged_row_sum(const StagedEntry *stage, int32_t lo, int32_t hi, uint32_t up, double acc = 0.0) {
  const uint4 *p = reinterpret_cast<const uint4 *>(stage) + lo, *const last = reinterpret_cast<const uint4 *>(stage) + (hi - 1);
#pragma unroll 1
  for (; p <= last; p += 2) {
    const uint4 x0 = p[0], x1 = p[1];  // the stage has a 33rd slot
    acc = __dadd_rn(acc, staged_entry(x0, up));
    if (p < last) acc = __dadd_rn(acc, staged_entry(x1, up));
  }
  return acc;
}

// Rows of 9..40 couplings (SK-type models, small full-basis models): the whole span of a task fits the wide stage.  The
// first 32 entries came with the pipeline; the other chunks are loaded here, ALL loads before the first use (values and
// columns, then the gathered spin words), so the task pays two memory round trips, not two per chunk; then the four rows
// are summed straight from the stage.  Not inlined: the register allocation of the short-row path stays as it is.
struct Sum4 {
  double v[4];
};
__device__ __noinline__ Sum4 wide_task_sums(const double *__restrict__ data, const int32_t *__restrict__ indices, const uint32_t *words, StagedEntry *stage,
                                            int64_t e_begin, int4 rows, double first_pv, uint32_t first_wv, uint32_t lane) {
  const uint32_t up = 31u - lane;
  double pvx[4];
  int32_t pix[4];
  uint32_t wvx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int32_t mine = 32 + 32 * c + static_cast<int32_t>(lane);
    pvx[c] = 0.0;
    pix[c] = -1;
    if (mine < rows.w) {
      pvx[c] = __ldg(&data[e_begin + mine]);
      pix[c] = __ldg(&indices[e_begin + mine]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) wvx[c] = pix[c] >= 0 ? __ldcg(&words[pix[c]]) : 0u;
  __syncwarp();
  StagedEntry se;
  se.val = first_pv;
  se.word = first_wv;
  se.pad = 0;
  stage[lane] = se;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    se.val = pvx[c];
    se.word = wvx[c];
    stage[32 + 32 * c + lane] = se;
  }
  __syncwarp();
  Sum4 out;
  out.v[0] = staged_row_sum(stage, 0, rows.x, up);
  out.v[1] = staged_row_sum(stage, rows.x, rows.y, up);
  out.v[2] = staged_row_sum(stage, rows.y, rows.z, up);
  out.v[3] = staged_row_sum(stage, rows.z, rows.w, up);
  return out;
}

// One class phase with ONE POSITION per warp-task -- for classes so small that the team has a warp for (nearly) every
// position: four times the warps of the 4-position tasks work at once and a task is a quarter as long, which is what
// counts when a phase is a single task deep (small models, few replicas: the time between two barriers is the latency
// of one task).  All the loads of a row are issued before the first use.  Same arithmetic per position, the same variate
// (component p % 4 of the Philox block of task p / 4), so the chain is unchanged.  Rows longer than the stage holds go
// through it piece by piece.  Not inlined: the register allocation of the main task loop stays as it is.
__device__ __noinline__ long long fine_class_phase(const int64_t *__restrict__ indptr, const int4 *__restrict__ bounds, const double *__restrict__ data,
                                                   const int32_t *__restrict__ indices, const double *__restrict__ field, uint32_t *words,
                                                   StagedEntry *stage, uint32_t stage_slots, uint32_t p_begin, uint32_t p_end, uint32_t my_warp,
                                                   uint32_t team_warps, uint32_t lane, uint32_t t, uint32_t stream_id, double beta, double escale,
                                                   uint2 key) {
  const uint32_t up = 31u - lane, lane_bit = 1u << lane;
  const int32_t room = static_cast<int32_t>(stage_slots) - 2;  // entries the stage holds: 32 or 160
  long long rel_delta = 0;
  for (uint32_t p = p_begin + my_warp; p < p_end; p += team_warps) {
    const uint32_t q = p >> 2, j = p & 3u;
    const int64_t e_task = __ldg(&indptr[static_cast<uint64_t>(q) * 4]);
    const int4 bd = __ldg(&bounds[q]);
    const uint32_t cur_w = __ldcg(&words[p]);
    const double f2 = field ? __dmul_rn(2.0, __ldg(&field[p])) : 0.0;
    const int32_t lo = j == 0 ? 0 : (j == 1 ? bd.x : (j == 2 ? bd.y : bd.z)), hi = j == 0 ? bd.x : (j == 1 ? bd.y : (j == 2 ? bd.z : bd.w));
    const int64_t e_row = e_task + lo;
    const int32_t len = hi - lo;
    double acc = 0.0;
    for (int32_t base = 0; base < len || base == 0; base += room) {
      const int32_t piece = min(len - base, room);
      double pvx[5];
      int32_t pix[5];
      uint32_t wvx[5];
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int32_t mine = 32 * c + static_cast<int32_t>(lane);
        pvx[c] = 0.0;
        pix[c] = -1;
        if (mine < piece) {
          pvx[c] = __ldg(&data[e_row + base + mine]);
          pix[c] = __ldg(&indices[e_row + base + mine]);
        }
      }
#pragma unroll
      for (int c = 0; c < 5; ++c) wvx[c] = pix[c] >= 0 ? __ldcg(&words[pix[c]]) : 0u;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (32 * c < room) {  // (uniform) the narrow stage holds one chunk
          StagedEntry se;
          se.val = pvx[c];
          se.word = wvx[c];
          se.pad = 0;
          stage[32 * c + lane] = se;
        }
      }
      __syncwarp();
      acc = staged_row_sum(stage, 0, piece, up, acc);
      if (len == 0) break;
    }
    const double gsum = __fma_rn(4.0, acc, f2);
    const uint32_t neg = (cur_w << up) & 0x80000000u;  // spin up: dE = -gsum
    const double dE = __hiloint2double(__double2hiint(gsum) ^ static_cast<int>(neg), __double2loint(gsum));
    const double x = __dmul_rn(beta, dE);
    const uint32_t uphill = __ballot_sync(0xffffffffu, dE > 0.0 && x < kRejectAbove);
    uint32_t accepted = __ballot_sync(0xffffffffu, !(dE > 0.0));
    if (uphill) {
      const uint4 rnd = philox4x32_10(make_uint4(q, t, stream_id, 0u), key);
      const uint32_t r = j == 0 ? rnd.x : (j == 1 ? rnd.y : (j == 2 ? rnd.z : rnd.w));
      accepted |= __ballot_sync(0xffffffffu, accept_uphill(x, r)) & uphill;
    }
    if (accepted & lane_bit) rel_delta += round_to_ll(__dmul_rn(dE, escale));
    if (lane == 0) __stcg(&words[p], cur_w ^ accepted);
  }
  return rel_delta;
}

constexpr int kStageSlots = 34;        // 32 staged entries + two: the row loop reads two entries per trip
constexpr int kStageSlotsWide = 162;   // 32 + 4 x 32 + two: models with a task span above 32 entries (SaArgs::stage_slots)

// kField = false: every field is zero, nothing of it is read.  kFine = true: a model so small that the team has a warp for
// (nearly) every position of its biggest class -- every class phase is fine_class_phase (one position per task); the host
// picks the instantiation, so the 4-position task loop below is compiled without the call.
template <bool kField, bool kFine>
__global__ void __launch_bounds__(kSaThreads, kSaCtasPerSm) sa_sweep_kernel(const SaArgs a) {
  extern __shared__ __align__(16) unsigned char sa_smem[];  // [kSaWarps][a.stage_slots] staged entries
  const uint32_t lane = threadIdx.x & 31;
  // broadcast from lane 0: tells the compiler that the warp index -- and with it every task loop below -- is warp-uniform
  const uint32_t warp_in_cta = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t team = blockIdx.x / a.team_size;
  if (team >= a.num_teams) return;
  const uint32_t member = blockIdx.x % a.team_size;
  const uint32_t team_warps = a.team_size * kSaWarps;
  const uint32_t my_warp = member * kSaWarps + warp_in_cta;
  const uint32_t up = 31u - lane, lane_bit = 1u << lane;
  TeamBarrier bar{a.barriers + team, 0ull, a.team_size};
  StagedEntry *stage = reinterpret_cast<StagedEntry *>(sa_smem) + static_cast<size_t>(warp_in_cta) * a.stage_slots;

  for (uint32_t g = team; g < a.groups; g += a.num_teams) {
    uint32_t *words = a.words + static_cast<uint64_t>(g) * a.n_padded;
    asm volatile("" : "+l"(words));  // opaque: the group's base pointer is not re-derived for every load of the task loop
    uint32_t *best = a.best_words + static_cast<uint64_t>(g) * a.n_padded;
    const uint32_t replica = g * 32 + lane;                  // local slot
    const uint32_t stream_id = a.replica_offset + replica;   // global replica: RNG stream
    long long best_rel = 0;  // every warp of the team tracks the same value
    for (uint32_t t = 0; t < a.num_sweeps; ++t) {
      const double beta = a.betas[t];
      long long rel_delta = 0;
      for (uint32_t c = 0; c < a.num_classes; ++c) {
        // positions < 2^31, so tasks (and tasks + 3 strides) fit 32 bits
        const uint32_t q_begin = static_cast<uint32_t>(a.class_ptr[c] >> 2), q_end = static_cast<uint32_t>(a.class_ptr[c + 1] >> 2);
        // Big classes hand their tasks out in chunks of kSaChunk consecutive tasks: the first chunk of a warp is
        // its own index, later ones come from the team's ticket counter of the class (reset after the class
        // barrier), so no warp waits at the barrier for a slower one with a fixed share.  Which warp runs a task
        // does not matter: the tasks of a class are independent and the variates are keyed by (task, sweep,
        // replica).  The ticket for the NEXT chunk is drawn when a chunk begins (lane 0; read by shuffle when
        // needed).  Small classes (fewer than four chunks per warp) are dealt out task by task with a fixed
        // stride instead -- every warp gets work, and no atomic sits on the path between two barriers.
        if constexpr (kFine) {
          rel_delta += fine_class_phase(a.indptr, a.bounds, a.data, a.indices, kField ? a.field : nullptr, words, stage, a.stage_slots, q_begin * 4u,
                                        q_end * 4u, my_warp, team_warps, lane, t, stream_id, beta, a.escale,
                                        make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)));
          bar.sync();
          continue;
        }
        unsigned int *const ticket_counter = a.tickets + static_cast<uint64_t>(team) * a.num_classes + c;
        const bool dealt = (q_end - q_begin) < 4u * kSaChunk * team_warps;
        const uint32_t chunk = dealt ? 1u : kSaChunk;
        uint32_t it_q = 0xFFFFFFFFu, it_end = 0xFFFFFFFFu, pending = my_warp;
        bool it_done = true;
        {
          const uint64_t start = static_cast<uint64_t>(q_begin) + static_cast<uint64_t>(my_warp) * chunk;
          if (start < q_end) {
            it_done = false;
            it_q = static_cast<uint32_t>(start);
            it_end = static_cast<uint32_t>(min(start + chunk, static_cast<uint64_t>(q_end)));
            if (!dealt && lane == 0) pending = atomicAdd(ticket_counter, 1u);
          }
        }
        auto next_task = [&]() -> uint32_t {  // warp-uniform; 0xFFFFFFFF when the class is used up
          if (it_q == it_end) {
            if (it_done) return 0xFFFFFFFFu;
            const uint32_t ticket = __shfl_sync(0xffffffffu, pending, 0);
            const uint64_t start = static_cast<uint64_t>(q_begin) + (static_cast<uint64_t>(ticket) + team_warps) * chunk;
            if (start >= q_end) {
              it_done = true;
              it_q = it_end = 0xFFFFFFFFu;
              return 0xFFFFFFFFu;
            }
            it_q = static_cast<uint32_t>(start);
            it_end = static_cast<uint32_t>(min(start + chunk, static_cast<uint64_t>(q_end)));
            if (dealt)
              pending += team_warps;
            else if (lane == 0)
              pending = atomicAdd(ticket_counter, 1u);
          }
          return it_q++;
        };
        // software pipeline over this warp's tasks of the class, three stages deep: task words (first
        // entry, row boundaries) three tasks ahead, first-chunk (value, column) two tasks ahead, the
        // gathered spin words one task ahead.  The CSR is read-only, and the neighbours of a class
        // lie in OTHER classes, whose words do not change before the next barrier -- so all of it is
        // safe to prefetch once the class has begun.  Every lane reads the same task word: one
        // transaction, no shuffles.
        //   stage A: (ebA, spanA) of task qA;  stage B: (pvB, piB) of qB;  stage C: (pvC, wvC) of q.
        // Little state rides along (a column of -1 marks a lane past the span; the full row boundaries are read
        // again one task ahead, when the sector is in L2), so rotating the pipeline is a handful of moves.
        uint32_t q = next_task(), qB = next_task(), qA = next_task();
        int64_t ebA = 0;
        int32_t spanA = 0;
        double pvB = 0.0, pvC = 0.0;
        int32_t piB = -1;
        uint32_t wvC = 0u;
        int4 rowsC = make_int4(0, 0, 0, 0);
        if (q != 0xFFFFFFFFu) {
          const int64_t eb = __ldg(&a.indptr[static_cast<uint64_t>(q) * 4]);
          rowsC = __ldg(&a.bounds[q]);
          if (static_cast<int32_t>(lane) < rowsC.w) {
            pvC = __ldg(&a.data[eb + lane]);
            wvC = __ldcg(&words[__ldg(&a.indices[eb + lane])]);
          }
        }
        if (qB != 0xFFFFFFFFu) {
          const int64_t eb = __ldg(&a.indptr[static_cast<uint64_t>(qB) * 4]);
          if (static_cast<int32_t>(lane) < __ldg(&a.bounds[qB].w)) {
            pvB = __ldg(&a.data[eb + lane]);
            piB = __ldg(&a.indices[eb + lane]);
          }
        }
        if (qA != 0xFFFFFFFFu) {
          ebA = __ldg(&a.indptr[static_cast<uint64_t>(qA) * 4]);
          spanA = __ldg(&a.bounds[qA].w);
        }
        while (q != 0xFFFFFFFFu) {
          const uint32_t p0 = q * 4, q_now = q;
          const int4 rows = rowsC;
          if (qB != 0xFFFFFFFFu) rowsC = __ldg(&a.bounds[qB]);
          const uint32_t wv = wvC;
          const double cur_pv = pvC;
          const uint4 cur = __ldcg(reinterpret_cast<const uint4 *>(words + p0));
          double field2 = 0.0;  // 2 h of position p0 + lane
          if (kField && lane < 4) field2 = __dmul_rn(2.0, __ldg(&a.field[p0 + lane]));
          // advance the pipeline
          wvC = 0u;
          if (piB >= 0) wvC = __ldcg(&words[piB]);
          pvC = pvB;
          pvB = 0.0;
          piB = -1;
          if (static_cast<int32_t>(lane) < spanA) {
            pvB = __ldg(&a.data[ebA + lane]);
            piB = __ldg(&a.indices[ebA + lane]);
          }
          q = qB;
          qB = qA;
          qA = next_task();
          spanA = 0;
          if (qA != 0xFFFFFFFFu) {
            ebA = __ldg(&a.indptr[static_cast<uint64_t>(qA) * 4]);
            spanA = __ldg(&a.bounds[qA].w);
          }

          double acc[4];
          if (rows.w <= 32) {  // the usual case: the four rows are one staged chunk
            __syncwarp();
            StagedEntry se;
            se.val = cur_pv;
            se.word = wv;
            se.pad = 0;
            stage[lane] = se;
            __syncwarp();
            acc[0] = staged_row_sum(stage, 0, rows.x, up);
            acc[1] = staged_row_sum(stage, rows.x, rows.y, up);
            acc[2] = staged_row_sum(stage, rows.y, rows.z, up);
            acc[3] = staged_row_sum(stage, rows.z, rows.w, up);
          } else if (rows.w <= 160 && a.stage_slots >= static_cast<uint32_t>(kStageSlotsWide)) {
            const Sum4 sums = wide_task_sums(a.data, a.indices, words, stage, __ldg(&a.indptr[static_cast<uint64_t>(q_now) * 4]), rows, cur_pv, wv, lane);
            acc[0] = sums.v[0];
            acc[1] = sums.v[1];
            acc[2] = sums.v[2];
            acc[3] = sums.v[3];
          } else {
            // longer rows still (dense models): the span goes through the stage in chunks of 32 entries; the
            // (value, column, spin word) of the NEXT chunk are on their way while this one is summed, and a row that
            // crosses a chunk boundary carries its partial sum on (same order of additions)
            const int64_t e_begin = __ldg(&a.indptr[static_cast<uint64_t>(q_now) * 4]);
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
            double pv_c = cur_pv;
            uint32_t wv_c = wv;
            for (int32_t cb = 0;; cb += 32) {
              double pv_n = 0.0;
              uint32_t wv_n = 0u;
              const int32_t mine = cb + 32 + static_cast<int32_t>(lane);
              if (mine < rows.w) {
                pv_n = __ldg(&a.data[e_begin + mine]);
                wv_n = __ldcg(&words[__ldg(&a.indices[e_begin + mine])]);
              }
              __syncwarp();
              StagedEntry se;
              se.val = pv_c;
              se.word = wv_c;
              se.pad = 0;
              stage[lane] = se;
              __syncwarp();
              const int32_t b[5] = {0, rows.x, rows.y, rows.z, rows.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int32_t lo = max(b[j], cb) - cb, hi = min(b[j + 1], cb + 32) - cb;
                if (lo < hi) acc[j] = staged_row_sum(stage, lo, hi, up, acc[j]);
              }
              if (cb + 32 >= rows.w) break;
              pv_c = pv_n;
              wv_c = wv_n;
            }
          }
          const uint32_t cur_w[4] = {cur.x, cur.y, cur.z, cur.w};
          // dE = -s (4 sum + 2 h): 4 acc is exact, so one fma rounds like the oracle's mul, mul, add;
          // uphill lanes (dE > 0, beta dE < 23) need a variate
          double dE[4], x[4];
          uint32_t uphill[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double f2 = kField ? __shfl_sync(0xffffffffu, field2, j) : 0.0;
            const double gsum = __fma_rn(4.0, acc[j], f2);
            const uint32_t neg = (cur_w[j] << up) & 0x80000000u;  // spin up: dE = -gsum
            dE[j] = __hiloint2double(__double2hiint(gsum) ^ static_cast<int>(neg), __double2loint(gsum));
            x[j] = __dmul_rn(beta, dE[j]);
            uphill[j] = __ballot_sync(0xffffffffu, dE[j] > 0.0 && x[j] < kRejectAbove);
          }
          // the variates and the four Metropolis tests in ONE warp-uniform block, so that the four
          // exponentials interleave; the test is garbage but harmless on the lanes that are not uphill
          uint32_t accepted[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) accepted[j] = __ballot_sync(0xffffffffu, !(dE[j] > 0.0));
          if (uphill[0] | uphill[1] | uphill[2] | uphill[3]) {
            const uint4 rnd = philox4x32_10_scheduled(make_uint4(q_now, t, stream_id, 0u), a.round_keys);
            const uint32_t rnd_w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) accepted[j] |= __ballot_sync(0xffffffffu, accept_uphill(x[j], rnd_w[j])) & uphill[j];
          }
          uint32_t new_w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long inc = round_to_ll(__dmul_rn(dE[j], a.escale));
            if (accepted[j] & lane_bit) rel_delta += inc;
            new_w[j] = cur_w[j] ^ accepted[j];
          }
          if (lane == 0) __stcg(reinterpret_cast<uint4 *>(words + p0), make_uint4(new_w[0], new_w[1], new_w[2], new_w[3]));
        }
        bar.sync();
        // every warp of the team has drawn its last ticket of this class; the counter is next used one sweep (>= 1 barrier) later
        if (!dealt && member == 0 && threadIdx.x == 0) *ticket_counter = 0u;
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

// ---- exact energies of the relabelled, replica-sliced configurations --------------------
// E_r = sum_p s_p (sum_{j != p} J_pj s_j + h_p) + sum_i J_ii: same task shape as the sweep
// (a CSR row is read once per 32 replicas).  Fixed grid + fixed summation order: deterministic.
constexpr int kEnWarps = 8;

__global__ void __launch_bounds__(kEnWarps * 32) energy_sliced_kernel(uint64_t n_padded, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                                      const double *__restrict__ data, const double *__restrict__ field,
                                                                      const uint32_t *__restrict__ all_words, double *__restrict__ partial) {
  __shared__ StagedEntry s_stage[kEnWarps][33];
  const uint32_t lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
  const uint32_t g = blockIdx.y;
  const uint32_t *words = all_words + static_cast<uint64_t>(g) * n_padded;
  const uint64_t total_warps = static_cast<uint64_t>(gridDim.x) * kEnWarps;
  const uint64_t warp_global = static_cast<uint64_t>(blockIdx.x) * kEnWarps + warp_in_cta;
  StagedEntry *stage = s_stage[warp_in_cta];
  double e = 0.0;
  for (uint64_t q = warp_global; q < n_padded / 4; q += total_warps) {
    const uint64_t p0 = q * 4;
    const int64_t ip = lane < 5 ? __ldg(&indptr[p0 + lane]) : 0;
    const int64_t e_begin = __shfl_sync(0xffffffffu, ip, 0);
    const int32_t rel_ip = static_cast<int32_t>(ip - e_begin);
    TaskRows rows;
    rows.b[0] = 0;
#pragma unroll
    for (int j = 1; j < 5; ++j) rows.b[j] = __shfl_sync(0xffffffffu, rel_ip, j);
    double pv = 0.0;
    uint32_t wv = 0u;
    if (static_cast<int32_t>(lane) < rows.b[4]) {
      pv = __ldg(&data[e_begin + lane]);
      wv = __ldg(&words[__ldg(&indices[e_begin + lane])]);
    }
    const uint4 cur = __ldg(reinterpret_cast<const uint4 *>(words + p0));
    const double my_field = lane < 4 ? __ldg(&field[p0 + lane]) : 0.0;
    double acc[4];
    accumulate_rows(rows, e_begin, pv, wv, indices, data, words, stage, lane, acc);
    const uint32_t cur_w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double local = __dadd_rn(acc[j], __shfl_sync(0xffffffffu, my_field, j));
      e = __dadd_rn(e, ((cur_w[j] >> lane) & 1) ? local : -local);
    }
  }
  partial[(static_cast<uint64_t>(g) * total_warps + warp_global) * 32 + lane] = e;
}

// one CTA per group: thread (w, lane) sums the partials of warps w, w+8, ...; the 8 strided
// sums are combined in a fixed order, then the diagonal constant is added.
__global__ void __launch_bounds__(256) energy_sliced_final_kernel(const double *__restrict__ partial, uint64_t total_warps, double diag_sum,
                                                                  uint32_t num_replicas, double *__restrict__ out) {
  __shared__ double s_sum[8][32];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = blockIdx.x;
  const double *p = partial + static_cast<uint64_t>(g) * total_warps * 32;
  double acc = 0.0;
  for (uint64_t k = w; k < total_warps; k += 8) acc = __dadd_rn(acc, p[k * 32 + lane]);
  s_sum[w][lane] = acc;
  __syncthreads();
  if (w == 0) {
    double total = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) total = __dadd_rn(total, s_sum[k][lane]);
    const uint32_t r = g * 32 + lane;
    if (r < num_replicas) out[r] = __dadd_rn(total, diag_sum);
  }
}

// largest possible single-flip energy change: max over rows of 4 sum |J| + 2 |h| (non-negative
// doubles order like their bit patterns, so an integer atomicMax does the reduction)
__global__ void __launch_bounds__(256) max_de_kernel(uint64_t n_padded, const int64_t *__restrict__ indptr, const double *__restrict__ data,
                                                     const double *__restrict__ field, unsigned long long *__restrict__ out) {
  const uint64_t p = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  double bound = 0.0;
  if (p < n_padded) {
    double sum = 0.0;
    for (int64_t k = indptr[p]; k < indptr[p + 1]; ++k) sum += fabs(data[k]);
    bound = 4.0 * sum + 2.0 * fabs(field[p]);
  }
  unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(bound));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));  // one atomic per warp
  if ((threadIdx.x & 31) == 0) atomicMax(out, bits);
}

// sum of the diagonal of the ORIGINAL model (dropped from the relabelled CSR): block partials
__global__ void __launch_bounds__(256) diag_partial_kernel(uint64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                           const double *__restrict__ data, double *__restrict__ partial) {
  __shared__ double smem[8];
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  double d = 0.0;
  if (i < n)
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
      if (static_cast<uint64_t>(indices[k]) == i) d = __dadd_rn(d, data[k]);
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t = __dadd_rn(t, smem[k]);
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) diag_final_kernel(const double *__restrict__ partial, uint64_t count, double *__restrict__ out) {
  __shared__ double smem[256];
  double acc = 0.0;
  for (uint64_t k = threadIdx.x; k < count; k += 256) acc = __dadd_rn(acc, partial[k]);
  smem[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 256; ++k) t = __dadd_rn(t, smem[k]);
    *out = t;
  }
}

// ---- outputs: relabelled [groups][n_padded] words -> original-order packed [R][words64] --
// A warp transposes 256 consecutive original spins of one replica group: lane = spin while
// gathering (position[] coalesced), lane = replica after 32 ballots, so every replica writes
// 32 contiguous bytes of its packed vector.
__global__ void __launch_bounds__(256) sa_unpermute_kernel(uint64_t n, uint64_t n_padded, uint32_t num_replicas, const int32_t *__restrict__ position,
                                                           const uint32_t *__restrict__ best_words, uint64_t *__restrict__ out) {
  const uint64_t warp_global = (static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31, g = blockIdx.y;
  const uint64_t i0 = warp_global * 256;
  if (i0 >= n) return;
  const uint32_t *src = best_words + static_cast<uint64_t>(g) * n_padded;
  uint32_t mine[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const uint64_t i = i0 + 32 * s + lane;
    const uint32_t w = i < n ? __ldg(&src[__ldg(&position[i])]) : 0u;
    uint32_t m = 0;
#pragma unroll
    for (uint32_t r = 0; r < 32; ++r) {
      const uint32_t b = __ballot_sync(0xffffffffu, (w >> r) & 1u);
      if (lane == r) m = b;
    }
    mine[s] = m;
  }
  const uint32_t r = g * 32 + lane;
  if (r >= num_replicas) return;
  const uint64_t words64 = (n + 63) / 64;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t wi = i0 / 64 + k;
    if (wi < words64) out[static_cast<uint64_t>(r) * words64 + wi] = static_cast<uint64_t>(mine[2 * k]) | (static_cast<uint64_t>(mine[2 * k + 1]) << 32);
  }
}

}  // namespace asp

using namespace asp;

static int g_sa_team_cap = 0;  // asp_debug_set_sa_team_ctas

extern "C" {

void asp_debug_set_sa_team_ctas(int max_ctas_per_team) { g_sa_team_cap = max_ctas_per_team > 0 ? max_ctas_per_team : 0; }

void asp_sa_plan_destroy(asp_sa_plan *plan) {
  if (!plan) return;
  // the buffers come from the stream-ordered pool; cudaFree waits for the device and hands them back to it
  for (void *p : {static_cast<void *>(plan->d_order), static_cast<void *>(plan->d_position), static_cast<void *>(plan->d_indptr),
                  static_cast<void *>(plan->d_indices), static_cast<void *>(plan->d_data), static_cast<void *>(plan->d_field),
                  static_cast<void *>(plan->d_class_ptr), static_cast<void *>(plan->d_bounds)})
    if (p) cudaFree(p);
  delete plan;
}

int asp_sa_plan_create(asp_sa_plan **out, uint64_t n, int64_t const *d_indptr, int32_t const *d_indices,
                       double const *d_data, double const *d_field, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  ASP_CUDA_CHECK(asp::keep_pool_memory());
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

  const bool timing = getenv("ASP_PLAN_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "plan: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  // Temporaries come from the stream-ordered pool (asp::keep_pool_memory keeps it warm): cudaMalloc / cudaFree of
  // buffers this size cost tens to hundreds of milliseconds each.
  // 1. colour the coupling graph on the device; the count of uncoloured spins is read back every fourth round
  //    (a round after the last one changes nothing)
  int32_t *d_colour[2] = {nullptr, nullptr};
  unsigned long long *d_remaining = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_colour[0]), n * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_colour[1]), n * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_remaining), sizeof(unsigned long long), s));
  ASP_CUDA_CHECK(cudaMemsetAsync(d_colour[0], 0xFF, n * sizeof(int32_t), s));
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  int cur = 0;
  for (int round = 0;; ++round) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_remaining, 0, sizeof(unsigned long long), s));
    colour_round_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), d_indptr, d_indices, d_colour[cur], d_colour[cur ^ 1], d_remaining);
    ASP_LAUNCH_CHECK();
    cur ^= 1;
    if (round % 4 != 3) continue;
    unsigned long long remaining = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&remaining, d_remaining, sizeof(remaining), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (remaining == 0) break;
    if (round > 100000) {
      set_error("colouring did not converge");
      return ASP_ERR_CUDA;
    }
  }
  const int32_t *d_col = d_colour[cur];
  lap("colouring");
  // 2. stable sort of the spins by colour on the device -> positions (classes padded to multiples of 4)
  unsigned int max_colour = 0;
  {
    unsigned int *d_max = reinterpret_cast<unsigned int *>(d_remaining);
    ASP_CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(unsigned int), s));
    colour_max_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), d_col, d_max);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(&max_colour, d_max, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  const uint32_t classes = max_colour + 1;
  int32_t *d_perm[2] = {nullptr, nullptr};
  int64_t *d_flag = nullptr, *d_zeros = nullptr, *d_start = nullptr;
  void *d_scan = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_perm[0]), n * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_perm[1]), n * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_flag), n * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_zeros), (n + 1) * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_start), (classes + 1) * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(&d_scan, scan_tmp_bytes(n), s));
  const int32_t *perm = nullptr;  // identity
  int side = 0;
  for (int bit = 0; (max_colour >> bit) != 0; ++bit) {
    split_flag_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), perm, d_col, bit, d_flag);
    ASP_LAUNCH_CHECK();
    int rc_scan = scan_exclusive_i64(d_flag, d_zeros, n, d_scan, s);
    if (rc_scan != ASP_OK) return rc_scan;
    split_scatter_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), perm, d_col, bit, d_zeros, d_perm[side]);
    ASP_LAUNCH_CHECK();
    perm = d_perm[side];
    side ^= 1;
  }
  class_start_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), classes, perm, d_col, d_start);
  ASP_LAUNCH_CHECK();
  std::vector<int64_t> start(classes + 1);
  ASP_CUDA_CHECK(cudaMemcpyAsync(start.data(), d_start, (classes + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  plan->class_ptr.assign(classes + 1, 0);
  for (uint32_t c = 0; c < classes; ++c) plan->class_ptr[c + 1] = plan->class_ptr[c] + (start[c + 1] - start[c] + 3) / 4 * 4;
  plan->num_classes = classes;
  plan->n_padded = static_cast<uint64_t>(plan->class_ptr[classes]);
  const uint64_t np = plan->n_padded;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_order), np * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_position), n * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_class_ptr), (classes + 1) * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_indptr), (np + 1) * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_field), np * sizeof(double), s));
  ASP_CUDA_CHECK(cudaMemcpyAsync(plan->d_class_ptr, plan->class_ptr.data(), (classes + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  ASP_CUDA_CHECK(cudaMemsetAsync(plan->d_order, 0xFF, np * sizeof(int32_t), s));
  place_kernel<<<blocks, 256, 0, s>>>(static_cast<uint32_t>(n), perm, d_col, d_start, plan->d_class_ptr, plan->d_order, plan->d_position);
  ASP_LAUNCH_CHECK();
  for (void *p : {static_cast<void *>(d_colour[0]), static_cast<void *>(d_colour[1]), static_cast<void *>(d_remaining), static_cast<void *>(d_perm[0]),
                  static_cast<void *>(d_perm[1]), static_cast<void *>(d_flag), static_cast<void *>(d_zeros), static_cast<void *>(d_start), d_scan})
    ASP_CUDA_CHECK(cudaFreeAsync(p, s));
  lap("positions");
  // 3. relabelled CSR without the diagonal (the diagonal never enters dE)
  int64_t *d_len = nullptr;
  void *d_tmp = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_len), np * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(&d_tmp, scan_tmp_bytes(np), s));
  const unsigned pblocks = static_cast<unsigned>((np + 255) / 256);
  relabel_count_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_order, d_indptr, d_indices, d_len);
  ASP_LAUNCH_CHECK();
  int rc = scan_exclusive_i64(d_len, plan->d_indptr, np, d_tmp, s);
  if (rc != ASP_OK) return rc;
  int64_t nnz = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&nnz, plan->d_indptr + np, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  ASP_CUDA_CHECK(cudaFreeAsync(d_len, s));
  ASP_CUDA_CHECK(cudaFreeAsync(d_tmp, s));
  plan->nnz = static_cast<uint64_t>(nnz);
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_indices), std::max<int64_t>(nnz, 1) * sizeof(int32_t), s));
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_data), std::max<int64_t>(nnz, 1) * sizeof(double), s));
  relabel_fill_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_order, plan->d_position, d_indptr, d_indices, d_data, d_field,
                                              plan->d_indptr, plan->d_indices, plan->d_data, plan->d_field);
  ASP_LAUNCH_CHECK();
  lap("relabelled CSR");
  {  // constant part of the energy: the diagonal, summed in a fixed order
    const unsigned dblocks = static_cast<unsigned>((n + 255) / 256);
    double *d_part = nullptr;
    ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_part), (dblocks + 1) * sizeof(double), s));
    diag_partial_kernel<<<dblocks, 256, 0, s>>>(n, d_indptr, d_indices, d_data, d_part);
    ASP_LAUNCH_CHECK();
    diag_final_kernel<<<1, 256, 0, s>>>(d_part, dblocks, d_part + dblocks);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(&plan->diag_sum, d_part + dblocks, sizeof(double), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaMemsetAsync(d_part, 0, sizeof(double), s));
    max_de_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_indptr, plan->d_data, plan->d_field, reinterpret_cast<unsigned long long *>(d_part));
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(&plan->max_de, d_part, sizeof(double), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    // the sweep kernel's task words; whether any field is non-zero
    uint32_t any_field = 0;
    ASP_CUDA_CHECK(cudaMemsetAsync(d_part, 0, sizeof(double), s));
    if (d_field) {
      any_field_kernel<<<pblocks, 256, 0, s>>>(np, plan->d_field, reinterpret_cast<uint32_t *>(d_part));
      ASP_LAUNCH_CHECK();
    }
    ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&plan->d_bounds), (np / 4) * sizeof(int4), s));
    task_bounds_kernel<<<static_cast<unsigned>((np / 4 + 255) / 256), 256, 0, s>>>(np / 4, plan->d_indptr, plan->d_bounds);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(&any_field, d_part, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    unsigned long long long_tasks = 0;
    unsigned long long *d_long = reinterpret_cast<unsigned long long *>(d_part);
    ASP_CUDA_CHECK(cudaMemsetAsync(d_long, 0, sizeof(unsigned long long), s));
    long_tasks_kernel<<<static_cast<unsigned>((np / 4 + 255) / 256), 256, 0, s>>>(np / 4, plan->d_bounds, d_long);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaMemcpyAsync(&long_tasks, d_long, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    plan->has_field = any_field != 0;
    plan->long_tasks = long_tasks;
    ASP_CUDA_CHECK(cudaFreeAsync(d_part, s));
  }
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  lap("diagonal, bounds, task words");
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
  ASP_REQUIRE(energy_scale * plan->max_de < 2251799813685248.0, "energy_scale too large: |dE| * energy_scale must stay below 2^51");
  ASP_REQUIRE(replica_offset % 32 == 0, "replica_offset must be a multiple of 32");
  const uint32_t groups = (num_replicas + 31) / 32;
  const uint64_t np = plan->n_padded;

  int device = 0, sms = 0, per_sm = 0, coop = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&device));
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  ASP_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  ASP_REQUIRE(coop != 0, "device does not support cooperative launches");
  // (chosen below, once the team size is known)
  void *sweep = nullptr;
  // the wide stage (20 KB of shared memory per CTA, taken from L1) pays when long tasks are common, not for a stray one
  const uint32_t stage_slots = plan->long_tasks * 64 > plan->n_padded / 4 ? kStageSlotsWide : kStageSlots;
  const size_t sweep_smem = static_cast<size_t>(kSaWarps) * stage_slots * sizeof(StagedEntry);
  {
    void (*const quad)(SaArgs) = plan->has_field ? sa_sweep_kernel<true, false> : sa_sweep_kernel<false, false>;
    void (*const fine)(SaArgs) = plan->has_field ? sa_sweep_kernel<true, true> : sa_sweep_kernel<false, true>;
    int per_sm_fine = 0;
    ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quad, kSaThreads, sweep_smem));
    ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fine, fine, kSaThreads, sweep_smem));
    per_sm = std::min(per_sm, per_sm_fine);
  }
  ASP_REQUIRE(per_sm >= 1, "sweep kernel does not fit on an SM");
  const uint32_t max_ctas = static_cast<uint32_t>(sms * per_sm);

  SaArgs a{};
  a.n_padded = np;
  a.indptr = plan->d_indptr;
  a.indices = plan->d_indices;
  a.data = plan->d_data;
  a.field = plan->d_field;
  a.class_ptr = plan->d_class_ptr;
  a.bounds = plan->d_bounds;
  a.stage_slots = stage_slots;
  a.num_classes = plan->num_classes;
  a.groups = groups;
  if (groups >= max_ctas) {
    a.team_size = 1;
    a.num_teams = max_ctas;
  } else {
    a.team_size = max_ctas / groups;
    a.num_teams = groups;
  }
  // A team larger than the work of the biggest class only adds barrier latency.  When the team can have a warp for
  // (nearly) every POSITION of the biggest class, the whole model runs one position per task (the fine kernel); otherwise
  // the team is sized for one 4-position task per warp.
  bool fine_model = false;
  {
    int64_t biggest = 4;
    for (uint32_t c = 0; c < plan->num_classes; ++c) biggest = std::max(biggest, plan->class_ptr[c + 1] - plan->class_ptr[c]);
    uint32_t team_fine = std::max(1u, std::min(a.team_size, static_cast<uint32_t>((biggest + kSaWarps - 1) / kSaWarps)));
    if (g_sa_team_cap > 0) team_fine = std::min(team_fine, static_cast<uint32_t>(g_sa_team_cap));
    fine_model = biggest <= 2ll * team_fine * kSaWarps;
    if (fine_model) {
      a.team_size = team_fine;
    } else {
      const uint32_t useful = static_cast<uint32_t>((biggest / 4 + kSaWarps - 1) / kSaWarps);
      a.team_size = std::max(1u, std::min(a.team_size, useful));
      if (g_sa_team_cap > 0) a.team_size = std::min(a.team_size, static_cast<uint32_t>(g_sa_team_cap));
    }
  }
  sweep = plan->has_field ? (fine_model ? reinterpret_cast<void *>(sa_sweep_kernel<true, true>) : reinterpret_cast<void *>(sa_sweep_kernel<true, false>))
                          : (fine_model ? reinterpret_cast<void *>(sa_sweep_kernel<false, true>) : reinterpret_cast<void *>(sa_sweep_kernel<false, false>));
  a.num_sweeps = num_sweeps;
  a.replica_offset = replica_offset;
  a.seed = seed;
  for (int r = 0; r < 10; ++r) {
    a.round_keys.x[r] = static_cast<uint32_t>(seed) + static_cast<uint32_t>(r) * 0x9E3779B9u;
    a.round_keys.y[r] = static_cast<uint32_t>(seed >> 32) + static_cast<uint32_t>(r) * 0xBB67AE85u;
  }
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
  const size_t ticket_bytes = static_cast<size_t>(a.num_teams) * plan->num_classes * sizeof(unsigned int);
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&a.tickets), ticket_bytes, s));
  ASP_CUDA_CHECK(cudaMemsetAsync(a.tickets, 0, ticket_bytes, s));

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
    ASP_CUDA_CHECK(cudaLaunchCooperativeKernel(sweep, dim3(grid), dim3(kSaThreads), params, sweep_smem, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  {
    const uint64_t warps = (plan->n + 255) / 256;
    sa_unpermute_kernel<<<dim3(static_cast<unsigned>((warps + 7) / 8), groups), 256, 0, s>>>(plan->n, np, num_replicas, plan->d_position, a.best_words, d_best_bits);
    ASP_LAUNCH_CHECK();
  }
  int rc = ASP_OK;
  if (d_best_energy) {
    // exact f64 energies straight from the replica-sliced best configurations
    const uint64_t tasks = np / 4;
    const unsigned en_ctas = static_cast<unsigned>(std::min<uint64_t>((tasks + kEnWarps - 1) / kEnWarps, static_cast<uint64_t>(kNumSMs) * 8));
    const uint64_t total_warps = static_cast<uint64_t>(en_ctas) * kEnWarps;
    double *partial = nullptr;
    ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&partial), groups * total_warps * 32 * sizeof(double), s));
    energy_sliced_kernel<<<dim3(en_ctas, groups), kEnWarps * 32, 0, s>>>(np, plan->d_indptr, plan->d_indices, plan->d_data, plan->d_field, a.best_words, partial);
    ASP_LAUNCH_CHECK();
    energy_sliced_final_kernel<<<groups, 256, 0, s>>>(partial, total_warps, plan->diag_sum, num_replicas, d_best_energy);
    ASP_LAUNCH_CHECK();
    ASP_CUDA_CHECK(cudaFreeAsync(partial, s));
  }
  ASP_CUDA_CHECK(cudaFreeAsync(d_betas, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.words, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.best_words, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.rel, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.best_rel, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.barriers, s));
  ASP_CUDA_CHECK(cudaFreeAsync(a.tickets, s));
  return rc;
}

}  // extern "C"
