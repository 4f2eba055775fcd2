// Single-pass Ising-model extraction on sm_100a (the headline kernel).
//
// Replaces, fused into ONE kernel launch:
//   lattice_symmetries Operator.batched_apply      (called at annealing_sign_problem/common.py:96)
//   _clipped_search_sorted + membership mask       (common.py:116-128, :173)
//   bsearch loop of build_matrix                   (cbits/build_matrix.c:33-51)
//   _make_ising_model_compute_elements             (common.py:71-82)
//   csr_matrix(...) + sort_indices                 (common.py:193-195)
//
// Differences from the two-pass kernels of extract.cu (kept for the count/fill API):
//   * every candidate is searched ONCE.  A tile (128 rows) keeps its hits in shared memory,
//     publishes its coupling count and obtains its CSR offset by decoupled look-back over
//     earlier tiles (tiles are handed out by an atomic ticket, so a tile only ever waits for
//     tiles that are already running), then writes indptr / indices / data in place;
//   * searches run on full warps: a warp walks the delta-sorted move list for its 32 rows,
//     pushes the (row, move) pairs that apply into a FIFO and pops 64 of them at a time, two
//     per lane with interleaved load chains.  The FIFO is move-major, so hits are discovered
//     in ascending column order per row and get their in-row rank on the fly;
//   * the radix index is a 4-byte first-position table with about one key per bucket: most
//     misses are decided by two adjacent table reads without touching the keys.
// Positions are exact indices into the sorted basis (bit-exact CSR), values are
// coef * (|psi_i| * |psi_j|) exactly as in extract.cu.
#include "fused.cuh"

namespace asp {

constexpr int kFxWarps = 4;
constexpr int kFxThreads = kFxWarps * 32;
constexpr int kFxTileRows = kFxThreads;
static_assert(kFxTileRows == kFusedTileRows, "fused.cuh out of sync");
constexpr int kFxQueue = 128;  // FIFO slots per warp (power of two, > 64 + 32 + 31)
constexpr uint32_t kFxMaxMoves = 2046;  // tag layout: move 11 bits in the FIFO / 16 bits in the hit list

constexpr unsigned long long kFlagAggregate = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

struct FusedArgs {
  const uint64_t *spins;   // [n_total] ascending, unique
  const uint32_t *starts;  // [num_buckets + 1] first position with key >= bucket << shift
  uint64_t num_buckets;
  int shift;
  uint32_t n_total;
  const double *psi;
  uint64_t row_begin, num_rows, num_tiles;
  const Move *moves;
  int n_moves, n_down;
  const DiagBond *diag;
  int n_diag;
  int list_cap;  // hit-list entries per warp
  unsigned long long *status;  // [num_tiles] look-back words (zeroed)
  unsigned int *ticket;        // zeroed
  uint64_t capacity;           // entries the caller's indices/data can hold
  int64_t *indptr;
  int32_t *indices;
  double *data;
  unsigned long long *nnz_out;          // running total after this launch (device)
  unsigned long long *nnz_mirror;       // optional copy of it in mapped host memory
  const unsigned long long *base_in;   // couplings emitted by earlier row chunks (NULL = 0)
};

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) build_starts_kernel(const uint64_t *__restrict__ spins, uint32_t n, int shift, uint64_t num_buckets, uint32_t *__restrict__ starts) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const uint64_t last = num_buckets;  // keys wider than the operator's word sort after every bucket
  uint64_t b = spins[i] >> shift;
  if (b > last) b = last;
  uint64_t prev = 0;  // first bucket this thread fills
  if (i > 0) {
    uint64_t pb = spins[i - 1] >> shift;
    if (pb > last) pb = last;
    prev = pb + 1;
  }
  for (uint64_t k = prev; k <= b && k <= last; ++k) starts[k] = i;
  if (i == n - 1)
    for (uint64_t k = b + 1; k <= last; ++k) starts[k] = n;
}

// Joint search of two candidates (independent load chains interleaved): position of cX in
// keys[loX, hiX) or -1.  Keys ascend inside a bucket.
__device__ __forceinline__ void search_two(const uint64_t *__restrict__ keys, uint64_t cA, uint32_t loA, uint32_t hiA, uint64_t cB, uint32_t loB,
                                           uint32_t hiB, int32_t &posA, int32_t &posB) {
  while (hiA - loA > 8) {  // pathological bucket: bisect down to a short scan
    const uint32_t mid = loA + ((hiA - loA) >> 1);
    if (__ldg(&keys[mid]) < cA)
      loA = mid + 1;
    else
      hiA = mid + 1;
  }
  while (hiB - loB > 8) {
    const uint32_t mid = loB + ((hiB - loB) >> 1);
    if (__ldg(&keys[mid]) < cB)
      loB = mid + 1;
    else
      hiB = mid + 1;
  }
  posA = -1;
  posB = -1;
  bool actA = loA < hiA, actB = loB < hiB;
  while (actA | actB) {
    const uint64_t kA = actA ? __ldg(&keys[loA]) : 0ull;
    const uint64_t kB = actB ? __ldg(&keys[loB]) : 0ull;
    if (actA) {
      if (kA >= cA) {
        if (kA == cA) posA = static_cast<int32_t>(loA);
        actA = false;
      } else if (++loA >= hiA) {
        actA = false;
      }
    }
    if (actB) {
      if (kB >= cB) {
        if (kB == cB) posB = static_cast<int32_t>(loB);
        actB = false;
      } else if (++loB >= hiB) {
        actB = false;
      }
    }
  }
}

__device__ __forceinline__ int32_t search_one(const FusedArgs &a, uint64_t c) {
  const uint64_t b = c >> a.shift;
  if (b >= a.num_buckets) return -1;
  uint32_t lo = __ldg(&a.starts[b]), hi = __ldg(&a.starts[b + 1]);
  while (hi - lo > 8) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&a.spins[mid]) < c)
      lo = mid + 1;
    else
      hi = mid + 1;
  }
  for (; lo < hi; ++lo) {
    const uint64_t k = __ldg(&a.spins[lo]);
    if (k >= c) return k == c ? static_cast<int32_t>(lo) : -1;
  }
  return -1;
}

// Per-warp shared-memory state.
struct WarpSmem {
  uint32_t *list_pos;  // [list_cap]
  uint32_t *list_tag;  // [list_cap]  move | row lane << 16 | in-row rank << 21
  uint16_t *queue;     // [kFxQueue]  move << 5 | row lane
  uint32_t *cnt;       // [32] couplings found so far per row
  double *abs_psi;     // [32]
  uint32_t *row_off;   // [32] row start relative to the tile's first coupling
};

__global__ void __launch_bounds__(kFxThreads) extract_csr_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- shared memory carve-up: move table (SoA), diagonal table, per-warp state ------------
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *s_need = s_mask + a.n_moves;
  uint64_t *s_flip = s_need + a.n_moves;
  double *s_coef = reinterpret_cast<double *>(s_flip + a.n_moves);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(s_coef + a.n_moves);
  unsigned char *cursor = reinterpret_cast<unsigned char *>(s_diag + a.n_diag);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  WarpSmem ws;
  {
    const size_t per_warp = static_cast<size_t>(a.list_cap) * 8 + kFxQueue * 2 + 32 * 4 + 32 * 8 + 32 * 4;
    unsigned char *base = cursor + per_warp * warp;
    ws.abs_psi = reinterpret_cast<double *>(base);
    ws.list_pos = reinterpret_cast<uint32_t *>(base + 32 * 8);
    ws.list_tag = ws.list_pos + a.list_cap;
    ws.cnt = ws.list_tag + a.list_cap;
    ws.row_off = ws.cnt + 32;
    ws.queue = reinterpret_cast<uint16_t *>(ws.row_off + 32);
  }
  __shared__ unsigned int s_tile;
  __shared__ unsigned int s_warp_total[kFxWarps];
  __shared__ unsigned long long s_tile_base;

  for (int k = threadIdx.x; k < a.n_moves; k += kFxThreads) {
    const Move mv = a.moves[k];
    s_mask[k] = mv.mask;
    s_need[k] = mv.need;
    s_flip[k] = mv.flip;
    s_coef[k] = mv.coef;
  }
  for (int k = threadIdx.x; k < a.n_diag; k += kFxThreads) s_diag[k] = a.diag[k];

  for (;;) {
    __syncthreads();  // previous tile fully written; tables loaded
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= a.num_tiles) break;

    const uint64_t r = tile * kFxTileRows + threadIdx.x;
    const bool live = r < a.num_rows;
    const uint64_t row = a.row_begin + r;
    const uint64_t s = live ? a.spins[row] : 0ull;
    const uint32_t s_lo = static_cast<uint32_t>(s), s_hi = static_cast<uint32_t>(s >> 32);

    // =========================== phase 1: generate, compact, search ==========================
    ws.cnt[lane] = 0;
    uint32_t list_count = 0;  // warp-uniform
    uint32_t qhead = 0, qcount = 0;
    int overflow_from = a.n_moves;  // first move handled by the lane-per-row fallback
    uint32_t down_cnt = 0;
    __syncwarp();

    // pops `nb` (<= 64) FIFO entries: lane handles entries lane and lane + 32
    auto process = [&](uint32_t nb) {
      const bool vA = lane < nb, vB = lane + 32 < nb;
      const uint32_t tagA = ws.queue[(qhead + lane) & (kFxQueue - 1)];
      const uint32_t tagB = ws.queue[(qhead + 32 + lane) & (kFxQueue - 1)];
      const uint32_t srcA = tagA & 31u, srcB = tagB & 31u;
      const uint32_t mA = vA ? (tagA >> 5) : 0u, mB = vB ? (tagB >> 5) : 0u;
      const uint64_t sA = (static_cast<uint64_t>(__shfl_sync(0xffffffffu, s_hi, srcA)) << 32) | __shfl_sync(0xffffffffu, s_lo, srcA);
      const uint64_t sB = (static_cast<uint64_t>(__shfl_sync(0xffffffffu, s_hi, srcB)) << 32) | __shfl_sync(0xffffffffu, s_lo, srcB);
      const uint64_t cA = sA ^ s_flip[mA], cB = sB ^ s_flip[mB];
      const uint64_t bA = cA >> a.shift, bB = cB >> a.shift;
      uint32_t loA = 0, hiA = 0, loB = 0, hiB = 0;
      if (vA && bA < a.num_buckets) {
        loA = __ldg(&a.starts[bA]);
        hiA = __ldg(&a.starts[bA + 1]);
      }
      if (vB && bB < a.num_buckets) {
        loB = __ldg(&a.starts[bB]);
        hiB = __ldg(&a.starts[bB + 1]);
      }
      int32_t posA, posB;
      search_two(a.spins, cA, loA, hiA, cB, loB, hiB, posA, posB);
      // record hits: A entries precede B entries in FIFO (= move-major) order
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int32_t pos = half ? posB : posA;
        const uint32_t src = half ? srcB : srcA, m = half ? mB : mA;
        const bool hit = pos >= 0;
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits == 0) continue;
        if (hit) {
          const uint32_t peers = __match_any_sync(hits, src);  // hits of the same row in this half-batch
          const uint32_t base = ws.cnt[src];
          __syncwarp(hits);
          if ((peers & lt_mask) == 0) ws.cnt[src] = base + __popc(peers);
          const uint32_t slot = list_count + __popc(hits & lt_mask);
          ws.list_pos[slot] = static_cast<uint32_t>(pos);
          ws.list_tag[slot] = m | (src << 16) | ((base + __popc(peers & lt_mask)) << 21);
        }
        list_count += __popc(hits);
        __syncwarp();
      }
      qhead += nb;
      qcount -= nb;
    };

    for (int m = 0; m < a.n_moves; ++m) {
      if (list_count + qcount + 32 > static_cast<uint32_t>(a.list_cap)) {  // hit list may not take this move
        overflow_from = m;
        break;
      }
      if (m == a.n_down) {  // the diagonal sits between the negative and the positive deltas
        while (qcount > 0) {
          __syncwarp();
          process(min(qcount, 64u));
        }
        __syncwarp();
        down_cnt = ws.cnt[lane];
        if (live) ws.cnt[lane] = down_cnt + 1;
        __syncwarp();
      }
      const bool app = live && ((s & s_mask[m]) == s_need[m]);
      const uint32_t bits = __ballot_sync(0xffffffffu, app);
      if (app) ws.queue[(qhead + qcount + __popc(bits & lt_mask)) & (kFxQueue - 1)] = static_cast<uint16_t>((m << 5) | lane);
      qcount += __popc(bits);
      if (qcount >= 64) {
        __syncwarp();
        process(64);
      }
    }
    while (qcount > 0) {
      __syncwarp();
      process(min(qcount, 64u));
    }
    __syncwarp();
    if (a.n_down >= a.n_moves && overflow_from == a.n_moves) {  // no positive-delta move: diagonal goes last
      down_cnt = ws.cnt[lane];
      if (live) ws.cnt[lane] = down_cnt + 1;
      __syncwarp();
    }
    // lane-per-row fallback for the moves the hit list could not take: count now, write later
    const uint32_t listed_cnt = ws.cnt[lane];
    uint32_t my_cnt = listed_cnt;
    if (overflow_from < a.n_moves) {
      if (live) {
        for (int m = overflow_from; m < a.n_moves; ++m) {
          if (m == a.n_down) down_cnt = my_cnt++;
          if ((s & s_mask[m]) != s_need[m]) continue;
          if (search_one(a, s ^ s_flip[m]) >= 0) ++my_cnt;
        }
        if (a.n_down >= a.n_moves) down_cnt = my_cnt++;
      }
    }

    // =========================== tile offset by decoupled look-back ==========================
    uint32_t incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_total[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kFxWarps; ++w) {
      if (w < static_cast<int>(warp)) warp_base += s_warp_total[w];
      tile_total += s_warp_total[w];
    }
    if (warp == 0) {
      unsigned long long exclusive = 0;
      if (tile == 0) {
        if (a.base_in) exclusive = *a.base_in;
        if (lane == 0) st_status(&a.status[0], kFlagPrefix | (exclusive + tile_total));
      } else {
        if (lane == 0) st_status(&a.status[tile], kFlagAggregate | tile_total);
        int64_t look = static_cast<int64_t>(tile) - 1;  // window [look - 31, look]
        for (;;) {
          const int64_t idx = look - lane;
          unsigned long long st = kFlagPrefix;  // virtual tiles before tile 0: prefix 0
          if (idx >= 0) {
            do {
              st = ld_status(&a.status[idx]);
            } while ((st >> 62) == 0);
          }
          const uint32_t has_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2);
          const uint32_t first = has_prefix ? static_cast<uint32_t>(__ffs(has_prefix)) - 1u : 32u;  // nearest tile with a prefix
          unsigned long long contrib = lane <= first ? (st & kValueMask) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
          exclusive += contrib;
          if (has_prefix) break;
          look -= 32;
        }
        if (lane == 0) st_status(&a.status[tile], kFlagPrefix | (exclusive + tile_total));
      }
      if (lane == 0) {
        s_tile_base = exclusive;
        if (tile == a.num_tiles - 1) {
          a.indptr[a.num_rows] = static_cast<int64_t>(exclusive + tile_total);
          *a.nnz_out = exclusive + tile_total;
          if (a.nnz_mirror) {
            *a.nnz_mirror = exclusive + tile_total;
            __threadfence_system();
          }
        }
      }
    }
    __syncthreads();
    const uint64_t tile_base = s_tile_base;

    // =========================== phase 2: write the tile's CSR rows ==========================
    const uint32_t my_off = warp_base + incl - my_cnt;  // row start relative to the tile
    const double a_i = live ? fabs(a.psi[row]) : 0.0;
    ws.row_off[lane] = my_off;
    ws.abs_psi[lane] = a_i;
    if (live) a.indptr[r] = static_cast<int64_t>(tile_base + my_off);
    __syncwarp();
    for (uint32_t k = lane; k < list_count; k += 32) {
      const uint32_t tag = ws.list_tag[k], pos = ws.list_pos[k];
      const uint32_t m = tag & 0xFFFFu, src = (tag >> 16) & 31u, rank = tag >> 21;
      const uint64_t dest = tile_base + ws.row_off[src] + rank;
      if (dest < a.capacity) {
        a.indices[dest] = static_cast<int32_t>(pos);
        a.data[dest] = s_coef[m] * (ws.abs_psi[src] * fabs(__ldg(&a.psi[pos])));
      }
    }
    if (live) {
      double d = 0.0;
      for (int k = 0; k < a.n_diag; ++k) {
        const DiagBond db = s_diag[k];
        const int idx = static_cast<int>(((s >> db.i) & 1) * 2 + ((s >> db.j) & 1));
        d += db.d[idx];
      }
      const uint64_t dest = tile_base + my_off + down_cnt;
      if (dest < a.capacity) {
        a.indices[dest] = static_cast<int32_t>(row);
        a.data[dest] = d * (a_i * a_i);
      }
      if (overflow_from < a.n_moves) {  // redo the fallback moves, now writing in place
        uint32_t c = listed_cnt;
        for (int m = overflow_from; m < a.n_moves; ++m) {
          if (m == a.n_down) ++c;
          if ((s & s_mask[m]) != s_need[m]) continue;
          const int32_t pos = search_one(a, s ^ s_flip[m]);
          if (pos < 0) continue;
          const uint64_t at = tile_base + my_off + c;
          if (at < a.capacity) {
            a.indices[at] = pos;
            a.data[at] = s_coef[m] * (a_i * fabs(__ldg(&a.psi[pos])));
          }
          ++c;
        }
      }
    }
  }
}

constexpr int kFxMaxChunks = kFusedMaxChunks;  // row chunks of one pipelined host call

struct FusedWorkspace {
  uint32_t *starts;
  unsigned long long *status;  // [num_tiles + kFxMaxChunks]: every row chunk has its own slice
  unsigned int *tickets;       // [kFxMaxChunks] one per chunk, 64 B apart
  unsigned long long *totals;  // [kFxMaxChunks] running totals
  uint64_t num_buckets;
  int shift;
  size_t bytes, zero_offset, zero_bytes;
};

static FusedWorkspace carve_fused(void *base, const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  FusedWorkspace w;
  int lg = 0;
  while ((1ull << lg) < n_total) ++lg;
  int bits = lg;  // about one key per bucket
  if (bits < 4) bits = 4;
  if (bits > 26) bits = 26;
  const int key_bits = static_cast<int>(op->number_spins);
  if (bits > key_bits) bits = key_bits;
  w.shift = key_bits - bits;
  w.num_buckets = 1ull << bits;
  const uint64_t tiles = (num_rows + kFxTileRows - 1) / kFxTileRows + kFxMaxChunks;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void *p = base ? static_cast<char *>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.starts = static_cast<uint32_t *>(take((w.num_buckets + 1) * sizeof(uint32_t)));
  w.zero_offset = off;
  w.status = static_cast<unsigned long long *>(take(tiles * sizeof(unsigned long long)));
  w.tickets = static_cast<unsigned int *>(take(kFxMaxChunks * 64));
  w.totals = static_cast<unsigned long long *>(take(kFxMaxChunks * sizeof(unsigned long long)));
  w.zero_bytes = off - w.zero_offset;
  w.bytes = off;
  return w;
}

static int g_list_cap_override = 0;

size_t fused_workspace_bytes(const asp_operator *op, uint64_t n_total, uint64_t num_rows) {
  return carve_fused(nullptr, op, n_total, num_rows).bytes;
}

int fused_check_operator(const asp_operator *op) {
  ASP_REQUIRE(op != nullptr, "operator is NULL");
  ASP_REQUIRE(op->d_moves != nullptr, "operator has no device mirror (created without a CUDA device)");
  if (!op->sorted_emitter() || op->moves.size() > kFxMaxMoves) {
    set_error("fused extraction needs an unsymmetrised operator with distinct moves (at most %u); use the apply + build_matrix + canonicalise path", kFxMaxMoves);
    return ASP_ERR_UNSUPPORTED;
  }
  return ASP_OK;
}

// Zero the look-back state and index the sorted basis (once per call, before the chunks).
int fused_prepare(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, uint64_t num_rows, void *d_workspace,
                  size_t workspace_bytes, cudaStream_t s) {
  FusedWorkspace w = carve_fused(d_workspace, op, n_total, num_rows);
  if (d_workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return ASP_ERR_WORKSPACE;
  }
  ASP_CUDA_CHECK(cudaMemsetAsync(static_cast<char *>(d_workspace) + w.zero_offset, 0, w.zero_bytes, s));
  build_starts_kernel<<<static_cast<unsigned>((n_total + 255) / 256), 256, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), w.shift, w.num_buckets, w.starts);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

// Rows [row_begin + chunk_begin, +chunk_rows) of the block that starts at row_begin: chunk
// number `chunk` (< kFxMaxChunks) of a call whose earlier chunks cover [0, chunk_begin).
// d_indptr is the BLOCK's indptr; offsets continue from the previous chunk's total.
// The running total lands in the workspace (fused_total()) and, when nnz_mirror != NULL, in
// that (mapped host) location too.
int fused_launch(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, const double *d_psi, uint64_t row_begin,
                 uint64_t num_rows, int chunk, uint64_t chunk_begin, uint64_t chunk_rows, void *d_workspace, uint64_t capacity,
                 int64_t *d_indptr, int32_t *d_indices, double *d_data, unsigned long long *nnz_mirror, cudaStream_t s) {
  ASP_REQUIRE(chunk >= 0 && chunk < kFxMaxChunks, "too many row chunks");
  ASP_REQUIRE(chunk_begin % kFxTileRows == 0, "row chunks start at tile boundaries");
  FusedWorkspace w = carve_fused(d_workspace, op, n_total, num_rows);
  FusedArgs a{};
  a.spins = d_spins;
  a.starts = w.starts;
  a.num_buckets = w.num_buckets;
  a.shift = w.shift;
  a.n_total = static_cast<uint32_t>(n_total);
  a.psi = d_psi;
  a.row_begin = row_begin + chunk_begin;
  a.num_rows = chunk_rows;
  a.num_tiles = (chunk_rows + kFxTileRows - 1) / kFxTileRows;
  a.moves = op->d_moves;
  a.n_moves = static_cast<int>(op->moves.size());
  a.n_down = static_cast<int>(op->n_down);
  a.diag = op->d_diag;
  a.n_diag = static_cast<int>(op->diag.size());
  a.status = w.status + chunk_begin / kFxTileRows + chunk;
  a.ticket = w.tickets + 16 * chunk;
  a.capacity = capacity;
  a.indptr = d_indptr + chunk_begin;
  a.indices = d_indices;
  a.data = d_data;
  a.nnz_out = w.totals + chunk;
  a.nnz_mirror = nnz_mirror;
  a.base_in = chunk == 0 ? nullptr : w.totals + (chunk - 1);
  // hit list: room for every candidate of 32 rows when that is small, else 40 per row (more is
  // handled by the in-kernel fallback)
  int cap = static_cast<int>(std::min<uint64_t>(32ull * op->max_candidates(), 1280));
  if (g_list_cap_override > 0) cap = g_list_cap_override;
  if (cap < 96) cap = 96;
  cap = (cap + 31) / 32 * 32;
  a.list_cap = cap;
  const size_t tables = op->moves.size() * 32 + op->diag.size() * sizeof(DiagBond);
  const size_t per_warp = static_cast<size_t>(cap) * 8 + kFxQueue * 2 + 32 * 4 + 32 * 8 + 32 * 4;
  const size_t smem = align_up(tables, 16) + per_warp * kFxWarps + 16;
  ASP_REQUIRE(smem <= 200 * 1024, "operator too large for the fused kernel's shared-memory tables");
  ASP_CUDA_CHECK(cudaFuncSetAttribute(extract_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  ASP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extract_csr_kernel, kFxThreads, smem));
  ASP_REQUIRE(per_sm >= 1, "fused extraction kernel does not fit on an SM");
  const uint64_t resident = static_cast<uint64_t>(kNumSMs) * per_sm;
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(a.num_tiles, resident));
  extract_csr_kernel<<<grid, kFxThreads, smem, s>>>(a);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

const unsigned long long *fused_total(const asp_operator *op, uint64_t n_total, uint64_t num_rows, void *d_workspace, int chunk) {
  return carve_fused(d_workspace, op, n_total, num_rows).totals + chunk;
}

}  // namespace asp

using namespace asp;

extern "C" {

void asp_debug_set_hit_list_capacity(int entries_per_warp) { g_list_cap_override = entries_per_warp; }

size_t asp_extract_csr_workspace_bytes(asp_operator const *op, uint64_t n_total, uint64_t num_rows) {
  if (!op) return 0;
  return fused_workspace_bytes(op, n_total, num_rows);
}

int asp_extract_csr(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                    uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                    uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                    void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  int rc = fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(n_total < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n_total, "row block exceeds the basis");
  ASP_REQUIRE(d_indptr != nullptr, "d_indptr is NULL");
  ASP_REQUIRE(capacity == 0 || (d_indices && d_data), "NULL output buffer");
  if (num_rows == 0 || n_total == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), s));
    if (h_nnz) {
      ASP_CUDA_CHECK(cudaStreamSynchronize(s));
      *h_nnz = 0;
    }
    return ASP_OK;
  }
  ASP_REQUIRE(d_spins && d_psi, "NULL input buffer");
  rc = fused_prepare(op, n_total, d_spins, num_rows, d_workspace, workspace_bytes, s);
  if (rc != ASP_OK) return rc;
  rc = fused_launch(op, n_total, d_spins, d_psi, row_begin, num_rows, 0, 0, num_rows, d_workspace, capacity, d_indptr, d_indices,
                    d_data, nullptr, s);
  if (rc != ASP_OK) return rc;
  if (h_nnz) {
    unsigned long long total = 0;
    ASP_CUDA_CHECK(cudaMemcpyAsync(&total, fused_total(op, n_total, num_rows, d_workspace, 0), sizeof(total), cudaMemcpyDeviceToHost, s));
    ASP_CUDA_CHECK(cudaStreamSynchronize(s));
    *h_nnz = total;
    if (total > capacity) {
      set_error("output capacity too small: %llu couplings, room for %llu (indptr is complete; call again with the larger capacity)",
                total, static_cast<unsigned long long>(capacity));
      return ASP_ERR_WORKSPACE;
    }
  }
  return ASP_OK;
}

}  // extern "C"
