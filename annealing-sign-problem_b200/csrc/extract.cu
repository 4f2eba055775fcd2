// Ising-model extraction on sm_100a.
//
// Replaces, fused into one path:
//   lattice_symmetries Operator.batched_apply      (called at annealing_sign_problem/common.py:96)
//   _clipped_search_sorted + membership mask       (common.py:116-128, :173)
//   bsearch loop of build_matrix                   (cbits/build_matrix.c:33-51)
//   _make_ising_model_compute_elements             (common.py:71-82)
//   csr_matrix(...) + sort_indices                 (common.py:193-195)
//
// Design (DESIGN.md section "Extraction"):
//   * the sorted basis is indexed once by a radix table over the top bits of the keys
//     (uint2 {first,last+1} per bucket); a lookup is ONE 8-byte table read plus a short
//     search inside the bucket instead of ~log2(n) dependent probes;
//   * one lane per row, a warp covers 32 consecutive sorted rows: for a given move the 32
//     candidates land in neighbouring buckets/keys, so table and key reads of a warp fall
//     into a handful of sectors (the "warp-cooperative" part is the access pattern);
//   * moves are pre-sorted by their signed key delta, so every row is emitted in ascending
//     column order and needs no sort pass;
//   * two passes: count (row counts + tile totals) -> scan of tile totals -> fill.
#include "operator.cuh"

namespace asp {

// ===================================================================================
// Exclusive scan of int64 (3 phases; tiles of 2048 elements).
// ===================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ tile_sums, uint64_t m) {
  __shared__ int64_t smem[33];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
  int64_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t idx = base + threadIdx.x + static_cast<uint64_t>(k) * kScanThreads;
    if (idx < m) acc += in[idx];
  }
  int64_t total;
  block_exclusive_scan_i64(acc, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of `count` values, total -> sums[count]
__global__ void __launch_bounds__(1024) scan_spine_kernel(int64_t *__restrict__ sums, uint64_t count) {
  __shared__ int64_t smem[33];
  int64_t carry = 0;
  for (uint64_t start = 0; start < count; start += blockDim.x) {
    const uint64_t idx = start + threadIdx.x;
    const int64_t v = idx < count ? sums[idx] : 0;
    int64_t total;
    const int64_t ex = block_exclusive_scan_i64(v, &total, smem);
    if (idx < count) sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) sums[count] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ out, const int64_t *__restrict__ tile_sums, uint64_t m, uint64_t num_tiles) {
  __shared__ int64_t smem[33];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
  int64_t v[kScanItems];
  int64_t acc = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < m) ? in[base + k] : 0;
    acc += v[k];
  }
  int64_t total;
  int64_t ex = block_exclusive_scan_i64(acc, &total, smem) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < m) out[base + k] = ex;
    ex += v[k];
  }
  if (blockIdx.x == num_tiles - 1 && threadIdx.x == 0) out[m] = tile_sums[num_tiles];
}

size_t scan_tmp_bytes(uint64_t m) {
  const uint64_t tiles = (m + kScanTile - 1) / kScanTile;
  return align_up((tiles + 2) * sizeof(int64_t), 256);
}

int scan_exclusive_i64(const int64_t *d_in, int64_t *d_out, uint64_t m, void *d_tmp, cudaStream_t s) {
  if (m == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(int64_t), s));
    return ASP_OK;
  }
  const uint64_t tiles = (m + kScanTile - 1) / kScanTile;
  auto *sums = static_cast<int64_t *>(d_tmp);
  scan_reduce_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(d_in, sums, m);
  ASP_LAUNCH_CHECK();
  scan_spine_kernel<<<1, 1024, 0, s>>>(sums, tiles);
  ASP_LAUNCH_CHECK();
  scan_apply_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(d_in, d_out, sums, m, tiles);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

// ===================================================================================
// Radix index of the sorted basis.
// ===================================================================================
struct BasisIndex {
  const uint64_t *spins;  // [n_total] ascending, unique
  const uint2 *table;     // [num_buckets] {first, last+1}; {0,0} = empty
  uint32_t n_total;
  int shift;              // bucket = key >> shift
  uint64_t num_buckets;
};

__global__ void __launch_bounds__(256) build_table_kernel(const uint64_t *__restrict__ spins, uint32_t n, int shift, uint64_t num_buckets, uint2 *__restrict__ table) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const uint64_t b = spins[i] >> shift;
  if (b >= num_buckets) return;  // key wider than the operator's word: never a candidate
  if (i == 0 || (spins[i - 1] >> shift) != b) table[b].x = i;
  if (i == n - 1 || (spins[i + 1] >> shift) != b) table[b].y = i + 1;
}

// Position of key c in the basis, or -1.
__device__ __forceinline__ int64_t basis_find(const BasisIndex &ix, uint64_t c) {
  const uint64_t b = c >> ix.shift;
  if (b >= ix.num_buckets) return -1;
  const uint2 se = __ldg(&ix.table[b]);
  uint32_t lo = se.x, hi = se.y;
  const uint32_t end = se.y;
  while (hi - lo > 4) {  // invariant: first key >= c lies in [lo, hi]
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&ix.spins[mid]) < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  for (; lo < end; ++lo) {  // at most 5 steps: spins[hi] >= c when hi < end
    const uint64_t k = __ldg(&ix.spins[lo]);
    if (k >= c) return k == c ? static_cast<int64_t>(lo) : -1;
  }
  return -1;
}

static int choose_table_bits(uint64_t n_total) {
  int lg = 0;
  while ((1ull << lg) < n_total) ++lg;
  int bits = lg - 2;  // ~4 keys per bucket on average
  if (bits < 4) bits = 4;
  if (bits > 26) bits = 26;
  return bits;
}

// ===================================================================================
// Fused sorted-emitter extraction (no symmetries): lane per row.
// ===================================================================================
constexpr int kTileRows = 256;

struct ExtractArgs {
  BasisIndex ix;
  const double *psi;
  uint64_t row_begin, num_rows;
  const Move *moves;
  int n_moves, n_down;
  const DiagBond *diag;
  int n_diag;
  uint16_t *row_counts;  // [num_rows]
  int64_t *tile_sums;    // count: totals out; fill: exclusive bases in
  int64_t *indptr;
  int32_t *indices;
  double *data;
};

template <bool kFill>
__global__ void __launch_bounds__(kTileRows) extract_sorted_kernel(const ExtractArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Move *s_moves = reinterpret_cast<Move *>(smem_raw);
  DiagBond *s_diag = reinterpret_cast<DiagBond *>(s_moves + a.n_moves);
  __shared__ int64_t s_scan[33];

  for (int k = threadIdx.x; k < a.n_moves; k += blockDim.x) s_moves[k] = a.moves[k];
  if (kFill)
    for (int k = threadIdx.x; k < a.n_diag; k += blockDim.x) s_diag[k] = a.diag[k];
  __syncthreads();

  const uint64_t num_tiles = (a.num_rows + kTileRows - 1) / kTileRows;
  for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint64_t r = tile * kTileRows + threadIdx.x;
    const bool live = r < a.num_rows;
    const uint64_t row = a.row_begin + r;
    uint64_t s = 0;
    if (live) s = a.ix.spins[row];

    int64_t out = 0;
    double a_i = 0.0;
    if (kFill) {
      const int64_t mine = live ? a.row_counts[r] : 0;
      int64_t total;
      out = block_exclusive_scan_i64(mine, &total, s_scan) + a.tile_sums[tile];
      if (live) {
        a.indptr[r] = out;
        if (r == a.num_rows - 1) a.indptr[a.num_rows] = out + mine;
        a_i = fabs(a.psi[row]);
      }
    }
    int count = 0;
    if (live) {
      for (int m = 0; m < a.n_down; ++m) {
        const Move mv = s_moves[m];
        if ((s & mv.mask) != mv.need) continue;
        const int64_t pos = basis_find(a.ix, s ^ mv.flip);
        if (pos < 0) continue;
        if (kFill) {
          a.indices[out + count] = static_cast<int32_t>(pos);
          a.data[out + count] = __dmul_rn(__dmul_rn(mv.coef, fabs(__ldg(&a.psi[pos]))), a_i);  // (c |psi_j|) |psi_i|: common.py:71-82
        }
        ++count;
      }
      // the diagonal entry sits between the negative and the positive deltas
      if (kFill) {
        double d = 0.0;
        for (int k = 0; k < a.n_diag; ++k) {
          const DiagBond db = s_diag[k];
          const int idx = static_cast<int>(((s >> db.i) & 1) * 2 + ((s >> db.j) & 1));
          d += db.d[idx];
        }
        a.indices[out + count] = static_cast<int32_t>(row);
        a.data[out + count] = __dmul_rn(__dmul_rn(d, a_i), a_i);
      }
      ++count;
      for (int m = a.n_down; m < a.n_moves; ++m) {
        const Move mv = s_moves[m];
        if ((s & mv.mask) != mv.need) continue;
        const int64_t pos = basis_find(a.ix, s ^ mv.flip);
        if (pos < 0) continue;
        if (kFill) {
          a.indices[out + count] = static_cast<int32_t>(pos);
          a.data[out + count] = __dmul_rn(__dmul_rn(mv.coef, fabs(__ldg(&a.psi[pos]))), a_i);  // (c |psi_j|) |psi_i|: common.py:71-82
        }
        ++count;
      }
    }
    if (!kFill) {
      if (live) a.row_counts[r] = static_cast<uint16_t>(count);
      int64_t total;
      block_exclusive_scan_i64(count, &total, s_scan);
      if (threadIdx.x == 0) a.tile_sums[tile] = total;
    }
  }
}

// Workspace layout for the fused path.
struct Workspace {
  uint2 *table;
  uint16_t *row_counts;
  int64_t *tile_sums;  // [tiles + 1]
  int shift;
  uint64_t num_buckets;
  uint64_t tiles;
  size_t bytes;
};

static Workspace carve_workspace(void *base, uint64_t n_total, uint64_t num_rows) {
  Workspace w;
  const int bits = choose_table_bits(n_total);
  w.num_buckets = 1ull << bits;
  w.shift = 0;  // set from the operator's word length by the callers
  w.tiles = (num_rows + kTileRows - 1) / kTileRows;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void *p = base ? static_cast<char *>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.table = static_cast<uint2 *>(take(w.num_buckets * sizeof(uint2)));
  w.row_counts = static_cast<uint16_t *>(take(std::max<uint64_t>(num_rows, 1) * sizeof(uint16_t)));
  w.tile_sums = static_cast<int64_t *>(take((w.tiles + 2) * sizeof(int64_t)));
  w.bytes = off;
  return w;
}

static int grid_for(uint64_t tiles, int ctas_per_sm) {
  const uint64_t cap = static_cast<uint64_t>(kNumSMs) * ctas_per_sm;
  return static_cast<int>(tiles < cap ? (tiles ? tiles : 1) : cap);
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_scan_tmp_bytes(uint64_t m) { return scan_tmp_bytes(m); }

int asp_exclusive_scan_i64(int64_t const *d_in, int64_t *d_out, uint64_t m, void *d_tmp, void *stream) {
  ASP_REQUIRE(d_out && (m == 0 || (d_in && d_tmp)), "NULL buffer");
  return scan_exclusive_i64(d_in, d_out, m, d_tmp, static_cast<cudaStream_t>(stream));
}

size_t asp_extract_workspace_bytes(asp_operator const *op, uint64_t n_total, uint64_t num_rows) {
  (void)op;
  return carve_workspace(nullptr, n_total, num_rows).bytes;
}

static int fused_args(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, uint64_t row_begin,
                      uint64_t num_rows, void *d_workspace, size_t workspace_bytes, Workspace &w, ExtractArgs &a) {
  ASP_REQUIRE(op != nullptr, "operator is NULL");
  ASP_REQUIRE(op->d_moves != nullptr, "operator has no device mirror (created without a CUDA device)");
  ASP_REQUIRE(n_total < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n_total, "row block exceeds the basis");
  ASP_REQUIRE(op->max_candidates() < 65535, "too many candidates per row for 16-bit row counts");
  if (!op->sorted_emitter()) {
    set_error("fused sorted extraction needs an unsymmetrised operator with distinct moves; use the apply + build_matrix + canonicalise path");
    return ASP_ERR_UNSUPPORTED;
  }
  w = carve_workspace(d_workspace, n_total, num_rows);
  if (d_workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return ASP_ERR_WORKSPACE;
  }
  a.ix.spins = d_spins;
  a.ix.table = w.table;
  a.ix.n_total = static_cast<uint32_t>(n_total);
  a.ix.num_buckets = w.num_buckets;
  a.row_begin = row_begin;
  a.num_rows = num_rows;
  a.moves = op->d_moves;
  a.n_moves = static_cast<int>(op->moves.size());
  a.n_down = static_cast<int>(op->n_down);
  a.diag = op->d_diag;
  a.n_diag = static_cast<int>(op->diag.size());
  a.row_counts = w.row_counts;
  a.tile_sums = w.tile_sums;
  return ASP_OK;
}

static size_t fused_smem(asp_operator const *op) {
  return op->moves.size() * sizeof(Move) + op->diag.size() * sizeof(DiagBond) + 16;
}

int asp_extract_count(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, uint64_t row_begin,
                      uint64_t num_rows, void *d_workspace, size_t workspace_bytes, uint64_t *h_nnz, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  Workspace w;
  ExtractArgs a{};
  int rc = fused_args(op, n_total, d_spins, row_begin, num_rows, d_workspace, workspace_bytes, w, a);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(h_nnz != nullptr, "h_nnz is NULL");
  if (num_rows == 0 || n_total == 0) {
    *h_nnz = 0;
    return ASP_OK;
  }
  // the shift is fixed by the operator's word length so that count and fill agree
  const int key_bits = static_cast<int>(op->number_spins);
  int bits = 0;
  while ((1ull << bits) < w.num_buckets) ++bits;
  a.ix.shift = key_bits > bits ? key_bits - bits : 0;
  ASP_CUDA_CHECK(cudaMemsetAsync(w.table, 0, w.num_buckets * sizeof(uint2), s));
  build_table_kernel<<<static_cast<unsigned>((n_total + 255) / 256), 256, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), a.ix.shift, w.num_buckets, w.table);
  ASP_LAUNCH_CHECK();
  const size_t smem = fused_smem(op);
  ASP_CUDA_CHECK(cudaFuncSetAttribute(extract_sorted_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  extract_sorted_kernel<false><<<grid_for(w.tiles, 8), kTileRows, smem, s>>>(a);
  ASP_LAUNCH_CHECK();
  scan_spine_kernel<<<1, 1024, 0, s>>>(w.tile_sums, w.tiles);
  ASP_LAUNCH_CHECK();
  int64_t total = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&total, w.tile_sums + w.tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  *h_nnz = static_cast<uint64_t>(total);
  return ASP_OK;
}

int asp_extract_fill(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins, double const *d_psi,
                     uint64_t row_begin, uint64_t num_rows, void *d_workspace, size_t workspace_bytes,
                     int64_t *d_indptr, int32_t *d_indices, double *d_data, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  Workspace w;
  ExtractArgs a{};
  int rc = fused_args(op, n_total, d_spins, row_begin, num_rows, d_workspace, workspace_bytes, w, a);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(d_indptr != nullptr, "d_indptr is NULL");
  if (num_rows == 0 || n_total == 0) {
    ASP_CUDA_CHECK(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), s));
    return ASP_OK;
  }
  ASP_REQUIRE(d_psi != nullptr, "d_psi is NULL");
  const int key_bits = static_cast<int>(op->number_spins);
  int bits = 0;
  while ((1ull << bits) < w.num_buckets) ++bits;
  a.ix.shift = key_bits > bits ? key_bits - bits : 0;
  a.psi = d_psi;
  a.indptr = d_indptr;
  a.indices = d_indices;
  a.data = d_data;
  const size_t smem = fused_smem(op);
  ASP_CUDA_CHECK(cudaFuncSetAttribute(extract_sorted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  extract_sorted_kernel<true><<<grid_for(w.tiles, 8), kTileRows, smem, s>>>(a);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

}  // extern "C"
