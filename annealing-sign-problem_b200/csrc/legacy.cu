// Drop-in kernels for the reference's C ABI (cbits/build_matrix.h:7-14): build_matrix on an
// explicit candidate list and extract_signs.  Same outputs as the C file: COO triplets in
// generation order, field[] accumulated in candidate order with the C association
// ((counts*coeff)*|psi_i|)*psi'_j  (cbits/build_matrix.c:38-49) -- explicit _rn intrinsics
// keep nvcc from contracting the accumulation into an FMA, so the result is bitwise the C one.
#include <algorithm>
#include <vector>

#include <type_traits>

#include "common.cuh"

namespace asp {

struct Key64 {
  uint64_t w;
};
__device__ __forceinline__ int key_cmp(const uint64_t &a, const uint64_t &b) { return a < b ? -1 : (a > b ? 1 : 0); }
// lexicographic from words[0] upward (cbits/build_matrix.c:7-20)
__device__ __forceinline__ int key_cmp(const asp_bits512 &a, const asp_bits512 &b) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (a.words[k] < b.words[k]) return -1;
    if (a.words[k] > b.words[k]) return 1;
  }
  return 0;
}

template <typename Key>
__device__ __forceinline__ int64_t plain_find(const Key *__restrict__ hay, uint64_t n, const Key &needle) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (key_cmp(hay[mid], needle) < 0)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < n && key_cmp(hay[lo], needle) == 0) ? static_cast<int64_t>(lo) : -1;
}

// First-position table over the leading bits of sorted 64-bit keys (device path only): a search is one table read and a
// bisection inside a bucket of ~2 keys instead of ~log2(n) dependent reads.  table == nullptr: plain bisection.
struct KeyTable {
  const uint2 *table;  // [num_buckets] {first, last + 1} of the keys with these leading bits; {0, 0} = none
  int shift;
  uint64_t num_buckets;
};

__global__ void __launch_bounds__(256) legacy_table_kernel(const uint64_t *__restrict__ spins, uint32_t n, int shift, uint64_t num_buckets,
                                                           uint2 *__restrict__ table) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const uint64_t b = spins[i] >> shift;
  if (b >= num_buckets) return;
  if (i == 0 || (spins[i - 1] >> shift) != b) table[b].x = i;
  if (i == n - 1 || (spins[i + 1] >> shift) != b) table[b].y = i + 1;
}

template <typename Key>
__device__ __forceinline__ int64_t find_key(const Key *__restrict__ hay, uint64_t n, const Key &needle, const KeyTable &t) {
  if constexpr (std::is_same<Key, uint64_t>::value) {
    if (t.table) {
      const uint64_t b = needle >> t.shift;
      if (b >= t.num_buckets) return -1;
      const uint2 se = __ldg(&t.table[b]);
      uint64_t lo = se.x, hi = se.y;  // the keys of the bucket are hay[lo .. hi)
      const uint64_t end = se.y;
      while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (hay[mid] < needle)
          lo = mid + 1;
        else
          hi = mid;
      }
      return (lo < end && hay[lo] == needle) ? static_cast<int64_t>(lo) : -1;
    }
  }
  return plain_find(hay, n, needle);
}

// One lane per row.  kFill == false: row hit counts + field; kFill == true: COO output.
template <typename Key, bool kFill>
__global__ void __launch_bounds__(128) legacy_build_kernel(
    uint64_t n_total, const Key *__restrict__ spins, uint64_t row_begin, uint64_t num_rows,
    const int64_t *__restrict__ counts, const double *__restrict__ psi, const Key *__restrict__ other_spins,
    const double *__restrict__ other_coeffs, const int64_t *__restrict__ offsets,
    const double *__restrict__ other_psi, int64_t *__restrict__ row_nnz_or_offsets,
    uint32_t *__restrict__ row_indices, uint32_t *__restrict__ col_indices, double *__restrict__ elements,
    double *__restrict__ field, const KeyTable key_table) {
  const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= num_rows) return;
  const uint64_t row = row_begin + r;
  const double mult = static_cast<double>(counts ? counts[r] : 1);
  const double a_i = fabs(psi[row]);
  int64_t out = kFill ? row_nnz_or_offsets[r] : 0;
  int64_t cnt = 0;
  double f = 0.0;
  for (int64_t k = offsets[r]; k < offsets[r + 1]; ++k) {
    const Key needle = other_spins[k];
    const int64_t pos = find_key(spins, n_total, needle, key_table);
    const double w = __dmul_rn(__dmul_rn(mult, other_coeffs[k]), a_i);
    if (pos >= 0) {
      if (kFill) {
        row_indices[out] = static_cast<uint32_t>(row);
        col_indices[out] = static_cast<uint32_t>(pos);
        elements[out] = __dmul_rn(w, fabs(other_psi ? other_psi[k] : psi[pos]));
        ++out;
      }
      ++cnt;
    } else if (!kFill && other_psi) {
      f = __dadd_rn(f, __dmul_rn(w, other_psi[k]));
    }
  }
  if (!kFill) {
    row_nnz_or_offsets[r] = cnt;
    if (field) field[r] = f;
  }
}

__global__ void __launch_bounds__(256) extract_signs_kernel(uint64_t n, const double *__restrict__ psi, uint32_t *__restrict__ halves, uint64_t num_halves) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool up = i < n && psi[i] > 0.0;  // strict: 0, -0 and NaN give bit 0 (build_matrix.c:72)
  const uint32_t ballot = __ballot_sync(0xffffffffu, up);
  const uint64_t h = i >> 5;
  if ((threadIdx.x & 31) == 0 && h < num_halves) halves[h] = ballot;
}

template <typename Key>
static int legacy_build_dev(uint64_t n_total, const Key *d_spins, uint64_t row_begin, uint64_t num_rows,
                            const int64_t *d_counts, const double *d_psi, const Key *d_other_spins,
                            const double *d_other_coeffs, const int64_t *d_offsets, const double *d_other_psi,
                            int64_t *d_row_offsets, uint32_t *d_row_indices, uint32_t *d_col_indices,
                            double *d_elements, double *d_field, uint64_t capacity, uint64_t *h_nnz, cudaStream_t s) {
  ASP_REQUIRE(h_nnz != nullptr && d_row_offsets != nullptr, "NULL output");
  ASP_CUDA_CHECK(asp::keep_pool_memory());
  if (num_rows == 0) {
    *h_nnz = 0;
    ASP_CUDA_CHECK(cudaMemsetAsync(d_row_offsets, 0, sizeof(int64_t), s));
    return ASP_OK;
  }
  const unsigned blocks = static_cast<unsigned>((num_rows + 127) / 128);
  KeyTable key_table{nullptr, 0, 0};
  uint2 *d_table = nullptr;
  if constexpr (std::is_same<Key, uint64_t>::value) {
    if (n_total >= 4096 && n_total < (1ull << 32)) {  // worth a table: ~2 keys per bucket
      uint64_t last_key = 0;  // the keys are ascending: the last one has the most bits
      ASP_CUDA_CHECK(cudaMemcpyAsync(&last_key, d_spins + (n_total - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
      ASP_CUDA_CHECK(cudaStreamSynchronize(s));
      int key_bits = 1;
      while (key_bits < 64 && (last_key >> key_bits) != 0) ++key_bits;
      int lg = 0;
      while ((1ull << lg) < n_total) ++lg;
      const int bits = std::min(std::min(std::max(lg - 1, 4), 26), key_bits);
      key_table.shift = key_bits - bits;
      key_table.num_buckets = 1ull << bits;
      ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d_table), key_table.num_buckets * sizeof(uint2), s));
      ASP_CUDA_CHECK(cudaMemsetAsync(d_table, 0, key_table.num_buckets * sizeof(uint2), s));
      legacy_table_kernel<<<static_cast<unsigned>((n_total + 255) / 256), 256, 0, s>>>(d_spins, static_cast<uint32_t>(n_total), key_table.shift,
                                                                                     key_table.num_buckets, d_table);
      ASP_LAUNCH_CHECK();
      key_table.table = d_table;
    }
  }
  struct TableGuard {  // released in stream order on every way out
    uint2 *p;
    cudaStream_t s;
    ~TableGuard() {
      if (p) cudaFreeAsync(p, s);
    }
  } table_guard{d_table, s};
  // pass 1 writes per-row counts into d_row_offsets[0..num_rows) ...
  legacy_build_kernel<Key, false><<<blocks, 128, 0, s>>>(n_total, d_spins, row_begin, num_rows, d_counts, d_psi,
                                                         d_other_spins, d_other_coeffs, d_offsets, d_other_psi,
                                                         d_row_offsets, nullptr, nullptr, nullptr, d_field, key_table);
  ASP_LAUNCH_CHECK();
  // ... which are scanned in place (the scan reads each tile before it writes it)
  void *tmp = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(&tmp, scan_tmp_bytes(num_rows), s));
  int64_t *counts_copy = nullptr;
  ASP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&counts_copy), num_rows * sizeof(int64_t), s));
  ASP_CUDA_CHECK(cudaMemcpyAsync(counts_copy, d_row_offsets, num_rows * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  int rc = scan_exclusive_i64(counts_copy, d_row_offsets, num_rows, tmp, s);
  if (rc != ASP_OK) return rc;
  int64_t total = 0;
  ASP_CUDA_CHECK(cudaMemcpyAsync(&total, d_row_offsets + num_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  ASP_CUDA_CHECK(cudaFreeAsync(counts_copy, s));
  ASP_CUDA_CHECK(cudaFreeAsync(tmp, s));
  ASP_CUDA_CHECK(cudaStreamSynchronize(s));
  *h_nnz = static_cast<uint64_t>(total);
  if (static_cast<uint64_t>(total) > capacity) {
    set_error("output capacity %llu < nnz %lld", static_cast<unsigned long long>(capacity), static_cast<long long>(total));
    return ASP_ERR_WORKSPACE;
  }
  if (total == 0 || d_col_indices == nullptr) return ASP_OK;
  legacy_build_kernel<Key, true><<<blocks, 128, 0, s>>>(n_total, d_spins, row_begin, num_rows, d_counts, d_psi,
                                                        d_other_spins, d_other_coeffs, d_offsets, d_other_psi,
                                                        d_row_offsets, d_row_indices, d_col_indices, d_elements, nullptr, key_table);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

// RAII device buffer for the host-pointer drop-ins.
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    ASP_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return ASP_OK;
  }
  int upload(const void *src, size_t bytes) {
    int rc = alloc(bytes);
    if (rc != ASP_OK) return rc;
    if (bytes) ASP_CUDA_CHECK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
    return ASP_OK;
  }
  template <typename T>
  T *as() { return static_cast<T *>(p); }
};

}  // namespace asp

using namespace asp;

extern "C" {

int asp_extract_signs_dev(uint64_t num_spins, double const *d_psi, uint64_t *d_signs, void *stream) {
  auto s = static_cast<cudaStream_t>(stream);
  const uint64_t words = (num_spins + 63) / 64;
  if (words == 0) return ASP_OK;
  ASP_REQUIRE(d_psi && d_signs, "NULL buffer");
  const uint64_t threads = words * 64;
  extract_signs_kernel<<<static_cast<unsigned>(threads / 256 + (threads % 256 != 0)), 256, 0, s>>>(
      num_spins, d_psi, reinterpret_cast<uint32_t *>(d_signs), words * 2);
  ASP_LAUNCH_CHECK();
  return ASP_OK;
}

void asp_extract_signs(uint64_t num_spins, double const *psi, uint64_t *signs) {
  const uint64_t words = (num_spins + 63) / 64;
  if (words == 0) return;
  DevBuf d_psi, d_bits;
  if (d_psi.upload(psi, num_spins * sizeof(double)) != ASP_OK) return;
  if (d_bits.alloc(words * sizeof(uint64_t)) != ASP_OK) return;
  if (asp_extract_signs_dev(num_spins, d_psi.as<double>(), d_bits.as<uint64_t>(), nullptr) != ASP_OK) return;
  if (cudaMemcpy(signs, d_bits.p, words * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess)
    set_error("asp_extract_signs: D2H copy failed");
}

int asp_build_matrix_dev(uint64_t n_total, uint64_t const *d_spins, uint64_t row_begin, uint64_t num_rows,
                         int64_t const *d_counts, double const *d_psi, uint64_t const *d_other_spins,
                         double const *d_other_coeffs, int64_t const *d_offsets, double const *d_other_psi,
                         int64_t *d_row_offsets, uint32_t *d_row_indices, uint32_t *d_col_indices,
                         double *d_elements, double *d_field, uint64_t capacity, uint64_t *h_nnz, void *stream) {
  ASP_REQUIRE(n_total < (1ull << 32), "uint32 indices need n_total < 2^32 (cbits/build_matrix.c:27)");
  return legacy_build_dev<uint64_t>(n_total, d_spins, row_begin, num_rows, d_counts, d_psi, d_other_spins,
                                    d_other_coeffs, d_offsets, d_other_psi, d_row_offsets, d_row_indices,
                                    d_col_indices, d_elements, d_field, capacity, h_nnz,
                                    static_cast<cudaStream_t>(stream));
}

uint64_t asp_build_matrix(uint64_t num_spins, asp_bits512 const spins[], int64_t const *counts, double const *psi,
                          asp_bits512 const *other_spins, double const *other_coeffs, int64_t const *other_counts,
                          double const *other_psi, uint32_t *row_indices, uint32_t *col_indices, double *elements,
                          double *field) {
  const uint64_t n = num_spins;
  if (n >= (1ull << 32)) {
    set_error("asp_build_matrix: num_spins >= 2^32");
    return UINT64_MAX;
  }
  // offsets = exclusive scan of other_counts (host; O(n) bookkeeping of the binding layer)
  std::vector<int64_t> offsets(n + 1, 0);
  for (uint64_t r = 0; r < n; ++r) offsets[r + 1] = offsets[r] + other_counts[r];
  const uint64_t T = static_cast<uint64_t>(offsets[n]);
  DevBuf d_spins, d_counts, d_psi, d_os, d_oc, d_off, d_op, d_rowoff, d_rows, d_cols, d_vals, d_field;
  int rc = ASP_OK;
  auto ok = [&](int r) { rc = r; return r == ASP_OK; };
  if (!ok(d_spins.upload(spins, n * sizeof(asp_bits512))) || !ok(d_counts.upload(counts, n * sizeof(int64_t))) ||
      !ok(d_psi.upload(psi, n * sizeof(double))) || !ok(d_os.upload(other_spins, T * sizeof(asp_bits512))) ||
      !ok(d_oc.upload(other_coeffs, T * sizeof(double))) || !ok(d_off.upload(offsets.data(), (n + 1) * sizeof(int64_t))) ||
      !ok(d_op.upload(other_psi, T * sizeof(double))) || !ok(d_rowoff.alloc((n + 1) * sizeof(int64_t))) ||
      !ok(d_rows.alloc(T * sizeof(uint32_t))) || !ok(d_cols.alloc(T * sizeof(uint32_t))) ||
      !ok(d_vals.alloc(T * sizeof(double))) || !ok(d_field.alloc(n * sizeof(double))))
    return UINT64_MAX;
  uint64_t nnz = 0;
  rc = legacy_build_dev<asp_bits512>(n, d_spins.as<asp_bits512>(), 0, n, d_counts.as<int64_t>(), d_psi.as<double>(),
                                     d_os.as<asp_bits512>(), d_oc.as<double>(), d_off.as<int64_t>(), d_op.as<double>(),
                                     d_rowoff.as<int64_t>(), d_rows.as<uint32_t>(), d_cols.as<uint32_t>(),
                                     d_vals.as<double>(), d_field.as<double>(), T, &nnz, nullptr);
  if (rc != ASP_OK) return UINT64_MAX;
  if (cudaMemcpy(field, d_field.p, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(row_indices, d_rows.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(col_indices, d_cols.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(elements, d_vals.p, nnz * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("asp_build_matrix: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return UINT64_MAX;
  }
  return nnz;
}

// The reference's own symbol names (cbits/build_matrix.h:7-14): its literal cdef binds this library unchanged.
static_assert(sizeof(ls_bits512) == sizeof(asp_bits512), "key types must match");
uint64_t build_matrix(uint64_t num_spins, ls_bits512 const spins[], int64_t const *counts, double const *psi,
                      ls_bits512 const *other_spins, double const *other_coeffs, int64_t const *other_counts,
                      double const *other_psi, uint32_t *row_indices, uint32_t *col_indices, double *elements, double *field) {
  return asp_build_matrix(num_spins, reinterpret_cast<asp_bits512 const *>(spins), counts, psi,
                          reinterpret_cast<asp_bits512 const *>(other_spins), other_coeffs, other_counts, other_psi, row_indices,
                          col_indices, elements, field);
}

void extract_signs(uint64_t num_spins, double const *psi, uint64_t *signs) { asp_extract_signs(num_spins, psi, signs); }

}  // extern "C"
