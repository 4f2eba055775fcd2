// HOST-buffer entry points: the end-to-end path a binding without device memory management
// uses (H2D of basis + amplitudes, single-pass extraction, D2H of the CSR).
//
//   asp_extract_host             one call, caller-sized outputs (the reference's C contract,
//                                cbits/build_matrix.c:22-28).  The row block is cut into chunks:
//                                while chunk c+1 is extracted, chunk c's rows are already on
//                                their way back over PCIe (separate copy stream), so the call
//                                costs about H2D + one chunk + D2H instead of the sum.
//   asp_extract_host_begin/finish  two calls with exact-size outputs (begin returns nnz).
//
// Device buffers, streams and events live in a process-wide arena that only grows, so repeated
// calls pay no cudaMalloc/cudaFree (each costs a device synchronisation).
#include <mutex>
#include <string>
#include <thread>

#include "fused.cuh"

namespace {

using asp::kFusedMaxChunks;
using asp::kFusedTileRows;

struct Arena {
  std::mutex mu;
  bool busy = false;
  int device = -1;
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_spins = nullptr, ev_psi = nullptr, ev_chunk[kFusedMaxChunks] = {};
  unsigned long long *h_totals = nullptr;  // pinned + mapped: running totals written by the kernels
  struct Buf {
    void *p = nullptr;
    size_t cap = 0;
  } spins, psi, workspace, indptr, indptr32, indices, data;

  int reserve(Buf &b, size_t bytes, bool headroom = true) {
    if (bytes <= b.cap) return ASP_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = headroom ? bytes + bytes / 8 + 256 : bytes + 256;  // nnz drifts a little between calls
    ASP_CUDA_CHECK(cudaMalloc(&b.p, want));
    b.cap = want;
    return ASP_OK;
  }
  int init(int dev) {
    if (device == dev && compute) return ASP_OK;
    release();
    device = dev;
    ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
    ASP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_spins, cudaEventDisableTiming));
    ASP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_psi, cudaEventDisableTiming));
    for (auto &e : ev_chunk) ASP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ASP_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&h_totals), kFusedMaxChunks * sizeof(unsigned long long), cudaHostAllocMapped));
    return ASP_OK;
  }
  void release() {
    for (Buf *b : {&spins, &psi, &workspace, &indptr, &indptr32, &indices, &data}) {
      if (b->p) cudaFree(b->p);
      b->p = nullptr;
      b->cap = 0;
    }
    for (cudaStream_t *s : {&compute, &copy_in, &copy_out}) {
      if (*s) cudaStreamDestroy(*s);
      *s = nullptr;
    }
    for (cudaEvent_t *e : {&ev_spins, &ev_psi}) {
      if (*e) cudaEventDestroy(*e);
      *e = nullptr;
    }
    for (auto &e : ev_chunk) {
      if (e) cudaEventDestroy(e);
      e = nullptr;
    }
    if (h_totals) cudaFreeHost(h_totals);
    h_totals = nullptr;
    device = -1;
  }
};

// Two arenas: two host extractions can be in flight, so that the upload of call k+1 rides under the download of
// call k on the full-duplex host link (asp_extract_host_i32_submit / asp_extract_host_join).
constexpr int kArenas = 2;
Arena g_arenas[kArenas];
std::mutex g_arenas_mu;

int acquire(Arena *&out) {
  std::lock_guard<std::mutex> pick(g_arenas_mu);
  int device = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&device));
  Arena *chosen = nullptr;
  for (Arena &A : g_arenas)  // prefer an arena that already holds buffers for this device
    if (!A.busy && A.device == device) {
      chosen = &A;
      break;
    }
  if (!chosen)
    for (Arena &A : g_arenas)
      if (!A.busy) {
        chosen = &A;
        break;
      }
  ASP_REQUIRE(chosen != nullptr, "two host extractions are already in flight (at most two jobs per process)");
  std::lock_guard<std::mutex> lock(chosen->mu);
  int rc = chosen->init(device);
  if (rc != ASP_OK) return rc;
  chosen->busy = true;
  out = chosen;
  return ASP_OK;
}

void release_busy(Arena &A) {
  std::lock_guard<std::mutex> lock(A.mu);
  A.busy = false;
}

#define HOST_CUDA(expr)                                                      \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      ::asp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                       cudaGetErrorString(_e));                              \
      rc = ASP_ERR_CUDA;                                                     \
      goto out;                                                              \
    }                                                                        \
  } while (0)

}  // namespace

// The chunked extraction + drain shared by the host entry points: chunk c+1 is extracted while chunk c's
// rows travel back over PCIe.  The workspace is already zeroed and indexed; d_indptr/d_indices/d_data are
// the arena's output buffers.
// int64 -> int32 row starts (the index type scipy itself picks for nnz < 2^31, and what the reference's model dump
// stores: common.py:762-763): halves the indptr bytes that cross PCIe.
__global__ void narrow_indptr_kernel(const int64_t *__restrict__ in, int32_t *__restrict__ out, uint64_t count) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < count) out[i] = static_cast<int32_t>(in[i]);
}

static int extract_chunks_to_host(Arena &A, asp_operator const *op, uint64_t n, const uint64_t *d_spins, const double *d_psi,
                                  uint64_t row_begin, uint64_t num_rows, uint64_t chunk_rows, int chunks, void *d_workspace,
                                  uint64_t dev_capacity, uint64_t capacity, int64_t *d_indptr, int32_t *d_indices, double *d_data,
                                  unsigned long long *d_totals, void *h_indptr_any, bool narrow, int32_t *h_indices, double *h_data,
                                  uint64_t *h_nnz) {
  int rc = ASP_OK;
  int64_t *const h_indptr = narrow ? nullptr : static_cast<int64_t *>(h_indptr_any);
  int32_t *const h_indptr32 = narrow ? static_cast<int32_t *>(h_indptr_any) : nullptr;
  int32_t *d_indptr32 = nullptr;
  if (narrow) {
    rc = A.reserve(A.indptr32, (num_rows + 1) * sizeof(int32_t));
    if (rc != ASP_OK) return rc;
    d_indptr32 = static_cast<int32_t *>(A.indptr32.p);
  }
  {
    for (int c = 0; c < chunks; ++c) {
      const uint64_t begin = c * chunk_rows, rows = std::min(chunk_rows, num_rows - begin);
      rc = asp::fused_launch(op, n, d_spins, d_psi, row_begin, num_rows, c, begin, rows, d_workspace, dev_capacity, d_indptr,
                             d_indices, d_data, d_totals + c, A.compute);
      if (rc != ASP_OK) goto out;
      if (narrow) {
        narrow_indptr_kernel<<<static_cast<unsigned>((rows + 1 + 255) / 256), 256, 0, A.compute>>>(d_indptr + begin, d_indptr32 + begin, rows + 1);
        asp::g_launches.fetch_add(1, std::memory_order_relaxed);
      }
      HOST_CUDA(cudaEventRecord(A.ev_chunk[c], A.compute));
    }
    // drain: as soon as a chunk is done its rows go back while the next chunk is extracted
    unsigned long long done = 0;
    bool overflow = false;
    for (int c = 0; c < chunks; ++c) {
      HOST_CUDA(cudaEventSynchronize(A.ev_chunk[c]));
      const unsigned long long total = *static_cast<volatile unsigned long long *>(A.h_totals + c);
      const uint64_t begin = c * chunk_rows, rows = std::min(chunk_rows, num_rows - begin);
      const uint64_t extra = (c == chunks - 1) ? 1 : 0;  // the closing indptr entry
      if (narrow)
        HOST_CUDA(cudaMemcpyAsync(h_indptr32 + begin, d_indptr32 + begin, (rows + extra) * sizeof(int32_t), cudaMemcpyDeviceToHost, A.copy_out));
      else
        HOST_CUDA(cudaMemcpyAsync(h_indptr + begin, d_indptr + begin, (rows + extra) * sizeof(int64_t), cudaMemcpyDeviceToHost, A.copy_out));
      if (total > capacity) overflow = true;
      if (!overflow && total > done) {
        HOST_CUDA(cudaMemcpyAsync(h_indices + done, d_indices + done, (total - done) * sizeof(int32_t), cudaMemcpyDeviceToHost, A.copy_out));
        HOST_CUDA(cudaMemcpyAsync(h_data + done, d_data + done, (total - done) * sizeof(double), cudaMemcpyDeviceToHost, A.copy_out));
      }
      done = total;
    }
    HOST_CUDA(cudaStreamSynchronize(A.copy_out));
    *h_nnz = done;
    if (narrow && done > 0x7FFFFFFFull) {
      asp::set_error("%llu couplings do not fit int32 row starts: use the int64 entry point", done);
      rc = ASP_ERR_UNSUPPORTED;
      goto out;
    }
    if (overflow) {
      asp::set_error("output capacity too small: %llu couplings, room for %llu (h_indptr is complete; call again with the larger capacity)",
                     done, static_cast<unsigned long long>(capacity));
      rc = ASP_ERR_WORKSPACE;
    }
  }
out:
  return rc;
}

struct asp_host_job {
  uint64_t num_rows = 0, nnz = 0;
  Arena *arena = nullptr;  // begin / finish: the arena the job holds
  std::thread worker;      // submit / join: the thread that runs the blocking call
  int rc = ASP_OK;
  std::string error;
};

extern "C" {

void asp_host_release(void) {
  std::lock_guard<std::mutex> pick(g_arenas_mu);
  for (Arena &A : g_arenas) {
    std::lock_guard<std::mutex> lock(A.mu);
    if (!A.busy) A.release();
  }
}

static int extract_host_impl(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi, uint64_t row_begin,
                             uint64_t num_rows, uint64_t capacity, void *h_indptr, bool narrow, int32_t *h_indices, double *h_data,
                             uint64_t *h_nnz) {
  int rc = asp::fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(h_nnz && h_indptr, "NULL argument");
  ASP_REQUIRE(n == 0 || (h_spins && h_psi), "NULL input buffer");
  ASP_REQUIRE(capacity == 0 || (h_indices && h_data), "NULL output buffer");
  ASP_REQUIRE(n < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n, "row block exceeds the basis");
  *h_nnz = 0;
  if (num_rows == 0 || n == 0) {
    if (narrow)
      static_cast<int32_t *>(h_indptr)[0] = 0;
    else
      static_cast<int64_t *>(h_indptr)[0] = 0;
    return ASP_OK;
  }
  Arena *arena = nullptr;
  rc = acquire(arena);
  if (rc != ASP_OK) return rc;
  Arena &A = *arena;
  {
    // row chunks: multiples of the tile, at most kFusedMaxChunks, about 2^20 rows each
    uint64_t chunk_rows = (num_rows + 7) / 8;
    if (chunk_rows < (1u << 17)) chunk_rows = 1u << 17;
    chunk_rows = (chunk_rows + kFusedTileRows - 1) / kFusedTileRows * kFusedTileRows;
    const int chunks = static_cast<int>((num_rows + chunk_rows - 1) / chunk_rows);
    const size_t ws_bytes = asp::fused_workspace_bytes(op, n, num_rows);
    const uint64_t worst = num_rows * op->max_candidates();
    const uint64_t dev_capacity = std::min(capacity, worst);
    rc = A.reserve(A.spins, (n + 1) * sizeof(uint64_t));
    if (rc == ASP_OK) rc = A.reserve(A.psi, (n + 1) * sizeof(double));
    if (rc == ASP_OK) rc = A.reserve(A.workspace, ws_bytes);
    if (rc == ASP_OK) rc = A.reserve(A.indptr, (num_rows + 1) * sizeof(int64_t));
    if (rc == ASP_OK) rc = A.reserve(A.indices, (dev_capacity + 1) * sizeof(int32_t), false);
    if (rc == ASP_OK) rc = A.reserve(A.data, (dev_capacity + 1) * sizeof(double), false);
    if (rc != ASP_OK) goto out;
    auto *d_spins = static_cast<uint64_t *>(A.spins.p);
    auto *d_psi = static_cast<double *>(A.psi.p);
    auto *d_indptr = static_cast<int64_t *>(A.indptr.p);
    auto *d_indices = static_cast<int32_t *>(A.indices.p);
    auto *d_data = static_cast<double *>(A.data.p);
    unsigned long long *d_totals = nullptr;
    HOST_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_totals), A.h_totals, 0));

    HOST_CUDA(cudaMemcpyAsync(d_spins, h_spins, n * sizeof(uint64_t), cudaMemcpyHostToDevice, A.copy_in));
    HOST_CUDA(cudaEventRecord(A.ev_spins, A.copy_in));
    HOST_CUDA(cudaMemcpyAsync(d_psi, h_psi, n * sizeof(double), cudaMemcpyHostToDevice, A.copy_in));
    HOST_CUDA(cudaEventRecord(A.ev_psi, A.copy_in));
    // the index only needs the basis words: it is built while the amplitudes are still in flight
    HOST_CUDA(cudaStreamWaitEvent(A.compute, A.ev_spins, 0));
    rc = asp::fused_prepare(op, n, d_spins, num_rows, A.workspace.p, A.workspace.cap, A.compute);
    if (rc != ASP_OK) goto out;
    HOST_CUDA(cudaStreamWaitEvent(A.compute, A.ev_psi, 0));
    rc = extract_chunks_to_host(A, op, n, d_spins, d_psi, row_begin, num_rows, chunk_rows, chunks, A.workspace.p, dev_capacity, capacity,
                                d_indptr, d_indices, d_data, d_totals, h_indptr, narrow, h_indices, h_data, h_nnz);
  }
out:
  if (rc == ASP_ERR_CUDA) cudaDeviceSynchronize();  // leave no work in flight behind a failed call
  release_busy(A);
  return rc;
}

int asp_extract_host(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi, uint64_t row_begin,
                     uint64_t num_rows, uint64_t capacity, int64_t *h_indptr, int32_t *h_indices, double *h_data,
                     uint64_t *h_nnz) {
  return extract_host_impl(op, n, h_spins, h_psi, row_begin, num_rows, capacity, h_indptr, false, h_indices, h_data, h_nnz);
}

int asp_extract_host_i32(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi, uint64_t row_begin,
                         uint64_t num_rows, uint64_t capacity, int32_t *h_indptr, int32_t *h_indices, double *h_data,
                         uint64_t *h_nnz) {
  return extract_host_impl(op, n, h_spins, h_psi, row_begin, num_rows, capacity, h_indptr, true, h_indices, h_data, h_nnz);
}

static int extract_indexed_to_host_impl(asp_operator const *op, uint64_t n, uint64_t const *d_spins, double const *d_psi, uint64_t row_begin,
                                        uint64_t num_rows, void *d_workspace, size_t workspace_bytes, uint64_t capacity, void *h_indptr,
                                        bool narrow, int32_t *h_indices, double *h_data, uint64_t *h_nnz, void *stream) {
  int rc = asp::fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(h_nnz && h_indptr, "NULL argument");
  ASP_REQUIRE(capacity == 0 || (h_indices && h_data), "NULL output buffer");
  ASP_REQUIRE(n < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n, "row block exceeds the basis");
  *h_nnz = 0;
  if (num_rows == 0 || n == 0) {
    if (narrow)
      static_cast<int32_t *>(h_indptr)[0] = 0;
    else
      static_cast<int64_t *>(h_indptr)[0] = 0;
    return ASP_OK;
  }
  ASP_REQUIRE(d_spins && d_psi && d_workspace, "NULL device buffer");
  ASP_REQUIRE(workspace_bytes >= asp::fused_workspace_bytes(op, n, num_rows), "workspace too small");
  ASP_REQUIRE(asp::fused_consume_indexed(d_workspace),
              "the workspace holds no fresh index: call asp_gather_index before every indexed extraction (the index is single-use)");
  Arena *arena = nullptr;
  rc = acquire(arena);
  if (rc != ASP_OK) return rc;
  Arena &A = *arena;
  {
    uint64_t chunk_rows = (num_rows + 7) / 8;
    if (chunk_rows < (1u << 17)) chunk_rows = 1u << 17;
    chunk_rows = (chunk_rows + kFusedTileRows - 1) / kFusedTileRows * kFusedTileRows;
    const int chunks = static_cast<int>((num_rows + chunk_rows - 1) / chunk_rows);
    const uint64_t worst = num_rows * op->max_candidates();
    const uint64_t dev_capacity = std::min(capacity, worst);
    rc = A.reserve(A.indptr, (num_rows + 1) * sizeof(int64_t));
    if (rc == ASP_OK) rc = A.reserve(A.indices, (dev_capacity + 1) * sizeof(int32_t), false);
    if (rc == ASP_OK) rc = A.reserve(A.data, (dev_capacity + 1) * sizeof(double), false);
    if (rc != ASP_OK) goto out;
    unsigned long long *d_totals = nullptr;
    HOST_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_totals), A.h_totals, 0));
    // the caller's stream gathered and indexed the basis: the arena's compute stream continues behind it
    HOST_CUDA(cudaEventRecord(A.ev_spins, static_cast<cudaStream_t>(stream)));
    HOST_CUDA(cudaStreamWaitEvent(A.compute, A.ev_spins, 0));
    rc = extract_chunks_to_host(A, op, n, d_spins, d_psi, row_begin, num_rows, chunk_rows, chunks, d_workspace, dev_capacity, capacity,
                                static_cast<int64_t *>(A.indptr.p), static_cast<int32_t *>(A.indices.p), static_cast<double *>(A.data.p),
                                d_totals, h_indptr, narrow, h_indices, h_data, h_nnz);
    // the caller's stream must not reuse the workspace or the basis before the chunks are done
    if (rc == ASP_OK) {
      HOST_CUDA(cudaEventRecord(A.ev_psi, A.compute));
      HOST_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), A.ev_psi, 0));
    }
  }
out:
  if (rc == ASP_ERR_CUDA) cudaDeviceSynchronize();
  release_busy(A);
  return rc;
}

int asp_extract_indexed_to_host(asp_operator const *op, uint64_t n, uint64_t const *d_spins, double const *d_psi, uint64_t row_begin,
                                uint64_t num_rows, void *d_workspace, size_t workspace_bytes, uint64_t capacity, int64_t *h_indptr,
                                int32_t *h_indices, double *h_data, uint64_t *h_nnz, void *stream) {
  return extract_indexed_to_host_impl(op, n, d_spins, d_psi, row_begin, num_rows, d_workspace, workspace_bytes, capacity, h_indptr, false,
                                      h_indices, h_data, h_nnz, stream);
}

int asp_extract_indexed_to_host_i32(asp_operator const *op, uint64_t n, uint64_t const *d_spins, double const *d_psi, uint64_t row_begin,
                                    uint64_t num_rows, void *d_workspace, size_t workspace_bytes, uint64_t capacity, int32_t *h_indptr,
                                    int32_t *h_indices, double *h_data, uint64_t *h_nnz, void *stream) {
  return extract_indexed_to_host_impl(op, n, d_spins, d_psi, row_begin, num_rows, d_workspace, workspace_bytes, capacity, h_indptr, true,
                                      h_indices, h_data, h_nnz, stream);
}

int asp_extract_host_begin(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi,
                           uint64_t row_begin, uint64_t num_rows, uint64_t *h_nnz, asp_host_job **out_job) {
  int rc = asp::fused_check_operator(op);
  if (rc != ASP_OK) return rc;
  ASP_REQUIRE(h_nnz && out_job, "NULL argument");
  ASP_REQUIRE(n == 0 || (h_spins && h_psi), "NULL input buffer");
  ASP_REQUIRE(n < (1ull << 31), "int32 column indices need n_total < 2^31 (scipy picks int32 the same way)");
  ASP_REQUIRE(row_begin + num_rows <= n, "row block exceeds the basis");
  Arena *arena = nullptr;
  rc = acquire(arena);
  if (rc != ASP_OK) return rc;
  Arena &A = *arena;
  auto *job = new asp_host_job();
  job->num_rows = num_rows;
  job->arena = arena;
  {
    const size_t ws_bytes = asp::fused_workspace_bytes(op, n, num_rows);
    const uint64_t worst = num_rows * op->max_candidates();
    rc = A.reserve(A.spins, (n + 1) * sizeof(uint64_t));
    if (rc == ASP_OK) rc = A.reserve(A.psi, (n + 1) * sizeof(double));
    if (rc == ASP_OK) rc = A.reserve(A.workspace, ws_bytes);
    if (rc == ASP_OK) rc = A.reserve(A.indptr, (num_rows + 1) * sizeof(int64_t));
    if (rc != ASP_OK) goto out;
    HOST_CUDA(cudaMemcpyAsync(A.spins.p, h_spins, n * sizeof(uint64_t), cudaMemcpyHostToDevice, A.compute));
    HOST_CUDA(cudaMemcpyAsync(A.psi.p, h_psi, n * sizeof(double), cudaMemcpyHostToDevice, A.compute));
    // outputs: what the arena already holds, at least 8 couplings per row; a short guess is
    // answered with the exact count and repeated once
    uint64_t capacity = std::max<uint64_t>(std::min<uint64_t>(worst, 8 * num_rows), std::min(A.indices.cap / sizeof(int32_t), A.data.cap / sizeof(double)));
    capacity = std::min(capacity, worst);
    for (int attempt = 0; attempt < 2; ++attempt) {
      rc = A.reserve(A.indices, (capacity + 1) * sizeof(int32_t));
      if (rc == ASP_OK) rc = A.reserve(A.data, (capacity + 1) * sizeof(double));
      if (rc != ASP_OK) goto out;
      uint64_t nnz = 0;
      rc = asp_extract_csr(op, n, static_cast<uint64_t *>(A.spins.p), static_cast<double *>(A.psi.p), row_begin, num_rows, A.workspace.p,
                           A.workspace.cap, capacity, static_cast<int64_t *>(A.indptr.p), static_cast<int32_t *>(A.indices.p),
                           static_cast<double *>(A.data.p), &nnz, A.compute);
      job->nnz = nnz;
      if (rc == ASP_ERR_WORKSPACE && nnz > capacity && attempt == 0) {
        capacity = nnz;
        continue;
      }
      break;
    }
    if (rc != ASP_OK) goto out;
    *h_nnz = job->nnz;
    *out_job = job;
    return ASP_OK;  // the arena stays busy until asp_extract_host_finish
  }
out:
  delete job;
  release_busy(A);
  return rc;
}

int asp_extract_host_finish(asp_host_job *job, int64_t *h_indptr, int32_t *h_indices, double *h_data) {
  ASP_REQUIRE(job != nullptr, "job is NULL");
  ASP_REQUIRE(job->arena != nullptr, "the job was not started by asp_extract_host_begin");
  Arena &A = *job->arena;
  int rc = ASP_OK;
  if (!h_indptr || (job->nnz != 0 && (!h_indices || !h_data))) {
    asp::set_error("asp_extract_host_finish: NULL output buffer");
    rc = ASP_ERR_ARG;
    goto out;
  }
  if (job->num_rows == 0) {
    h_indptr[0] = 0;
    goto out;
  }
  HOST_CUDA(cudaMemcpyAsync(h_indptr, A.indptr.p, (job->num_rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, A.compute));
  if (job->nnz) {
    HOST_CUDA(cudaMemcpyAsync(h_indices, A.indices.p, job->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, A.copy_out));
    HOST_CUDA(cudaMemcpyAsync(h_data, A.data.p, job->nnz * sizeof(double), cudaMemcpyDeviceToHost, A.compute));
  }
  HOST_CUDA(cudaStreamSynchronize(A.compute));
  HOST_CUDA(cudaStreamSynchronize(A.copy_out));
out:
  delete job;
  release_busy(A);
  return rc;
}

// Asynchronous form of asp_extract_host_i32: the blocking call runs on a worker thread with its own arena (device
// buffers, streams), so a second submit can upload its basis while this job's CSR is still travelling back.
int asp_extract_host_i32_submit(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi, uint64_t row_begin,
                                uint64_t num_rows, uint64_t capacity, int32_t *h_indptr, int32_t *h_indices, double *h_data,
                                asp_host_job **out_job) {
  ASP_REQUIRE(out_job != nullptr, "out_job is NULL");
  int device = 0;
  ASP_CUDA_CHECK(cudaGetDevice(&device));
  auto *job = new asp_host_job();
  job->num_rows = num_rows;
  job->worker = std::thread([=]() {
    if (cudaSetDevice(device) != cudaSuccess) {
      job->rc = ASP_ERR_CUDA;
      job->error = "cudaSetDevice failed on the worker thread";
      return;
    }
    uint64_t nnz = 0;
    job->rc = extract_host_impl(op, n, h_spins, h_psi, row_begin, num_rows, capacity, h_indptr, true, h_indices, h_data, &nnz);
    job->nnz = nnz;
    if (job->rc != ASP_OK) job->error = asp_last_error();  // the error text is per thread: carry it to the joiner
  });
  *out_job = job;
  return ASP_OK;
}

int asp_extract_host_join(asp_host_job *job, uint64_t *h_nnz) {
  ASP_REQUIRE(job != nullptr, "job is NULL");
  ASP_REQUIRE(job->worker.joinable(), "the job was not started by asp_extract_host_i32_submit");
  job->worker.join();
  const int rc = job->rc;
  if (h_nnz) *h_nnz = job->nnz;
  if (rc != ASP_OK) asp::set_error("%s", job->error.c_str());
  delete job;
  return rc;
}

}  // extern "C"
