// HOST-buffer convenience entry points: the end-to-end path a binding without device
// memory management uses (H2D of basis + amplitudes, extraction, D2H of the CSR).
// Device buffers and the copy stream live in a process-wide arena that only grows, so
// repeated calls pay no cudaMalloc/cudaFree (each costs a device synchronisation).
#include <mutex>

#include "operator.cuh"

namespace {

struct Arena {
  std::mutex mu;
  bool busy = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  struct Buf {
    void *p = nullptr;
    size_t cap = 0;
  } spins, psi, workspace, indptr, indices, data;

  int reserve(Buf &b, size_t bytes) {
    if (bytes <= b.cap) return ASP_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;  // headroom: nnz drifts a little between calls
    ASP_CUDA_CHECK(cudaMalloc(&b.p, want));
    b.cap = want;
    return ASP_OK;
  }
  void release() {
    for (Buf *b : {&spins, &psi, &workspace, &indptr, &indices, &data}) {
      if (b->p) cudaFree(b->p);
      b->p = nullptr;
      b->cap = 0;
    }
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
    device = -1;
  }
};

Arena g_arena;

}  // namespace

struct asp_host_job {
  const asp_operator *op = nullptr;
  uint64_t n = 0, row_begin = 0, num_rows = 0, nnz = 0;
  size_t workspace_bytes = 0;
};

extern "C" {

void asp_host_release(void) {
  std::lock_guard<std::mutex> lock(g_arena.mu);
  if (!g_arena.busy) g_arena.release();
}

int asp_extract_host_begin(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi,
                           uint64_t row_begin, uint64_t num_rows, uint64_t *h_nnz, asp_host_job **out) {
  ASP_REQUIRE(op && h_nnz && out, "NULL argument");
  ASP_REQUIRE(n == 0 || (h_spins && h_psi), "NULL input buffer");
  ASP_REQUIRE(row_begin + num_rows <= n, "row block exceeds the basis");
  Arena &A = g_arena;
  {
    std::lock_guard<std::mutex> lock(A.mu);
    ASP_REQUIRE(!A.busy, "another host extraction is in flight (one job at a time per process)");
    int device = 0;
    ASP_CUDA_CHECK(cudaGetDevice(&device));
    if (A.device != device) {
      A.release();
      A.device = device;
    }
    if (!A.stream) ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking));
    A.busy = true;
  }
  auto *job = new asp_host_job();
  job->op = op;
  job->n = n;
  job->row_begin = row_begin;
  job->num_rows = num_rows;
  job->workspace_bytes = asp_extract_workspace_bytes(op, n, num_rows);
  auto fail = [&](int rc) {
    delete job;
    std::lock_guard<std::mutex> lock(A.mu);
    A.busy = false;
    return rc;
  };
  int rc = A.reserve(A.spins, (n + 1) * sizeof(uint64_t));
  if (rc == ASP_OK) rc = A.reserve(A.psi, (n + 1) * sizeof(double));
  if (rc == ASP_OK) rc = A.reserve(A.workspace, job->workspace_bytes);
  if (rc != ASP_OK) return fail(rc);
  cudaError_t e = cudaMemcpyAsync(A.spins.p, h_spins, n * sizeof(uint64_t), cudaMemcpyHostToDevice, A.stream);
  // the amplitudes ride behind the count pass (same stream, consumed only by the fill)
  if (e == cudaSuccess) e = cudaMemcpyAsync(A.psi.p, h_psi, n * sizeof(double), cudaMemcpyHostToDevice, A.stream);
  if (e != cudaSuccess) {
    asp::set_error("asp_extract_host_begin: %s", cudaGetErrorString(e));
    return fail(ASP_ERR_CUDA);
  }
  rc = asp_extract_count(op, n, static_cast<uint64_t *>(A.spins.p), row_begin, num_rows, A.workspace.p, A.workspace.cap, &job->nnz, A.stream);
  if (rc != ASP_OK) return fail(rc);
  *h_nnz = job->nnz;
  *out = job;
  return ASP_OK;
}

int asp_extract_host_finish(asp_host_job *job, int64_t *h_indptr, int32_t *h_indices, double *h_data) {
  ASP_REQUIRE(job != nullptr, "job is NULL");
  Arena &A = g_arena;
  auto done = [&](int rc) {
    delete job;
    std::lock_guard<std::mutex> lock(A.mu);
    A.busy = false;
    return rc;
  };
  if (!h_indptr || (job->nnz != 0 && (!h_indices || !h_data))) {
    asp::set_error("asp_extract_host_finish: NULL output buffer");
    return done(ASP_ERR_ARG);
  }
  int rc = A.reserve(A.indptr, (job->num_rows + 1) * sizeof(int64_t));
  if (rc == ASP_OK) rc = A.reserve(A.indices, (job->nnz + 1) * sizeof(int32_t));
  if (rc == ASP_OK) rc = A.reserve(A.data, (job->nnz + 1) * sizeof(double));
  if (rc != ASP_OK) return done(rc);
  rc = asp_extract_fill(job->op, job->n, static_cast<uint64_t *>(A.spins.p), static_cast<double *>(A.psi.p), job->row_begin,
                        job->num_rows, A.workspace.p, A.workspace.cap, static_cast<int64_t *>(A.indptr.p),
                        static_cast<int32_t *>(A.indices.p), static_cast<double *>(A.data.p), A.stream);
  if (rc != ASP_OK) return done(rc);
  cudaError_t e = cudaMemcpyAsync(h_indptr, A.indptr.p, (job->num_rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, A.stream);
  if (e == cudaSuccess && job->nnz)
    e = cudaMemcpyAsync(h_indices, A.indices.p, job->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, A.stream);
  if (e == cudaSuccess && job->nnz)
    e = cudaMemcpyAsync(h_data, A.data.p, job->nnz * sizeof(double), cudaMemcpyDeviceToHost, A.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(A.stream);
  if (e != cudaSuccess) {
    asp::set_error("asp_extract_host_finish: %s", cudaGetErrorString(e));
    return done(ASP_ERR_CUDA);
  }
  return done(ASP_OK);
}

}  // extern "C"
