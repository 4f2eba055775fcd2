// HOST-buffer convenience entry points: the end-to-end path a binding without device
// memory management uses (H2D of basis + amplitudes, extraction, D2H of the CSR).
#include "operator.cuh"

struct asp_host_job {
  const asp_operator *op = nullptr;
  uint64_t n = 0, row_begin = 0, num_rows = 0, nnz = 0;
  uint64_t *d_spins = nullptr;
  double *d_psi = nullptr;
  void *d_workspace = nullptr;
  size_t workspace_bytes = 0;
  cudaStream_t stream = nullptr;
};

static void free_job(asp_host_job *job) {
  if (!job) return;
  if (job->d_spins) cudaFree(job->d_spins);
  if (job->d_psi) cudaFree(job->d_psi);
  if (job->d_workspace) cudaFree(job->d_workspace);
  if (job->stream) cudaStreamDestroy(job->stream);
  delete job;
}

extern "C" {

int asp_extract_host_begin(asp_operator const *op, uint64_t n, uint64_t const *h_spins, double const *h_psi,
                           uint64_t row_begin, uint64_t num_rows, uint64_t *h_nnz, asp_host_job **out) {
  ASP_REQUIRE(op && h_nnz && out, "NULL argument");
  ASP_REQUIRE(n == 0 || (h_spins && h_psi), "NULL input buffer");
  auto *job = new asp_host_job();
  job->op = op;
  job->n = n;
  job->row_begin = row_begin;
  job->num_rows = num_rows;
  ASP_REQUIRE(row_begin + num_rows <= n, "row block exceeds the basis");
  struct Guard {
    asp_host_job *j;
    ~Guard() { free_job(j); }
  } guard{job};
  ASP_CUDA_CHECK(cudaStreamCreateWithFlags(&job->stream, cudaStreamNonBlocking));
  job->workspace_bytes = asp_extract_workspace_bytes(op, n, num_rows);
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&job->d_spins), (n + 1) * sizeof(uint64_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&job->d_psi), (n + 1) * sizeof(double)));
  ASP_CUDA_CHECK(cudaMalloc(&job->d_workspace, job->workspace_bytes));
  ASP_CUDA_CHECK(cudaMemcpyAsync(job->d_spins, h_spins, n * sizeof(uint64_t), cudaMemcpyHostToDevice, job->stream));
  // the amplitudes ride behind the count pass
  ASP_CUDA_CHECK(cudaMemcpyAsync(job->d_psi, h_psi, n * sizeof(double), cudaMemcpyHostToDevice, job->stream));
  int rc = asp_extract_count(op, n, job->d_spins, row_begin, num_rows, job->d_workspace, job->workspace_bytes, &job->nnz, job->stream);
  if (rc != ASP_OK) return rc;
  *h_nnz = job->nnz;
  *out = job;
  guard.j = nullptr;
  return ASP_OK;
}

int asp_extract_host_finish(asp_host_job *job, int64_t *h_indptr, int32_t *h_indices, double *h_data) {
  ASP_REQUIRE(job != nullptr, "job is NULL");
  struct Guard {
    asp_host_job *j;
    ~Guard() { free_job(j); }
  } guard{job};
  ASP_REQUIRE(h_indptr && (job->nnz == 0 || (h_indices && h_data)), "NULL output buffer");
  int64_t *d_indptr = nullptr;
  int32_t *d_indices = nullptr;
  double *d_data = nullptr;
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_indptr), (job->num_rows + 1) * sizeof(int64_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_indices), (job->nnz + 1) * sizeof(int32_t)));
  ASP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_data), (job->nnz + 1) * sizeof(double)));
  int rc = asp_extract_fill(job->op, job->n, job->d_spins, job->d_psi, job->row_begin, job->num_rows, job->d_workspace, job->workspace_bytes,
                            d_indptr, d_indices, d_data, job->stream);
  if (rc == ASP_OK) {
    cudaError_t e = cudaMemcpyAsync(h_indptr, d_indptr, (job->num_rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, job->stream);
    if (e == cudaSuccess && job->nnz)
      e = cudaMemcpyAsync(h_indices, d_indices, job->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, job->stream);
    if (e == cudaSuccess && job->nnz)
      e = cudaMemcpyAsync(h_data, d_data, job->nnz * sizeof(double), cudaMemcpyDeviceToHost, job->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(job->stream);
    if (e != cudaSuccess) {
      asp::set_error("asp_extract_host_finish: %s", cudaGetErrorString(e));
      rc = ASP_ERR_CUDA;
    }
  }
  cudaFree(d_indptr);
  cudaFree(d_indices);
  cudaFree(d_data);
  return rc;
}

}  // extern "C"
