// Internal interface of the single-pass extraction kernel (extract_fused.cu), shared with the
// host-buffer pipeline (host.cu).
#pragma once
#include "operator.cuh"

namespace asp {

constexpr int kFusedTileRows = 32;   // rows per tile; row chunks start at multiples of this
constexpr int kFusedMaxChunks = 32;

size_t fused_workspace_bytes(const asp_operator *op, uint64_t n_total, uint64_t num_rows);
int fused_check_operator(const asp_operator *op);
int fused_prepare(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, uint64_t num_rows, void *d_workspace,
                  size_t workspace_bytes, cudaStream_t s);
int fused_prepare_gather(const asp_operator *op, uint32_t world, uint32_t rank, const uint64_t *shard_begin,
                         const uint64_t *const *d_shard_spins, const double *const *d_shard_psi, const uint64_t *d_ready,
                         uint64_t epoch, uint64_t *d_spins, double *d_psi, uint64_t num_rows, void *d_workspace,
                         size_t workspace_bytes, cudaStream_t s, bool tma);
int fused_launch(const asp_operator *op, uint64_t n_total, const uint64_t *d_spins, const double *d_psi, uint64_t row_begin,
                 uint64_t num_rows, int chunk, uint64_t chunk_begin, uint64_t chunk_rows, void *d_workspace, uint64_t capacity,
                 int64_t *d_indptr, int32_t *d_indices, double *d_data, unsigned long long *nnz_mirror, cudaStream_t s);
void fused_mark_indexed(const void *workspace);     // asp_gather_index left a fresh index in this workspace
bool fused_consume_indexed(const void *workspace);  // true once per index: an extraction uses it up
const unsigned long long *fused_total(const asp_operator *op, uint64_t n_total, uint64_t num_rows, void *d_workspace, int chunk);

}  // namespace asp
