/* asp_b200.h -- C ABI of the B200-native Ising-extraction + annealing hot path.
 *
 * Drop-in boundary for twesterhout/annealing-sign-problem.  Every entry point takes plain
 * pointers and sizes (no torch / C++ types) so the reference's cffi layer
 * (annealing_sign_problem/build_extension.py:5-21) can bind it unchanged; INTEGRATION.md
 * shows the cdef a maintainer would add.
 *
 * Conventions (those of cbits/build_matrix.h:7-14 unless noted):
 *   - the caller owns every buffer; the callee never frees caller memory;
 *   - `*_host` / legacy entry points take HOST pointers and do their own H2D/D2H;
 *     all other entry points take DEVICE pointers valid on the current CUDA device and
 *     enqueue on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - ADDITION to the reference (which has no error channel): functions returning `int`
 *     give ASP_OK or a negative ASP_ERR_* and leave a message in asp_last_error();
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     ASP_ERR_CUDA.
 *
 * Bit layout of packed sign vectors: word i/64, bit i%64, 1 <=> s_i = +1
 * (cbits/build_matrix.c:67-76).
 */
#ifndef ASP_B200_H
#define ASP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASP_OK 0
#define ASP_ERR_CUDA (-1)     /* CUDA runtime error or no device */
#define ASP_ERR_ARG (-2)      /* invalid argument */
#define ASP_ERR_WORKSPACE (-3) /* workspace too small */
#define ASP_ERR_UNSUPPORTED (-4)

/* == struct ls_bits512, cbits/build_matrix.h:3-5 */
typedef struct asp_bits512 {
  uint64_t words[8];
} asp_bits512;

typedef struct asp_operator asp_operator; /* Hamiltonian bond list compiled to XOR moves */
typedef struct asp_sa_plan asp_sa_plan;   /* Ising model prepared for replica annealing  */
typedef struct asp_host_job asp_host_job; /* in-flight host-buffer extraction            */

int asp_version(void);
const char *asp_last_error(void);
/* Number of CUDA devices visible (0 when none / no driver). */
int asp_device_count(void);

/* ------------------------------------------------------------------------------------
 * 1. Legacy drop-ins (HOST pointers).  Replace cbits/build_matrix.h:7-14 one to one:
 *    same argument list, same outputs (COO triplets in generation order, field[]
 *    accumulated in candidate order, returns nnz).  On failure returns UINT64_MAX.
 * ---------------------------------------------------------------------------------- */
uint64_t asp_build_matrix(uint64_t num_spins, asp_bits512 const spins[], int64_t const *counts,
                          double const *psi, asp_bits512 const *other_spins,
                          double const *other_coeffs, int64_t const *other_counts,
                          double const *other_psi, uint32_t *row_indices, uint32_t *col_indices,
                          double *elements, double *field);

void asp_extract_signs(uint64_t num_spins, double const *psi, uint64_t *signs);

/* The same two under the reference's own names and key type (cbits/build_matrix.h:3-14), so that the
 * reference's literal cdef (annealing_sign_problem/build_extension.py:5-21) binds this library with
 * ffi.dlopen and no edit: build_matrix == asp_build_matrix, extract_signs == asp_extract_signs. */
typedef struct ls_bits512 {
  uint64_t words[8];
} ls_bits512;
uint64_t build_matrix(uint64_t num_spins, ls_bits512 const spins[], int64_t const *counts, double const *psi,
                      ls_bits512 const *other_spins, double const *other_coeffs, int64_t const *other_counts,
                      double const *other_psi, uint32_t *row_indices, uint32_t *col_indices, double *elements,
                      double *field);
void extract_signs(uint64_t num_spins, double const *psi, uint64_t *signs);

/* Same two on DEVICE pointers with 64-bit keys (the reference's glue only fills words[0],
 * common.py:58-68,100), for the row block [row_begin, row_begin+num_rows) of the basis.
 * d_spins, d_psi: the FULL basis [n_total].  d_counts (NULL = all 1), d_offsets, d_field:
 * per row of the block; d_offsets = exclusive scan of other_counts, [num_rows+1]
 * (asp_exclusive_scan_i64).  d_other_psi == NULL: a hit uses |d_psi[column]| and no field
 * is accumulated (the live path, common.py:71-82).  d_row_offsets[num_rows+1] receives the
 * row starts of the emitted triplets.  Synchronises; *h_nnz = emitted triplets. */
int asp_build_matrix_dev(uint64_t n_total, uint64_t const *d_spins, uint64_t row_begin,
                         uint64_t num_rows, int64_t const *d_counts, double const *d_psi,
                         uint64_t const *d_other_spins, double const *d_other_coeffs,
                         int64_t const *d_offsets, double const *d_other_psi,
                         int64_t *d_row_offsets /* [num_rows+1] out */, uint32_t *d_row_indices,
                         uint32_t *d_col_indices, double *d_elements, double *d_field,
                         uint64_t capacity, uint64_t *h_nnz, void *stream);
int asp_extract_signs_dev(uint64_t num_spins, double const *d_psi, uint64_t *d_signs, void *stream);
/* out[0..m] = exclusive prefix sums of in[0..m) (out has m+1 entries). tmp: >= asp_scan_tmp_bytes(m). */
size_t asp_scan_tmp_bytes(uint64_t m);
int asp_exclusive_scan_i64(int64_t const *d_in, int64_t *d_out, uint64_t m, void *d_tmp, void *stream);

/* ------------------------------------------------------------------------------------
 * 2. Fused extraction: neighbour generation (replaces lattice_symmetries'
 *    Operator.batched_apply as called at common.py:85-106) + search in the sorted sampled
 *    set (common.py:116-128,173 / build_matrix.c:36-37) + coupling
 *    J_ij = c_ij |psi_i| |psi_j| (build_matrix.c:38-43 / common.py:71-82) + two-pass CSR.
 *    Output is canonical CSR: columns ascending inside a row, duplicates summed in
 *    generation order, diagonal kept (what common.py:193-196 ends with, before its
 *    0.5(M+M^T) which asp_csr_symmetrize performs).
 * ---------------------------------------------------------------------------------- */

/* matrices: [num_terms][16] row-major 4x4 two-site matrices; term t acts on the bonds
 * sites[2*b], sites[2*b+1] for b in [term_offsets[t], term_offsets[t+1]).
 * Local state index a = 2*bit(s, site0) + bit(s, site1); coefficient of s -> s' is
 * matrix[a][a'].  hamming_weight < 0: unrestricted.  spin_inversion in {0, +1, -1}.
 * perms: [num_perms][number_spins] -- ALL non-identity elements of the permutation group
 * (new bit k = old bit perm[k]) with real characters[num_perms]; num_perms = 0 for none. */
int asp_operator_create(asp_operator **out, uint32_t number_spins, int32_t hamming_weight,
                        int32_t spin_inversion, uint32_t num_terms, double const *matrices,
                        uint32_t const *term_offsets, uint32_t const *sites, uint32_t num_perms,
                        uint32_t const *perms, double const *characters);
void asp_operator_destroy(asp_operator *op);
/* Upper bound on candidates per row (off-diagonal moves + the diagonal). */
uint32_t asp_operator_max_candidates(asp_operator const *op);
/* 1 if rows come out column-sorted and duplicate-free without a canonicalisation pass. */
int asp_operator_is_sorted_emitter(asp_operator const *op);

/* batched_apply on device (common.py:96 contract, 64-bit keys): writes up to
 * max_candidates per row in row-major order; d_counts[num_rows]; returns total in *h_total. */
int asp_operator_apply_dev(asp_operator const *op, uint64_t num_rows, uint64_t const *d_spins,
                           uint64_t *d_other_spins, double *d_other_coeffs, int64_t *d_counts,
                           uint64_t capacity, uint64_t *h_total, void *stream);

size_t asp_extract_workspace_bytes(asp_operator const *op, uint64_t n_total, uint64_t num_rows);

/* Pass 1: index the sorted basis, count couplings of rows [row_begin, row_begin+num_rows)
 * against the FULL basis d_spins[0..n_total), scan.  Synchronises; *h_nnz = couplings. */
int asp_extract_count(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                      uint64_t row_begin, uint64_t num_rows, void *d_workspace,
                      size_t workspace_bytes, uint64_t *h_nnz, void *stream);

/* Pass 2: d_psi[n_total] amplitudes (sign ignored).  d_indptr[num_rows+1] int64 (starts at
 * 0 for this row block), d_indices[nnz] int32 GLOBAL column numbers, d_data[nnz] f64. */
int asp_extract_fill(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                     double const *d_psi, uint64_t row_begin, uint64_t num_rows,
                     void *d_workspace, size_t workspace_bytes, int64_t *d_indptr,
                     int32_t *d_indices, double *d_data, void *stream);

/* Single pass (the fast path): index + generate + search + couplings + CSR in one kernel;
 * every candidate is searched once and a tile of rows obtains its CSR offset by decoupled
 * look-back.  Like the reference's C contract (cbits/build_matrix.c:22-28: outputs sized by
 * the caller to a worst case) the caller passes the room of d_indices/d_data in `capacity`
 * (entries); an upper bound is num_rows * asp_operator_max_candidates(op).  d_indptr
 * [num_rows+1] is always complete.  h_nnz != NULL: synchronises, *h_nnz = couplings; when
 * that exceeds `capacity` nothing past the capacity was written and ASP_ERR_WORKSPACE is
 * returned (call again with the larger capacity).  h_nnz == NULL: no synchronisation, the
 * count is d_indptr[num_rows]. */
size_t asp_extract_csr_workspace_bytes(asp_operator const *op, uint64_t n_total, uint64_t num_rows);
int asp_extract_csr(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                    double const *d_psi, uint64_t row_begin, uint64_t num_rows, void *d_workspace,
                    size_t workspace_bytes, uint64_t capacity, int64_t *d_indptr,
                    int32_t *d_indices, double *d_data, uint64_t *h_nnz, void *stream);
/* Test hooks.  Survivor-list entries per warp of the single-pass kernel (0 = automatic, 512):
 * small values force many exact-search rounds per tile.  Tuning: change the log2 size of the
 * Bloom filter / first-position table relative to the automatic choice, stage_a_mode 1 = test
 * move applicability lane by lane instead of on bit planes. */
void asp_debug_set_hit_list_capacity(int entries_per_warp);
void asp_debug_set_extract_tuning(int filter_bits_delta, int table_bits_delta, int stage_a_mode);
/* Measurement hook: when enabled, every launch of the single-pass extraction kernel is bracketed by
 * CUDA events on its own stream; the second call returns the device time of the LAST such launch
 * (milliseconds; synchronises on it; -1 if none). */
void asp_debug_set_apply_mode(int mode); /* asp_operator_apply_dev: 1 = always the general kernels, 2 = lane-per-row fill instead of warp-per-row */
void asp_debug_set_sa_team_ctas(int max_ctas_per_team); /* asp_sa_anneal: cap the CTAs that share a replica group (0 = no cap); tests use it to reach the ticketed task hand-out on small models */
void asp_debug_time_extract_kernel(int enable);
float asp_debug_last_extract_kernel_ms(void);
/* Same for the launch `back` launches before the last one (0 = last; the last 64 timed launches are kept). */
float asp_debug_extract_kernel_ms(int back);

/* Canonical CSR of generation-order rows (raw output of asp_build_matrix_dev): inside each
 * row a stable sort by column, duplicates summed in generation order -- what scipy's
 * csr_matrix + sort_indices yield at common.py:193-195.  max_row_len: upper bound on raw
 * row length (<= 1024; 0 = 1024).  d_indices/d_data need capacity nnz_in. */
int asp_csr_canonicalize(uint64_t num_rows, uint32_t max_row_len, int64_t const *d_row_offsets,
                         uint32_t const *d_cols, double const *d_vals, uint64_t nnz_in,
                         int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                         void *stream);

/* In-place J <- 0.5 (J + J^T) for a structurally symmetric canonical CSR (Hermitian H);
 * *h_asymmetric = number of entries whose transpose is missing (then nothing is changed). */
int asp_csr_symmetrize(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices,
                       double *d_data, uint64_t *h_asymmetric, void *stream);

/* HOST-buffer entry points (the end-to-end path) for rows [row_begin, row_begin+num_rows).
 * One call, caller-sized outputs (`capacity` entries of h_indices/h_data, as in
 * cbits/build_matrix.c:22-28): H2D of the full basis, single-pass extraction in row chunks,
 * each chunk's rows copied back while the next chunk is extracted.  *h_nnz = couplings;
 * ASP_ERR_WORKSPACE when that exceeds `capacity` (h_indptr is complete, call again).
 * Pinned host buffers make the copies asynchronous. */
int asp_extract_host(asp_operator const *op, uint64_t n_total, uint64_t const *h_spins,
                     double const *h_psi, uint64_t row_begin, uint64_t num_rows, uint64_t capacity,
                     int64_t *h_indptr, int32_t *h_indices, double *h_data, uint64_t *h_nnz);
/* Same with int32 row starts -- the index type scipy itself picks for both CSR index arrays when
 * nnz < 2^31 and what the reference's model dump stores (common.py:762-763): csr_matrix((data,
 * indices, indptr)) then adopts the buffers without a copy, and 4 bytes per row less cross PCIe.
 * ASP_ERR_UNSUPPORTED when the couplings do not fit. */
int asp_extract_host_i32(asp_operator const *op, uint64_t n_total, uint64_t const *h_spins,
                         double const *h_psi, uint64_t row_begin, uint64_t num_rows,
                         uint64_t capacity, int32_t *h_indptr, int32_t *h_indices, double *h_data,
                         uint64_t *h_nnz);
/* Two calls, exact-size outputs: begin = H2D + extraction (returns nnz), finish = D2H into
 * caller buffers (h_indptr[num_rows+1], h_indices/h_data[nnz]), then frees the job. */
int asp_extract_host_begin(asp_operator const *op, uint64_t n_total, uint64_t const *h_spins,
                           double const *h_psi, uint64_t row_begin, uint64_t num_rows,
                           uint64_t *h_nnz, asp_host_job **job);
int asp_extract_host_finish(asp_host_job *job, int64_t *h_indptr, int32_t *h_indices, double *h_data);
/* asp_extract_host_i32 without blocking the caller: the call runs on a worker thread with its own
 * device buffers and streams; asp_extract_host_join waits for it, returns its code (the error text
 * moves to the joining thread) and frees the job.  Two jobs may be in flight: submitting call k+1
 * before joining call k lets its upload and extraction ride under call k's download (the host link
 * is full duplex).  The caller's buffers must stay untouched until the join. */
int asp_extract_host_i32_submit(asp_operator const *op, uint64_t n_total, uint64_t const *h_spins,
                                double const *h_psi, uint64_t row_begin, uint64_t num_rows,
                                uint64_t capacity, int32_t *h_indptr, int32_t *h_indices,
                                double *h_data, asp_host_job **job);
int asp_extract_host_join(asp_host_job *job, uint64_t *h_nnz);
/* The host entry points keep their device buffers in two process-wide arenas that only grow
 * (at most two jobs in flight); this frees the idle ones. */
void asp_host_release(void);

/* ------------------------------------------------------------------------------------
 * 3. Energies and overlaps (replace sa.Hamiltonian.energy, full_hilbert_space.py:144, and
 *    compute_accuracy_and_overlap, common.py:211-229), batched over R packed replicas.
 *    d_bits: [R][ceil(n/64)].  d_field / d_weights may be NULL (zeros / ones).
 * ---------------------------------------------------------------------------------- */
int asp_energy(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices, double const *d_data,
               double const *d_field, uint32_t num_replicas, uint64_t const *d_bits,
               double *d_energy, void *stream);
int asp_accuracy_overlap(uint64_t n, uint32_t num_replicas, uint64_t const *d_predicted,
                         uint64_t const *d_exact, double const *d_weights, double *d_accuracy,
                         double *d_overlap, void *stream);

/* ------------------------------------------------------------------------------------
 * 4. Replica simulated annealing (replaces ising_glass_annealer.anneal as called at
 *    common.py:242-248).  The plan colours the coupling graph, relabels spins so every
 *    colour class is a contiguous range (classes padded to multiples of 4), and keeps the
 *    relabelled CSR on the device.  A sweep visits positions 0..n_padded-1 in order; spins
 *    of one class do not interact, so the kernel updates a class in parallel and the
 *    result equals the sequential sweep (DESIGN.md "SA chain definition").
 *    The CSR must be SYMMETRIC (J = J^T; dE is taken from row p alone and the colouring looks
 *    at row adjacency only) -- the caller's job, like the reference's preconditions; the Python
 *    mirror checks it.  The plan BORROWS d_indptr / d_indices / d_data / d_field (the greedy
 *    solver and the energy pass read the original model): they must outlive the plan.
 *    Everything is built on the device in `stream` (the plan's own buffers come from the
 *    device's stream-ordered memory pool); the call returns when the plan is complete, and
 *    asp_sa_plan_destroy waits for the device before it hands the buffers back.
 * ---------------------------------------------------------------------------------- */
int asp_sa_plan_create(asp_sa_plan **out, uint64_t n, int64_t const *d_indptr,
                       int32_t const *d_indices, double const *d_data, double const *d_field,
                       void *stream);
void asp_sa_plan_destroy(asp_sa_plan *plan);
/* sizes: n_padded positions, number of colour classes, nnz of the relabelled CSR. */
int asp_sa_plan_info(asp_sa_plan const *plan, uint64_t *n_padded, uint32_t *num_classes, uint64_t *nnz);
/* Copy the relabelled model to HOST buffers (tests feed it to the CPU oracle):
 * order[n_padded] (position -> original spin, -1 = padding), class_ptr[num_classes+1],
 * indptr[n_padded+1], indices[nnz], data[nnz], field[n_padded]. Any pointer may be NULL. */
int asp_sa_plan_export(asp_sa_plan const *plan, int32_t *h_order, int64_t *h_class_ptr,
                       int64_t *h_indptr, int32_t *h_indices, double *h_data, double *h_field);
/* h_betas[num_sweeps] inverse temperatures (host).  d_x0: packed start configuration in
 * ORIGINAL spin order or NULL (random start).  replica_offset: global index of this call's
 * first replica (a multiple of 32) -- replica r draws the random stream of global replica
 * replica_offset + r, so R replicas split over G GPUs reproduce one G*R-replica run.
 * Outputs, original spin order: d_best_bits [R][ceil(n/64)], d_best_energy[R] (exact,
 * recomputed in f64; may be NULL). */
int asp_sa_anneal(asp_sa_plan *plan, uint32_t num_replicas, uint32_t replica_offset,
                  uint32_t num_sweeps, double const *h_betas, uint64_t seed, uint64_t const *d_x0,
                  double energy_scale, uint64_t *d_best_bits, double *d_best_energy, void *stream);
/* ------------------------------------------------------------------------------------
 * 5. Greedy solver (replaces ising_glass_annealer.greedy_solve as called at common.py:249-250;
 *    algorithm restated from the Python preserved at common.py:298-438, see csrc/greedy.cu):
 *    strongest couplings first -- clusters merge with the joining edge satisfied (= maximum
 *    spanning forest by |J|, built with Boruvka rounds) -- then local-descent sweeps in the
 *    plan's position order until no flip lowers the energy.  Deterministic.
 *    d_bits [ceil(n/64)] packed signs in ORIGINAL spin order; d_energy (may be NULL) the exact
 *    energy; *h_rounds / *h_sweeps (may be NULL) merge rounds and descent sweeps performed.
 * ---------------------------------------------------------------------------------- */
int asp_greedy_solve(asp_sa_plan *plan, uint64_t *d_bits, double *d_energy, uint32_t *h_rounds,
                     uint32_t *h_sweeps, void *stream);

/* ------------------------------------------------------------------------------------
 * 6. Cluster sparsification helpers (callers around the extraction).
 *    asp_csr_strongest_offdiag: out[i] = max_{j != i} |J_ij| (get_strongest_off_diag,
 *    common.py:525-541).  asp_cutoff_components: connected components of the couplings that
 *    survive the global cutoff of sparsify_using_global_cutoff (common.py:621-662): an entry
 *    survives when it is non-zero and (|J_ij| >= reltol * max|J| or both spins are frozen;
 *    d_frozen [n] bytes, may be NULL); d_labels[i] = smallest vertex of i's component.
 * ---------------------------------------------------------------------------------- */
int asp_csr_strongest_offdiag(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices,
                              double const *d_data, double *d_out, void *stream);
int asp_cutoff_components(uint64_t n, int64_t const *d_indptr, int32_t const *d_indices,
                          double const *d_data, uint64_t nnz, double reltol,
                          unsigned char const *d_frozen, int32_t *d_labels, void *stream);

/* ------------------------------------------------------------------------------------
 * 7. Multi-GPU exchange X1 over NVLink peer memory (one process per GPU).
 *    The reference is single-process: every caller holds the whole basis (common.py:146).
 *    Sharded, every rank needs all row blocks of the sorted basis.  The blocks live in
 *    peer buffers (device memory other processes map through a 64-byte CUDA IPC handle
 *    that travels over any host channel); asp_gather_index pulls them over NVLink, writes
 *    the rank's private full copy AND indexes it in ONE kernel (replaces
 *    ncclAllGather x2 + the index pass of asp_extract_csr); asp_extract_csr_indexed then
 *    runs the single-pass extraction on the indexed workspace.
 *    Epoch flags (uint64 arrays in peer memory, zeroed by asp_peer_alloc) order the ranks on
 *    the device: asp_peer_signal stores `value` into d_flags[p][slot] of every rank p
 *    (system-scope release, stream-ordered after the caller's earlier work);
 *    asp_peer_wait / asp_gather_index spin until flags[q] >= value (system-scope acquire;
 *    10 s without progress traps instead of hanging the device).
 * ---------------------------------------------------------------------------------- */
int asp_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle /* [64] out */);
int asp_peer_open(unsigned char const *handle /* [64] */, void **d_ptr);
int asp_peer_close(void *d_ptr);
int asp_peer_free(void *d_ptr);
int asp_peer_signal(uint32_t world, uint64_t *const *d_flags /* host [world] of device ptrs */,
                    uint32_t slot, uint64_t value, void *stream);
int asp_peer_wait(uint32_t world, uint64_t const *d_flags /* device [world] */, uint64_t value,
                  void *stream);
/* shard_begin: host [world+1], global index of every block's first key (shard_begin[world] =
 * n_total); d_shard_spins / d_shard_psi: host [world] of device pointers (16-byte aligned; own
 * or peer-mapped).  d_ready: this rank's flag array [world] (NULL = blocks are already
 * complete); block q is read once d_ready[q] >= epoch.  d_spins / d_psi [n_total]: private
 * full copy (out).  d_workspace: asp_extract_csr_workspace_bytes(op, n_total, num_rows). */
int asp_gather_index(asp_operator const *op, uint32_t world, uint32_t rank,
                     uint64_t const *shard_begin, uint64_t const *const *d_shard_spins,
                     double const *const *d_shard_psi, uint64_t const *d_ready, uint64_t epoch,
                     uint64_t *d_spins, double *d_psi, uint64_t num_rows, void *d_workspace,
                     size_t workspace_bytes, void *stream);
/* X1 without the index (flags and arguments as for asp_gather_index).  Gather mode 0: the copy
 * engines pull every block into the private full copy -- no SM is used, so the call overlaps
 * completely with an extraction running on another stream: a pipeline over independent
 * extractions (one per cluster in the reference's experiment) gathers basis k+1 this way while
 * basis k is indexed and extracted by the ordinary asp_extract_csr.  Gather mode 1/2: one thread per
 * CTA drives bulk copies peer -> shared -> private copy (cp.async.bulk both ways); faster alone,
 * but beside a resident extraction it costs more than it gains. */
int asp_gather_blocks(uint32_t world, uint32_t rank, uint64_t const *shard_begin,
                      uint64_t const *const *d_shard_spins, double const *const *d_shard_psi,
                      uint64_t const *d_ready, uint64_t epoch, uint64_t *d_spins, double *d_psi,
                      void *stream);
/* The sharded end-to-end path: asp_gather_index on `stream`, then this call -- the single-pass
 * extraction of rows [row_begin, row_begin+num_rows) in row chunks on the indexed workspace, every
 * chunk's rows copied to the caller's HOST buffers (pinned: asynchronously) while the next chunk is
 * extracted.  Same output contract as asp_extract_host. */
int asp_extract_indexed_to_host(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                                double const *d_psi, uint64_t row_begin, uint64_t num_rows,
                                void *d_workspace, size_t workspace_bytes, uint64_t capacity,
                                int64_t *h_indptr, int32_t *h_indices, double *h_data, uint64_t *h_nnz,
                                void *stream);
int asp_extract_indexed_to_host_i32(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                                    double const *d_psi, uint64_t row_begin, uint64_t num_rows,
                                    void *d_workspace, size_t workspace_bytes, uint64_t capacity,
                                    int32_t *h_indptr, int32_t *h_indices, double *h_data,
                                    uint64_t *h_nnz, void *stream);
/* asp_gather_index has three implementations of the same contract (measured within 5 % of each
 * other and of NCCL's all-gather alone, ~510 GB/s pulled per rank on 4 GPUs -- the fabric, not the
 * kernel, sets the pace): 2 (default) = ONE persistent kernel, cp.async.bulk (TMA) keeps 128 KB per
 * SM of peer memory in flight into shared memory while the threads index the chunk that landed and
 * write the private copy; 1 = ONE kernel with plain 16-byte loads; 0 = the copy engines pull the
 * blocks (cudaMemcpyAsync on an internal stream) and every block is indexed on the SMs as soon as
 * it has landed. */
void asp_set_gather_mode(int mode);
/* Copy streams (= copy engines, 1..8, default 2) the copy-engine gather (mode 0, asp_gather_blocks) deals the row
  * blocks over: that many peers are pulled at a time (default 2: measured best on 8 GPUs). */
void asp_set_copy_streams(int streams);
/* asp_extract_csr without its zero + index pass: the workspace was prepared by asp_gather_index
 * for the same (op, n_total, num_rows) on the same stream.  The index is SINGLE-USE (the extraction
 * consumes its tickets and look-back words): a second indexed extraction on the same workspace without
 * a new asp_gather_index returns ASP_ERR_ARG. */
int asp_extract_csr_indexed(asp_operator const *op, uint64_t n_total, uint64_t const *d_spins,
                            double const *d_psi, uint64_t row_begin, uint64_t num_rows,
                            void *d_workspace, size_t workspace_bytes, uint64_t capacity,
                            int64_t *d_indptr, int32_t *d_indices, double *d_data, uint64_t *h_nnz,
                            void *stream);

/* ------------------------------------------------------------------------------------
 * 8. Sampling front-end: the data-parallel pieces that feed n and psi
 *    into the extraction.
 *    asp_batched_index: basis.batched_index / ls.batched_index as called at common.py:283,
 *    :817 and sampled_connected_components.py:720,:730 -- d_index[j] = position of
 *    d_needles[j] in the ascending (unsigned) unique array d_sorted[n], -1 when absent;
 *    h_missing (may be NULL; non-NULL synchronises) = number of absent needles.
 *    asp_sample_indices: monte_carlo_sampling (common.py:269-278), i.e. legacy
 *    np.random.choice(n, m, replace=True, p = |psi|^power / sum): cdf = cumsum(|psi|^power),
 *    cdf /= cdf[-1], d_index[j] = searchsorted(cdf, d_uniform[j], side="right") for the
 *    caller's uniform draws in [0, 1).  d_cdf: optional [n] buffer that receives the cdf.
 * ---------------------------------------------------------------------------------- */
int asp_batched_index(uint64_t n, uint64_t const *d_sorted, uint64_t m, uint64_t const *d_needles,
                      int64_t *d_index, uint64_t *h_missing, void *stream);
int asp_sample_indices(uint64_t n, double const *d_psi, double power, uint64_t m,
                       double const *d_uniform, int64_t *d_index, double *d_cdf, void *stream);

/* Number of kernel launches the library has issued in this process (bench accounting). */
uint64_t asp_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ASP_B200_H */
