"""The reference's hot-path entry points (annealing_sign_problem/common.py:46-261, 544-548,
782-787) with the same names, arguments and error behaviour, running on the B200 through
libasp_b200.so.  ``import annealing_sign_problem_b200.common as common`` is the drop-in for
``import annealing_sign_problem.common as common`` as far as this path is concerned.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np
import scipy.sparse
import torch

from . import annealer as sa
from . import symmetries as ls
from ._lib import AspError, check, ffi, lib, ptr, require_cuda, stream

try:  # the reference logs through loguru (common.py:15); stay quiet if it is absent
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("annealing_sign_problem_b200")

_SIGN_BIT = -0x8000000000000000


@dataclass
class IsingModel:  # common.py:46-55
    spins: np.ndarray
    quantum_hamiltonian: object
    ising_hamiltonian: sa.Hamiltonian
    initial_signs: np.ndarray

    @property
    def size(self):
        return self.spins.shape[0]


def load_hamiltonian(filename: str) -> ls.Operator:  # common.py:782-787
    config = ls.load_config(filename)
    basis = ls.SpinBasis.load_from_yaml(config["basis"])
    return ls.Operator.load_from_yaml(config["hamiltonian"], basis)


def _normalize_spins_1d(spins) -> np.ndarray:
    """common.py:58-68 accepts [n] or [n,8]; every caller keeps only column 0 (:100)."""
    spins = np.asarray(spins, dtype=np.uint64, order="C")
    if spins.ndim <= 1:
        return np.ascontiguousarray(spins.reshape(-1))
    if spins.ndim == 2:
        if spins.shape[1] != 8:
            raise ValueError("'spins' has wrong shape: {}; expected (?, 8)".format(spins.shape))
        if spins[:, 1:].any():
            raise ValueError("only works with up to 64 bits")
        return np.ascontiguousarray(spins[:, 0])
    raise ValueError("'spins' has wrong shape: {}; expected a 2D array".format(spins.shape))


def _sort_unique_device(x: torch.Tensor):
    """Ascending UNSIGNED order of int64 bit patterns; -> (unique sorted, index of first
    occurrence, counts), i.e. np.unique(..., return_index=True, return_counts=True)."""
    keys = x ^ _SIGN_BIT
    sorted_keys, perm = torch.sort(keys, stable=True)
    n = sorted_keys.shape[0]
    if n == 0:
        return x, perm, torch.zeros(0, dtype=torch.int64, device=x.device)
    head = torch.ones(n, dtype=torch.bool, device=x.device)
    head[1:] = sorted_keys[1:] != sorted_keys[:-1]
    starts = torch.nonzero(head).reshape(-1)
    counts = torch.diff(starts, append=torch.tensor([n], device=x.device))
    return sorted_keys[starts] ^ _SIGN_BIT, perm[starts], counts


# ---------------------------------------------------------------------------------------
# Device-resident extraction (what bench.py times as `value`; host wrappers below add I/O)
# ---------------------------------------------------------------------------------------
def extract_csr_two_pass_device(operator: ls.Operator, spins: torch.Tensor, psi: torch.Tensor,
                                row_begin: int = 0, num_rows: Optional[int] = None, workspace: Optional[torch.Tensor] = None):
    """asp_extract_count + asp_extract_fill: the exact-allocation two-call API (every candidate
    is searched twice).  Same result as :func:`extract_csr_device`."""
    dev = require_cuda()
    n_total = int(spins.shape[0])
    if num_rows is None:
        num_rows = n_total - row_begin
    need = int(lib().asp_extract_workspace_bytes(operator.handle, n_total, num_rows))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    nnz = ffi.new("uint64_t *")
    check(lib().asp_extract_count(operator.handle, n_total, ptr(spins, "uint64_t *"), row_begin, num_rows,
                                  ptr(workspace, "void *"), workspace.numel(), nnz, stream()))
    indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(int(nnz[0]), dtype=torch.int32, device=dev)
    data = torch.empty(int(nnz[0]), dtype=torch.float64, device=dev)
    check(lib().asp_extract_fill(operator.handle, n_total, ptr(spins, "uint64_t *"), ptr(psi, "double *"),
                                 row_begin, num_rows, ptr(workspace, "void *"), workspace.numel(),
                                 ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"), stream()))
    return indptr, indices, data


_WORST_CASE_BYTES = 1 << 30  # allocate num_rows * max_candidates outright below this many bytes


def extract_csr_device(operator: ls.Operator, spins: torch.Tensor, psi: torch.Tensor,
                       row_begin: int = 0, num_rows: Optional[int] = None, workspace: Optional[torch.Tensor] = None,
                       nnz_hint: Optional[int] = None):
    """Rows [row_begin, row_begin+num_rows) of J against the full sorted basis ``spins``.

    spins: int64 bit patterns, ascending (unsigned), unique, CUDA.  psi: f64 CUDA, same length.
    -> (indptr int64 [num_rows+1], indices int32 [nnz] global columns, data f64 [nnz]).

    Unsymmetrised operators take the single-pass kernel (asp_extract_csr).  Its outputs are
    sized by the caller like the reference's C contract (cbits/build_matrix.c:22-28): the worst
    case num_rows * max_candidates when that is small, else ``nnz_hint`` (default 8 per row);
    if the guess was short the call reports the exact count and is repeated once.
    """
    dev = require_cuda()
    n_total = int(spins.shape[0])
    if num_rows is None:
        num_rows = n_total - row_begin
    if operator.is_sorted_emitter:
        need = int(lib().asp_extract_csr_workspace_bytes(operator.handle, n_total, num_rows))
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        worst = num_rows * operator.max_candidates
        if nnz_hint is not None:
            capacity = min(worst, int(nnz_hint))
        elif worst * 12 <= _WORST_CASE_BYTES:
            capacity = worst
        else:
            capacity = min(worst, 8 * num_rows)
        indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
        nnz = ffi.new("uint64_t *")
        for _ in range(2):
            indices = torch.empty(capacity, dtype=torch.int32, device=dev)
            data = torch.empty(capacity, dtype=torch.float64, device=dev)
            rc = lib().asp_extract_csr(operator.handle, n_total, ptr(spins, "uint64_t *"), ptr(psi, "double *"),
                                       row_begin, num_rows, ptr(workspace, "void *"), workspace.numel(), capacity,
                                       ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"), nnz, stream())
            if rc == lib().ASP_ERR_UNSUPPORTED:  # more distinct moves than the kernel's tag layout holds
                return extract_csr_two_pass_device(operator, spins, psi, row_begin, num_rows)
            if rc == lib().ASP_ERR_WORKSPACE and int(nnz[0]) > capacity:
                capacity = int(nnz[0])  # the guess was short: exact size, second (last) attempt
                del indices, data
                continue
            check(rc)
            break
        m = int(nnz[0])
        if m == capacity:
            return indptr, indices, data
        if 2 * m < capacity:  # do not pin an oversized buffer behind a small view
            return indptr, indices[:m].clone(), data[:m].clone()
        return indptr, indices[:m], data[:m]
    # symmetrised / non-sorted operators: neighbour lists on the device, then the
    # explicit-candidate kernel, then canonicalisation
    rows = spins[row_begin:row_begin + num_rows]
    other_spins, other_coeffs, other_counts = operator.batched_apply_device(rows)
    return build_csr_from_candidates_device(spins, psi, row_begin, other_spins, other_coeffs, other_counts,
                                            max_row_len=operator.max_candidates)


def extract_csr_indexed_device(operator: ls.Operator, spins: torch.Tensor, psi: torch.Tensor, row_begin: int, num_rows: int,
                               workspace: torch.Tensor, capacity: int):
    """Single-pass extraction on a workspace already indexed by asp_gather_index (the sharded path,
    distributed.PeerBasis.gather_index).  -> (indptr, indices[:nnz], data[:nnz]); raises when
    ``capacity`` is too small (the caller sized it from an earlier pass, like the C contract)."""
    dev = require_cuda()
    indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(capacity, dtype=torch.int32, device=dev)
    data = torch.empty(capacity, dtype=torch.float64, device=dev)
    nnz = ffi.new("uint64_t *")
    check(lib().asp_extract_csr_indexed(operator.handle, int(spins.shape[0]), ptr(spins, "uint64_t *"), ptr(psi, "double *"),
                                        row_begin, num_rows, ptr(workspace, "void *"), workspace.numel(), capacity,
                                        ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"), nnz, stream()))
    m = int(nnz[0])
    return indptr, indices[:m], data[:m]


def extract_indexed_to_host(operator: ls.Operator, spins: torch.Tensor, psi: torch.Tensor, row_begin: int, num_rows: int,
                            workspace: torch.Tensor, h_indptr: torch.Tensor, h_indices: torch.Tensor, h_data: torch.Tensor) -> int:
    """The sharded end-to-end call (asp_extract_indexed_to_host): extraction on a workspace indexed by
    asp_gather_index, CSR rows written to the caller's (pinned) HOST tensors chunk by chunk while the next
    chunk is extracted.  -> number of couplings."""
    require_cuda()
    assert h_indptr.dtype in (torch.int64, torch.int32) and h_indices.dtype == torch.int32 and h_data.dtype == torch.float64
    assert h_indptr.numel() >= num_rows + 1 and h_indices.numel() == h_data.numel()
    nnz = ffi.new("uint64_t *")
    narrow = h_indptr.dtype == torch.int32  # scipy's own index type below 2^31 couplings
    call = lib().asp_extract_indexed_to_host_i32 if narrow else lib().asp_extract_indexed_to_host
    check(call(operator.handle, int(spins.shape[0]), ptr(spins, "uint64_t *"), ptr(psi, "double *"),
               row_begin, num_rows, ptr(workspace, "void *"), workspace.numel(), h_indices.numel(),
               ffi.cast("int32_t *" if narrow else "int64_t *", h_indptr.data_ptr()), ffi.cast("int32_t *", h_indices.data_ptr()),
               ffi.cast("double *", h_data.data_ptr()), nnz, stream()))
    return int(nnz[0])


def build_csr_from_candidates_device(spins, psi, row_begin, other_spins, other_coeffs, other_counts, max_row_len=0):
    """Explicit-candidate path (what cbits/build_matrix.c does) + canonical CSR."""
    dev = require_cuda()
    n_total = int(spins.shape[0])
    num_rows = int(other_counts.shape[0])
    T = int(other_spins.shape[0])
    other_counts = other_counts.contiguous()
    offsets = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    tmp = torch.empty(int(lib().asp_scan_tmp_bytes(num_rows)), dtype=torch.uint8, device=dev)
    check(lib().asp_exclusive_scan_i64(ptr(other_counts, "int64_t *"), ptr(offsets, "int64_t *"),
                                       num_rows, ptr(tmp, "void *"), stream()))
    row_offsets = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    rows = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    cols = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    vals = torch.empty(max(T, 1), dtype=torch.float64, device=dev)
    nnz = ffi.new("uint64_t *")
    # other_psi = NULL: a hit takes |psi[column]| (live-path semantics, common.py:71-82)
    check(lib().asp_build_matrix_dev(
        n_total, ptr(spins, "uint64_t *"), row_begin, num_rows, ffi.NULL, ptr(psi, "double *"),
        ptr(other_spins.contiguous(), "uint64_t *"), ptr(other_coeffs.contiguous(), "double *"),
        ptr(offsets, "int64_t *"), ffi.NULL, ptr(row_offsets, "int64_t *"), ptr(rows, "uint32_t *"),
        ptr(cols, "uint32_t *"), ptr(vals, "double *"), ffi.NULL, T, nnz, stream()))
    m = int(nnz[0])
    if max_row_len == 0:
        max_row_len = int(other_counts.max()) if num_rows else 1
    return canonicalize_csr_device(num_rows, row_offsets, cols, vals, m, max_row_len)


def canonicalize_csr_device(num_rows, row_offsets, cols, vals, nnz_in, max_row_len):
    """Stable column sort inside each row + duplicates summed in generation order
    (asp_csr_canonicalize)."""
    dev = cols.device
    indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(max(nnz_in, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(nnz_in, 1), dtype=torch.float64, device=dev)
    nnz = ffi.new("uint64_t *")
    check(lib().asp_csr_canonicalize(num_rows, max_row_len, ptr(row_offsets, "int64_t *"), ptr(cols, "uint32_t *"),
                                     ptr(vals, "double *"), nnz_in, ptr(indptr, "int64_t *"),
                                     ptr(indices, "int32_t *"), ptr(data, "double *"), nnz, stream()))
    m = int(nnz[0])
    return indptr, indices[:m].contiguous(), data[:m].contiguous()


def symmetrize_csr_device(n: int, indptr, indices, data) -> None:
    """In place 0.5 (J + J^T) (common.py:194) for a structurally symmetric J."""
    missing = ffi.new("uint64_t *")
    check(lib().asp_csr_symmetrize(n, ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"),
                                   missing, stream()))
    if missing[0]:
        raise AspError("the extracted matrix is not structurally symmetric ({} unmatched entries): "
                       "a non-Hermitian operator is outside the hot path".format(int(missing[0])))


_STAGING = {}


def _upload(array: np.ndarray, dev) -> torch.Tensor:
    """Host array -> device through a cached pinned staging buffer: a pageable numpy array goes up at ~8 GB/s through the
    driver's own bounce buffers; a memcpy into pinned memory (~20 GB/s on one core) plus a DMA at link speed is quicker for
    the 240 MB a 10^7-state call uploads."""
    src = torch.from_numpy(array)
    if array.nbytes < (1 << 22):
        return src.to(dev)
    key = src.dtype
    buf = _STAGING.get(key)
    if buf is None or buf.numel() < src.numel():
        buf = torch.empty(src.numel(), dtype=src.dtype, pin_memory=True)
        _STAGING[key] = buf
    view = buf[: src.numel()].view(src.shape)
    view.copy_(src)
    out = torch.empty(src.shape, dtype=src.dtype, device=dev)
    out.copy_(view, non_blocking=True)
    torch.cuda.current_stream().synchronize()  # the staging buffer is reused by the next upload
    return out


def _coo_from_valid_arrays(n: int, rows: np.ndarray, cols: np.ndarray, data: np.ndarray) -> scipy.sparse.coo_matrix:
    """A real scipy COO matrix around arrays that are valid by construction (row-sorted, columns ascending inside a row,
    no duplicates: the kernel's output).  The checked constructor would walk both index arrays four times for their
    minima and maxima -- 36 ms on 4.5e7 entries, half of what make_ising_model takes."""
    matrix = scipy.sparse.coo_matrix((n, n), dtype=np.float64)
    matrix.data = data
    try:
        matrix.coords = (rows, cols)
    except AttributeError:  # scipy < 1.13 keeps .row / .col
        matrix.row, matrix.col = rows, cols
    matrix.has_canonical_format = True
    return matrix


# ---------------------------------------------------------------------------------------
# make_ising_model -- common.py:131-208
# ---------------------------------------------------------------------------------------
def make_ising_model(
    spins,
    quantum_hamiltonian,
    log_psi=None,
    log_psi_fn: Optional[Callable] = None,
    external_field: bool = False,
    debug: bool = False,
):
    start_time = time.time()
    if log_psi is None and log_psi_fn is None:
        raise ValueError("at least one of log_psi or log_psi_fn should be specified")
    if external_field and log_psi_fn is None:
        raise ValueError("log_psi_fn should be specified when external_field=True")
    if external_field:
        assert False  # common.py:199-200: the reference's branch is `assert False`
    assert quantum_hamiltonian.basis.number_spins <= 64, "TODO: only works with up to 64 bits"
    dev = require_cuda()

    spins = _normalize_spins_1d(spins)
    d_spins_in = _upload(spins.view(np.int64), dev)
    d_spins, first, counts = _sort_unique_device(d_spins_in)
    if bool((counts != 1).any()):
        logger.warning("'spins' were not unique, are you sure this is what you want?")
    h_sorted = torch.empty(d_spins.shape, dtype=torch.int64, pin_memory=True)
    h_sorted.copy_(d_spins, non_blocking=True)
    if log_psi is not None:
        # first occurrence, sorted order (common.py:149-151) -- gathered on the device
        d_log_psi = _upload(np.ascontiguousarray(log_psi, dtype=np.complex128), dev)[first]
    torch.cuda.synchronize()
    spins = h_sorted.numpy().view(np.uint64)
    if log_psi is None:
        d_log_psi = torch.from_numpy(np.ascontiguousarray(log_psi_fn(spins), dtype=np.complex128)).to(dev)
    n = spins.shape[0]

    psi = torch.exp(d_log_psi)
    if not bool(torch.all(psi.imag.abs() <= 1e-6)):
        raise ValueError("expected all wavefunction coefficients to be real")
    psi = psi.real.contiguous()
    psi = psi / torch.linalg.norm(psi)

    tick = time.time()
    if isinstance(quantum_hamiltonian, ls.Operator):
        indptr, indices, data = extract_csr_device(quantum_hamiltonian, d_spins, psi)
    else:
        # a foreign ls.Operator-like object: its own batched_apply generates the neighbours
        # (chunked exactly as common.py:85-106), the search/coupling/CSR still run on the GPU
        out_s, out_c, out_k = [], [], []
        for start in range(0, n, 10000):
            x = np.zeros((min(start + 10000, n) - start, 8), dtype=np.uint64)
            x[:, 0] = spins[start:start + 10000]
            s, c, k = quantum_hamiltonian.batched_apply(x)
            if not np.allclose(c.imag, 0, atol=1e-6):
                raise ValueError("expected all Hamiltonian matrix elements to be real")
            out_s.append(np.ascontiguousarray(s[:, 0]))
            out_c.append(np.ascontiguousarray(c.real))
            out_k.append(np.asarray(k, dtype=np.int64))
        indptr, indices, data = build_csr_from_candidates_device(
            d_spins, psi, 0,
            torch.from_numpy(np.hstack(out_s).view(np.int64)).to(dev),
            torch.from_numpy(np.hstack(out_c)).to(dev),
            torch.from_numpy(np.hstack(out_k)).to(dev))
    symmetrize_csr_device(n, indptr, indices, data)
    if bool((data == 0).any()):  # scipy's binop drops explicit zeros (common.py:194)
        keep = data != 0
        rows = torch.repeat_interleave(torch.arange(n, device=dev), indptr[1:] - indptr[:-1])[keep]
        indices, data = indices[keep].contiguous(), data[keep].contiguous()
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    torch.cuda.synchronize()
    tock = time.time()
    logger.debug("extraction took {:.4f} seconds", tock - tick)

    # COO as the reference returns it (common.py:195-196: csr -> sort_indices -> tocoo): the row of every entry is made
    # on the device and the three arrays cross PCIe once, into pinned host memory (no CSR -> COO pass on the CPU)
    index_dtype = torch.int32 if max(n, int(indices.shape[0])) < 2 ** 31 else torch.int64
    rows = torch.repeat_interleave(torch.arange(n, dtype=index_dtype, device=dev), indptr[1:] - indptr[:-1])
    h_rows = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
    h_cols = torch.empty(indices.shape, dtype=index_dtype, pin_memory=True)
    h_data = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
    h_rows.copy_(rows, non_blocking=True)
    h_cols.copy_(indices if index_dtype == torch.int32 else indices.to(index_dtype), non_blocking=True)
    h_data.copy_(data, non_blocking=True)
    del rows
    torch.cuda.synchronize()
    matrix = _coo_from_valid_arrays(n, h_rows.numpy(), h_cols.numpy(), h_data.numpy())
    field = np.zeros(n, dtype=np.float64)
    ising_hamiltonian = sa.Hamiltonian(matrix, field, _device_csr=(indptr, indices, data, None))
    x0 = sa.signs_to_bits_device(psi).cpu().numpy().view(np.uint64)
    logger.debug("Took {:.2} seconds", time.time() - start_time)
    return IsingModel(spins, quantum_hamiltonian, ising_hamiltonian, x0)


# ---------------------------------------------------------------------------------------
# compute_accuracy_and_overlap -- common.py:211-229
# ---------------------------------------------------------------------------------------
def compute_accuracy_and_overlap(predicted, exact, weights=None, number_spins: Optional[int] = None) -> Tuple[float, float]:
    if weights is None and number_spins is None:
        raise ValueError("'weights' and 'number_spins' cannot be both None")
    if number_spins is None:
        number_spins = len(weights)
    acc, ov = accuracy_and_overlap_batched(np.asarray(predicted, dtype=np.uint64).reshape(1, -1), exact, weights, number_spins)
    return float(acc[0]), float(ov[0])


def accuracy_and_overlap_batched(predicted, exact, weights, number_spins: int):
    """All replicas at once (full_hilbert_space.py:164-186 loops over repetitions on the CPU)."""
    dev = require_cuda()
    p = torch.from_numpy(np.ascontiguousarray(predicted, dtype=np.uint64).view(np.int64)).to(dev)
    e = torch.from_numpy(np.ascontiguousarray(exact, dtype=np.uint64).view(np.int64)).to(dev)
    w = None if weights is None else torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
    acc = torch.empty(p.shape[0], dtype=torch.float64, device=dev)
    ov = torch.empty(p.shape[0], dtype=torch.float64, device=dev)
    check(lib().asp_accuracy_overlap(number_spins, p.shape[0], ptr(p, "uint64_t *"), ptr(e, "uint64_t *"),
                                     ptr(w, "double *"), ptr(acc, "double *"), ptr(ov, "double *"), stream()))
    return acc.cpu().numpy(), ov.cpu().numpy()


# ---------------------------------------------------------------------------------------
# solve_ising_model -- common.py:232-261
# ---------------------------------------------------------------------------------------
def binary_search(haystack, needles):  # common.py:544-548
    haystack = np.asarray(haystack)
    assert np.all(np.sort(haystack) == haystack)
    indices = np.searchsorted(haystack, needles)
    assert np.all(haystack[indices] == needles)
    return indices


# ---------------------------------------------------------------------------------------
# Cluster extension and sparsification -- the callers around the extraction
# (common.py:516-541, 621-692; experiments/sampled_connected_components.py:737-749)
# ---------------------------------------------------------------------------------------
def make_hamiltonian_extension(model: IsingModel, log_psi_fn: Callable) -> IsingModel:
    """common.py:516-522: one batched_apply shell around the model's states, unique, new model."""
    dev = require_cuda()
    op = model.quantum_hamiltonian
    if isinstance(op, ls.Operator):
        d_spins = torch.from_numpy(np.ascontiguousarray(model.spins).view(np.int64)).to(dev)
        other, _, _ = op.batched_apply_device(d_spins)
        spins = _sort_unique_device(other)[0].cpu().numpy().view(np.uint64)
    else:  # a foreign operator: chunks of 10 000 rows exactly as common.py:85-106
        parts = []
        for start in range(0, model.size, 10000):
            x = np.zeros((min(start + 10000, model.size) - start, 8), dtype=np.uint64)
            x[:, 0] = model.spins[start:start + 10000]
            s, c, _ = op.batched_apply(x)
            if not np.allclose(c.imag, 0, atol=1e-6):
                raise ValueError("expected all Hamiltonian matrix elements to be real")
            parts.append(np.ascontiguousarray(s[:, 0]))
        spins = np.unique(np.hstack(parts))
    return make_ising_model(spins, op, log_psi_fn=log_psi_fn)


def _device_csr_of(matrix_or_hamiltonian):
    if isinstance(matrix_or_hamiltonian, sa.Hamiltonian):
        return matrix_or_hamiltonian.device_csr()[:3]
    dev = require_cuda()
    m = scipy.sparse.csr_matrix(matrix_or_hamiltonian)
    m.sort_indices()
    return (torch.from_numpy(m.indptr.astype(np.int64)).to(dev), torch.from_numpy(m.indices.astype(np.int32)).to(dev),
            torch.from_numpy(np.ascontiguousarray(m.data, dtype=np.float64)).to(dev))


def get_strongest_off_diag(matrix) -> np.ndarray:
    """common.py:539-541: max_{j != i} |J_ij| per row (0 for rows without off-diagonal entries).
    Accepts a scipy matrix or a :class:`sa.Hamiltonian` (then the device copy is used)."""
    indptr, indices, data = _device_csr_of(matrix)
    n = int(indptr.shape[0]) - 1
    out = torch.zeros(n, dtype=torch.float64, device=indptr.device)
    check(lib().asp_csr_strongest_offdiag(n, ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"),
                                          ptr(out, "double *"), stream()))
    return out.cpu().numpy()


def sparsify_using_global_cutoff(model: IsingModel, reltol: float, frozen_spins) -> IsingModel:
    """common.py:634-692: couplings below ``reltol * max|J|`` are dropped unless both spins are frozen;
    the connected component that holds the frozen spins survives, with the ORIGINAL couplings among
    its members (common.py:671)."""
    dev = require_cuda()
    frozen_indices = binary_search(model.spins, frozen_spins)
    n = model.size
    ham = model.ising_hamiltonian
    indptr, indices, data, fld = ham.device_csr()
    original_nnz = int(data.shape[0])
    frozen = torch.zeros(n, dtype=torch.uint8, device=dev)
    frozen[torch.from_numpy(np.asarray(frozen_indices, dtype=np.int64)).to(dev)] = 1
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib().asp_cutoff_components(n, ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"), original_nnz,
                                      float(reltol), ptr(frozen, "unsigned char *"), ptr(labels, "int32_t *"), stream()))
    frozen_labels = labels[frozen.bool()]
    magic = int(labels[int(frozen_indices[0])])
    assert bool((frozen_labels == magic).all())  # common.py:667
    keep_spin = labels == magic
    new_index = torch.cumsum(keep_spin.to(torch.int64), 0) - 1
    rows = torch.repeat_interleave(torch.arange(n, device=dev), indptr[1:] - indptr[:-1])
    cols = indices.to(torch.int64)
    keep = keep_spin[rows] & keep_spin[cols]
    m = int(keep_spin.sum())
    new_rows = new_index[rows[keep]]
    new_indices = new_index[cols[keep]].to(torch.int32).contiguous()
    new_data = data[keep].contiguous()
    new_indptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
    new_indptr[1:] = torch.cumsum(torch.bincount(new_rows, minlength=m), 0)
    h_keep = keep_spin.cpu().numpy()
    spins = model.spins[h_keep]
    initial_signs = sa.signs_to_bits(sa.bits_to_signs(model.initial_signs, n)[h_keep])
    matrix = scipy.sparse.csr_matrix((new_data.cpu().numpy(), new_indices.cpu().numpy(), new_indptr.cpu().numpy().astype(np.int32)),
                                     shape=(m, m))
    field = (fld.cpu().numpy() if fld is not None else np.zeros(n))[h_keep]
    new_field = None if fld is None else fld[keep_spin].contiguous()
    new_model = IsingModel(spins, model.quantum_hamiltonian,
                           sa.Hamiltonian(matrix, field, _device_csr=(new_indptr, new_indices, new_data, new_field)), initial_signs)
    logger.info("number of spins: {} -> {}; number of connections: {} -> {}", n, new_model.size, original_nnz, int(new_data.shape[0]))
    return new_model


def solve_ising_model(
    model: IsingModel,
    mode: str = "sa",
    frozen_spins=None,
    seed: int = 12345,
    number_sweeps: int = 5120,
    repetitions: int = 64,
    only_best: bool = True,
):
    """Signs of the model's spins by replica annealing (``mode="sa"``) or by the greedy solver (``"greedy"``), as
    packed bits; with ``frozen_spins`` only the signs of those states, in their order (reference: common.py:232-261,
    same defaults and the same ValueError for an unknown mode)."""
    solvers = {
        "sa": lambda: sa.anneal(model.ising_hamiltonian, seed=seed, number_sweeps=number_sweeps, repetitions=repetitions,
                                only_best=only_best),
        "greedy": lambda: sa.greedy_solve(model.ising_hamiltonian),
    }
    if mode not in solvers:
        raise ValueError("invalid mode specified: '{}'; expected either 'sa' or 'greedy'".format(mode))
    x, _ = solvers[mode]()
    if frozen_spins is None:
        return x
    positions = binary_search(model.spins, frozen_spins)  # asserts that every frozen spin is part of the model
    return sa.signs_to_bits(sa.bits_to_signs(x, count=model.spins.size)[positions])


# ---------------------------------------------------------------------------------------
# Sampling front-end and frozen-spin glue -- common.py:262-285, 481-513,
# 806-838; experiments/sampled_connected_components.py:653-669, 719-723
# ---------------------------------------------------------------------------------------
@dataclass
class SamplingResult:  # common.py:262-265
    spins: np.ndarray
    weights: Optional[np.ndarray]


def sample_indices_device(ground_state: torch.Tensor, uniform: torch.Tensor, sampled_power: float = 2) -> torch.Tensor:
    """Index draws of ``np.random.choice(n, m, replace=True, p=|psi|^power/sum)`` for the given
    uniform numbers: cumulative sum, normalisation and the m searches run on the device."""
    dev = require_cuda()
    out = torch.empty(uniform.shape[0], dtype=torch.int64, device=dev)
    check(lib().asp_sample_indices(ground_state.shape[0], ptr(ground_state, "double *"), float(sampled_power), uniform.shape[0],
                                   ptr(uniform, "double *"), ptr(out, "int64_t *"), ffi.NULL, stream()))
    return out


def monte_carlo_sampling(states, ground_state, number_samples: int, sampled_power: float = 2) -> SamplingResult:
    """common.py:268-278.  The uniform numbers come from numpy's GLOBAL legacy stream exactly as
    ``np.random.choice`` would draw them (``random_sample(number_samples)``), so after
    ``np.random.seed(k)`` the sampled states are the reference's."""
    dev = require_cuda()
    states = np.asarray(states)
    psi = torch.from_numpy(np.ascontiguousarray(ground_state, dtype=np.float64)).to(dev)
    if psi.ndim != 1 or psi.shape[0] != len(states):
        raise ValueError("'states' and 'ground_state' must be one-dimensional and of equal length")
    uniform = torch.from_numpy(np.random.random_sample(int(number_samples))).to(dev)
    indices = sample_indices_device(psi, uniform, sampled_power).cpu().numpy()
    return SamplingResult(spins=states[indices], weights=None)


def determine_exact_solution(spins, quantum_hamiltonian, ground_state):  # common.py:282-285
    indices = quantum_hamiltonian.basis.batched_index(spins)
    psi = np.asarray(ground_state)[indices]
    return sa.signs_to_bits(np.sign(psi))


def create_small_cluster_around_point(s0: int, hamiltonian, required_size: int = 20, keep_probability: float = 0.5):
    """Random breadth-first growth of a cluster around ``s0`` (reference: common.py:481-513).  Sequential by
    construction: ONE ``np.random.rand()`` is consumed per neighbour that is not yet a member, in the order in
    which the operator lists the neighbours, and the frontier of the next generation is a Python ``set`` -- both
    are part of the behaviour (a seeded run must grow the reference's cluster), so they are kept exactly."""
    if hamiltonian.basis.number_spins > 64:
        raise AssertionError("only works with up to 64 spins")
    members = {int(s0)}

    def kept_neighbours(state):
        images, _ = hamiltonian.apply(state)
        images = images[:, 0] if images.ndim > 1 else images
        kept = []
        for image in images:
            if image not in members and np.random.rand() <= keep_probability:
                kept.append(int(image))
        return kept

    frontier = kept_neighbours(int(s0))
    while frontier and len(members) < required_size:
        upcoming = set()
        for state in frontier:
            members.add(state)
            if len(members) >= required_size:
                break
            upcoming |= set(kept_neighbours(state))
        frontier = upcoming
    return sorted(members)


def ground_state_to_log_coeff_fn(ground_state: np.ndarray, basis: ls.SpinBasis):
    """Closure ``spins -> log|psi| + i*pi*[psi < 0]`` over a ground state given on the full basis (reference:
    common.py:806-823); the lookup of the states is the device search ``asp_batched_index``."""
    psi = np.ascontiguousarray(ground_state, dtype=np.float64)
    if psi.ndim != 1:
        raise AssertionError("'ground_state' must be a vector")
    with np.errstate(divide="ignore"):
        table = np.log(np.abs(psi)) + 1j * np.where(psi >= 0, 0.0, np.pi)

    def log_coeff_fn(spins: np.ndarray) -> np.ndarray:
        words = np.asarray(spins, dtype=np.uint64, order="C")
        if words.ndim > 1:  # the [n, 8] form: only word 0 is used below 65 spins
            words = words[:, 0]
        return table[ls.batched_index(basis, words)]

    return log_coeff_fn


def add_noise_to_amplitudes(ground_state, eps: float):
    """|psi_i| -> |psi_i| * exp(eps * u_i), u_i uniform in (-1, 1) from numpy's global stream, signs kept,
    renormalised (reference: common.py:826-838)."""
    psi = np.ascontiguousarray(ground_state, dtype=np.float64)
    if psi.ndim != 1:
        raise AssertionError("'ground_state' must be a vector")
    kick = eps * 2 * (np.random.rand(psi.size) - 0.5)
    noisy = np.sign(psi) * np.exp(np.log(np.abs(psi)) + kick)
    return noisy / np.linalg.norm(noisy)


def amplitude_overlap(cluster, ground_state, noisy_ground_state, basis):  # sampled_connected_components.py:719-723
    indices = basis.batched_index(cluster)
    a = np.abs(np.asarray(ground_state)[indices])
    b = np.abs(np.asarray(noisy_ground_state)[indices])
    return np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b)
