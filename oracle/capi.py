"""ctypes bindings of the CPU oracle libraries (TEST INFRASTRUCTURE ONLY).

* ``oracle/_lib/liboracle.so``          -- extract_port.c + anneal_port.c + greedy_port.c + colour_port.c (our restatements)
* ``oracle/_ref/libref_build_matrix.so`` -- the reference's cbits/build_matrix.c, compiled
  where it lies by ``oracle/Makefile``; exports the two symbols of cbits/build_matrix.h:7-14.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_lib", "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_build_matrix.so")

_p = C.c_void_p
_u64, _i64, _u32, _f64 = C.c_uint64, C.c_int64, C.c_uint32, C.c_double


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_PORT) or (os.path.isdir("/root/reference") and not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _load(path):
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        lib = _load(_PORT)
        lib.oracle_build_matrix.restype = _u64
        lib.oracle_build_matrix.argtypes = [_u64] + [_p] * 11
        lib.oracle_build_matrix_u64.restype = _u64
        lib.oracle_build_matrix_u64.argtypes = [_u64] + [_p] * 11
        lib.oracle_extract_signs.restype = None
        lib.oracle_extract_signs.argtypes = [_u64, _p, _p]
        lib.oracle_coo_to_canonical_csr.restype = _u64
        lib.oracle_coo_to_canonical_csr.argtypes = [_u64, _u64] + [_p] * 6
        lib.oracle_energy.restype = _f64
        lib.oracle_energy.argtypes = [_u64] + [_p] * 5
        lib.oracle_philox4x32_10.restype = None
        lib.oracle_philox4x32_10.argtypes = [_p, _p, _p]
        lib.oracle_exp_neg.restype = _f64
        lib.oracle_exp_neg.argtypes = [_f64]
        lib.oracle_greedy.restype = _u32
        lib.oracle_greedy.argtypes = [_u64, _p, _p, _p, _p, _p]
        lib.oracle_anneal.restype = None
        lib.oracle_anneal.argtypes = [_u64, _p, _p, _p, _p, _u32, _u32, _p, _u64, _p, _f64, _p, _p, _p, _u32]
        lib.oracle_colour.restype = _u32
        lib.oracle_colour.argtypes = [_u64, _p, _p, _p]
        lib.oracle_positions.restype = _u64
        lib.oracle_positions.argtypes = [_u64, _u32, _p, _p, _p, _p]
        lib.oracle_relabel.restype = _u64
        lib.oracle_relabel.argtypes = [_u64, _u64] + [_p] * 10
        _port = lib
    return _port


def have_ref() -> bool:
    if not os.path.exists(_REF) and os.path.isdir("/root/reference"):
        build()
    return os.path.exists(_REF)


def ref():
    """The reference's own compiled C (cbits/build_matrix.h:7-14)."""
    global _ref
    if _ref is None:
        lib = _load(_REF)
        lib.build_matrix.restype = _u64
        lib.build_matrix.argtypes = [_u64] + [_p] * 11
        lib.extract_signs.restype = None
        lib.extract_signs.argtypes = [_u64, _p, _p]
        _ref = lib
    return _ref


def pad512(x: np.ndarray) -> np.ndarray:
    """[n] u64 -> [n,8] u64, word 0 = the key (what common.py:58-68 does)."""
    x = np.asarray(x, dtype=np.uint64)
    out = np.zeros((x.shape[0], 8), dtype=np.uint64)
    out[:, 0] = x
    return out


def build_matrix(spins, counts, psi, other_spins, other_coeffs, other_counts, other_psi, impl="port"):
    """Run build_matrix (reference arg list) -> (rows u32, cols u32, vals f64, field f64).

    impl: "port" (oracle_build_matrix, 512-bit keys), "port64" (64-bit keys) or "ref"
    (the reference's compiled C).
    """
    n = int(spins.shape[0])
    T = int(other_spins.shape[0])
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    other_coeffs = np.ascontiguousarray(other_coeffs, dtype=np.float64)
    other_counts = np.ascontiguousarray(other_counts, dtype=np.int64)
    other_psi = np.ascontiguousarray(other_psi, dtype=np.float64)
    rows = np.zeros(T, dtype=np.uint32)
    cols = np.zeros(T, dtype=np.uint32)
    vals = np.zeros(T, dtype=np.float64)
    field = np.zeros(n, dtype=np.float64)
    if impl == "port64":
        s = np.ascontiguousarray(spins, dtype=np.uint64)
        o = np.ascontiguousarray(other_spins, dtype=np.uint64)
        assert s.ndim == 1 and o.ndim == 1
        fn = port().oracle_build_matrix_u64
    else:
        s = pad512(spins) if spins.ndim == 1 else np.ascontiguousarray(spins, dtype=np.uint64)
        o = pad512(other_spins) if other_spins.ndim == 1 else np.ascontiguousarray(other_spins, dtype=np.uint64)
        fn = port().oracle_build_matrix if impl == "port" else ref().build_matrix
    nnz = fn(n, _ptr(s), _ptr(counts), _ptr(psi), _ptr(o), _ptr(other_coeffs), _ptr(other_counts),
             _ptr(other_psi), _ptr(rows), _ptr(cols), _ptr(vals), _ptr(field))
    nnz = int(nnz)
    return rows[:nnz], cols[:nnz], vals[:nnz], field


def extract_signs(psi, impl="port"):
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    n = psi.shape[0]
    out = np.full((n + 63) // 64, 0xDEADBEEFDEADBEEF, dtype=np.uint64)  # callee must zero it
    fn = port().oracle_extract_signs if impl == "port" else ref().extract_signs
    fn(n, _ptr(psi), _ptr(out))
    return out


def canonical_csr(n, rows, cols, vals):
    """Row-monotone COO -> (indptr i64[n+1], indices i32, data f64): stable column sort,
    duplicates summed in generation order."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    nnz = rows.shape[0]
    indptr = np.zeros(n + 1, dtype=np.int64)
    oc = np.zeros(max(nnz, 1), dtype=np.int32)
    ov = np.zeros(max(nnz, 1), dtype=np.float64)
    m = port().oracle_coo_to_canonical_csr(n, nnz, _ptr(rows), _ptr(cols), _ptr(vals), _ptr(indptr), _ptr(oc), _ptr(ov))
    return indptr, oc[:m].copy(), ov[:m].copy()


def energy(indptr, indices, data, field, bits) -> float:
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.float64)
    n = indptr.shape[0] - 1
    return float(port().oracle_energy(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(field), _ptr(bits)))


def greedy(indptr, indices, data, field):
    """oracle/greedy_port.c on a CSR model (diagonal ignored) -> (spins int8 [+1/-1], sweeps)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.float64)
    n = indptr.shape[0] - 1
    spin = np.zeros(n, dtype=np.int8)
    sweeps = int(port().oracle_greedy(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(field), _ptr(spin)))
    return spin, sweeps


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    port().oracle_philox4x32_10(_ptr(ctr), _ptr(key), _ptr(out))
    return out


def exp_neg(x: float) -> float:
    return float(port().oracle_exp_neg(float(x)))


def anneal(indptr, indices, data, field, repetitions, betas, seed, x0=None, escale=2.0 ** 40, threads=None):
    """Sequential-order Metropolis SA on the model as given.

    -> (best_bits u64[R, ceil(n/64)], best_rel i64[R], final_rel i64[R])
    """
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    betas = np.ascontiguousarray(betas, dtype=np.float64)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.float64)
    x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.uint64)
    n = indptr.shape[0] - 1
    words = (n + 63) // 64
    R = int(repetitions)
    best = np.zeros((R, words), dtype=np.uint64)
    best_rel = np.zeros(R, dtype=np.int64)
    final_rel = np.zeros(R, dtype=np.int64)
    if threads is None:
        threads = os.cpu_count() or 1
    port().oracle_anneal(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(field), R, betas.shape[0],
                         _ptr(betas), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(x0), float(escale),
                         _ptr(best), _ptr(best_rel), _ptr(final_rel), int(threads))
    return best, best_rel, final_rel


def plan(indptr, indices, data, field):
    """oracle/colour_port.c: the annealing plan of a CSR model (greedy colouring in (hash, index) priority, classes
    padded to multiples of 4, relabelled CSR without the diagonal) -> dict like AnnealPlan.export()."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    field = None if field is None else np.ascontiguousarray(field, dtype=np.float64)
    n = indptr.shape[0] - 1
    colour = np.empty(n, dtype=np.int32)
    classes = int(port().oracle_colour(n, _ptr(indptr), _ptr(indices), _ptr(colour)))
    class_ptr = np.empty(classes + 1, dtype=np.int64)
    n_padded = int(port().oracle_positions(n, classes, _ptr(colour), _ptr(class_ptr), None, None))
    order = np.empty(n_padded, dtype=np.int32)
    position = np.empty(n, dtype=np.int32)
    port().oracle_positions(n, classes, _ptr(colour), _ptr(class_ptr), _ptr(order), _ptr(position))
    out_indptr = np.empty(n_padded + 1, dtype=np.int64)
    out_cols = np.empty(max(indices.shape[0], 1), dtype=np.int32)
    out_vals = np.empty(max(indices.shape[0], 1), dtype=np.float64)
    out_field = np.empty(n_padded, dtype=np.float64)
    nnz = int(port().oracle_relabel(n, n_padded, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(field), _ptr(order), _ptr(position),
                                    _ptr(out_indptr), _ptr(out_cols), _ptr(out_vals), _ptr(out_field)))
    return dict(order=order, position=position, class_ptr=class_ptr, indptr=out_indptr, indices=out_cols[:nnz], data=out_vals[:nnz],
                field=out_field, colour=colour)
