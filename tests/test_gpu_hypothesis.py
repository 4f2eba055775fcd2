"""-m gpu property tests (hypothesis): make_ising_model on the CUDA path against the numpy restatement
of the reference's live path (oracle/live_path.py, itself pinned bitwise by golden vectors) on RANDOM
operators and RANDOM sampled subsets -- KAT-4 of SURVEY.md 8c widened from the shipped lattices to
arbitrary bond lists: 2..40 spins, with and without a Hamming-weight sector, repeated bonds (duplicate
candidates that must be summed), single-spin-flip matrix elements, couplings that cancel to exactly
zero (dropped by scipy's 0.5 (M + M^T), common.py:194), zero amplitudes, empty and one-state subsets."""
import numpy as np
import pytest
import scipy.sparse

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)
hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

import annealing_sign_problem_b200 as asp  # noqa: E402
from oracle import live_path  # noqa: E402
from oracle.operator_np import OperatorNP  # noqa: E402

from _strategies import problems, random_subset as _random_subset  # noqa: E402


@settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(problems())
def test_make_ising_model_equals_the_live_path_restatement_on_random_problems(problem):
    cfg, seed, m = problem
    spins, psi = _random_subset(cfg, seed, m)
    if spins.shape[0] == 0 or not np.any(psi):
        return  # the reference divides by the norm of psi (common.py:181)
    with np.errstate(divide="ignore"):
        log_psi = np.log(psi.astype(np.complex128))
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
    ours = asp.make_ising_model(spins, op, log_psi=log_psi)
    ref = live_path.make_ising_model(spins, OperatorNP.from_config(cfg), log_psi=log_psi)
    n = spins.shape[0]
    assert np.array_equal(ours.spins, ref.spins) and np.array_equal(ours.initial_signs, ref.initial_signs)
    a = ours.ising_hamiltonian.exchange.tocsr()
    b = ref.exchange.tocsr()
    a.sort_indices()
    b.sort_indices()
    assert a.shape == b.shape == (n, n)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    np.testing.assert_allclose(a.data, b.data, rtol=1e-12, atol=0)
    assert isinstance(ours.ising_hamiltonian.exchange, scipy.sparse.coo_matrix)


@st.composite
def legacy_inputs(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    n = draw(st.sampled_from([0, 1, 2, 7, 64, 65, 300]))
    wide = draw(st.booleans())          # keys that use more than words[0] of the 512-bit struct
    max_count = draw(st.sampled_from([0, 1, 3, 40]))
    hit_rate = draw(st.sampled_from([0.0, 0.3, 1.0]))
    return seed, n, wide, max_count, hit_rate


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(legacy_inputs())
def test_legacy_build_matrix_dropin_equals_the_reference_c_on_random_inputs(oracle_capi, inputs):
    """asp_build_matrix (the reference's argument list: host pointers, 512-bit keys, multiplicities, amplitudes of
    candidates outside the set, caller-sized outputs) against cbits/build_matrix.c compiled where it lies
    (oracle/_ref; the C restatement when that is absent): rows, columns, values and field bit for bit."""
    from annealing_sign_problem_b200._lib import ffi, lib

    seed, n, wide, max_count, hit_rate = inputs
    rng = np.random.default_rng(seed)
    keys = np.zeros((n, 8), dtype=np.uint64)
    keys[:, 0] = rng.integers(0, 1 << 40, size=n, dtype=np.uint64)
    if wide:
        keys[:, 1] = rng.integers(0, 3, size=n, dtype=np.uint64)
        keys[:, 7] = rng.integers(0, 2, size=n, dtype=np.uint64)
    # sorted ascending by ls_bits512_cmp (build_matrix.c:7-20): the comparison starts at words[0]
    keys = np.unique(keys, axis=0)
    order = np.lexsort([keys[:, w] for w in range(7, -1, -1)])  # primary key = words[0]
    spins = np.ascontiguousarray(keys[order])
    n = spins.shape[0]
    other_counts = rng.integers(0, max_count + 1, size=n).astype(np.int64)
    T = int(other_counts.sum())
    inside = rng.random(T) < hit_rate
    other = np.zeros((T, 8), dtype=np.uint64)
    if n:
        other[inside] = spins[rng.integers(0, n, size=int(inside.sum()))]
    other[~inside, 0] = rng.integers(1 << 41, 1 << 42, size=int((~inside).sum()), dtype=np.uint64)
    counts = rng.integers(1, 4, size=n).astype(np.int64)
    psi = rng.standard_normal(n)
    other_coeffs = rng.standard_normal(T)
    other_psi = rng.standard_normal(T)
    impl = "ref" if oracle_capi.have_ref() else "port"
    ref_rows, ref_cols, ref_vals, ref_field = oracle_capi.build_matrix(spins, counts, psi, other, other_coeffs, other_counts, other_psi, impl=impl)
    rows = np.zeros(max(T, 1), dtype=np.uint32)
    cols = np.zeros(max(T, 1), dtype=np.uint32)
    vals = np.zeros(max(T, 1), dtype=np.float64)
    field = np.full(max(n, 1), 7.0)
    c = lambda a, t: ffi.cast(t, a.ctypes.data)  # noqa: E731
    nnz = lib().asp_build_matrix(n, c(spins, "asp_bits512 *"), c(counts, "int64_t *"), c(psi, "double *"), c(other, "asp_bits512 *"),
                                 c(other_coeffs, "double *"), c(other_counts, "int64_t *"), c(other_psi, "double *"),
                                 c(rows, "uint32_t *"), c(cols, "uint32_t *"), c(vals, "double *"), c(field, "double *"))
    assert nnz != 2 ** 64 - 1, ffi.string(lib().asp_last_error()).decode()
    assert nnz == ref_rows.shape[0]
    assert np.array_equal(rows[:nnz], ref_rows) and np.array_equal(cols[:nnz], ref_cols)
    assert np.array_equal(vals[:nnz], ref_vals)
    assert np.array_equal(field[:n], ref_field)
