"""numpy/scipy restatement of the reference's LIVE extraction path (TEST INFRASTRUCTURE).

Follows annealing_sign_problem/common.py:
  make_ising_model                         :131-208
  _batched_apply                           :85-106
  _clipped_search_sorted                   :116-128  (searchsorted left, clip to [0, n-1])
  membership mask                          :173
  psi = Re exp(log psi), psi /= ||psi||    :177-181
  _make_ising_model_compute_elements       :71-82    (association (c*|psi_j|)*|psi_i|)
  csr -> 0.5*(M + M^T) -> sort -> COO      :193-196
  compute_accuracy_and_overlap             :211-229
  binary_search                            :544-548
Pinned against the reference file itself (imported with stub third-party modules) by
tests/golden/make_golden.py -> tests/golden/*.npz -> tests/test_oracle.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import scipy.sparse


def signs_to_bits(signs) -> np.ndarray:
    """Packed LSB-first bit vector, bit 1 <=> sign > 0 (cbits/build_matrix.c:67-76; the
    third-party ``sa.signs_to_bits`` is used the same way at common.py:205)."""
    signs = np.asarray(signs)
    n = signs.shape[0]
    bits = np.zeros((n + 63) // 64, dtype=np.uint64)
    idx = np.nonzero(signs > 0)[0]
    np.bitwise_or.at(bits, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    return bits


def bits_to_signs(bits, count: int) -> np.ndarray:
    """+1/-1 float array from packed bits: ``(bits[i//64] >> (i%64)) & 1``
    (annealing_sign_problem/train.py:247-252, square_4x4.py:180-181)."""
    bits = np.asarray(bits, dtype=np.uint64)
    i = np.arange(count, dtype=np.uint64)
    b = (bits[(i // np.uint64(64)).astype(np.int64)] >> (i % np.uint64(64))) & np.uint64(1)
    return 2.0 * b.astype(np.float64) - 1.0


@dataclass
class OracleIsingModel:
    spins: np.ndarray
    exchange: scipy.sparse.coo_matrix
    field: np.ndarray
    initial_signs: np.ndarray
    psi: np.ndarray


def make_ising_model(spins, quantum_hamiltonian, log_psi=None, log_psi_fn=None) -> OracleIsingModel:
    """common.py:131-208 (external_field=False branch; the True branch is ``assert False``)."""
    if log_psi is None and log_psi_fn is None:
        raise ValueError("at least one of log_psi or log_psi_fn should be specified")
    spins = np.asarray(spins, dtype=np.uint64)
    if spins.ndim == 2:
        spins = spins[:, 0]
    spins, first, counts = np.unique(spins, return_index=True, return_counts=True)
    if log_psi is not None and np.any(counts != 1):
        log_psi = np.asarray(log_psi)[first]
    if log_psi is None:
        log_psi = log_psi_fn(spins)
    n = spins.shape[0]

    # :85-106 -- chunked batched_apply, real part, column 0
    out_s, out_c, out_k = [], [], []
    for start in range(0, n, 10000):
        x = np.zeros((min(start + 10000, n) - start, 8), dtype=np.uint64)
        x[:, 0] = spins[start:start + 10000]
        s, c, k = quantum_hamiltonian.batched_apply(x)
        if not np.allclose(c.imag, 0, atol=1e-6):
            raise ValueError("expected all Hamiltonian matrix elements to be real")
        out_s.append(np.ascontiguousarray(s[:, 0]))
        out_c.append(np.ascontiguousarray(c.real))
        out_k.append(k)
    other_spins = np.hstack(out_s)
    other_coeffs = np.hstack(out_c)
    other_counts = np.hstack(out_k)

    # :116-128, :173
    idx = np.clip(np.searchsorted(spins, other_spins), 0, n - 1)
    belong = other_spins == spins[idx]

    # :177-181
    psi = np.exp(np.asarray(log_psi), dtype=np.complex128)
    if not np.allclose(psi.imag, 0, atol=1e-6):
        raise ValueError("expected all wavefunction coefficients to be real")
    psi = np.ascontiguousarray(psi.real)
    psi /= np.linalg.norm(psi)

    # :71-82
    other_psi = np.where(belong, psi[idx], 0)
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(other_counts)
    elements = other_coeffs * np.abs(other_psi)
    elements *= np.repeat(np.abs(psi), other_counts)

    # :193-196
    matrix = scipy.sparse.csr_matrix((elements, idx, offsets), shape=(n, n))
    matrix = 0.5 * (matrix + matrix.T)
    matrix.sort_indices()
    matrix = matrix.tocoo()
    field = np.zeros(n, dtype=np.float64)
    return OracleIsingModel(spins, matrix, field, signs_to_bits(np.sign(psi)), psi)


def compute_accuracy_and_overlap(predicted, exact, weights=None, number_spins=None) -> Tuple[float, float]:
    """common.py:211-229."""
    if weights is None and number_spins is None:
        raise ValueError("'weights' and 'number_spins' cannot be both None")
    if number_spins is None:
        number_spins = len(weights)
    if weights is None:
        weights = np.ones(number_spins, dtype=np.float64)
    p = bits_to_signs(predicted, number_spins)
    e = bits_to_signs(exact, number_spins)
    accuracy = np.mean(e == p)
    accuracy = max(accuracy, 1 - accuracy)
    overlap = abs(np.dot(e * p, weights / np.sum(weights)))
    return float(accuracy), float(overlap)


def binary_search(haystack, needles):
    """common.py:544-548."""
    assert np.all(np.sort(haystack) == haystack)
    indices = np.searchsorted(haystack, needles)
    assert np.all(haystack[indices] == needles)
    return indices


def default_betas(indptr, indices, data, field, number_sweeps: int, beta0=None, beta1=None):
    """Geometric inverse-temperature ladder (OUR definition -- the reference leaves
    beta0/beta1 to the absent annealer, common.py:242-248).  Hot end: the largest
    single-flip barrier is accepted with probability 1/2; cold end: the smallest non-zero
    coupling barrier with probability 1/100 (capped at 1e4 x hot), then number_sweeps//50
    zero-temperature sweeps (beta = inf)."""
    indptr = np.asarray(indptr)
    n = indptr.shape[0] - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    off = rows != np.asarray(indices)
    a = np.abs(np.asarray(data))[off]
    f = np.zeros(n) if field is None else np.abs(np.asarray(field))
    row_sum = np.bincount(rows[off], weights=a, minlength=n)
    max_de = float(np.max(4.0 * row_sum + 2.0 * f)) if n else 1.0
    nz = a[a > 0]
    fz = f[f > 0]
    cands = []
    if nz.size:
        cands.append(4.0 * float(nz.min()))
    if fz.size:
        cands.append(2.0 * float(fz.min()))
    min_de = min(cands) if cands else 1.0
    if max_de <= 0:
        max_de = 1.0
    b0 = np.log(2.0) / max_de if beta0 is None else float(beta0)
    # cold end: the smallest barrier, but at most 4 decades above the hot end -- amplitudes
    # span many decades and a ladder reaching 1/min|J| would spend its sweeps frozen
    b1 = min(np.log(100.0) / min_de, b0 * 1e4) if beta1 is None else float(beta1)
    # final zero-temperature sweeps: a fiftieth of the run, but at least 8 (the spins of tiny amplitude only settle
    # there; with 2 such sweeps a 100-sweep run of kagome_16 never reached accuracy > 0.995) and at most a quarter
    quench = min(max(number_sweeps // 50, 8), number_sweeps // 4) if beta1 is None else 0
    ladder = number_sweeps - quench
    if ladder <= 1:
        betas = np.full(number_sweeps, b1, dtype=np.float64)
    else:
        t = np.arange(ladder, dtype=np.float64) / (ladder - 1)
        betas = np.concatenate([b0 * (b1 / b0) ** t, np.full(quench, np.inf)])
    return np.ascontiguousarray(betas, dtype=np.float64)


# ---------------------------------------------------------------------------------------
# Sampling front-end, pinned by tests/golden/n3_*.npz (the reference's own
# functions run in the build container, tests/golden/make_golden.py:sampling_case)
# ---------------------------------------------------------------------------------------
def sample_indices(ground_state, uniform, sampled_power: float = 2) -> np.ndarray:
    """monte_carlo_sampling, annealing_sign_problem/common.py:274-277, with the uniform numbers made
    explicit: legacy ``np.random.choice(n, m, replace=True, p)`` is ``cdf = p.cumsum(); cdf /=
    cdf[-1]; cdf.searchsorted(random_sample(m), side="right")``."""
    p = np.abs(np.asarray(ground_state, dtype=np.float64)) ** sampled_power
    p /= np.sum(p)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(uniform, dtype=np.float64), side="right")


def batched_index(states, spins) -> np.ndarray:
    """``basis.batched_index`` on a sorted basis (call sites common.py:283, :817): position of every
    state, ValueError when one is absent."""
    states = np.asarray(states, dtype=np.uint64)
    spins = np.asarray(spins, dtype=np.uint64)
    idx = np.searchsorted(states, spins)
    if np.any(idx >= states.shape[0]) or np.any(states[np.minimum(idx, states.shape[0] - 1)] != spins):
        raise ValueError("state not in the basis")
    return idx


def log_coeff(ground_state, states, spins) -> np.ndarray:
    """ground_state_to_log_coeff_fn, common.py:806-823."""
    ground_state = np.asarray(ground_state, dtype=np.float64)
    idx = batched_index(states, spins)
    return np.log(np.abs(ground_state))[idx] + 1j * np.where(ground_state >= 0, 0, np.pi)[idx]
