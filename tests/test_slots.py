"""CPU checks of the host-side tables of the single-pass extraction kernel (csrc/extract_fused.cu):
the operator's |delta|-sorted slots and the index word {first position, presence bits}.

A numpy emulation walks the slots exactly as a lane of the kernel does (same applicability test,
same linear bit positions, same lower-bound guess from the order bits, same placement "outward
from the diagonal") and must reproduce the oracle's canonical CSR -- the reference's
cbits/build_matrix.c output, sorted and merged -- bit for bit.  No GPU needed: an
operator can be created without a device.
"""
import numpy as np
import pytest

import annealing_sign_problem_b200 as asp
from annealing_sign_problem_b200._lib import ffi, lib
from oracle import capi
from oracle.operator_np import OperatorNP

ABSENT = np.uint64(0xFFFFFFFFFFFFFFFF)


def _slots(op):
    n = int(lib().asp_debug_operator_slots(op.handle, 0, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL))
    arrs = [np.zeros(n, dtype=np.uint64) for _ in range(4)] + [np.zeros(n, dtype=np.float64) for _ in range(2)]
    got = lib().asp_debug_operator_slots(op.handle, n, *[ffi.cast("uint64_t *", a.ctypes.data) for a in arrs[:4]],
                                         *[ffi.cast("double *", a.ctypes.data) for a in arrs[4:]])
    assert got == n
    return arrs


def _filter_hash(key):
    """csrc/extract_fused.cu: filter_hash (GF(2)-linear)."""
    lo, hi = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    y = lo ^ (((hi << 7) | (hi >> 25)) & 0xFFFFFFFF)
    y ^= y >> 15
    y ^= (y << 11) & 0xFFFFFFFF
    y ^= y >> 7
    y ^= (y << 3) & 0xFFFFFFFF
    y ^= y >> 17
    return y


def _sub(key, oshift):
    return ((key >> oshift) & 15) | ((_filter_hash(key) & 15) << 8)


def _emulate(op, spins, psi, bucket_bits_delta=0):
    """Per row what a lane of extract_csr_kernel computes; returns (indptr, indices, data, false_positives)."""
    flip, mask, need_down, need_up, coef_down, coef_up = _slots(op)
    n = spins.shape[0]
    key_bits = op.basis.number_spins
    lg = 0
    while (1 << lg) < n:
        lg += 1
    bbits = min(max(2, min(lg - 2 + bucket_bits_delta, 27)), key_bits)
    bshift = key_bits - bbits
    oshift = max(bshift - 4, 0)
    nb = 1 << bbits
    keys = [int(k) for k in spins]
    start = np.zeros(nb + 1, dtype=np.int64)
    bits = np.zeros(nb + 1, dtype=np.int64)
    buckets = np.array([k >> bshift for k in keys], dtype=np.int64)
    start[:] = np.searchsorted(buckets, np.arange(nb + 1), side="left")
    for k in keys:
        sub = _sub(k, oshift)
        bits[k >> bshift] |= (1 << (sub & 15)) | (1 << (16 + ((sub >> 8) & 15)))
    amp = np.abs(psi)
    indptr, indices, data = [0], [], []
    false_positives = 0
    for row, s in enumerate(keys):
        s_idx, s_sub = s >> bshift, _sub(s, oshift) | (16 << 8)
        down, up = [], []
        for k in range(flip.shape[0]):
            t = s & int(mask[k])
            if t == int(need_down[k]):
                is_down = True
            elif t == int(need_up[k]):
                is_down = False
            else:
                continue
            f = int(flip[k])
            b = s_idx ^ (f >> bshift)
            sub = s_sub ^ _sub(f, oshift)
            assert b == (s ^ f) >> bshift and (sub & 15) == ((s ^ f) >> oshift) & 15  # linearity of the probe
            word = int(bits[b])
            if not ((word >> (sub & 15)) & (word >> ((sub >> 8) & 31)) & 1):
                continue
            c = s ^ f
            p = int(start[b]) + bin(word & ((1 << (sub & 15)) - 1) & 0xFFFF).count("1")
            assert p <= np.searchsorted(spins, np.uint64(c))  # the guess never overshoots
            while p < int(start[b + 1]) and keys[p] < c:
                p += 1
            if p >= n or keys[p] != c:
                false_positives += 1
                continue
            value = ((coef_down[k] if is_down else coef_up[k]) * amp[p]) * amp[row]
            (down if is_down else up).append((p, value))
        entries = down[::-1] + [(row, None)] + up
        cols = [e[0] for e in entries]
        assert cols == sorted(cols) and len(set(cols)) == len(cols), (row, cols)
        indices += cols
        data += [e[1] for e in entries]
        indptr.append(len(indices))
    return np.array(indptr), np.array(indices), data, false_positives


def _u1(system):
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
    op.config = cfg
    return op


def _subset(op, n, seed):
    """Cluster-closed subset made with the oracle's numpy operator (no GPU)."""
    from oracle import synthetic_np

    op_np = OperatorNP.from_config(op.config)
    spins = synthetic_np.cluster_closed_states(op_np, n, seed)
    return spins, synthetic_np.synthetic_amplitudes(spins.shape[0], seed)


@pytest.mark.parametrize("system,n,delta", [("heisenberg_kagome_16", 600, 0), ("j1j2_square_4x4", 900, 0), ("sk_16_1", 400, 0),
                                            ("heisenberg_kagome_16", 600, -6), ("heisenberg_kagome_36", 500, 0)])
def test_slot_walk_reproduces_the_reference_rows(system, n, delta):
    op = _u1(system)
    spins, psi = _subset(op, n, 3)
    indptr, indices, data, _ = _emulate(op, spins, psi, delta)
    ref_indptr, ref_indices, ref_data = _oracle_csr_u1(op, spins, psi)
    assert np.array_equal(indptr, ref_indptr) and np.array_equal(indices, ref_indices)
    off = np.array([v is not None for v in data])
    ours = np.array([v for v in data if v is not None])
    np.testing.assert_allclose(ours, ref_data[off], rtol=1e-14, atol=0)


def _oracle_csr_u1(op, spins, psi):
    """The reference's C path (oracle port, 64-bit keys) on the U(1)-only operator, canonical CSR."""
    op_np = OperatorNP.from_config(op.config)
    other_spins, other_coeffs, other_counts = op_np.apply_u64(spins)
    idx = np.clip(np.searchsorted(spins, other_spins), 0, spins.shape[0] - 1)
    other_psi = np.where(spins[idx] == other_spins, psi[idx], 0.0)
    counts = np.ones(spins.shape[0], dtype=np.int64)
    rows, cols, vals, _ = capi.build_matrix(spins, counts, psi, other_spins, other_coeffs, other_counts, other_psi, impl="port64")
    return capi.canonical_csr(spins.shape[0], rows, cols, vals)


def test_slots_pair_the_two_directions_of_every_exchange():
    """Heisenberg bonds: one slot per bond, both directions present, needs complementary inside the mask,
    |delta| ascending; SK (every pair of 16 spins): 120 slots."""
    for system, bonds in [("heisenberg_kagome_16", None), ("sk_16_1", 120)]:
        op = _u1(system)
        flip, mask, need_down, need_up, coef_down, coef_up = _slots(op)
        if bonds is not None:
            assert flip.shape[0] == bonds
        assert 2 * flip.shape[0] == op.max_candidates - 1
        assert np.all(flip == mask) and np.all(need_down ^ need_up == mask)
        assert np.all(need_down != ABSENT) and np.all(need_up != ABSENT)
        # the down move removes the higher bit: need_down > need_up as integers, |delta| = difference of the two bits
        assert np.all(need_down > need_up)
        delta = need_down.astype(np.int64) - need_up.astype(np.int64)
        assert np.all(np.diff(delta) >= 0)
        assert np.array_equal(coef_down, coef_up)


def test_index_filter_rejects_most_misses():
    """The presence bits reject most absent candidates.  At this small size (bucket = top 10 of 36 key bits) most
    flips stay inside the row's own bucket AND order bit, so only the hashed bit filters them (the row itself set
    the order bit); at 10^7 states (bucket = top 22 bits) that holds for the ~8 % of the flips below bit 10."""
    op = _u1("heisenberg_kagome_36")
    spins, psi = _subset(op, 4000, 5)
    indptr, indices, _, false_positives = _emulate(op, spins, psi)
    candidates = spins.shape[0] * 37
    assert false_positives < 0.5 * candidates, (false_positives, candidates)
    _, _, _, finer = _emulate(op, spins, psi, bucket_bits_delta=2)
    assert finer < false_positives
