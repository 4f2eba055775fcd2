"""Generate golden vectors by running THE REFERENCE ITSELF in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; not run on the GPU box)

Two reference implementations are exercised, both unmodified:

* the live path ``annealing_sign_problem/common.py:make_ising_model`` (:131-208), imported
  by file path with stub modules for its absent third-party imports (h5py,
  lattice_symmetries, ising_glass_annealer) -- the stubs provide only a container class
  ``Hamiltonian(exchange, field)`` and a ``signs_to_bits`` following
  cbits/build_matrix.c:67-76; the operator is the oracle's duck-typed ``OperatorNP``
  (the reference only calls ``.basis.number_spins`` and ``.batched_apply``);
* the legacy C path ``cbits/build_matrix.c`` compiled where it lies (oracle/_ref).

Outputs ``tests/golden/*.npz`` (inputs + outputs, small) which the CPU tests use to pin
the oracle and the GPU tests use to check the CUDA path.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import capi  # noqa: E402
from oracle.live_path import bits_to_signs, signs_to_bits  # noqa: E402
from oracle.operator_np import OperatorNP, ground_state, system_path  # noqa: E402


def import_reference_common():
    class Hamiltonian:
        def __init__(self, exchange, field):
            self.exchange, self.field = exchange, field
            self.shape = exchange.shape

    sa = types.ModuleType("ising_glass_annealer")
    sa.Hamiltonian = Hamiltonian
    sa.signs_to_bits = signs_to_bits
    sa.bits_to_signs = bits_to_signs
    ls = types.ModuleType("lattice_symmetries")
    ls.Operator = object
    ls.SpinBasis = object
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.modules["ising_glass_annealer"] = sa
    sys.modules["lattice_symmetries"] = ls
    spec = importlib.util.spec_from_file_location(
        "reference_common", "/root/reference/annealing_sign_problem/common.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cluster_closed_subset(op, n_target, rng):
    """Seeds + one batched_apply shell, truncated (mirrors make_hamiltonian_extension,
    common.py:516-522): gives hit rates of tens of percent."""
    states = op.basis.states
    seeds = rng.choice(states, size=max(4, n_target // 8), replace=False)
    shell, _, _ = op.apply_u64(np.sort(seeds))
    pool = np.unique(np.concatenate([seeds, shell]))
    if pool.shape[0] > n_target:
        pool = np.sort(rng.choice(pool, size=n_target, replace=False))
    return pool


def synthetic_log_psi(n, rng, sigma=2.0):
    """psi_i = +-exp(sigma z_i), as log|psi| + i*pi*[psi<0] (common.py:791-803 format)."""
    z = rng.standard_normal(n)
    neg = rng.random(n) < 0.5
    return sigma * z + 1j * np.pi * neg


def one_case(ref_common, name, n_target, seed, out):
    rng = np.random.default_rng(seed)
    op = OperatorNP.load(system_path(name))
    spins = cluster_closed_subset(op, n_target, rng)
    log_psi = synthetic_log_psi(spins.shape[0], rng)

    model = ref_common.make_ising_model(spins, op, log_psi=log_psi)
    coo = model.ising_hamiltonian.exchange
    assert np.array_equal(model.spins, spins)

    # legacy C path on the same candidates, amplitudes known for ALL candidates
    psi = np.exp(log_psi).real
    psi = psi / np.linalg.norm(psi)
    other_spins, other_coeffs, other_counts = op.apply_u64(spins)
    # amplitudes outside the set: deterministic pseudo-amplitudes keyed on the state
    idx = np.clip(np.searchsorted(spins, other_spins), 0, spins.shape[0] - 1)
    inside = spins[idx] == other_spins
    h = (other_spins * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
    outside_psi = (h.astype(np.float64) / 2.0 ** 24 - 0.5) * 1e-2
    other_psi = np.where(inside, psi[idx], outside_psi)
    counts = np.ones(spins.shape[0], dtype=np.int64)
    rows, cols, vals, field = capi.build_matrix(
        spins, counts, psi, other_spins, other_coeffs, other_counts, other_psi, impl="ref")
    signs = capi.extract_signs(psi, impl="ref")

    np.savez_compressed(
        out,
        system=name,
        spins=spins,
        log_psi=log_psi,
        live_row=coo.row.astype(np.int32),
        live_col=coo.col.astype(np.int32),
        live_data=coo.data,
        live_x0=model.initial_signs,
        other_spins=other_spins,
        other_coeffs=other_coeffs,
        other_counts=other_counts,
        other_psi=other_psi,
        c_rows=rows,
        c_cols=cols,
        c_vals=vals,
        c_field=field,
        c_signs=signs,
    )
    print("%-28s n=%d T=%d hits=%d live_nnz=%d -> %s (%d bytes)" % (
        name, spins.shape[0], other_spins.shape[0], rows.shape[0], coo.nnz,
        os.path.basename(out), os.path.getsize(out)))


def hashed_log_psi(spins):
    """Deterministic log-amplitude of a basis state (stands in for the noisy ground-state lookup of
    sampled_connected_components.py:699-722): log|psi| + i*pi*[psi < 0], a function of the state only."""
    h = (np.asarray(spins, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)
    u = h.astype(np.float64) / 2.0 ** 53
    v = ((np.asarray(spins, dtype=np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(63)).astype(np.float64)
    return 4.0 * (u - 0.5) + 1j * np.pi * v


def extension_case(ref_common, name, n_target, seed, reltol, out):
    """Row N2 of the scope table, produced by the reference itself: make_hamiltonian_extension
    (common.py:516-522), get_strongest_off_diag (common.py:539-541) and
    sparsify_using_global_cutoff (common.py:634-692) on a sampled cluster."""
    import scipy.sparse
    from scipy.sparse.csgraph import connected_components

    rng = np.random.default_rng(seed)
    op = OperatorNP.load(system_path(name))
    cluster = cluster_closed_subset(op, n_target, rng)
    model0 = ref_common.make_ising_model(cluster, op, log_psi_fn=hashed_log_psi)
    model1 = ref_common.make_hamiltonian_extension(model0, hashed_log_psi)
    strongest = ref_common.get_strongest_off_diag(model1.ising_hamiltonian.exchange)
    # the reference slices `exchange[mask][:, mask]` (common.py:671): its annealer's Hamiltonian must hold
    # a sliceable matrix; give the stub's the CSR form of the same COO matrix
    model1.ising_hamiltonian.exchange = model1.ising_hamiltonian.exchange.tocsr()
    # frozen spins: members of the original cluster that stay connected under the cutoff
    m = model1.ising_hamiltonian.exchange.tocsr().copy()
    big = np.abs(m.data).max()
    m.data[np.abs(m.data) < reltol * big] = 0
    m.eliminate_zeros()
    _, comp = connected_components(m, directed=False)
    idx = np.searchsorted(model1.spins, cluster)
    magic = np.bincount(comp[idx]).argmax()
    frozen = cluster[comp[idx] == magic][:40]
    model2 = ref_common.sparsify_using_global_cutoff(model1, reltol, frozen)
    c1 = model1.ising_hamiltonian.exchange.tocoo()
    c2 = model2.ising_hamiltonian.exchange.tocoo()
    np.savez_compressed(
        out, system=name, cluster=cluster, reltol=reltol, frozen=frozen,
        ext_spins=model1.spins, ext_row=c1.row.astype(np.int32), ext_col=c1.col.astype(np.int32), ext_data=c1.data,
        ext_x0=model1.initial_signs, strongest=strongest,
        sp_spins=model2.spins, sp_row=c2.row.astype(np.int32), sp_col=c2.col.astype(np.int32), sp_data=c2.data,
        sp_x0=model2.initial_signs, sp_field=model2.ising_hamiltonian.field)
    print("%-28s cluster=%d extension=%d (nnz %d) sparsified=%d (nnz %d) frozen=%d -> %s (%d bytes)" % (
        name, cluster.shape[0], model1.spins.shape[0], c1.nnz, model2.spins.shape[0], c2.nnz, frozen.shape[0],
        os.path.basename(out), os.path.getsize(out)))


def sampling_case(ref_common, name, seed, out):
    """Row N3 of the scope table, produced by the reference's own functions: monte_carlo_sampling
    (common.py:268-278), ground_state_to_log_coeff_fn (:806-823), determine_exact_solution (:282-285),
    add_noise_to_amplitudes (:826-838), create_small_cluster_around_point (:481-513).  The stubs stand
    in for lattice_symmetries only: ``ls.batched_index`` = numpy searchsorted on the sorted basis;
    ``hamiltonian.apply(s)`` = the oracle operator's neighbours of s in ASCENDING key order (the
    library's own order is unpinned; ascending is what this repo's operator emits)."""
    rng = np.random.default_rng(seed)
    op = OperatorNP.load(system_path(name))
    states = op.basis.states
    n = states.shape[0]
    z = rng.standard_normal(n)
    psi = np.where(rng.random(n) < 0.5, -1.0, 1.0) * np.exp(2.0 * z)
    psi /= np.linalg.norm(psi)

    class Basis:
        number_spins = op.basis.number_spins

        def batched_index(self, spins):
            return np.searchsorted(states, np.asarray(spins, dtype=np.uint64))

    class Hamiltonian:
        basis = Basis()

        def apply(self, s):
            xs, cs, _ = op.apply_u64(np.array([s], dtype=np.uint64))
            order = np.argsort(xs, kind="stable")
            return xs[order], cs[order]

    sys.modules["lattice_symmetries"].batched_index = lambda basis, spins: basis.batched_index(spins)
    np.random.seed(seed)
    mc2 = ref_common.monte_carlo_sampling(states, psi, 3000, 2).spins
    mc1 = ref_common.monte_carlo_sampling(states, psi, 1000, 1).spins
    fn = ref_common.ground_state_to_log_coeff_fn(psi, Basis())
    log_coeff = fn(mc2[:500])
    exact_bits = ref_common.determine_exact_solution(mc2, Hamiltonian(), psi)
    np.random.seed(seed + 1)
    noisy = ref_common.add_noise_to_amplitudes(psi, 0.3)
    np.random.seed(seed + 2)
    clusters = [np.asarray(ref_common.create_small_cluster_around_point(int(s0), Hamiltonian(), required_size=size, keep_probability=0.5),
                           dtype=np.uint64) for s0, size in zip(mc2[:3], [20, 77, 300])]
    np.savez_compressed(out, system=name, seed=seed, psi=psi, mc2=mc2, mc1=mc1, log_coeff=log_coeff, exact_bits=exact_bits,
                        noisy_stride=noisy[::37], cluster0=clusters[0], cluster1=clusters[1], cluster2=clusters[2])
    print("%-28s n=%d samples=%d/%d clusters=%s -> %s (%d bytes)" % (
        name, n, mc2.shape[0], mc1.shape[0], [c.shape[0] for c in clusters], os.path.basename(out), os.path.getsize(out)))


def known_answers(out):
    """Full-basis known-answer table, recomputed by the oracle's ED."""
    table = {}
    for name in ["j1j2_square_4x4", "heisenberg_kagome_16", "heisenberg_kagome_18",
                 "sk_16_1", "sk_16_2", "sk_16_3"]:
        op = OperatorNP.load(system_path(name))
        n = op.basis.number_states
        s, c, k = op.apply_u64(op.basis.states)
        e0, psi, e1 = ground_state(op)
        table[name] = {"n": n, "T": int(s.shape[0]), "E0": e0, "gap": e1 - e0}
        print(name, table[name])
    with open(out, "w") as f:
        json.dump(table, f, indent=1)


def main():
    if not os.path.isdir("/root/reference"):
        sys.exit("needs /root/reference (build container only)")
    capi.build()
    ref_common = import_reference_common()
    one_case(ref_common, "j1j2_square_4x4", 700, 1, os.path.join(HERE, "live_j1j2_square_4x4.npz"))
    one_case(ref_common, "heisenberg_kagome_18", 900, 2, os.path.join(HERE, "live_heisenberg_kagome_18.npz"))
    one_case(ref_common, "sk_16_1", 300, 3, os.path.join(HERE, "live_sk_16_1.npz"))
    known_answers(os.path.join(HERE, "known_answers.json"))
    extension_case(ref_common, "heisenberg_kagome_16", 120, 4, 2e-2, os.path.join(HERE, "n2_heisenberg_kagome_16.npz"))
    extension_case(ref_common, "j1j2_square_4x4", 80, 5, 5e-2, os.path.join(HERE, "n2_j1j2_square_4x4.npz"))
    extension_case(ref_common, "heisenberg_kagome_18", 100, 7, 2e-2, os.path.join(HERE, "n2_heisenberg_kagome_18.npz"))  # spin-inversion basis


if __name__ == "__main__":
    main()
    sampling_case(ref_common, "heisenberg_kagome_16", 6, os.path.join(HERE, "n3_heisenberg_kagome_16.npz"))
