"""-m gpu parity tests of the extraction path: CUDA (through the C ABI) vs the CPU oracle,
vs golden vectors produced by the reference itself, and size-independent properties."""
import json
import os

import numpy as np
import pytest
import scipy.sparse

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from annealing_sign_problem_b200._lib import AspError, ffi, lib  # noqa: E402
from oracle import live_path  # noqa: E402
from oracle.operator_np import OperatorNP  # noqa: E402

GOLDEN = ["live_j1j2_square_4x4.npz", "live_heisenberg_kagome_18.npz", "live_sk_16_1.npz"]
DEV = torch.device("cuda")


def _coo_to_csr(row, col, data, n):
    m = scipy.sparse.coo_matrix((data, (row, col)), shape=(n, n)).tocsr()
    m.sort_indices()
    return m


def _assert_same_matrix(ours: scipy.sparse.spmatrix, ref: scipy.sparse.spmatrix, rtol=1e-12):
    ours, ref = ours.tocsr(), ref.tocsr()
    ours.sort_indices()
    ref.sort_indices()
    assert ours.shape == ref.shape
    assert np.array_equal(ours.indptr, ref.indptr)
    assert np.array_equal(ours.indices, ref.indices)  # bit-exact structure
    np.testing.assert_allclose(ours.data, ref.data, rtol=rtol, atol=0)  # 1e-12 relative (north_star)


@pytest.mark.parametrize("name", GOLDEN)
def test_make_ising_model_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    op = asp.load_hamiltonian(asp.ls.system_path(str(g["system"])))
    model = asp.make_ising_model(g["spins"], op, log_psi=g["log_psi"])
    n = g["spins"].shape[0]
    assert np.array_equal(model.spins, g["spins"])
    ex = model.ising_hamiltonian.exchange
    assert isinstance(ex, scipy.sparse.coo_matrix) and ex.shape == (n, n)
    _assert_same_matrix(ex, _coo_to_csr(g["live_row"], g["live_col"], g["live_data"], n))
    assert np.array_equal(model.initial_signs, g["live_x0"])
    assert not model.ising_hamiltonian.field.any()


@pytest.mark.parametrize("name", GOLDEN)
def test_legacy_build_matrix_dropin_is_bitwise_the_reference_c(golden_dir, name):
    """asp_build_matrix / asp_extract_signs with the reference's argument list (HOST pointers,
    512-bit keys) against outputs of the reference's own compiled C."""
    g = np.load(os.path.join(golden_dir, name))
    spins = g["spins"]
    n, T = spins.shape[0], g["other_spins"].shape[0]
    psi = np.exp(g["log_psi"]).real
    psi = np.ascontiguousarray(psi / np.linalg.norm(psi))
    s512 = np.zeros((n, 8), dtype=np.uint64)
    s512[:, 0] = spins
    o512 = np.zeros((T, 8), dtype=np.uint64)
    o512[:, 0] = g["other_spins"]
    counts = np.ones(n, dtype=np.int64)
    rows = np.zeros(T, dtype=np.uint32)
    cols = np.zeros(T, dtype=np.uint32)
    vals = np.zeros(T, dtype=np.float64)
    field = np.full(n, 7.0)
    other_coeffs = np.ascontiguousarray(g["other_coeffs"])  # keep the buffers alive across the call
    other_counts = np.ascontiguousarray(g["other_counts"])
    other_psi = np.ascontiguousarray(g["other_psi"])
    c = lambda a, t: ffi.cast(t, a.ctypes.data)  # noqa: E731
    nnz = lib().asp_build_matrix(n, c(s512, "asp_bits512 *"), c(counts, "int64_t *"), c(psi, "double *"),
                                 c(o512, "asp_bits512 *"), c(other_coeffs, "double *"), c(other_counts, "int64_t *"),
                                 c(other_psi, "double *"),
                                 c(rows, "uint32_t *"), c(cols, "uint32_t *"), c(vals, "double *"), c(field, "double *"))
    assert nnz == g["c_rows"].shape[0]
    assert np.array_equal(rows[:nnz], g["c_rows"])
    assert np.array_equal(cols[:nnz], g["c_cols"])
    assert np.array_equal(vals[:nnz], g["c_vals"])  # same association, no FMA: bitwise
    assert np.array_equal(field, g["c_field"])
    signs = np.full((n + 63) // 64, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    lib().asp_extract_signs(n, c(psi, "double *"), c(signs, "uint64_t *"))
    assert np.array_equal(signs, g["c_signs"])


def test_reference_binding_computes_through_our_library(golden_dir):
    """The reference's own cdef (annealing_sign_problem/build_extension.py:5-21) dlopen'ed on libasp_b200.so: calling
    build_matrix / extract_signs BY THE REFERENCE'S NAMES reproduces the outputs of the reference's compiled C."""
    from cffi import FFI

    from test_boundary import ROOT, reference_cdef

    rffi = FFI()
    rffi.cdef(reference_cdef())
    rlib = rffi.dlopen(os.path.join(ROOT, "annealing-sign-problem_b200", "libasp_b200.so"))
    g = np.load(os.path.join(golden_dir, GOLDEN[0]))
    spins = g["spins"]
    n, T = spins.shape[0], g["other_spins"].shape[0]
    psi = np.exp(g["log_psi"]).real
    psi = np.ascontiguousarray(psi / np.linalg.norm(psi))
    s512 = np.zeros((n, 8), dtype=np.uint64)
    s512[:, 0] = spins
    o512 = np.zeros((T, 8), dtype=np.uint64)
    o512[:, 0] = g["other_spins"]
    counts = np.ones(n, dtype=np.int64)
    rows, cols = np.zeros(T, dtype=np.uint32), np.zeros(T, dtype=np.uint32)
    vals, field = np.zeros(T, dtype=np.float64), np.full(n, 7.0)
    other_coeffs, other_counts, other_psi = (np.ascontiguousarray(g[k]) for k in ("other_coeffs", "other_counts", "other_psi"))
    c = lambda a, t: rffi.cast(t, a.ctypes.data)  # noqa: E731  (the reference casts the same way)
    nnz = rlib.build_matrix(n, c(s512, "ls_bits512 *"), c(counts, "int64_t *"), c(psi, "double *"), c(o512, "ls_bits512 *"),
                            c(other_coeffs, "double *"), c(other_counts, "int64_t *"), c(other_psi, "double *"),
                            c(rows, "uint32_t *"), c(cols, "uint32_t *"), c(vals, "double *"), c(field, "double *"))
    assert nnz == g["c_rows"].shape[0]
    assert np.array_equal(rows[:nnz], g["c_rows"]) and np.array_equal(cols[:nnz], g["c_cols"])
    assert np.array_equal(vals[:nnz], g["c_vals"]) and np.array_equal(field, g["c_field"])
    signs = np.zeros((n + 63) // 64, dtype=np.uint64)
    rlib.extract_signs(n, c(psi, "double *"), c(signs, "uint64_t *"))
    assert np.array_equal(signs, g["c_signs"])


def test_extract_signs_edge_cases(oracle_capi):
    rng = np.random.default_rng(0)
    for n in [1, 31, 32, 63, 64, 65, 130, 1000, 4097]:
        psi = rng.standard_normal(n)
        if n > 4:
            psi[1], psi[2], psi[3] = 0.0, -0.0, np.nan
        assert np.array_equal(asp.sa.signs_to_bits(psi), oracle_capi.extract_signs(psi))


def _oracle_csr(op_np, spins, psi):
    m = live_path.make_ising_model(spins, op_np, log_psi=np.log(psi.astype(np.complex128)))
    return m.exchange.tocsr()


@pytest.mark.parametrize("system,n,seed", [
    ("j1j2_square_4x4", 3000, 0), ("heisenberg_kagome_16", 5000, 1), ("sk_16_2", 1500, 2),
    ("heisenberg_kagome_18", 4000, 3), ("heisenberg_kagome_36", 20000, 4), ("heisenberg_pyrochlore_2x2x2", 20000, 5),
])
def test_fused_extraction_vs_oracle_on_random_subsets(system, n, seed):
    """U(1)-only operators (symmetries stripped for the 32/36-spin shapes)."""
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    if system in ("heisenberg_kagome_36", "heisenberg_pyrochlore_2x2x2"):
        cfg["basis"]["symmetries"] = []
        cfg["basis"]["spin_inversion"] = None
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
    op_np = OperatorNP.from_config(cfg)
    if basis.is_symmetrised:
        pool = torch.from_numpy(op_np.basis.states.view(np.int64))
        spins = pool[torch.sort(torch.randperm(pool.shape[0], generator=torch.Generator().manual_seed(seed))[:n]).values].to(DEV)
    else:
        spins = synthetic.cluster_closed_states(op, n, seed, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], seed, device=DEV)
    indptr, indices, data = common.extract_csr_device(op, spins, psi)
    n = spins.shape[0]
    ours = scipy.sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n, n))
    assert ours.has_sorted_indices
    ref = _oracle_csr(op_np, spins.cpu().numpy().view(np.uint64), psi.cpu().numpy())
    sym = (0.5 * (ours + ours.T)).tocsr()
    _assert_same_matrix(sym, ref)
    assert ours.nnz > n  # more than the diagonal: the subsets are cluster-closed


def test_foreign_operator_goes_through_its_own_batched_apply(golden_dir):
    g = np.load(os.path.join(golden_dir, GOLDEN[0]))
    foreign = OperatorNP.load(asp.ls.system_path(str(g["system"])))  # only .basis.number_spins + .batched_apply
    model = asp.make_ising_model(g["spins"], foreign, log_psi=g["log_psi"])
    n = g["spins"].shape[0]
    _assert_same_matrix(model.ising_hamiltonian.exchange, _coo_to_csr(g["live_row"], g["live_col"], g["live_data"], n))


def test_input_normalisation_duplicates_unsorted_and_n8(golden_dir):
    g = np.load(os.path.join(golden_dir, GOLDEN[0]))
    op = asp.load_hamiltonian(asp.ls.system_path(str(g["system"])))
    n = g["spins"].shape[0]
    rng = np.random.default_rng(0)
    perm = rng.permutation(n)
    dup = np.concatenate([perm, perm[:50]])
    spins8 = np.zeros((dup.shape[0], 8), dtype=np.uint64)
    spins8[:, 0] = g["spins"][dup]
    model = asp.make_ising_model(spins8, op, log_psi=g["log_psi"][dup])
    assert np.array_equal(model.spins, g["spins"])
    _assert_same_matrix(model.ising_hamiltonian.exchange, _coo_to_csr(g["live_row"], g["live_col"], g["live_data"], n))
    with pytest.raises(ValueError):
        asp.make_ising_model(g["spins"], op)
    with pytest.raises(ValueError):
        asp.make_ising_model(np.zeros((4, 3), dtype=np.uint64), op, log_psi=np.zeros(4))
    with pytest.raises(ValueError):
        asp.make_ising_model(g["spins"], op, log_psi=g["log_psi"] + 0.3j)  # complex amplitudes


def test_tiny_and_empty_inputs():
    op = asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_16"))
    one = torch.tensor([0b0000000011111111], dtype=torch.int64, device=DEV)
    psi = torch.ones(1, dtype=torch.float64, device=DEV)
    indptr, indices, data = common.extract_csr_device(op, one, psi)
    assert indptr.tolist() == [0, 1] and indices.tolist() == [0]
    empty = torch.zeros(0, dtype=torch.int64, device=DEV)
    indptr, indices, data = common.extract_csr_device(op, empty, torch.zeros(0, dtype=torch.float64, device=DEV))
    assert indptr.tolist() == [0] and indices.numel() == 0


def test_full_basis_known_answers(golden_dir, oracle_capi):
    """KAT-1 (E(sign psi_ED) = E0), KAT-3 (nnz incl. one diagonal per row), KAT-6 (J = J^T)."""
    from oracle.operator_np import ground_state

    table = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    for name in ["j1j2_square_4x4", "heisenberg_kagome_18"]:
        op_np = OperatorNP.load(asp.ls.system_path(name))
        e0, psi, _ = ground_state(op_np)
        op = asp.load_hamiltonian(asp.ls.system_path(name))
        assert np.array_equal(op.basis.states, op_np.basis.states)
        with np.errstate(divide="ignore"):
            log_psi = np.log(psi.astype(np.complex128))
        model = asp.make_ising_model(op.basis.states, op, log_psi=log_psi)
        ex = model.ising_hamiltonian.exchange.tocsr()
        if name == "j1j2_square_4x4":
            assert ex.nnz == table[name]["T"] == 452166
        assert np.all(ex.diagonal() != 0) or name != "j1j2_square_4x4"
        assert abs(ex - ex.T).max() == 0.0
        e = model.ising_hamiltonian.energy(model.initial_signs)
        assert abs(e - table[name]["E0"]) < 1e-10
        assert abs(e - oracle_capi.energy(ex.indptr, ex.indices, ex.data, None, model.initial_signs)) < 1e-10


def test_row_block_shards_concatenate_to_the_unsharded_result():
    """KAT-7: what each of G GPUs would build from its row block, bit for bit."""
    op = asp.load_hamiltonian(asp.ls.system_path("sk_16_3"))
    spins = synthetic.cluster_closed_states(op, 6000, 11, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], 11, device=DEV)
    n = spins.shape[0]
    indptr, indices, data = common.extract_csr_device(op, spins, psi)
    for shards in [2, 3, 8]:
        bounds = [n * k // shards for k in range(shards + 1)]
        parts = [common.extract_csr_device(op, spins, psi, bounds[k], bounds[k + 1] - bounds[k]) for k in range(shards)]
        assert torch.equal(torch.cat([p[1] for p in parts]), indices)
        assert torch.equal(torch.cat([p[2] for p in parts]), data)
        offset, glued = 0, [torch.zeros(1, dtype=torch.int64, device=DEV)]
        for p in parts:
            glued.append(p[0][1:] + offset)
            offset += int(p[0][-1])
        assert torch.equal(torch.cat(glued), indptr)


def test_gather_index_kernel_over_emulated_shards_equals_the_unsharded_build():
    """X1 fused with the index build (asp_gather_index): the basis handed over as row blocks of odd,
    even and zero length (here all on one device) must give the same private copy, the same index and
    hence the same CSR, bit for bit, as asp_extract_csr on the whole array."""
    op = _u1_operator("heisenberg_kagome_36")
    spins = synthetic.cluster_closed_states(op, 50_001, 5, DEV)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 5, device=DEV)
    ref = common.extract_csr_device(op, spins, psi)
    for bounds in ([0, n], [0, 1, n], [0, 12_345, 12_345, 30_000, n], [0, 0, 7, 4_096, 20_001, 20_002, n, n]):
        world = len(bounds) - 1
        parts_s = [spins[bounds[q]:bounds[q + 1]].clone() if bounds[q + 1] > bounds[q] else torch.zeros(2, dtype=torch.int64, device=DEV) for q in range(world)]
        parts_p = [psi[bounds[q]:bounds[q + 1]].clone() if bounds[q + 1] > bounds[q] else torch.zeros(2, dtype=torch.float64, device=DEV) for q in range(world)]
        for rank, mode in [(0, 0), (world - 1, 2)]:  # copy only (asp_gather_blocks): copy engines / one TMA thread per CTA
            lib().asp_set_gather_mode(mode)
            full_s = torch.zeros(n, dtype=torch.int64, device=DEV)
            full_p = torch.zeros(n, dtype=torch.float64, device=DEV)
            common.check(lib().asp_gather_blocks(
                world, rank, ffi.new("uint64_t[]", bounds),
                ffi.new("uint64_t const *[]", [common.ptr(t, "uint64_t const *") for t in parts_s]),
                ffi.new("double const *[]", [common.ptr(t, "double const *") for t in parts_p]),
                ffi.NULL, 0, common.ptr(full_s, "uint64_t *"), common.ptr(full_p, "double *"), common.stream()))
            assert torch.equal(full_s, spins) and torch.equal(full_p, psi), (bounds, rank, mode)
        for rank, mode in [(0, 0), (0, 2), (world - 1, 1), (world - 1, 2), (world // 2, 0), (world // 2, 1), (world // 2, 2)]:
            lib().asp_set_gather_mode(mode)  # 0 = copy engines + per-block index kernels, 1 = one SM kernel, 2 = one TMA kernel
            row_begin, num_rows = bounds[rank], bounds[rank + 1] - bounds[rank]
            need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n, num_rows))
            workspace = torch.empty(need, dtype=torch.uint8, device=DEV)
            full_s = torch.zeros(n, dtype=torch.int64, device=DEV)
            full_p = torch.zeros(n, dtype=torch.float64, device=DEV)
            common.check(lib().asp_gather_index(
                op.handle, world, rank, ffi.new("uint64_t[]", bounds),
                ffi.new("uint64_t const *[]", [common.ptr(t, "uint64_t const *") for t in parts_s]),
                ffi.new("double const *[]", [common.ptr(t, "double const *") for t in parts_p]),
                ffi.NULL, 0, common.ptr(full_s, "uint64_t *"), common.ptr(full_p, "double *"), num_rows,
                common.ptr(workspace, "void *"), need, common.stream()))
            assert torch.equal(full_s, spins) and torch.equal(full_p, psi)
            if num_rows == 0:
                continue
            indptr, indices, data = common.extract_csr_indexed_device(op, full_s, full_p, row_begin, num_rows, workspace,
                                                                      capacity=int(ref[1].numel()))
            lo, hi = int(ref[0][row_begin]), int(ref[0][row_begin + num_rows])
            assert torch.equal(indptr, ref[0][row_begin:row_begin + num_rows + 1] - lo)
            assert torch.equal(indices, ref[1][lo:hi]) and torch.equal(data, ref[2][lo:hi])
            if mode == 2:  # the index is single-use: a second extraction without a new asp_gather_index is refused
                with pytest.raises(AspError, match="single-use"):
                    common.extract_csr_indexed_device(op, full_s, full_p, row_begin, num_rows, workspace, capacity=int(ref[1].numel()))
            if mode == 2 and world == 4:  # the sharded host-buffer path: gather again, CSR straight into pinned host memory
                common.check(lib().asp_gather_index(
                    op.handle, world, rank, ffi.new("uint64_t[]", bounds),
                    ffi.new("uint64_t const *[]", [common.ptr(t, "uint64_t const *") for t in parts_s]),
                    ffi.new("double const *[]", [common.ptr(t, "double const *") for t in parts_p]),
                    ffi.NULL, 0, common.ptr(full_s, "uint64_t *"), common.ptr(full_p, "double *"), num_rows,
                    common.ptr(workspace, "void *"), need, common.stream()))
                h_indptr = torch.empty(num_rows + 1, dtype=torch.int32 if rank else torch.int64).pin_memory()
                h_indices = torch.empty(hi - lo + 5, dtype=torch.int32).pin_memory()
                h_data = torch.empty(hi - lo + 5, dtype=torch.float64).pin_memory()
                m = common.extract_indexed_to_host(op, full_s, full_p, row_begin, num_rows, workspace, h_indptr, h_indices, h_data)
                assert m == hi - lo and torch.equal(h_indptr.to(torch.int64), indptr.cpu())
                assert torch.equal(h_indices[:m], indices.cpu()) and torch.equal(h_data[:m], data.cpu())
    lib().asp_set_gather_mode(2)


def test_peer_memory_exchange_on_two_gpus():
    """X1 over NVLink peer memory with real processes: tests/_peer_sharding.py under torchrun, 2 ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k peer_memory)")
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(__file__), "_peer_sharding.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PEER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_batched_apply_device_matches_oracle_including_symmetry_groups():
    """Neighbour generation incl. orbit representatives / norms for the full kagome_36
    (|G| = 144 x 2) and pyrochlore (|G| = 384 x 2) groups; canonical per-row comparison."""
    for system, m in [("heisenberg_kagome_36", 300), ("heisenberg_pyrochlore_2x2x2", 200), ("heisenberg_kagome_18", 500),
                      ("j1j2_square_4x4", 500)]:
        cfg = asp.ls.load_config(asp.ls.system_path(system))
        op_np = OperatorNP.from_config(cfg)
        op = asp.load_hamiltonian(asp.ls.system_path(system))
        raw = synthetic.random_sector_states(op.basis.number_spins, op.basis.hamming_weight, m, 5).numpy().view(np.uint64)
        rep, _, norm = op_np.basis.state_info(raw)
        rows = np.unique(rep[norm > 0])
        s_ref, c_ref, k_ref = op_np.apply_u64(rows)
        s, c, k = op.batched_apply(rows)
        assert np.array_equal(k, k_ref)
        off = np.concatenate([[0], np.cumsum(k)])
        for r in range(rows.shape[0]):
            a = np.argsort(s[off[r]:off[r + 1], 0], kind="stable")
            b = np.argsort(s_ref[off[r]:off[r + 1]], kind="stable")
            assert np.array_equal(s[off[r]:off[r + 1], 0][a], s_ref[off[r]:off[r + 1]][b])
            got = np.bincount(np.unique(s[off[r]:off[r + 1], 0][a], return_inverse=True)[1], weights=c[off[r]:off[r + 1]].real[a])
            exp = np.bincount(np.unique(s_ref[off[r]:off[r + 1]][b], return_inverse=True)[1], weights=c_ref[off[r]:off[r + 1]][b])
            np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-300)


def test_symmetrised_apply_fast_kernels_equal_the_general_ones_bitwise():
    """All characters +1: counts without visiting orbits + the warp-per-row orbit kernel (apply_fill_orbit_kernel)
    against the lane-per-row walk (apply_fill_positive_kernel) and the general move-by-move kernels."""
    for system, m in [("heisenberg_kagome_36", 20000), ("heisenberg_pyrochlore_2x2x2", 5000), ("heisenberg_kagome_18", 24310),
                      ("j1j2_square_4x4", 3000), ("heisenberg_kagome_16", 2000)]:  # the last two: 16 spins, the top bit in the low word
        op = asp.load_hamiltonian(asp.ls.system_path(system))
        rows = op.basis.states if system == "heisenberg_kagome_18" else None
        if rows is None:
            rows = synthetic.cluster_closed_states(op, m, 9, DEV)
        else:
            rows = torch.from_numpy(np.ascontiguousarray(rows).view(np.int64)).to(DEV)
        fast = op.batched_apply_device(rows)  # warp per row: g(s ^ flip) = g(s) ^ g(flip), orbit minimum by warp reduction
        for mode in (1, 2):  # 1 = general move-by-move kernels, 2 = lane per row (apply_fill_positive_kernel)
            lib().asp_debug_set_apply_mode(mode)
            try:
                other = op.batched_apply_device(rows)
            finally:
                lib().asp_debug_set_apply_mode(0)
            for a, b in zip(fast, other):
                assert torch.equal(a, b), (system, mode)


def test_symmetrised_kagome_36_extraction_vs_oracle():
    """Full symmetrised path (orbit representatives + canonicalisation) on a subset of
    representatives of the 36-spin kagome basis."""
    system = "heisenberg_kagome_36"
    op_np = OperatorNP.load(asp.ls.system_path(system))
    op = asp.load_hamiltonian(asp.ls.system_path(system))
    raw = synthetic.random_sector_states(36, 18, 15, 9).numpy().view(np.uint64)
    rep, _, norm = op_np.basis.state_info(raw)
    seeds = np.unique(rep[norm > 0])
    shell, _, _ = op_np.apply_u64(seeds)
    spins = np.unique(np.concatenate([seeds, shell]))
    psi = synthetic.synthetic_amplitudes(spins.shape[0], 3).numpy()
    d_spins = torch.from_numpy(spins.view(np.int64)).to(DEV)
    indptr, indices, data = common.extract_csr_device(op, d_spins, torch.from_numpy(psi).to(DEV))
    n = spins.shape[0]
    ours = scipy.sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n, n))
    ref = _oracle_csr(op_np, spins, psi)
    _assert_same_matrix((0.5 * (ours + ours.T)).tocsr(), ref)


@pytest.mark.parametrize("system,states", [
    ("heisenberg_kagome_36", 10_000_000),       # BASELINE.json configs[3] at full size (one GPU's row block)
    ("heisenberg_pyrochlore_2x2x2", 10_000_000),  # configs[4]
    ("sk_32_1", 1_000_000),                      # configs[2]
])
def test_properties_at_scale(system, states):
    """Size-independent properties at BASELINE.json's full sizes (U(1) bases): rows sorted and
    duplicate-free, one diagonal per row, structurally symmetric, values symmetric to one rounding,
    random rows BITWISE equal to numpy's (c |psi_j|) |psi_i| (the reference's association)."""
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    op_np = OperatorNP.from_config(cfg)
    spins = synthetic.cluster_closed_states(op, states, 21, DEV)
    n = spins.shape[0]
    psi = synthetic.synthetic_amplitudes(n, 21, device=DEV)
    indptr, indices, data = common.extract_csr_device(op, spins, psi)
    lens = indptr[1:] - indptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), lens)
    assert int(indptr[0]) == 0 and int(indptr[-1]) == indices.numel() and bool((lens >= 1).all())
    same_row = rows[1:] == rows[:-1]
    assert bool((indices[1:][same_row] > indices[:-1][same_row]).all())  # strictly ascending
    assert int((indices.to(torch.int64) == rows).sum()) == n  # exactly one diagonal per row
    before = data.clone()
    common.symmetrize_csr_device(n, indptr, indices, data)  # raises if the pattern is asymmetric
    # raw entries are (c |psi_j|) |psi_i| (the reference's association, common.py:71-82): M and M^T differ by at most
    # one rounding, 0.5 (M + M^T) lies between them
    assert bool(((data - before).abs() <= 2.3e-16 * before.abs()).all())
    data = before
    # spot check 200 random rows against the oracle
    pick = np.sort(np.random.default_rng(0).choice(n, 200, replace=False))
    h_spins = spins.cpu().numpy().view(np.uint64)
    h_psi = psi.cpu().numpy()
    s, c, k = op_np.apply_u64(h_spins[pick])
    pos = np.clip(np.searchsorted(h_spins, s), 0, n - 1)
    hit = h_spins[pos] == s
    off = np.concatenate([[0], np.cumsum(k)])
    h_indptr, h_indices, h_data = indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy()
    for q, r in enumerate(pick):
        sel = slice(off[q], off[q + 1])
        cols = pos[sel][hit[sel]]
        vals = c[sel][hit[sel]] * np.abs(h_psi[cols]) * abs(h_psi[r])
        order = np.argsort(cols)
        assert np.array_equal(h_indices[h_indptr[r]:h_indptr[r + 1]], cols[order])
        assert np.array_equal(h_data[h_indptr[r]:h_indptr[r + 1]], vals[order])  # same association: bitwise


def test_full_matrix_equals_the_reference_c_at_a_million_states(oracle_capi):
    """EVERY row of a 10^6-state extraction (kagome_36-shaped, 4.6e6 couplings) against the reference's own compiled C
    (oracle/_ref = cbits/build_matrix.c, 512-bit keys), fed row chunk by row chunk with the oracle's neighbour lists:
    row starts and columns bit for bit, values to 1e-12 (the C path multiplies in another order)."""
    if not oracle_capi.have_ref():
        pytest.skip("oracle/_ref was not built (no /root/reference at build time)")
    from oracle.capi import pad512

    cfg = asp.ls.load_config(asp.ls.system_path("heisenberg_kagome_36"))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    op_np = OperatorNP.from_config(cfg)
    spins = synthetic.cluster_closed_states(op, 1_000_000, 31, DEV)
    n = spins.shape[0]
    psi = synthetic.synthetic_amplitudes(n, 31, device=DEV)
    indptr, indices, data = (t.cpu().numpy() for t in common.extract_csr_device(op, spins, psi))
    h_spins, h_psi = spins.cpu().numpy().view(np.uint64), psi.cpu().numpy()
    s512 = pad512(h_spins)
    counts = np.ones(n, dtype=np.int64)
    step = 125_000
    for lo in range(0, n, step):
        hi = min(lo + step, n)
        other, coeffs, k = op_np.apply_u64(h_spins[lo:hi])
        pos = np.clip(np.searchsorted(h_spins, other), 0, n - 1)
        other_psi = np.where(h_spins[pos] == other, h_psi[pos], 0.0)
        other_counts = np.zeros(n, dtype=np.int64)
        other_counts[lo:hi] = k
        rows, cols, vals, _ = oracle_capi.build_matrix(s512, counts, h_psi, pad512(other), coeffs, other_counts, other_psi, impl="ref")
        order = np.lexsort((cols, rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
        assert np.array_equal(np.bincount(rows.astype(np.int64) - lo, minlength=hi - lo), np.diff(indptr[lo:hi + 1]))
        assert np.array_equal(cols.astype(np.int32), indices[indptr[lo]:indptr[hi]])
        np.testing.assert_allclose(data[indptr[lo]:indptr[hi]], vals, rtol=1e-12, atol=0)


def test_host_buffer_entry_points_match_device_path():
    op = asp.load_hamiltonian(asp.ls.system_path("j1j2_square_4x4"))
    spins = synthetic.cluster_closed_states(op, 5000, 2, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], 2, device=DEV)
    indptr, indices, data = common.extract_csr_device(op, spins, psi)
    h_spins = spins.cpu().numpy().view(np.uint64)
    h_psi = psi.cpu().numpy()
    n = h_spins.shape[0]
    nnz = ffi.new("uint64_t *")
    job = ffi.new("asp_host_job **")
    rc = lib().asp_extract_host_begin(op.handle, n, ffi.cast("uint64_t *", h_spins.ctypes.data),
                                      ffi.cast("double *", h_psi.ctypes.data), 0, n, nnz, job)
    assert rc == 0 and nnz[0] == indices.numel()
    hp = np.zeros(n + 1, dtype=np.int64)
    hi = np.zeros(nnz[0], dtype=np.int32)
    hd = np.zeros(nnz[0], dtype=np.float64)
    rc = lib().asp_extract_host_finish(job[0], ffi.cast("int64_t *", hp.ctypes.data), ffi.cast("int32_t *", hi.ctypes.data),
                                       ffi.cast("double *", hd.ctypes.data))
    assert rc == 0
    assert np.array_equal(hp, indptr.cpu().numpy()) and np.array_equal(hi, indices.cpu().numpy())
    assert np.array_equal(hd, data.cpu().numpy())


def test_two_host_extractions_in_flight_match_the_device_path():
    """asp_extract_host_i32_submit / asp_extract_host_join: two jobs on different bases run concurrently (two arenas),
    each equals the device path."""
    op = asp.load_hamiltonian(asp.ls.system_path("j1j2_square_4x4"))
    cases = []
    for seed, n in [(3, 4000), (4, 9000)]:
        spins = synthetic.cluster_closed_states(op, n, seed, DEV)
        psi = synthetic.synthetic_amplitudes(spins.shape[0], seed, device=DEV)
        ref = common.extract_csr_device(op, spins, psi)
        m = int(ref[1].numel())
        h = dict(spins=spins.cpu().pin_memory(), psi=psi.cpu().pin_memory(), indptr=torch.zeros(spins.shape[0] + 1, dtype=torch.int32).pin_memory(),
                 indices=torch.zeros(m + 7, dtype=torch.int32).pin_memory(), data=torch.zeros(m + 7, dtype=torch.float64).pin_memory())
        cases.append((ref, m, h))
    jobs = []
    for ref, m, h in cases:
        job = ffi.new("asp_host_job **")
        n = h["spins"].shape[0]
        common.check(lib().asp_extract_host_i32_submit(op.handle, n, ffi.cast("uint64_t *", h["spins"].data_ptr()), ffi.cast("double *", h["psi"].data_ptr()),
                                                       0, n, m + 7, ffi.cast("int32_t *", h["indptr"].data_ptr()),
                                                       ffi.cast("int32_t *", h["indices"].data_ptr()), ffi.cast("double *", h["data"].data_ptr()), job))
        jobs.append(job[0])
    for job, (ref, m, h) in zip(jobs, cases):
        nnz = ffi.new("uint64_t *")
        common.check(lib().asp_extract_host_join(job, nnz))
        assert int(nnz[0]) == m
        assert torch.equal(h["indptr"].to(torch.int64), ref[0].cpu())
        assert torch.equal(h["indices"][:m], ref[1].cpu()) and torch.equal(h["data"][:m], ref[2].cpu())
    assert lib().asp_extract_host_join(ffi.NULL, ffi.NULL) == lib().ASP_ERR_ARG


def _u1_operator(system):
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    return asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))


@pytest.mark.parametrize("system,n,seed", [
    ("j1j2_square_4x4", 12870, 0), ("sk_16_1", 4000, 1), ("heisenberg_kagome_36", 300000, 2), ("sk_32_1", 30000, 3),
    ("heisenberg_kagome_16", 1, 4), ("heisenberg_kagome_16", 127, 5), ("heisenberg_kagome_16", 129, 6),
])
def test_single_pass_kernel_equals_two_pass_bitwise(system, n, seed):
    """asp_extract_csr (single pass, decoupled look-back, compacted searches) against
    asp_extract_count/fill (lane per row, two passes): identical indptr, indices AND values, for
    sparse subsets, a full basis (every candidate hits) and partial tiles."""
    op = _u1_operator(system)
    spins = synthetic.cluster_closed_states(op, n, seed, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], seed, device=DEV)
    ref = common.extract_csr_two_pass_device(op, spins, psi)
    # automatic survivor lists; tiny ones (1, 3 and 8 slots per lane) force many exact-search rounds per
    # tile; a coarse filter lets most misses through to the exact search, a fine one almost none;
    # stage A lane by lane instead of on bit planes
    for cap, tuning in [(0, (0, 0, 0)), (32, (0, 0, 0)), (96, (-6, -3, 0)), (256, (-4, 0, 0)), (0, (3, 2, 1)), (0, (-30, -30, 1))]:
        lib().asp_debug_set_hit_list_capacity(cap)
        lib().asp_debug_set_extract_tuning(*tuning)
        try:
            got = common.extract_csr_device(op, spins, psi)
        finally:
            lib().asp_debug_set_hit_list_capacity(0)
            lib().asp_debug_set_extract_tuning(0, 0, 0)
        for a, b in zip(got, ref):
            assert torch.equal(a, b), (system, cap, tuning)
    # row blocks (what each GPU of a sharded run builds)
    m = spins.shape[0]
    lo, hi = m // 3, m - m // 5
    part = common.extract_csr_device(op, spins, psi, lo, hi - lo)
    base = int(ref[0][lo])
    assert torch.equal(part[0], ref[0][lo:hi + 1] - base)
    assert torch.equal(part[1], ref[1][base:int(ref[0][hi])]) and torch.equal(part[2], ref[2][base:int(ref[0][hi])])


def test_single_pass_capacity_contract():
    """Caller-sized outputs (cbits/build_matrix.c:22-28 convention): a short guess reports the
    exact count with ASP_ERR_WORKSPACE and complete indptr; the retry succeeds."""
    op = _u1_operator("heisenberg_kagome_36")
    spins = synthetic.cluster_closed_states(op, 50000, 8, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], 8, device=DEV)
    n = spins.shape[0]
    ref = common.extract_csr_two_pass_device(op, spins, psi)
    nnz_true = int(ref[1].numel())
    got = common.extract_csr_device(op, spins, psi, nnz_hint=nnz_true // 3)  # forces the retry
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    ws = torch.empty(int(lib().asp_extract_csr_workspace_bytes(op.handle, n, n)), dtype=torch.uint8, device=DEV)
    indptr = torch.empty(n + 1, dtype=torch.int64, device=DEV)
    small = nnz_true // 2
    indices = torch.full((small + 16,), -7, dtype=torch.int32, device=DEV)
    data = torch.zeros(small + 16, dtype=torch.float64, device=DEV)
    nnz = ffi.new("uint64_t *")
    rc = lib().asp_extract_csr(op.handle, n, common.ptr(spins, "uint64_t *"), common.ptr(psi, "double *"), 0, n,
                               common.ptr(ws, "void *"), ws.numel(), small, common.ptr(indptr, "int64_t *"),
                               common.ptr(indices, "int32_t *"), common.ptr(data, "double *"), nnz, common.stream())
    assert rc == lib().ASP_ERR_WORKSPACE and int(nnz[0]) == nnz_true
    assert torch.equal(indptr, ref[0])
    assert torch.equal(indices[:small], ref[1][:small]) and bool((indices[small:] == -7).all())  # nothing past the capacity
    rc = lib().asp_extract_csr(op.handle, n, common.ptr(spins, "uint64_t *"), common.ptr(psi, "double *"), 0, n,
                               common.ptr(ws, "void *"), 10, small, common.ptr(indptr, "int64_t *"),
                               common.ptr(indices, "int32_t *"), common.ptr(data, "double *"), nnz, common.stream())
    assert rc == lib().ASP_ERR_WORKSPACE


def test_one_call_host_pipeline_with_row_chunks():
    """asp_extract_host: several row chunks (> 2^17 rows each), D2H overlapped with extraction;
    equal to the device path, for the whole basis and for a row block; capacity contract."""
    op = _u1_operator("heisenberg_kagome_36")
    spins = synthetic.cluster_closed_states(op, 400000, 13, DEV)
    psi = synthetic.synthetic_amplitudes(spins.shape[0], 13, device=DEV)
    n = spins.shape[0]
    h_spins = spins.cpu().numpy().view(np.uint64)
    h_psi = psi.cpu().numpy()
    for lo, rows in [(0, n), (n // 7, n - n // 3)]:
        indptr, indices, data = common.extract_csr_device(op, spins, psi, lo, rows)
        m = int(indices.numel())
        cap = m + 100
        hp = np.full(rows + 1, -1, dtype=np.int64)
        hi = np.full(cap, -1, dtype=np.int32)
        hd = np.zeros(cap, dtype=np.float64)
        nnz = ffi.new("uint64_t *")
        c = lambda a, t: ffi.cast(t, a.ctypes.data)  # noqa: E731
        rc = lib().asp_extract_host(op.handle, n, c(h_spins, "uint64_t *"), c(h_psi, "double *"), lo, rows, cap,
                                    c(hp, "int64_t *"), c(hi, "int32_t *"), c(hd, "double *"), nnz)
        assert rc == 0 and int(nnz[0]) == m
        assert np.array_equal(hp, indptr.cpu().numpy())
        assert np.array_equal(hi[:m], indices.cpu().numpy()) and np.array_equal(hd[:m], data.cpu().numpy())
        assert np.all(hi[m:] == -1)
        rc = lib().asp_extract_host(op.handle, n, c(h_spins, "uint64_t *"), c(h_psi, "double *"), lo, rows, m // 2,
                                    c(hp, "int64_t *"), c(hi, "int32_t *"), c(hd, "double *"), nnz)
        assert rc == lib().ASP_ERR_WORKSPACE and int(nnz[0]) == m and np.array_equal(hp, indptr.cpu().numpy())
        # int32 row starts (scipy's own index type below 2^31 couplings): the buffers drop into csr_matrix without a copy
        hp32 = np.full(rows + 1, -1, dtype=np.int32)
        hi[:] = -1
        rc = lib().asp_extract_host_i32(op.handle, n, c(h_spins, "uint64_t *"), c(h_psi, "double *"), lo, rows, cap,
                                        c(hp32, "int32_t *"), c(hi, "int32_t *"), c(hd, "double *"), nnz)
        assert rc == 0 and int(nnz[0]) == m and np.array_equal(hp32, indptr.cpu().numpy().astype(np.int32))
        assert np.array_equal(hi[:m], indices.cpu().numpy()) and np.array_equal(hd[:m], data.cpu().numpy())
        block = scipy.sparse.csr_matrix((hd[:m], hi[:m], hp32), shape=(rows, n))
        assert block.indptr.dtype == np.int32 and np.shares_memory(block.indptr, hp32) and np.shares_memory(block.data, hd)


def _golden_log_psi(spins):
    """tests/golden/make_golden.py:hashed_log_psi (the amplitude model the N2 vectors were made with)."""
    s = np.asarray(spins, dtype=np.uint64)
    u = ((s * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)).astype(np.float64) / 2.0 ** 53
    v = ((s * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(63)).astype(np.float64)
    return 4.0 * (u - 0.5) + 1j * np.pi * v


@pytest.mark.parametrize("name", ["n2_heisenberg_kagome_16", "n2_j1j2_square_4x4", "n2_heisenberg_kagome_18"])
def test_cluster_extension_and_sparsification_match_the_reference(golden_dir, name):
    """SURVEY 8f N2 against vectors produced by the reference's OWN make_hamiltonian_extension
    (common.py:516-522), get_strongest_off_diag (:539-541) and sparsify_using_global_cutoff
    (:634-692): same states, same sparsity pattern, couplings to 1e-12, same packed signs."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    op = asp.load_hamiltonian(asp.ls.system_path(str(g["system"])))
    model0 = asp.make_ising_model(g["cluster"], op, log_psi_fn=_golden_log_psi)
    model1 = asp.make_hamiltonian_extension(model0, _golden_log_psi)
    assert np.array_equal(model1.spins, g["ext_spins"])
    n1 = model1.size
    ref1 = scipy.sparse.coo_matrix((g["ext_data"], (g["ext_row"], g["ext_col"])), shape=(n1, n1)).tocsr()
    _assert_same_matrix(model1.ising_hamiltonian.exchange.tocsr(), ref1)
    assert np.array_equal(model1.initial_signs, g["ext_x0"])
    for arg in (model1.ising_hamiltonian, model1.ising_hamiltonian.exchange):
        np.testing.assert_allclose(asp.get_strongest_off_diag(arg), g["strongest"], rtol=1e-12, atol=0)
    model2 = asp.sparsify_using_global_cutoff(model1, float(g["reltol"]), g["frozen"])
    assert np.array_equal(model2.spins, g["sp_spins"])
    n2 = model2.size
    ref2 = scipy.sparse.coo_matrix((g["sp_data"], (g["sp_row"], g["sp_col"])), shape=(n2, n2)).tocsr()
    _assert_same_matrix(model2.ising_hamiltonian.exchange.tocsr(), ref2)
    assert np.array_equal(model2.initial_signs, g["sp_x0"]) and np.array_equal(model2.ising_hamiltonian.field, g["sp_field"])
    # the sparsified model is a working Ising model: its device copy anneals
    x = asp.solve_ising_model(model2, mode="sa", seed=1, number_sweeps=50, repetitions=32)
    assert x.shape == ((n2 + 63) // 64,)
    with pytest.raises(AssertionError):  # frozen spins that are not part of the model (common.py:544-548)
        asp.sparsify_using_global_cutoff(model1, float(g["reltol"]), np.array([int(g["ext_spins"][0]) + 1], dtype=np.uint64))


def test_cutoff_components_equal_scipy_on_random_graphs():
    from scipy.sparse.csgraph import connected_components

    rng = np.random.default_rng(3)
    for n, density, reltol in [(1, 0.0, 0.5), (500, 0.004, 0.0), (20000, 0.00008, 0.3), (5000, 0.0006, 0.6)]:
        a = scipy.sparse.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal).tocsr()
        a = (a + a.T).tocsr()
        a.sort_indices()
        frozen = (rng.random(n) < 0.05).astype(np.uint8)
        data = a.data.copy()
        if data.size:
            rows = np.repeat(np.arange(n), np.diff(a.indptr))
            big = np.abs(data).max()
            drop = (np.abs(data) < reltol * big) & ~((frozen[rows] == 1) & (frozen[a.indices] == 1))
            data[drop] = 0
        b = scipy.sparse.csr_matrix((data, a.indices.copy(), a.indptr.copy()), shape=(n, n))  # eliminate_zeros works in place
        b.eliminate_zeros()
        _, ref = connected_components(b, directed=False)
        labels = torch.empty(n, dtype=torch.int32, device=DEV)
        t = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(DEV)  # noqa: E731
        indptr, indices, vals, fr = t(a.indptr, np.int64), t(a.indices, np.int32), t(a.data, np.float64), t(frozen, np.uint8)
        common.check(lib().asp_cutoff_components(n, common.ptr(indptr, "int64_t *"), common.ptr(indices, "int32_t *"),
                                                 common.ptr(vals, "double *"), a.nnz, reltol, common.ptr(fr, "unsigned char *"),
                                                 common.ptr(labels, "int32_t *"), common.stream()))
        got = labels.cpu().numpy()
        # same partition; our label is the smallest vertex of the component
        smallest = np.full(ref.max() + 1, n, dtype=np.int64)
        np.minimum.at(smallest, ref, np.arange(n))
        assert np.array_equal(got, smallest[ref])
