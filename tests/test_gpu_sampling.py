"""-m gpu parity tests of the sampling front-end: CUDA through the C ABI vs golden
vectors produced by the reference's own functions and vs the numpy oracle at scale."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from oracle import live_path  # noqa: E402

DEV = torch.device("cuda")


def _setup(golden_dir):
    g = np.load(os.path.join(golden_dir, "n3_heisenberg_kagome_16.npz"))
    op = asp.load_hamiltonian(asp.ls.system_path(str(g["system"])))
    return g, op, op.basis.states


def test_monte_carlo_sampling_draws_the_reference_states(golden_dir):
    g, op, states = _setup(golden_dir)
    np.random.seed(int(g["seed"]))  # the reference consumes numpy's global legacy stream (common.py:276)
    mc2 = asp.monte_carlo_sampling(states, g["psi"], g["mc2"].shape[0], sampled_power=2)
    mc1 = asp.monte_carlo_sampling(states, g["psi"], g["mc1"].shape[0], sampled_power=1)
    assert isinstance(mc2, asp.SamplingResult) and mc2.weights is None
    assert np.array_equal(mc2.spins, g["mc2"]) and np.array_equal(mc1.spins, g["mc1"])


def test_log_coeff_fn_exact_solution_noise_and_amplitude_overlap(golden_dir):
    g, op, states = _setup(golden_dir)
    fn = asp.ground_state_to_log_coeff_fn(g["psi"], op.basis)
    assert np.array_equal(fn(g["mc2"][:500]), g["log_coeff"])
    padded = np.zeros((500, 8), dtype=np.uint64)  # the [n, 8] form (common.py:814-815)
    padded[:, 0] = g["mc2"][:500]
    assert np.array_equal(fn(padded), g["log_coeff"])
    assert np.array_equal(asp.determine_exact_solution(g["mc2"], op, g["psi"]), g["exact_bits"])
    np.random.seed(int(g["seed"]) + 1)
    noisy = asp.add_noise_to_amplitudes(g["psi"], 0.3)
    assert np.array_equal(noisy[::37], g["noisy_stride"])
    ov = asp.amplitude_overlap(g["mc2"], g["psi"], noisy, op.basis)
    idx = live_path.batched_index(states, g["mc2"])
    a, b = np.abs(g["psi"][idx]), np.abs(noisy[idx])
    assert ov == np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b)
    with pytest.raises(ValueError):  # lattice_symmetries raises for a state outside the basis
        op.basis.batched_index(np.array([states[5], 0b111], dtype=np.uint64))


def test_small_cluster_growth_equals_the_reference_function(golden_dir):
    """create_small_cluster_around_point (common.py:481-513) driven by the device operator's apply():
    same clusters as the reference's function driven by the oracle operator (ascending neighbours)."""
    g, op, states = _setup(golden_dir)
    np.random.seed(int(g["seed"]) + 2)
    for k, size in enumerate([20, 77, 300]):
        cluster = asp.create_small_cluster_around_point(int(g["mc2"][k]), op, required_size=size, keep_probability=0.5)
        assert np.array_equal(np.asarray(cluster, dtype=np.uint64), g["cluster%d" % k])


def test_batched_index_and_sampling_at_scale():
    """3e7-state basis (the size of the kagome_36 representative list, common.py:817 call site):
    every position equals numpy's searchsorted; 10^6 draws equal the numpy restatement (a draw
    may move to the neighbouring bin only when u sits within an ulp of a bin edge)."""
    n, m = 30_000_000, 1_000_000
    gen = torch.Generator(device=DEV)
    gen.manual_seed(3)
    states = synthetic._sorted_unique_unsigned(torch.randint(0, 1 << 36, (n,), generator=gen, device=DEV, dtype=torch.int64))
    n = int(states.shape[0])
    basis = asp.ls.SpinBasis(36).build(states.cpu().numpy().view(np.uint64))
    pick = torch.randint(0, n, (m,), generator=gen, device=DEV)
    needles = states[pick]
    assert torch.equal(basis.batched_index_device(needles), pick)
    absent = needles.clone()
    absent[::1000] |= 1 << 40  # 1000 needles outside the basis
    with pytest.raises(ValueError, match="1000 state"):
        basis.batched_index_device(absent)
    psi = synthetic.synthetic_amplitudes(n, 9, device=DEV)
    u = torch.rand(m, generator=gen, device=DEV, dtype=torch.float64)
    for power in (2, 1, 1.5):
        got = common.sample_indices_device(psi, u, power).cpu().numpy()
        ref = live_path.sample_indices(psi.cpu().numpy(), u.cpu().numpy(), power)
        diff = np.abs(got - ref)
        assert diff.max() <= 1 and np.count_nonzero(diff) <= (2 if power != 1.5 else 50)
    # all the weight on one state; zero-weight states are never drawn
    spike = torch.zeros(1000, dtype=torch.float64, device=DEV)
    spike[[17, 400]] = torch.tensor([3.0, -4.0], dtype=torch.float64, device=DEV)
    draws = common.sample_indices_device(spike, u[:10000], 2).cpu().numpy()
    assert set(np.unique(draws)) == {17, 400}
    assert abs(np.mean(draws == 400) - 16 / 25) < 0.02


def test_cluster_experiment_pipeline_end_to_end(tmp_path):
    """sampled_connected_components.py:646-750 on kagome_16 with its exact ground state: sampling ->
    clusters -> extraction -> extension + sparsification -> greedy / SA -> accuracy, overlap, CSV; and
    the model dump (common.py:750-768) with its int32 index arrays."""
    from annealing_sign_problem_b200 import experiments, formats
    from oracle.operator_np import OperatorNP, ground_state, system_path

    e0, psi, _ = ground_state(OperatorNP.load(system_path("heisenberg_kagome_16")))
    op = asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_16"))
    path = str(tmp_path / "heisenberg_kagome_16.npz")
    formats.save_ground_state(path, psi, e0, op.basis.states)
    ground, energy, reps = formats.load_ground_state(path)
    op.basis.build(reps)  # common.py:801
    assert energy == e0

    class Args:
        seed, order, noise, global_cutoff, sampled_power = 11, 1, 0.0, 2e-2, 2
        number_samples, min_cluster_size, max_cluster_size, keep_probability = 3, 30, 60, 0.5
        annealing, number_sweeps, repetitions = True, 300, 16

    np.random.seed(Args.seed)
    fn = asp.ground_state_to_log_coeff_fn(ground, op.basis)
    clusters = experiments.generate_clusters(op, ground, Args)
    assert len(clusters) == 3 and all(30 <= c.shape[0] <= 60 or c.shape[0] < 30 for c in clusters)
    out = str(tmp_path / "result.csv")
    experiments.write_csv_header(out, Args)
    for cluster in clusters:
        columns = experiments.process_cluster(cluster, op, ground, ground, fn, Args, number_sweeps=Args.number_sweeps,
                                              repetitions=Args.repetitions)
        experiments.append_csv_row(out, columns)
        assert len(columns) == 2 and columns[0].size == cluster.shape[0] and columns[1].size >= columns[0].size
        for r in columns:
            print(r.to_csv_str())
            assert 0.5 <= r.greedy_accuracy <= 1 and 0.5 <= r.sa_accuracy <= 1
            assert 0 <= r.greedy_overlap <= 1 + 1e-12 and 0 <= r.sa_overlap <= 1 + 1e-12
            assert abs(r.amplitude_overlap - 1) < 1e-12  # no noise
        # the extension sees every neighbour of the cluster: SA recovers (almost) all its signs
        assert columns[1].sa_accuracy >= columns[0].sa_accuracy - 0.1
        assert min(r.sa_overlap for r in columns) > 0.99 and min(r.greedy_overlap for r in columns) > 0.99
    assert len(open(out).read().splitlines()) == 12 + 3
    # model dump of the full-basis model: int32 CSR, energy of the exact signs = E0
    model = asp.make_ising_model(op.basis.states, op, log_psi_fn=fn)
    dump = str(tmp_path / "model.npz")
    formats.dump_ising_model_to_hdf5(model, ground, dump)
    d = np.load(dump)
    assert d["indices"].dtype == np.int32 and d["indptr"].dtype == np.int32 and d["signs"].dtype == np.uint64
    # 177 606 candidates minus the diagonal entries that are exactly zero (12 parallel +
    # 12 antiparallel bonds), which scipy's 0.5 (M + M^T) drops in the reference too (common.py:194)
    assert d["indptr"][-1] == d["elements"].shape[0] == d["indices"].shape[0] and 170_000 < d["elements"].shape[0] <= 177_606
    assert np.count_nonzero(d["elements"]) == d["elements"].shape[0]
    assert abs(float(d["energy"]) - e0) < 1e-10 and not d["field"].any()
