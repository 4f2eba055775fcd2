#!/usr/bin/env python
"""Times the BASELINE.json configurations other than the headline one (development tool; the
numbers quoted in DESIGN.md 6 come from here).  python tools/run_configs.py [cfg2 cfg3 cfg5]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402

DEV = torch.device("cuda", 0)


def u1(system):
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    return asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def sampled(system, states, replicas, sweeps, symmetrised=False):
    op = asp.load_hamiltonian(asp.ls.system_path(system)) if symmetrised else u1(system)
    spins = synthetic.cluster_closed_states(op, states, 5, DEV)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 5, device=DEV)
    dt, (indptr, indices, data) = timed(lambda: common.extract_csr_device(op, spins, psi, nnz_hint=None))
    nnz = int(indices.numel())
    print(("symmetrised " if symmetrised else "") + "%s: n=%d nnz=%d  extraction %.2f ms (%.3g couplings/s, %.3g candidates/s)" % (
        system, n, nnz, 1e3 * dt, nnz / dt, n * (op.max_candidates / 4.0 + 1) / dt), flush=True)

    class _Shape:
        shape = (n, n)

    ham = asp.sa.Hamiltonian(_Shape(), np.zeros(n), _device_csr=(indptr, indices, data, None))
    t0 = time.perf_counter()
    plan = asp.sa.AnnealPlan(ham)
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    betas = asp.sa.default_betas(ham, max(64, sweeps * 8))[:sweeps]
    escale = asp.sa.energy_scale(ham)
    dt, (bits, energies) = timed(lambda: plan.anneal_device(replicas, betas, 1, escale=escale), reps=2)
    gdt, (gbits, genergy, rounds, gsweeps) = timed(lambda: plan.greedy_device(), reps=2)
    print("   greedy: %.1f ms (%d merge rounds, %d descent sweeps), E %.6f" % (1e3 * gdt, rounds, gsweeps, float(genergy)), flush=True)
    print("   SA: %d replicas x %d sweeps: %.1f ms = %.3g proposals/s (plan %.0f ms, %d colour classes), best E %.6f" % (
        replicas, sweeps, 1e3 * dt, replicas * sweeps * n / dt, 1e3 * plan_s, plan.num_classes, float(energies.min())), flush=True)


def cfg2():
    from oracle.operator_np import OperatorNP, ground_state  # checker only: exact eigenvector for the model

    name = "heisenberg_kagome_18"
    e0, psi, _ = ground_state(OperatorNP.load(asp.ls.system_path(name)))
    op = asp.load_hamiltonian(asp.ls.system_path(name))
    with np.errstate(divide="ignore"):
        t0 = time.perf_counter()
        model = asp.make_ising_model(op.basis.states, op, log_psi=np.log(psi.astype(np.complex128)))
        dt_model = time.perf_counter() - t0
    ham = model.ising_hamiltonian
    for sweeps in (100, 1600):
        t0 = time.perf_counter()
        xs, es = asp.sa.anneal(ham, seed=0, number_sweeps=sweeps, repetitions=1024, only_best=False)
        dt = time.perf_counter() - t0
        ok = np.abs((es - e0) / e0) <= 1e-12
        print("kagome_18 symmetrised: n=%d nnz=%d make_ising_model %.0f ms; SA 1024 x %d sweeps %.1f ms = %.3g proposals/s; "
              "best E - E0 = %.2e, replicas at E0: %.3f" % (model.size, ham.exchange.nnz, 1e3 * dt_model, sweeps, 1e3 * dt,
                                                            1024 * sweeps * model.size / dt, es.min() - e0, ok.mean()), flush=True)


def dense(number_spins=26):
    """Every candidate is a hit: the FULL U(1) basis of a J1-J2 Heisenberg ring (C(26,13) = 1.04e7 states).
    Shows how the roofline fraction of the extraction kernel moves with the hit rate."""
    import json

    from annealing_sign_problem_b200._lib import lib

    heis = [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]]
    cfg = {"basis": {"number_spins": number_spins, "hamming_weight": number_spins // 2, "symmetries": []},
           "hamiltonian": {"name": "J1-J2 ring", "terms": [
               {"matrix": heis, "sites": [[i, (i + 1) % number_spins] for i in range(number_spins)]},
               {"matrix": (0.5 * np.array(heis)).tolist(), "sites": [[i, (i + 2) % number_spins] for i in range(number_spins)]}]}}
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
    spins = torch.from_numpy(basis.states.view(np.int64)).to(DEV)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 5, device=DEV)
    lib().asp_debug_time_extract_kernel(1)
    dt, (indptr, indices, data) = timed(lambda: common.extract_csr_device(op, spins, psi, nnz_hint=n * op.max_candidates))
    kernel_ms = float(lib().asp_debug_last_extract_kernel_ms())
    lib().asp_debug_time_extract_kernel(0)
    nnz = int(indices.numel())
    algo = 24.0 * n + 20.0 * nnz
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print("dense (full basis, %d-spin J1-J2 ring): n=%d nnz=%d (%.1f per row, every candidate a hit)  extraction call %.2f ms, kernel %.3f ms "
          "= %.3g couplings/s; algorithmic bytes %.2f GB -> %.0f GB/s = %.1f %% of the HBM roofline (%.0f GB/s)" % (
              number_spins, n, nnz, nnz / n, 1e3 * dt, kernel_ms, nnz / (kernel_ms * 1e-3), algo / 1e9, algo / kernel_ms / 1e6,
              100 * algo / kernel_ms / 1e6 / peak, peak), flush=True)


def n3():
    """Sampling front-end at the size of the kagome_36 representative list."""
    n, m = 31_527_894, 10_000_000
    gen = torch.Generator(device=DEV)
    gen.manual_seed(3)
    states = synthetic._sorted_unique_unsigned(torch.randint(0, 1 << 36, (n,), generator=gen, device=DEV, dtype=torch.int64))
    n = int(states.shape[0])
    basis = asp.ls.SpinBasis(36).build(states.cpu().numpy().view(np.uint64))
    needles = states[torch.randint(0, n, (m,), generator=gen, device=DEV)]
    basis.batched_index_device(needles[:10])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    basis.batched_index_device(needles)
    ev[1].record()
    psi = synthetic.synthetic_amplitudes(n, 9, device=DEV)
    u = torch.rand(m, generator=gen, device=DEV, dtype=torch.float64)
    common.sample_indices_device(psi, u[:10])
    ev[2].record()
    common.sample_indices_device(psi, u)
    ev[3].record()
    torch.cuda.synchronize()
    print("N3: batched_index of %d needles in a %d-state basis: %.2f ms (%.3g lookups/s); monte_carlo_sampling of %d states from it "
          "(cumulative sum + searches): %.2f ms" % (m, n, ev[0].elapsed_time(ev[1]), m / ev[0].elapsed_time(ev[1]) * 1e3, m,
                                                   ev[2].elapsed_time(ev[3])), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]
    if "cfg2" in which:
        cfg2()
    if "cfg3" in which:
        sampled("sk_32_1", 1_000_000, 4096, 4)
    if "cfg5" in which:
        sampled("heisenberg_pyrochlore_2x2x2", 10_000_000, 64, 16)
    if "cfg4sym" in which:  # symmetrised kagome_36 (|G| = 144 x spin inversion): integer-ALU-bound orbit representatives
        sampled("heisenberg_kagome_36", 1_000_000, 64, 16, symmetrised=True)
    if "dense" in which:
        dense()
    if "n3" in which:
        n3()
    if "cfg4" in which:
        sampled("heisenberg_kagome_36", 10_000_000, 64, 16)
