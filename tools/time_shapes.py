#!/usr/bin/env python
"""Kernel-only time of the single-pass extraction on the BASELINE shapes (development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from annealing_sign_problem_b200._lib import lib  # noqa: E402

DEV = torch.device("cuda", 0)


def u1(system):
    cfg = asp.ls.load_config(asp.ls.system_path(system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    return asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))


def run(name, op, spins):
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 5, device=DEV)
    lib().asp_debug_time_extract_kernel(1)
    times = []
    for _ in range(5):
        indptr, indices, data = common.extract_csr_device(op, spins, psi, nnz_hint=None if name != "dense" else n * op.max_candidates)
        torch.cuda.synchronize()
        times.append(float(lib().asp_debug_last_extract_kernel_ms()))
    lib().asp_debug_time_extract_kernel(0)
    nnz = int(indices.numel())
    ms = float(np.median(times[1:]))
    algo = 24.0 * n + 20.0 * nnz
    print("%-28s n=%d nnz=%d kernel %.3f ms  %.3g couplings/s  %.1f %% of 6451.8 GB/s" % (name, n, nnz, ms, nnz / ms * 1e3, 100 * algo / ms / 1e6 / 6451.8), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kagome", "pyro", "sk", "dense"]
    if "kagome" in which:
        op = u1("heisenberg_kagome_36")
        run("kagome_36 1e7", op, synthetic.cluster_closed_states(op, 10_000_000, 1000, DEV))
    if "pyro" in which:
        op = u1("heisenberg_pyrochlore_2x2x2")
        run("pyrochlore_2x2x2 1e7", op, synthetic.cluster_closed_states(op, 10_000_000, 5, DEV))
    if "sk" in which:
        op = u1("sk_32_1")
        run("sk_32_1 1e6", op, synthetic.cluster_closed_states(op, 1_000_000, 5, DEV))
    if "dense" in which:
        heis = [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]]
        ns = 26
        cfg = {"basis": {"number_spins": ns, "hamming_weight": ns // 2, "symmetries": []},
               "hamiltonian": {"name": "J1-J2 ring", "terms": [
                   {"matrix": heis, "sites": [[i, (i + 1) % ns] for i in range(ns)]},
                   {"matrix": (0.5 * np.array(heis)).tolist(), "sites": [[i, (i + 2) % ns] for i in range(ns)]}]}}
        basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
        op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
        run("dense", op, torch.from_numpy(basis.states.view(np.int64)).to(DEV))
