#!/usr/bin/env python
"""The annealing step of the bench (kagome_36-shaped 10^7-spin model, 64 replicas, 16 strided sweeps) three times:
profiling target for ncu (-k regex:sa_sweep_kernel).

    python tools/prof_anneal.py [--states N] [--replicas R] [--sweeps S]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--states", type=int, default=10_000_000)
    p.add_argument("--system", default="heisenberg_kagome_36")
    p.add_argument("--replicas", type=int, default=64)
    p.add_argument("--sweeps", type=int, default=16)
    p.add_argument("--calls", type=int, default=3)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    cfg = asp.ls.load_config(asp.ls.system_path(args.system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    spins = synthetic.cluster_closed_states(op, args.states, 1000, dev)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 77, device=dev)
    indptr, indices, data = common.extract_csr_device(op, spins, psi)

    class _Shape:
        shape = (n, n)

    ham = asp.sa.Hamiltonian(_Shape(), np.zeros(n), _device_csr=(indptr, indices, data, None))
    plan = asp.sa.AnnealPlan(ham)
    if os.environ.get("ASP_PRINT_CLASSES"):
        ex = plan.export()
        print("class sizes:", np.diff(ex["class_ptr"]).tolist(), "row length histogram:", np.bincount(np.diff(ex["indptr"]))[:16].tolist(), flush=True)
    betas = np.ascontiguousarray(asp.sa.default_betas(ham, args.sweeps * 8)[3::8])
    escale = asp.sa.energy_scale(ham)
    for k in range(args.calls):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        bits, energies = plan.anneal_device(args.replicas, betas, 100 + k, escale=escale)
        ev[1].record()
        torch.cuda.synchronize()
        print("anneal call %d: %.2f ms, best E %.9f" % (k, ev[0].elapsed_time(ev[1]), float(energies.min())), flush=True)


if __name__ == "__main__":
    main()
