#!/usr/bin/env python
"""Annealing time on a small full-basis model (development tool): j1j2_square_4x4, 64 replicas, many sweeps -- the
latency-bound regime (a few hundred tasks per colour class), where the team barrier and the start-up of every class
phase dominate.

    python tools/time_small_anneal.py [--system S] [--replicas R] [--sweeps N]
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
    p.add_argument("--system", default="j1j2_square_4x4")
    p.add_argument("--replicas", type=int, default=64)
    p.add_argument("--sweeps", type=int, default=1024)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    cfg = asp.ls.load_config(asp.ls.system_path(args.system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis)
    basis.build()
    spins = torch.from_numpy(np.ascontiguousarray(basis.states).view(np.int64)).to(dev)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 3, device=dev)
    indptr, indices, data = common.extract_csr_device(op, spins, psi)

    class _Shape:
        shape = (n, n)

    ham = asp.sa.Hamiltonian(_Shape(), np.zeros(n), _device_csr=(indptr, indices, data, None))
    plan = asp.sa.AnnealPlan(ham)
    betas = asp.sa.default_betas(ham, args.sweeps)
    escale = asp.sa.energy_scale(ham)
    for k in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        bits, energies = plan.anneal_device(args.replicas, betas, 100 + k, escale=escale)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        print("%s n=%d classes=%d R=%d sweeps=%d: %.2f ms  %.3g proposals/s  best E %.9f" % (
            args.system, n, plan.num_classes, args.replicas, args.sweeps, ms, n * args.replicas * args.sweeps / ms * 1e3, float(energies.min())), flush=True)


if __name__ == "__main__":
    main()
