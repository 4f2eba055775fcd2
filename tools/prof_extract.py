#!/usr/bin/env python
"""Three extractions of the bench workload (profiling target for ncu: -k regex:extract_csr_kernel).

    python tools/prof_extract.py [--states N] [--system S]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--states", type=int, default=10_000_000)
    p.add_argument("--system", default="heisenberg_kagome_36")
    p.add_argument("--calls", type=int, default=3)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    cfg = asp.ls.load_config(asp.ls.system_path(args.system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    spins = synthetic.cluster_closed_states(op, args.states, 1000, dev)
    psi = synthetic.synthetic_amplitudes(int(spins.shape[0]), 77, device=dev)
    for _ in range(args.calls):
        indptr, indices, data = common.extract_csr_device(op, spins, psi)
        torch.cuda.synchronize()
    print("nnz", int(indices.numel()))


if __name__ == "__main__":
    main()
