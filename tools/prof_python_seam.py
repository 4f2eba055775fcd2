#!/usr/bin/env python
"""cProfile of common.make_ising_model on the bench workload (where the Python seam spends its time)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
cfg = asp.ls.load_config(asp.ls.system_path("heisenberg_kagome_36"))
cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
spins = synthetic.cluster_closed_states(op, n, 1000, dev).cpu().numpy().view(np.uint64)
log_psi = np.log(synthetic.synthetic_amplitudes(spins.shape[0], 77).numpy().astype(np.complex128))
asp.make_ising_model(spins, op, log_psi=log_psi)
t0 = time.perf_counter()
asp.make_ising_model(spins, op, log_psi=log_psi)
print("make_ising_model: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
os.environ["CUDA_LAUNCH_BLOCKING"] = "0"
pr = cProfile.Profile()
pr.enable()
asp.make_ising_model(spins, op, log_psi=log_psi)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
