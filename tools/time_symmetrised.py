#!/usr/bin/env python
"""Stage times of the symmetrised extraction path (kagome_36, |G| = 144 x spin inversion) on 10^6 representatives."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
op = asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_36"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spins = synthetic.representative_cluster_states(op, n, 5, dev)
psi = synthetic.synthetic_amplitudes(int(spins.shape[0]), 5, device=dev)


def timed(fn, reps=3):
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        out = fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps, out


t_apply, (other, coeffs, counts) = timed(lambda: op.batched_apply_device(spins))
t_build, csr = timed(lambda: common.build_csr_from_candidates_device(spins, psi, 0, other, coeffs, counts, max_row_len=op.max_candidates))
t_all, _ = timed(lambda: common.extract_csr_device(op, spins, psi))
print("n=%d candidates=%d couplings=%d: batched_apply %.2f ms, build + canonicalise %.2f ms, whole extract_csr_device %.2f ms" % (
    spins.shape[0], other.shape[0], csr[1].numel(), t_apply, t_build, t_all))
