"""Small single-pass extraction against the two-pass path (debug helper: run under compute-sanitizer)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import annealing_sign_problem_b200 as asp
from annealing_sign_problem_b200 import common, synthetic

system = sys.argv[1] if len(sys.argv) > 1 else "j1j2_square_4x4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda")
cfg = asp.ls.load_config(asp.ls.system_path(system))
cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
spins = synthetic.cluster_closed_states(op, n, 0, dev)
psi = synthetic.synthetic_amplitudes(spins.shape[0], 0, device=dev)
ref = common.extract_csr_two_pass_device(op, spins, psi)
torch.cuda.synchronize()
print("two-pass ok", ref[1].numel(), flush=True)
got = common.extract_csr_device(op, spins, psi)
torch.cuda.synchronize()
print("single-pass ok", got[1].numel(), flush=True)
import numpy as np

indptr_ref = ref[0].cpu().numpy()
for name, a, b in zip(("indptr", "indices", "data"), got, ref):
    same = torch.equal(a, b)
    print(name, "equal" if same else "DIFFERENT", flush=True)
    if not same:
        bad = torch.nonzero(a != b).flatten().cpu().numpy()
        print("  %d mismatches, first at" % bad.shape[0], bad[:8].tolist(), a[bad[:8]].tolist(), b[bad[:8]].tolist())
        rows = bad if name == "indptr" else np.unique(np.searchsorted(indptr_ref, bad, side="right") - 1)
        print("  rows affected: %d; first rows (row, tile, lane):" % rows.shape[0], [(int(r), int(r) // 32, int(r) % 32) for r in rows[:24]])
