import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import annealing_sign_problem_b200 as asp
from annealing_sign_problem_b200 import common, synthetic
dev = torch.device("cuda", 0)
op = asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_36"))
spins = synthetic.cluster_closed_states(op, 200000, 5, dev)
psi = synthetic.synthetic_amplitudes(spins.shape[0], 5, device=dev)
torch.cuda.synchronize()
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print("  %-28s %.4f s" % (label, time.perf_counter() - t0), flush=True); return out
for it in range(3):
    print("iteration", it)
    o = T("batched_apply_device", lambda: op.batched_apply_device(spins))
    r = T("build_csr_from_candidates", lambda: common.build_csr_from_candidates_device(spins, psi, 0, o[0], o[1], o[2], max_row_len=op.max_candidates))
    T("extract_csr_device (both)", lambda: common.extract_csr_device(op, spins, psi))
