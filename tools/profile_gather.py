"""Single-GPU stand-in for the X1 kernels under ncu (ncu cannot follow a multi-rank job): four row blocks of
10^7 states on ONE device, pulled + indexed by asp_gather_index in each of its modes."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from annealing_sign_problem_b200._lib import ffi, lib  # noqa: E402

dev = torch.device("cuda", 0)
cfg = asp.ls.load_config(asp.ls.system_path("heisenberg_kagome_36"))
cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
world, rows = 4, 10_000_000
n = world * rows
spins = synthetic.random_sector_states(36, 18, n, 5, dev)
psi = synthetic.synthetic_amplitudes(n, 5, device=dev)
bounds = [rows * q for q in range(world + 1)]
parts_s = [spins[bounds[q]:bounds[q + 1]].clone() for q in range(world)]
parts_p = [psi[bounds[q]:bounds[q + 1]].clone() for q in range(world)]
need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n, rows))
ws = torch.empty(need, dtype=torch.uint8, device=dev)
full_s, full_p = torch.empty_like(spins), torch.empty_like(psi)
for mode in (2, 1, 0, 2, 1, 0):
    lib().asp_set_gather_mode(mode)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    common.check(lib().asp_gather_index(op.handle, world, 1, ffi.new("uint64_t[]", bounds),
                                        ffi.new("uint64_t const *[]", [common.ptr(t, "uint64_t const *") for t in parts_s]),
                                        ffi.new("double const *[]", [common.ptr(t, "double const *") for t in parts_p]), ffi.NULL, 0,
                                        common.ptr(full_s, "uint64_t *"), common.ptr(full_p, "double *"), rows, common.ptr(ws, "void *"), need,
                                        common.stream()))
    e[1].record()
    torch.cuda.synchronize()
    assert torch.equal(full_s, spins) and torch.equal(full_p, psi)
    print("mode %d: %.3f ms (all blocks local: %.0f GB/s read + written)" % (mode, e[0].elapsed_time(e[1]), 2 * n * 16 / e[0].elapsed_time(e[1]) / 1e6))
