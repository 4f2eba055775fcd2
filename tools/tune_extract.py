#!/usr/bin/env python
"""Sweep the debug knobs of the single-pass extraction kernel on the bench workload and print
kernel milliseconds (CUDA events, median of 5 after 2 warm-ups).  Development tool, not a bench.

    python tools/tune_extract.py [--states N] [--combos "f,t,a,slots;..."]
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
from annealing_sign_problem_b200._lib import ffi, lib  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--states", type=int, default=10_000_000)
    p.add_argument("--system", default="heisenberg_kagome_36")
    p.add_argument("--combos", default="0,0,0,0;0,0,1,0;-2,0,0,0;-1,0,0,0;1,0,0,0;0,-1,0,0;0,1,0,0;0,0,0,256;0,0,0,512;0,0,0,1024")
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    cfg = asp.ls.load_config(asp.ls.system_path(args.system))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    spins = synthetic.cluster_closed_states(op, args.states, 1000, dev)
    n = int(spins.shape[0])
    psi = synthetic.synthetic_amplitudes(n, 77, device=dev)
    ref = None
    for combo in args.combos.split(";"):
        f, t, a, slots = (int(v) for v in combo.split(","))
        lib().asp_debug_set_extract_tuning(f, t, a)
        lib().asp_debug_set_hit_list_capacity(slots)
        need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n, n))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        cap = max(8 * n, min(n * op.max_candidates, 400_000_000))
        indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(cap, dtype=torch.int32, device=dev)
        data = torch.empty(cap, dtype=torch.float64, device=dev)
        nnz = ffi.new("uint64_t *")
        times = []
        for it in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            common.check(lib().asp_extract_csr(op.handle, n, common.ptr(spins, "uint64_t *"), common.ptr(psi, "double *"), 0, n,
                                               common.ptr(ws, "void *"), ws.numel(), cap, common.ptr(indptr, "int64_t *"),
                                               common.ptr(indices, "int32_t *"), common.ptr(data, "double *"), nnz, common.stream()))
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1))
        m = int(nnz[0])
        sig = (m, int(indices[:m].sum(dtype=torch.int64)), float(data[:m].sum()))
        if ref is None:
            ref = sig
        print("filter%+d table%+d stageA=%d slots=%d: %.3f ms (min %.3f)  nnz=%d ws=%.0f MB %s" % (
            f, t, a, slots, float(np.median(times)), min(times), m, need / 1e6, "" if sig == ref else "MISMATCH"), flush=True)
        del ws, indices, data


if __name__ == "__main__":
    main()
